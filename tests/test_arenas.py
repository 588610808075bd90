"""Terrain generators (groundwork for SURVEY.md 8(f).1) against heightfields produced by the reference's own functions
(`flybody/tasks/arenas/hills.py`, tests/golden/make_terrain_goldens.py)."""
import os

import numpy as np

from flybody_b200 import arenas

G = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'terrain_goldens.npz'))
DIM, DENS = int(G['dim']), int(G['grid_density'])


def test_grid_shape_and_index_map():
    nrow, ncol = arenas.grid_shape(DIM, DENS)
    assert (nrow, ncol) == G['bowl'].shape
    assert arenas.grid_shape(20, 10) == (401, 401)                       # the arena of vision_guided_flight (hills.py:293-297)
    got = np.array([arenas.pos_to_terrain_idx(x, y, (DIM, DIM), nrow, ncol) for x, y in G['idx_points']])
    assert np.array_equal(got, G['idx'])


def test_bowl_bumps_trench_match_reference():
    nrow, ncol = arenas.grid_shape(DIM, DENS)
    bowl = arenas.terrain_bowl((DIM, DIM), nrow, ncol, elevation_z=4.3, random_state=np.random.RandomState(3))
    assert np.allclose(bowl, G['bowl'], atol=1e-12)
    assert np.allclose(arenas.add_sine_bumps(bowl, (DIM, DIM), wavelength=3.7, phase=0.9, height=0.8), G['bumps'], atol=1e-12)
    t, s = arenas.add_sine_trench(bowl, (DIM, DIM), wavelength=2.5, phase=1.1, amplitude=0.45, start_x=-2.0, end_x=2.5, width=1.3,
                                  height=1.3, sigma=0.2)
    assert np.allclose(t, G['trench'], atol=1e-12) and np.allclose(s, G['trench_sine'], atol=1e-14)


def test_arena_classes_draw_in_reference_order_and_are_reproducible():
    a = arenas.SineBumps(dim=DIM, grid_density=DENS)
    t1, t2 = a.generate(np.random.RandomState(5)), a.generate(np.random.RandomState(5))
    assert np.array_equal(t1, t2) and t1.shape == G['bowl'].shape and t1.min() >= 0 and 1.0 < t1.max() <= 5.0   # horizon mountains up to elevation_z
    # draw order (hills.py:442-462): elevation, bowl bumps, wavelength, phase, height
    rs = np.random.RandomState(5)
    elev = rs.uniform(4.0, 5.0)
    bowl = arenas.terrain_bowl((DIM, DIM), *arenas.grid_shape(DIM, DENS), elevation_z=elev, random_state=rs)
    want = arenas.add_sine_bumps(bowl, (DIM, DIM), wavelength=rs.uniform(10.0, 15.0), phase=rs.uniform(0.0, 2 * np.pi), height=rs.uniform(0.5, 1.0))
    assert np.array_equal(t1, want)
    tr = arenas.SineTrench(dim=DIM, grid_density=DENS, start_offset_range=(-3.0, -2.0), trench_len_range=(3.0, 4.0))
    t = tr.generate(np.random.RandomState(6))
    xs, ys = tr.trench_specs['x_coords'], tr.trench_specs['y_coords']
    assert len(xs) == len(ys) and -3.0 <= xs[0] <= -2.0 and 3.0 <= xs[-1] - xs[0] <= 4.0 and ys[0] == 0.0
    # the corridor floor is lower than the plateau next to it
    k = len(xs) // 2
    h_in = arenas.hfield_height(t, [xs[k]], [ys[k]], DIM)[0]
    h_wall = arenas.hfield_height(t, [xs[k]], [ys[k] + 1.2], DIM)[0]
    assert h_in < 0.5 and h_wall > 1.0


def test_hfield_height_nearest_grid_point():
    t = np.arange(25, dtype=np.float64).reshape(5, 5)                    # half size 2: grid points at -2, -1, 0, 1, 2
    assert arenas.hfield_height(t, [0.0, 1.4, -2.0, 0.6], [0.0, -1.6, 2.0, 0.4], 2.0).tolist() == [12.0, 3.0, 20.0, 13.0]
    batch = np.stack([t, t + 100])
    assert arenas.hfield_height(batch, [0.0, 1.4], [0.0, -1.6], 2.0).tolist() == [12.0, 103.0]
