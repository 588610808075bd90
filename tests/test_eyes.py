"""Eye-camera ray caster (fb_eye_program / fb_render_eyes; groundwork for SURVEY.md 8(f).1) on the host-emulation build:
MuJoCo's camera model (pinhole, -z forward, +y up, vertical fovy, row 0 at the top) checked analytically on flat ground and
against an independent numpy ray marcher on a terrain."""
import numpy as np
import pytest

import __graft_entry__ as ge
from flybody_b200 import arenas, fly_envs, stepper as st
from flybody_b200.compiler.quat import q2mat


import inspect
GROUND_RED = inspect.signature(st.BatchedStepper.eye_program).parameters['ground'].default[0]      # red channel of the default ground colour


@pytest.fixture(scope='module')
def emu():
    ge.build()
    return ge.EMU


def camera_rays(env, e, cam, size, fovy):
    """world origin [3] and unit directions [size, size, 3] of eye `cam` of env e from the stepper's body poses."""
    sim, m = env._sim, env.model
    head = m.body_id('walker/head')
    xpos = sim.get(st.XPOS).reshape(env.n_envs, -1, 3)[e, head].astype(np.float64)
    xmat = sim.get(st.XMAT).reshape(env.n_envs, -1, 3, 3)[e, head].astype(np.float64)
    _, pos, quat = env._EYE_CAMERAS[cam]
    Rc = q2mat(np.asarray(quat, np.float64) / np.linalg.norm(quat))
    o = xpos + xmat @ np.asarray(pos)
    t = np.tan(np.deg2rad(fovy) / 2)
    j, i = np.meshgrid(np.arange(size), np.arange(size))
    u, v = (2 * (j + 0.5) / size - 1) * t, (1 - 2 * (i + 0.5) / size) * t
    d = np.stack([u, v, -np.ones_like(u)], -1); d /= np.linalg.norm(d, axis=-1, keepdims=True)
    return o, d @ Rc.T @ xmat.T


def test_flat_ground_horizon_follows_the_camera_model(emu):
    env = fly_envs.flight_imitation(n_envs=2, lib_path=emu)
    env.reset()
    env.enable_eyes(size=32, fovy=150.0)
    img = env.render_eyes()
    assert list(img) == ['walker/right_eye', 'walker/left_eye'] and img['walker/left_eye'].shape == (2, 32, 32, 3)
    for cam, name in enumerate(img):
        o, d = camera_rays(env, 0, cam, 32, 150.0)
        assert o[2] > 0.5                                                   # the fly hovers 1 cm above the floor
        with np.errstate(divide='ignore'):
            sky = (d[..., 2] >= 0) | (-o[2] / np.minimum(d[..., 2], -1e-12) > 50.0)      # beyond the far plane (zfar = 50) the ground is clipped
        px = img[name][0].astype(np.int64)
        blue_dominant = px[..., 2] > px[..., 0]                             # sky is bluish, ground brownish
        clear = np.abs(d[..., 2] + o[2] / 50.0) > 0.004                     # leave the pixels on the clipped horizon itself aside
        assert np.array_equal(blue_dominant[clear], sky[clear])
        assert sky.any() and (~sky).any()
        # ground shading: ambient + diffuse * cos(incidence) on a 1 x 1 checker of albedo 1 / 0.75
        g = ~sky & clear
        t = -o[2] / d[..., 2]; hit = o + d * t[..., None]
        tex = np.where((np.floor(hit[..., 0]) + np.floor(hit[..., 1])).astype(np.int64) & 1, 1.0, 0.75)
        want = GROUND_RED * (0.4 + 0.8 * -d[..., 2]) * tex
        near = g & (t < 50.0)
        edge = np.abs(hit[..., 0] - np.round(hit[..., 0])) < 0.02
        edge |= np.abs(hit[..., 1] - np.round(hit[..., 1])) < 0.02          # fp32 vs fp64 may disagree on the checker cell at its edges
        ok = near & ~edge
        assert ok.sum() > 100 and np.abs(px[..., 0][ok] - np.round(np.clip(want[ok], 0, 1) * 255)).max() <= 1
    # the two eyes look to opposite sides: their images differ, the two envs (same pose) agree
    assert not np.array_equal(img['walker/right_eye'][0], img['walker/left_eye'][0])
    assert np.array_equal(img['walker/right_eye'][0], img['walker/right_eye'][1])
    env.close()


def test_terrain_hits_match_an_independent_ray_marcher(emu):
    env = fly_envs.flight_imitation(n_envs=2, lib_path=emu)
    env.reset()
    dim, dens = 6, 5
    nrow, ncol = arenas.grid_shape(dim, dens)
    env.enable_eyes(size=24, fovy=150.0, terrain_shape=(nrow, ncol), half_size=float(dim), z_offset=-0.01)
    terr = arenas.SineBumps(dim=dim, grid_density=dens, wavelength_range=(2.0, 3.0), height_range=(0.6, 0.9)).generate(np.random.RandomState(2))
    env.set_terrain(np.array([1]), terr[None])                              # env 0 stays flat
    img = env.render_eyes()
    assert not np.array_equal(img['walker/left_eye'][0], img['walker/left_eye'][1])
    S, cell = float(dim), 2.0 * dim / (ncol - 1)
    axis = np.linspace(-S, S, ncol)

    def height(x, y):                                                       # bilinear, as MuJoCo triangulates finer than we need here
        fx, fy = (x + S) / cell, (y + S) / cell
        ix, iy = np.clip(np.floor(fx).astype(int), 0, ncol - 2), np.clip(np.floor(fy).astype(int), 0, nrow - 2)
        tx, ty = fx - ix, fy - iy
        return ((terr[iy, ix] * (1 - tx) + terr[iy, ix + 1] * tx) * (1 - ty) + (terr[iy + 1, ix] * (1 - tx) + terr[iy + 1, ix + 1] * tx) * ty) - 0.01
    for cam, name in enumerate(img):
        o, d = camera_rays(env, 1, cam, 24, 150.0)
        ts = np.arange(0.0, 20.0, 0.004)                                    # brute force: fine uniform marching
        p = o[None, None, None] + d[:, :, None, :] * ts[None, None, :, None]
        inside = (np.abs(p[..., 0]) <= S) & (np.abs(p[..., 1]) <= S)
        hh = height(np.clip(p[..., 0], -S, S), np.clip(p[..., 1], -S, S))
        # (the device marches in half-cell steps: a ray that only grazes a crest by less than a few hundredths may pass it)
        hit_terrain = (inside & (p[..., 2] < hh - 0.03)).any(-1)
        clear_of_terrain = ~(inside & (p[..., 2] < hh + 0.03)).any(-1)
        px = img[name][1].astype(np.int64)
        is_sky = px[..., 2] > px[..., 0]
        # a pixel whose ray goes well into the terrain must not be sky; a ray that stays clear of it and points up must be
        assert not (is_sky & hit_terrain).any()
        assert np.all(is_sky[clear_of_terrain & (d[..., 2] > 0.02)])
        assert hit_terrain.sum() > 20                                       # the bumps are in view
    env.close()


def test_eye_program_validation(emu):
    env = fly_envs.flight_imitation(n_envs=1, lib_path=emu)
    with pytest.raises(RuntimeError):
        env.render_eyes()
    with pytest.raises(st.StepperError):
        env._sim.eye_program([0, 0], [(0, 0, 0)] * 2, [(1, 0, 0, 0)] * 2)  # the world body carries no eye
    with pytest.raises(st.StepperError):
        env._sim.eye_program([2, 2], [(0, 0, 0)] * 2, [(1, 0, 0, 0)] * 2, fovy_deg=190.0)
    env.close()


@pytest.mark.parametrize('arena,tol', [('trench', 8.0), ('bumps', 12.0)])
def test_eye_statistics_match_the_vision_network_normalisation(emu, arena, tol):
    """MuJoCo's GL pixels are out of reach, the statistics its consumer assumes are not.  The reference's VisNet
    normalises the gray image with mean 77 and std 56 "from the trench task" (network_factory_vis.py:158-162); the eyes of
    `vision_guided_flight` over a few flying steps land within `tol` gray levels of both (the default palette of
    stepper.eye_program is calibrated for this), so the network's first layer sees inputs of unit scale."""
    env = fly_envs.vision_guided_flight(n_envs=8, lib_path=emu, seed=3, terrain_bank=4, bumps_or_trench=arena)
    env.reset()
    rs = np.random.RandomState(0)
    imgs = []
    for k in range(4):
        ts = env.step(rs.uniform(-0.2, 0.2, (8, 12)))
        imgs.append(np.stack([ts.observation['walker/left_eye'], ts.observation['walker/right_eye']]))
    gray = np.stack(imgs).astype(np.float64).mean(-1)
    assert abs(gray.mean() - 77.0) < tol and abs(gray.std() - 56.0) < tol, (gray.mean(), gray.std())
    env.close()
