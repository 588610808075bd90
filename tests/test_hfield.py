"""Convex geom against a heightfield (MuJoCo mjc_ConvexHField; groundwork for SURVEY.md 8(f).1): the fp64 oracle
(`orc_convex_hfield`) against closed forms, the device code (`fb_hfield.h`, through the host-emulation hook) against the oracle."""
import ctypes as C

import numpy as np
import pytest

import __graft_entry__ as ge
from oracle import fly_oracle as fo
from test_oracle_invariants import SPH, CAP, ELL, CYL, _rot

dp = lambda a: np.ascontiguousarray(a, np.float64).ctypes.data_as(C.POINTER(C.c_double))
fp = lambda a: np.ascontiguousarray(a, np.float32).ctypes.data_as(C.POINTER(C.c_float))
EYE = np.eye(3)


def _small_rot(ax, ay, az):
    cx, sx, cy, sy, cz, sz = np.cos(ax), np.sin(ax), np.cos(ay), np.sin(ay), np.cos(az), np.sin(az)
    Rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]]); Ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]]); Rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]])
    return Rz @ Ry @ Rx


@pytest.fixture(scope='module')
def libs():
    ge.build()
    o = C.CDLL(fo.build())
    o.orc_convex_hfield.argtypes = [C.c_int] + [C.POINTER(C.c_double)] * 3 + [C.c_double] + [C.POINTER(C.c_double)] * 3 + [C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_int]
    e = C.CDLL(ge.EMU)
    e.fb_emu_convex_hfield.argtypes = [C.c_int] + [C.POINTER(C.c_float)] * 3 + [C.c_float] + [C.POINTER(C.c_float)] * 3 + [C.c_int, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_int]
    return o, e


def oracle_hf(o, t, p, R, s, margin, hp, hR, hsize, data, kmax=50):
    out = np.zeros((kmax, 7))
    n = o.orc_convex_hfield(t, dp(p), dp(np.asarray(R).reshape(9)), dp(s), margin, dp(hp), dp(np.asarray(hR).reshape(9)), dp(hsize), data.shape[0], data.shape[1],
                            dp(data), dp(out), kmax)
    return out[:n]


def device_hf(e, t, p, R, s, margin, hp, hR, hsize, data, kmax=50):
    out = np.zeros((kmax, 7), np.float32)
    n = e.fb_emu_convex_hfield(t, fp(p), fp(np.asarray(R).reshape(9)), fp(s), margin, fp(hp), fp(np.asarray(hR).reshape(9)), fp(hsize), data.shape[0], data.shape[1],
                               fp(data), out.ctypes.data_as(C.POINTER(C.c_float)), kmax)
    return out[:n].astype(np.float64)


HSIZE = np.array([2.0, 2.0, 1.0, 0.05])            # 4 x 4 arena, elevation scale 1, base 0.05
GRID = 41                                           # 0.1 spacing, as the flight arenas (hills.py:172-176)


def test_oracle_sphere_on_flat_and_tilted_terrain(libs):
    o, _ = libs
    flat = np.full((GRID, GRID), 0.3)
    r = 0.04
    for z, hit in ((0.3 + r - 0.004, True), (0.3 + r + 0.003, False)):
        c = oracle_hf(o, SPH, [0.13, -0.21, z], EYE, [r, 0, 0], 0.0, [0, 0, 0], EYE, HSIZE, flat)
        assert (len(c) > 0) == hit
        if hit:
            k = np.argmin(c[:, 0])
            assert abs(c[k, 0] - (z - r - 0.3)) < 2e-6 and np.allclose(c[k, 4:], [0, 0, 1], atol=1e-4)       # the plane-sphere answer
            assert np.all(c[:, 0] >= c[k, 0] - 1e-9) and np.all(c[:, 6] > 0.5)
    # margin: reported from margin away, dist stays the geometric distance
    c = oracle_hf(o, SPH, [0.13, -0.21, 0.3 + r + 0.003], EYE, [r, 0, 0], 0.01, [0, 0, 0], EYE, HSIZE, flat)
    assert len(c) > 0 and abs(c[:, 0].min() - 0.003) < 2e-6
    # a plane of slope b along x: distance (z - H(x)) cos(theta) - r, normal (-sin, 0, cos)
    b = 0.25
    xs = np.linspace(-2, 2, GRID)
    ramp = np.tile(0.5 + b * xs / 1.0, (GRID, 1)) / 1.0
    th = np.arctan(b)
    x0, z0 = 0.31, 0.5 + b * 0.31 + 0.035 / np.cos(th)                       # 0.005 cos(theta) deep
    c = oracle_hf(o, SPH, [x0, 0.17, z0], EYE, [r, 0, 0], 0.0, [0, 0, 0], EYE, HSIZE, ramp)
    k = np.argmin(c[:, 0])
    assert abs(c[k, 0] - ((z0 - (0.5 + b * x0)) * np.cos(th) - r)) < 5e-6
    assert np.allclose(c[k, 4:], [-np.sin(th), 0, np.cos(th)], atol=2e-3)
    # the heightfield's own frame: the same scene rotated and shifted as a whole gives the same distances
    R = _rot(np.random.RandomState(1))
    t = np.array([0.4, -0.7, 0.2])
    c2 = oracle_hf(o, SPH, R @ np.array([x0, 0.17, z0]) + t, R, [r, 0, 0], 0.0, t, R, HSIZE, ramp)
    assert len(c2) == len(c) and np.allclose(np.sort(c2[:, 0]), np.sort(c[:, 0]), atol=1e-6)
    k2 = np.argmin(c2[:, 0])
    assert np.allclose(c2[k2, 4:], R @ c[k, 4:], atol=1e-5) and np.allclose(c2[k2, 1:4], R @ c[k, 1:4] + t, atol=1e-5)
    # outside the footprint or above the tallest point: nothing
    assert len(oracle_hf(o, SPH, [2.2, 0, 0.3], EYE, [r, 0, 0], 0.0, [0, 0, 0], EYE, HSIZE, flat)) == 0
    assert len(oracle_hf(o, SPH, [0, 0, 1.2], EYE, [r, 0, 0], 0.0, [0, 0, 0], EYE, HSIZE, flat)) == 0


def test_device_matches_oracle_on_a_bumpy_terrain(libs):
    o, e = libs
    rs = np.random.RandomState(3)
    xs = np.linspace(-2, 2, GRID)
    terr = 0.3 + 0.15 * np.sin(2.1 * xs)[None, :] * np.cos(1.7 * xs)[:, None] + 0.02 * rs.uniform(size=(GRID, GRID))
    hp, hR = np.array([0.2, -0.1, -0.01]), _small_rot(0.15, -0.1, 0.4)

    def height(x, y):
        fx, fy = (x + 2) / 0.1, (y + 2) / 0.1
        ix, iy = int(np.clip(np.floor(fx), 0, GRID - 2)), int(np.clip(np.floor(fy), 0, GRID - 2))
        tx, ty = fx - ix, fy - iy
        return (terr[iy, ix] * (1 - tx) + terr[iy, ix + 1] * tx) * (1 - ty) + (terr[iy + 1, ix] * (1 - tx) + terr[iy + 1, ix + 1] * tx) * ty
    n_hit = n_cmp = n_close = 0
    for trial in range(120):
        t = [SPH, CAP, ELL, CYL][trial % 4]
        s = {SPH: [rs.uniform(0.02, 0.06), 0, 0], CAP: [rs.uniform(0.01, 0.03), rs.uniform(0.02, 0.08), 0],
             ELL: list(rs.uniform(0.015, 0.07, 3)), CYL: [rs.uniform(0.01, 0.04), rs.uniform(0.01, 0.05), 0]}[t]
        x, y = rs.uniform(-1.7, 1.7, 2)
        # centre a little above / into the surface: shallow contacts and near misses
        local = np.array([x, y, height(x, y) + max(s) * rs.uniform(0.2, 1.1)])
        R = _rot(rs)
        p, Rg = hR @ local + hp, hR @ R
        a = oracle_hf(o, t, p, Rg, s, 0.0, hp, hR, HSIZE, terr)
        b = device_hf(e, t, p, Rg, s, 0.0, hp, hR, HSIZE, terr)
        n_hit += len(a) > 0
        if len(a) == 0 and len(b) == 0:
            continue
        n_cmp += 1
        # fp32 MPR against fp64 MPR: the deepest contact agrees; grazing prisms at the edge of the footprint may differ
        if len(a) and len(b) and abs(a[:, 0].min() - b[:, 0].min()) < 5e-5 and abs(len(a) - len(b)) <= 1:
            ka, kb = np.argmin(a[:, 0]), np.argmin(b[:, 0])
            if np.allclose(a[ka, 4:], b[kb, 4:], atol=2e-2) and np.allclose(a[ka, 1:4], b[kb, 1:4], atol=2e-3):
                n_close += 1
    print("hfield parity: oracle hits", n_hit, "compared", n_cmp, "close", n_close)
    assert n_hit > 40 and n_cmp > 40
    assert n_close >= 0.95 * n_cmp, (n_close, n_cmp)


def test_vision_model_terrain_contacts_match_the_oracle_through_the_step():
    check_vision_terrain_contacts(ge.EMU)


def check_vision_terrain_contacts(lib):
    """the compiled `vision` variant (flight model, no ghost, ground contacts on, 401 x 401 terrain geom): stage parity of the whole
    forward pass against the oracle with the fly dipped into a bumpy terrain -- heightfield contacts, their rows and forces."""
    import __graft_entry__ as ge
    from flybody_b200 import arenas, stepper as st
    from flybody_b200.flymodel import load_model
    from parity_common import rel_err
    ge.build()
    m = load_model('vision')
    assert (m.nq, m.nv, m.nu) == (43, 42, 11) and m.meta['hf_nrow'] == 401 and len(m.hf_pair_geom) == 70
    terr = arenas.SineBumps().generate(np.random.RandomState(4)).astype(np.float32)
    rs = np.random.RandomState(0)
    for trial, dz in enumerate((0.13, 0.17, 0.6)):
        x, y = (-5.0, 0.0) if trial < 2 else (3.0, -2.0)
        ground = float(arenas.hfield_height(terr, [x], [y], 20.0)[0]) - 0.01
        q = m.qpos0.copy(); q[:3] = [x, y, ground + dz]
        v = np.zeros(m.nv); v[0] = 20.0
        sim = st.BatchedStepper(m, 2, lib_path=lib)
        sim.hfield_collision(m.meta['hf_geom'], m.hf_size, 401, 401, m.hf_pair_geom)
        sim.hfield_write(np.arange(2), np.stack([terr, terr]))
        o = fo.Oracle(m, tolerance=1e-12)
        o.set_hfield(m.meta['hf_geom'], m.hf_size, terr, m.hf_pair_geom)
        ctrl = rs.uniform(-0.2, 0.2, m.nu)
        o.reset(q, v); sim.reset(np.tile(q, (2, 1)), np.tile(v, (2, 1)))
        o.set(fo.CTRL, ctrl); sim.set_control(np.tile(ctrl, (2, 1)).astype(np.float32))
        o.forward(); sim.forward()
        ncon_o, ncon_d = int(o.get(fo.NCON)[0]), int(sim.get(st.NCON)[1, 0])
        co = o.get(fo.CONTACT).reshape(-1, 16)[:ncon_o]
        terrain_contacts = int((co[:, 7] == m.meta['hf_geom']).sum())
        assert (terrain_contacts > 0) == (dz < 0.3), (trial, terrain_contacts)
        assert ncon_o == ncon_d and int(o.get(fo.NEFC)[0]) == int(sim.get(st.NEFC)[1, 0])
        cd = sim.get(st.CONTACT)[1].reshape(-1, 16)[:ncon_d].astype(np.float64)
        assert np.array_equal(co[:, 7:9], cd[:, 7:9])                                   # same geom pairs in the same order
        assert np.abs(co[:, 0] - cd[:, 0]).max() < 5e-5                                 # distances
        for f in (fo.EFC_FORCE, fo.QFRC_CONSTRAINT, fo.QACC):
            a = o.get(f)
            if np.abs(a).max() > 0:
                assert rel_err(a, sim.get(f)[1]) < 2e-3, (trial, f)
        sim.close()


def test_vision_model_stage_parity_in_free_flight():
    check_vision_free_flight(ge.EMU)


def check_vision_free_flight(lib):
    """the `vision` variant away from the ground: every stage of the forward pass against the oracle (as the flight variant)."""
    import __graft_entry__ as ge
    from flybody_b200 import stepper as st
    from flybody_b200.flymodel import load_model
    from parity_common import compare_stage_fields
    ge.build()
    m = load_model('vision')
    res = compare_stage_fields(m, st.BatchedStepper(m, 2, lib_path=lib), seed=1, vel_scale=20.0)
    assert 'qM' in res and 'qfrc_passive' in res
