"""Golden vectors produced by the reference's own pure-Python helpers (tests/golden/make_python_goldens.py ran the
reference modules in the build container) against: the batched wing-beat pattern generator, the CoM<->root helpers and
quaternion utilities of the host side, and the oracle's ellipsoid fluid force."""
import ctypes as C
import os

import numpy as np

from flybody_b200 import fly_envs as fe

G = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'python_goldens.npz'))


def test_wbpg_matches_reference_sequences():
    """`WingBeatPatternGenerator.reset/step` (tasks/pattern_generators.py:131-203), 4 generators, 300 steps incl. a
    sweep to both ends of the frequency range."""
    w = fe.BatchedWingBeatPatternGenerator(4)
    q, v = w.reset(np.arange(4), G['wbpg_phases'])
    assert np.allclose(q, G['wbpg_reset_qpos'], atol=1e-12)
    assert np.allclose(v, G['wbpg_reset_qvel'], atol=1e-9)
    acts = G['wbpg_actions']
    for t in range(acts.shape[0]):
        out = w.step(w.base_beat_freq * (1 + w.rel_freq_range * acts[t]))
        assert np.allclose(out, G['wbpg_steps'][:, t], atol=1e-12), t


def test_wbpg_partial_reset_and_inactive_envs():
    w = fe.BatchedWingBeatPatternGenerator(4)
    w.reset(np.arange(4), G['wbpg_phases'])
    f = np.full(4, 218.0)
    a = w.step(f)
    pos = w.pos.copy()
    b = w.step(f, active=np.array([True, False, True, False]))
    assert np.array_equal(w.pos[[1, 3]], pos[[1, 3]]) and np.allclose(b[[1, 3]], a[[1, 3]])
    q, _ = w.reset(np.array([2]), np.array([0.3]))
    assert np.allclose(q[0], G['wbpg_reset_qpos'][1])


def test_com_root_and_quaternion_helpers():
    q, com = G['quat'], G['com']
    assert np.allclose(fe.com2root(com, q), G['com2root'], atol=1e-14)
    assert np.allclose(fe.root2com(np.concatenate([com, q], 1)), G['root2com'], atol=1e-14)
    assert np.allclose(fe.rotate_vec_with_quat(com, q), G['rotate_vec'], atol=1e-14)
    assert np.allclose(fe.mult_quat(fe.reciprocal_quat(q), G['quat2']), G['dquat_local'], atol=1e-14)
    assert np.allclose(fe.quat_dist_short_arc(q, G['quat2']), G['quat_dist_short_arc'], atol=1e-12)


def test_oracle_ellipsoid_fluid_force_matches_reference_python():
    """oracle K7 local wrench vs flybody/ellipsoid_fluid_model.py:88-209 evaluated by the reference itself."""
    from oracle import fly_oracle as fo
    lib = fo.lib()
    lib.orc_ellipsoid_local_force.argtypes = [C.POINTER(C.c_double)] * 4 + [C.c_double, C.c_double, C.POINTER(C.c_double)]
    lib.orc_ellipsoid_local_force.restype = None
    c5 = G['fluid_coefs']                          # blunt, slender, angular, kutta, magnus
    coef = np.concatenate([[1.0], c5, G['fluid_vmass'], G['fluid_vinert']])
    size = np.ascontiguousarray(G['fluid_size'])
    rho, eta = G['fluid_density_viscosity']
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    for lv, want in zip(G['fluid_local_vels'], G['fluid_local_force']):
        w, v, out = np.ascontiguousarray(lv[:3]), np.ascontiguousarray(lv[3:]), np.zeros(6)
        lib.orc_ellipsoid_local_force(dp(w), dp(v), dp(size), dp(coef), rho, eta, dp(out))
        assert np.allclose(out, want, rtol=1e-12, atol=1e-18), (lv, out, want)
