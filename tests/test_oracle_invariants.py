"""Physics invariants that pin the CPU oracle in the absence of a live MuJoCo."""
import numpy as np
import pytest

from flybody_b200.flymodel import load_model
from oracle import fly_oracle as fo
from conftest import walk_reset_qpos


def free_model(gravity=True):
    """walk model with everything dissipative / external switched off."""
    m = load_model('walk').copy()
    a = m.a
    a['opt_density'] = 0.0
    a['opt_viscosity'] = 0.0
    a['dof_damping'] = np.zeros_like(a['dof_damping'])
    a['jnt_stiffness'] = np.zeros_like(a['jnt_stiffness'])
    a['jnt_limited'] = np.zeros_like(a['jnt_limited'])
    a['npair'] = 0
    a['actuator_gainprm'] = np.zeros_like(a['actuator_gainprm'])
    a['actuator_biasprm'] = np.zeros_like(a['actuator_biasprm'])
    if not gravity:
        a['opt_gravity'] = np.zeros(3)
    a['opt_timestep'] = 2e-5
    return m.rebuild()


def energy(m, o):
    M = o.get(fo.QM_DENSE).reshape(m.nv, m.nv)
    v = o.qvel
    com = o.get(fo.SUBTREE_COM).reshape(-1, 3)
    # ghost (last body) and walker root are the two children of world
    xipos_w = com[1]
    pe = -m.body_subtreemass[1] * np.dot(m.opt_gravity, xipos_w) \
         - m.body_subtreemass[m.nbody - 1] * np.dot(m.opt_gravity, com[m.nbody - 1])
    return 0.5 * v @ M @ v, pe


def test_mass_matrix_spd_and_matches_compiler():
    from flybody_b200.compiler.compile_model import dense_mass_matrix
    m = load_model('walk')
    o = fo.Oracle(m)
    rs = np.random.RandomState(1)
    q = walk_reset_qpos(m)
    q[7:109] += rs.uniform(-0.2, 0.2, 102)
    o.reset(q)
    M = o.get(fo.QM_DENSE).reshape(m.nv, m.nv)
    Mc, _ = dense_mass_matrix(m.a, q)
    assert np.abs(M - M.T).max() == 0
    assert np.linalg.eigvalsh(M).min() > 0
    assert np.abs(M - Mc).max() <= 1e-12 * np.abs(Mc).max()


def test_energy_conservation_free_fall_tumbling():
    m = free_model()
    o = fo.Oracle(m)
    rs = np.random.RandomState(2)
    q = walk_reset_qpos(m)
    q[2] = 5.0
    v = np.zeros(m.nv)
    v[:3] = rs.uniform(-5, 5, 3)
    v[3:6] = rs.uniform(-20, 20, 3)
    v[6:108] = rs.uniform(-30, 30, 102)
    o.reset(q, v)
    ke0, pe0 = energy(m, o)
    for _ in range(400):
        o.step()
    ke1, pe1 = energy(m, o)
    e0, e1 = ke0 + pe0, ke1 + pe1
    # semi-implicit Euler: drift is O(h); the fall exchanged a sizeable part of KE
    assert abs(e1 - e0) < 2e-3 * max(abs(ke0), abs(ke1)), (e0, e1, ke0, ke1)


def test_momentum_conservation_no_gravity():
    m = free_model(gravity=False)
    o = fo.Oracle(m)
    rs = np.random.RandomState(3)
    q = walk_reset_qpos(m)
    v = np.zeros(m.nv)
    v[:3] = [1.0, -2.0, 0.5]
    v[3:6] = rs.uniform(-10, 10, 3)
    v[6:108] = rs.uniform(-30, 30, 102)
    o.reset(q, v)
    c0 = o.get(fo.SUBTREE_COM).reshape(-1, 3)[1].copy()
    o.control_step(1)
    c1 = o.get(fo.SUBTREE_COM).reshape(-1, 3)[1].copy()
    o.control_step(300)
    c2 = o.get(fo.SUBTREE_COM).reshape(-1, 3)[1].copy()
    o.control_step(1)
    c3 = o.get(fo.SUBTREE_COM).reshape(-1, 3)[1].copy()
    vel_a, vel_b = (c1 - c0), (c3 - c2)
    assert np.abs(vel_a - vel_b).max() < 2e-5 * np.abs(vel_a).max()   # O(h^2) per step drift of semi-implicit Euler


def test_gravity_bias_matches_potential_gradient():
    m = free_model()
    o = fo.Oracle(m)
    q = walk_reset_qpos(m)
    rs = np.random.RandomState(4)
    q[7:109] += rs.uniform(-0.3, 0.3, 102)
    o.reset(q)
    bias = o.get(fo.QFRC_BIAS)

    def pe(qq):
        o.reset(qq)
        return energy(m, o)[1]
    eps = 1e-6
    for dof in (6, 20, 40, 60, 100):         # hinge dofs: qpos index = dof + 1
        qp, qm = q.copy(), q.copy()
        qp[dof + 1] += eps
        qm[dof + 1] -= eps
        g = (pe(qp) - pe(qm)) / (2 * eps)
        assert abs(g - bias[dof]) < 1e-6 * max(1e-6, np.abs(bias).max()), (dof, g, bias[dof])


def test_standing_contact_force_supports_weight():
    m = load_model('walk')
    o = fo.Oracle(m)
    o.reset(walk_reset_qpos(m))
    # hold the reset pose with the position servos (ctrl = current joint angle)
    ctrl = np.zeros(m.nu)
    for i in range(m.nu):
        if m.actuator_trntype[i] == 0:
            ctrl[i] = o.qpos[m.jnt_qposadr[m.actuator_trnid[i]]]
    o.set(fo.CTRL, ctrl)
    fz = []
    for k in range(150):
        o.control_step(10)
        con = o.get(fo.CONTACT).reshape(-1, 16)
        f = o.get(fo.EFC_FORCE)
        tot = 0.0
        for c in con:
            if c[12] >= 0 and c[7] == 0:
                tot += f[int(c[12])] * c[6]
        fz.append(tot)
    weight = m.body_subtreemass[1] * 981.0
    assert o.get(fo.FLAGS)[0] == 0
    assert abs(np.mean(fz[-50:]) - weight) < 0.05 * weight, (np.mean(fz[-50:]), weight)


def test_solver_kkt_residual_small():
    m = load_model('walk')
    o = fo.Oracle(m, tolerance=1e-12)
    o.reset(walk_reset_qpos(m))
    rs = np.random.RandomState(5)
    for k in range(20):
        o.set(fo.CTRL, rs.uniform(-0.5, 0.5, m.nu))
        o.control_step(10)
    o.forward()
    # at the solution: M (qacc - qacc_smooth) = J^T f  (before noslip this is exact; noslip keeps it by construction)
    M = o.get(fo.QM_DENSE).reshape(m.nv, m.nv)
    r = M @ o.qacc - o.get(fo.QFRC_SMOOTH) - o.get(fo.QFRC_CONSTRAINT)
    assert np.abs(r).max() < 1e-6 * max(1.0, np.abs(o.get(fo.QFRC_SMOOTH)).max())
    e = o.efc()
    f = o.get(fo.EFC_FORCE)
    # unilateral rows push only
    for i, tp in enumerate(e['type']):
        if tp in (0, 1):
            assert f[i] >= 0


# ------------------------------------------------------------------------------------------ generic convex pairs (MPR)
def _convex(t1, p1, R1, s1, t2, p2, R2, s2, margin=0.0):
    import ctypes as C
    lib = fo.lib()
    dp = lambda a: np.ascontiguousarray(a, np.float64).ctypes.data_as(C.POINTER(C.c_double))
    lib.orc_convex_pair.argtypes = [C.c_int] + [C.POINTER(C.c_double)] * 3 + [C.c_int] + [C.POINTER(C.c_double)] * 3 + [C.c_double, C.POINTER(C.c_double)]
    out = np.zeros(7)
    n = lib.orc_convex_pair(t1, dp(p1), dp(np.asarray(R1).reshape(9)), dp(s1), t2, dp(p2), dp(np.asarray(R2).reshape(9)), dp(s2), margin,
                            out.ctypes.data_as(C.POINTER(C.c_double)))
    return n, out[0], out[1:4], out[4:7]


SPH, CAP, ELL, CYL = 2, 3, 4, 5


def _rot(rs):
    q = rs.normal(size=4); q /= np.linalg.norm(q); w, x, y, z = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                     [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])


def test_mpr_round_ellipsoids_match_the_sphere_formula():
    """two ellipsoids with equal semi-axes are spheres: dist = d - r1 - r2, normal along the centres, position mid-way
    (libccd MPR as linked by MuJoCo's mjc_Convex, tolerance 1e-6)"""
    rs = np.random.RandomState(0)
    for _ in range(20):
        r1, r2 = rs.uniform(0.01, 0.05, 2)
        u = rs.normal(size=3); u /= np.linalg.norm(u)
        d = (r1 + r2) * rs.uniform(0.5, 0.98)
        p1 = rs.normal(size=3) * 0.1; p2 = p1 + u * d
        n, dist, pos, nrm = _convex(ELL, p1, _rot(rs), [r1] * 3, ELL, p2, _rot(rs), [r2] * 3)
        assert n == 1
        assert abs(dist - (d - r1 - r2)) < 5e-6
        assert np.allclose(nrm, u, atol=2e-3)
        assert np.allclose(pos, p1 + u * (r1 + (d - r1 - r2) / 2), atol=2e-3 * (r1 + r2))


def test_mpr_separated_shapes_give_no_contact_and_margin_inflates():
    rs = np.random.RandomState(1)
    R = _rot(rs)
    n, *_ = _convex(ELL, [0, 0, 0], R, [0.02, 0.03, 0.05], CYL, [0.2, 0, 0], _rot(rs), [0.02, 0.03, 0])
    assert n == 0
    # just out of touch along x, brought into range by the margin (each shape is inflated by margin / 2)
    n0, *_ = _convex(ELL, [0, 0, 0], np.eye(3), [0.02, 0.03, 0.05], ELL, [0.0405, 0, 0], np.eye(3), [0.02, 0.03, 0.05])
    n1, dist, pos, nrm = _convex(ELL, [0, 0, 0], np.eye(3), [0.02, 0.03, 0.05], ELL, [0.0405, 0, 0], np.eye(3), [0.02, 0.03, 0.05], margin=0.001)
    assert n0 == 0 and n1 == 1
    assert abs(dist - 0.0005) < 5e-6 and np.allclose(nrm, [1, 0, 0], atol=1e-3)


def test_mpr_capsule_against_round_ellipsoid_matches_the_analytic_sphere_capsule():
    rs = np.random.RandomState(2)
    for _ in range(10):
        r, cr, ch = rs.uniform(0.01, 0.03), rs.uniform(0.005, 0.02), rs.uniform(0.02, 0.06)
        Rc = _rot(rs); axis = Rc[:, 2]
        pc = rs.normal(size=3) * 0.05
        t = rs.uniform(-0.8, 0.8) * ch
        side = np.cross(axis, rs.normal(size=3)); side /= np.linalg.norm(side)
        gap = (r + cr) * rs.uniform(0.6, 0.95)
        ps = pc + axis * t + side * gap
        n, dist, pos, nrm = _convex(CAP, pc, Rc, [cr, ch, 0], ELL, ps, _rot(rs), [r] * 3)
        assert n == 1 and abs(dist - (gap - r - cr)) < 5e-6
        assert np.allclose(nrm, side, atol=1e-2)          # the portal normal is only as good as the 1e-6 support tolerance


def test_sparse_and_dense_oracle_modes_agree():
    """The oracle's default mode (MuJoCo's tree-sparse L^T D L of M and of M + hD, Newton direction through the low-rank form of
    the Hessian) against its original dense statement (dense Cholesky everywhere): same forces / accelerations on hard random
    states (30-70 constraint rows, cones in all three zones) and the same trajectory, to round-off."""
    from flybody_b200.flymodel import load_model
    from parity_common import random_state, reset_qpos
    for variant, n_sub, scale in (('walk', 10, 0.5), ('flight', 4, 0.2)):
        m = load_model(variant)
        for seed in range(3):
            q, v = random_state(m, seed, vel_scale=1.0 if variant == 'walk' else 20.0)
            out = []
            for dense in (True, False):
                o = fo.Oracle(m, tolerance=1e-12, dense=dense)
                o.reset(q, v); o.set(fo.CTRL, np.random.RandomState(seed).uniform(-scale, scale, m.nu)); o.forward()
                out.append((o.get(fo.QACC), o.get(fo.EFC_FORCE), int(o.get(fo.NEFC)[0]), int(o.get(fo.SOLVER_NITER)[0])))
            assert out[0][2] == out[1][2] and out[0][3] == out[1][3]
            assert np.abs(out[0][0] - out[1][0]).max() <= 1e-10 * np.abs(out[0][0]).max()
            assert np.abs(out[0][1] - out[1][1]).max() <= 1e-10 * (np.abs(out[0][1]).max() + 1e-30)
        traj = []
        for dense in (True, False):
            o = fo.Oracle(m, dense=dense); o.reset(reset_qpos(m)); rs = np.random.RandomState(1)
            for k in range(15):
                o.set(fo.CTRL, rs.uniform(-scale, scale, m.nu)); o.control_step(n_sub)
            traj.append((o.qpos, o.qvel))
        assert np.abs(traj[0][0] - traj[1][0]).max() < 1e-10 and np.abs(traj[0][1] - traj[1][1]).max() < 1e-8
