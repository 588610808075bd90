"""Edge cases of the batched step through the C ABI (host-emulation build): ragged env counts (the library pads to whole
blocks), extreme / non-finite actions (clamped to the actuator ctrlrange, reference fruitfly.xml:11 ctrllimited), argument
validation (negative return codes with a message, no exceptions across the ABI), partial resets."""
import ctypes as C

import numpy as np
import pytest

import __graft_entry__ as ge
from flybody_b200 import stepper as st
from flybody_b200.flymodel import load_model
from conftest import walk_reset_qpos


@pytest.fixture(scope='module')
def emu():
    ge.build()
    return ge.EMU


@pytest.mark.parametrize('n', [1, 3, 5])
def test_ragged_env_counts_are_padded_and_independent(emu, n):
    m = load_model('walk')
    rs = np.random.RandomState(n)
    q = np.tile(walk_reset_qpos(m), (5, 1)); q[:, 7:109] += rs.uniform(-0.05, 0.05, (5, 102))
    c = rs.uniform(-0.5, 0.5, (2, 5, m.nu)).astype(np.float32)
    ref = st.BatchedStepper(m, 5, lib_path=emu); ref.reset(q)
    sim = st.BatchedStepper(m, n, lib_path=emu); sim.reset(q[:n])
    assert sim.n_envs_padded % 4 == 0 and sim.n_envs_padded >= n
    for k in range(2):
        ref.set_control(c[k]); ref.step(10)
        sim.set_control(c[k, :n]); sim.step(10)
    assert sim.get(st.QPOS).shape == (n, m.nq)
    assert np.array_equal(sim.get(st.QPOS), ref.get(st.QPOS)[:n])          # an env does not depend on the batch it sits in
    assert np.array_equal(sim.get(st.SENSOR_MEAN), ref.get(st.SENSOR_MEAN)[:n])


def test_extreme_actions_are_clamped_and_the_state_stays_finite(emu):
    m = load_model('walk')
    sim = st.BatchedStepper(m, 4, lib_path=emu)
    sim.reset(np.tile(walk_reset_qpos(m), (4, 1)))
    c = np.zeros((4, m.nu), np.float32)
    c[0] = 1e6; c[1] = -1e6; c[2] = np.inf; c[3, ::2] = 1e30
    for _ in range(3):
        sim.set_control(c); sim.step(10)
    lim = m.actuator_ctrlrange
    q, v, f = sim.get(st.QPOS), sim.get(st.QVEL), sim.get(st.QFRC_ACTUATOR)
    assert np.all(np.isfinite(q)) and np.all(np.isfinite(v)) and np.all(np.isfinite(f))
    # the same trajectory as commanding the range limits directly: clamping happens before the actuator model
    ref = st.BatchedStepper(m, 4, lib_path=emu)
    ref.reset(np.tile(walk_reset_qpos(m), (4, 1)))
    cc = np.clip(np.where(np.isfinite(c), c, np.sign(c) * 1e30), lim[:, 0], lim[:, 1]).astype(np.float32)
    for _ in range(3):
        ref.set_control(cc); ref.step(10)
    assert np.array_equal(q, ref.get(st.QPOS))


def test_partial_reset_leaves_the_other_envs_alone(emu):
    m = load_model('walk')
    rs = np.random.RandomState(2)
    sim = st.BatchedStepper(m, 4, lib_path=emu); sim.reset(np.tile(walk_reset_qpos(m), (4, 1)))
    sim.set_control(rs.uniform(-0.5, 0.5, (4, m.nu)).astype(np.float32)); sim.step(10)
    before = sim.get(st.QPOS).copy()
    q1 = walk_reset_qpos(m)[None]
    sim.reset(qpos=q1, env_ids=np.array([2], np.int32))
    after = sim.get(st.QPOS)
    assert np.array_equal(after[[0, 1, 3]], before[[0, 1, 3]])
    assert np.allclose(after[2], q1[0].astype(np.float32)) and np.all(sim.get(st.QVEL)[2] == 0) and np.all(sim.get(st.ACT)[2] == 0)


def test_abi_argument_validation_returns_codes_not_crashes(emu):
    m = load_model('walk')
    sim = st.BatchedStepper(m, 2, lib_path=emu)
    lib, h = sim._lib, sim._h
    assert lib.fb_step(h, 0) < 0 and lib.fb_step(None, 1) < 0                        # n_substeps must be positive; null handle
    bad = np.array([7], np.int32); q = np.zeros((1, m.nq), np.float32)
    assert lib.fb_reset(h, bad.ctypes.data, 1, q.ctypes.data, None) < 0              # env id out of range
    assert b'range' in lib.fb_last_error(h)
    idx = np.array([m.nq + 3], np.int32); v = np.zeros((2, 1), np.float32)
    assert lib.fb_write_state(h, st.QPOS, idx.ctypes.data, 1, v.ctypes.data) < 0     # index outside the field
    assert lib.fb_write_state(h, st.XPOS, idx.ctypes.data, 1, v.ctypes.data) < 0     # not a writable field
    assert lib.fb_field_size(h, 9999) < 0
    assert lib.fb_task_step(h, v.ctypes.data, 0, 10) < 0                             # no task program uploaded
    assert lib.fb_ref_slots(h, 10) < 0                                               # no observation program uploaded
    m2 = st.BatchedStepper(m, 2, lib_path=emu)                                       # the handle is still usable after errors
    sim.set_control(np.zeros((2, m.nu), np.float32)); sim.step(1)
    assert np.all(np.isfinite(sim.get(st.QPOS)))
    m2.close()
