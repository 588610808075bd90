"""Device-side task logic (fb_task_*): auto-reset, ghost placement, wing-beat pattern generator,
termination, reward and discount evaluated on the device, held step by step against the host-side task code of
`BatchedFlyEnv` ON THE SAME STEPPER (the physics is chaotic in the last bits of the solver, so two steppers cannot be compared
over an episode; one stepper with both logics can).  Host-emulation build; tests/test_gpu_parity.py repeats it on the B200."""
import numpy as np
import pytest

import __graft_entry__ as ge
from flybody_b200 import fly_envs, stepper as st
from flybody_b200.dm_env_shim import StepType


@pytest.fixture(scope='module')
def emu():
    ge.build()
    return ge.EMU


def drive(env, actions, wb_ref=None):
    """step a device-task env; after every step re-derive what the host-side task code would have produced from the same
    observation record and the same counters, and compare.  Returns the step types seen."""
    n = env.n_envs
    sim = env._sim
    ts = env.reset()
    assert np.all(np.asarray(ts.step_type) == int(StepType.FIRST)) and np.all(ts.discount == 1) and np.all(ts.reward == 0)
    seen = [np.asarray(ts.step_type, np.int64)]
    time = np.zeros(n)
    wing_q = env._rec[:, env._obs_slices['walker/joints_pos']][:, env._wing_in_obs].astype(np.float64) if wb_ref is not None else None
    for k in range(len(actions)):
        resetting = env._needs_reset.copy()
        step = np.where(resetting, 0, np.minimum(np.round(time / env._control_timestep).astype(np.int64), env._ref_len - 1))
        ghost = env._ref_at(step); ghost[:, :3] += env.task._ghost_offset; ghost = ghost.astype(np.float32); ghost[resetting, 7:] = 0
        rs_state = env._rs.get_state()
        ts = env.step(actions[k])
        # before_step: ghost on the reference row of the step, wing commands from the pattern generator
        q, v = sim.get(st.QPOS), sim.get(st.QVEL)
        # (the ghost is a free body without gravity or contacts: over the control step it coasts with the velocity it was given,
        # which takes it to the next row of a self-consistent reference; a held env stays on row 0)
        nxt = env._ref_at(np.minimum(step + 1, env._ref_len - 1)); nxt[:, :3] += env.task._ghost_offset
        want = np.where(resetting[:, None], ghost[:, :7], nxt[:, :7])
        assert np.allclose(q[:, env._ghost_q:env._ghost_q + 3], want[:, :3], atol=2e-4), (k, q[:, env._ghost_q:env._ghost_q + 3], want[:, :3])
        # (the synthetic reference's angular velocity is per control step, synthetic_trajectories.py:58-60: the heading barely coasts)
        assert np.allclose(q[:, env._ghost_q + 3:env._ghost_q + 7], ghost[:, 3:7], atol=5e-4), k
        assert np.allclose(v[:, env._ghost_v:env._ghost_v + 6], ghost[:, 7:], atol=1e-2), k
        if wb_ref is not None:
            if resetting.any():                                     # same phases as the device consumed
                probe = np.random.RandomState(); probe.set_state(rs_state)
                ids = np.nonzero(resetting)[0]
                wq, _ = wb_ref.reset(ids, probe.uniform(size=len(ids)))
                wing_q[ids] = wq
            a = np.nan_to_num(np.asarray(actions[k], np.float64), nan=0.0)
            target = wb_ref.step(wb_ref.base_beat_freq * (1 + wb_ref.rel_freq_range * a[:, -1]), active=~resetting)
            wi = env._action_indices['wings']
            want = a[:, wi] + np.where(resetting[:, None], 0.0, target - wing_q)
            ctrl = sim.get(st.CTRL)[:, env._ctrl_of_action[wi]]
            rng = env.model.actuator_ctrlrange[env._ctrl_of_action[wi]]
            live = ~resetting
            assert np.allclose(ctrl[live], want[live], atol=2e-5), (k, np.abs(ctrl[live] - want[live]).max())
            wing_q = env._rec[:, env._obs_slices['walker/joints_pos']][:, env._wing_in_obs].astype(np.float64)
        # after the step: termination / reward / discount
        time = np.where(resetting, 0.0, time + env._control_timestep)
        obs = env._observation(env._rec)
        step_type, reward, discount, needs = env._task_after(env._rec, obs, resetting, time, ghost)
        assert np.array_equal(np.asarray(step_type, np.int64), np.asarray(ts.step_type, np.int64)), (k, step_type, ts.step_type)
        assert np.allclose(reward, ts.reward, atol=2e-6), (k, reward, ts.reward)
        assert np.array_equal(discount, ts.discount), k
        assert np.array_equal(needs, env._needs_reset), k
        assert np.allclose(time, env._time)
        seen.append(np.asarray(ts.step_type, np.int64))
    return np.array(seen)


def test_walk_device_task_matches_host_task_code(emu):
    n = 4
    rs = np.random.RandomState(0)
    actions = rs.uniform(-0.5, 0.5, (40, n, 59)).astype(np.float32)
    actions[3, 1, 5] = np.nan                                              # NaN actions act as 0 (tasks/base.py:199)
    actions[20:, 2] = 3.0                                                  # env 2 thrashes
    env = fly_envs.walk_imitation(terminal_com_dist=0.05, n_envs=n, lib_path=emu, device_task=True)   # ghost leaves 0.05 cm after ~13 steps
    # a turning reference: every component of the ghost's pose and velocity changes from row to row
    from flybody_b200.synthetic import constant_speed_trajectory
    env.task._traj_generator.set_next_trajectory(*constant_speed_trajectory(n_steps=300, speed=2, yaw_speed=4.0, init_pos=(0, 0, 0.1278),
                                                                            control_timestep=2e-3))
    seen = drive(env, actions)
    assert (seen == int(StepType.LAST)).any() and (seen[1:] == int(StepType.FIRST)).any()
    # the NaN action reached the actuators as 0
    assert np.all(np.isfinite(env._sim.get(st.CTRL)))
    # after an auto-reset the env is back on the start pose
    first = np.nonzero(seen[1:] == int(StepType.FIRST))
    assert len(first[0]) > 0
    env.close()


def test_walk_device_task_episode_end_is_a_good_termination(emu):
    from flybody_b200.synthetic import constant_speed_trajectory
    q, v = constant_speed_trajectory(n_steps=64 + 6, speed=0.0, init_pos=(0, 0, 0.1278), control_timestep=2e-3)
    env = fly_envs.walk_imitation(terminal_com_dist=float('inf'), n_envs=2, lib_path=emu, device_task=True)
    env.task._traj_generator.set_next_trajectory(q, v)
    seen = drive(env, np.zeros((8, 2, 59), np.float32))
    assert np.all(seen[5] == int(StepType.LAST))                          # step == episode_steps = 70 - 64 - 1
    assert np.all(seen[6] == int(StepType.FIRST))
    env.close()


def test_switching_the_shared_trajectory_restarts_every_env(emu):
    """shared-reference mode + device task: a trajectory set between episodes is uploaded when the next env resets; the device then
    restarts all envs and the host's counters follow; an in-place edit of the loader's arrays counts as a change (content digest)"""
    from flybody_b200.synthetic import constant_speed_trajectory
    env = fly_envs.walk_imitation(terminal_com_dist=float('inf'), n_envs=3, lib_path=emu, device_task=True)
    env.reset()
    for k in range(3):
        ts = env.step(np.zeros((3, 59), np.float32))
    assert np.all(np.asarray(ts.step_type) == int(StepType.MID)) and np.all(env._step_counter == 3)
    q, v = constant_speed_trajectory(n_steps=200, speed=1.0, init_pos=(0, 0, 0.1278), control_timestep=2e-3)
    env.task._traj_generator.set_next_trajectory(q, v)
    env.request_reset([1])
    ts = env.step(np.zeros((3, 59), np.float32))
    assert np.all(np.asarray(ts.step_type) == int(StepType.FIRST)) and np.all(env._step_counter == 0) and np.all(env._time == 0)
    first_digest = env._program_ref_id
    ts = env.step(np.zeros((3, 59), np.float32))
    assert np.all(np.asarray(ts.step_type) == int(StepType.MID))
    q[:, 0] += 0.01                                                        # same arrays, edited in place
    env.task._traj_generator.set_next_trajectory(q, v)
    env.request_reset([0])
    ts = env.step(np.zeros((3, 59), np.float32))
    assert env._program_ref_id != first_digest and np.all(np.asarray(ts.step_type) == int(StepType.FIRST))
    env.close()


def test_walk_device_task_reset_noise_is_bounded_and_varies(emu):
    env = fly_envs.walk_imitation(terminal_com_dist=float('inf'), n_envs=8, lib_path=emu, device_task=True, reset_noise=0.05, seed=3)
    env.reset()
    q = env._sim.get(st.QPOS)[:, env._leg_act_qadr].astype(np.float64) - env.model.qpos0[env._leg_act_qadr]
    assert np.all(np.abs(q) <= 0.05 + 1e-6) and np.std(q) > 0.02           # U(-0.05, 0.05) from the device's counter hash
    assert np.abs(q[0] - q[1]).max() > 1e-3
    env.close()


def test_flight_device_task_matches_host_task_code(emu):
    n = 3
    rs = np.random.RandomState(1)
    actions = rs.uniform(-0.2, 0.2, (60, n, 12)).astype(np.float32)
    actions[:, :, -1] = rs.uniform(-1, 1, (60, n))                         # beat-frequency action sweeps the pattern tables
    actions[25:, 0, :] = 1.0                                               # env 0 leaves the reference -> termination -> auto-reset
    env = fly_envs.flight_imitation(n_envs=n, lib_path=emu, seed=4, terminal_com_dist=0.02, device_task=True)
    wb_ref = fly_envs.BatchedWingBeatPatternGenerator(n)
    # the reset() of drive() consumes the first n phases
    probe = np.random.RandomState(4)
    wb_ref.reset(np.arange(n), probe.uniform(size=n))
    seen = drive(env, actions, wb_ref=wb_ref)
    assert (seen[1:] == int(StepType.FIRST)).any()
    env.close()


def test_device_task_requires_shared_reference(emu, tmp_path):
    import test_rewards_loaders as trl
    path = str(tmp_path / 'w.npz')
    trl._write_walking_dataset(path, np.random.RandomState(0), nj=0, ns=0)
    with pytest.raises(NotImplementedError):
        fly_envs.walk_imitation(ref_path=path, n_envs=2, lib_path=emu, device_task=True)


def test_request_reset_restarts_single_envs(emu):
    """fb_task_request_reset: the listed envs restart at the next step (FIRST, action dropped), the others carry on; the device's
    episode counters (fb_task_episodes) count it."""
    env = fly_envs.walk_imitation(terminal_com_dist=float('inf'), n_envs=4, lib_path=emu, device_task=True)
    env.reset()
    rs = np.random.RandomState(0)
    for _ in range(3):
        env.step(rs.uniform(-0.5, 0.5, (4, 59)).astype(np.float32))
    ep0 = env._sim.task_episodes().copy()
    env.request_reset([1, 3])
    ts = env.step(rs.uniform(-0.5, 0.5, (4, 59)).astype(np.float32))
    assert list(np.asarray(ts.step_type)) == [int(StepType.MID), int(StepType.FIRST), int(StepType.MID), int(StepType.FIRST)]
    assert list(env._sim.task_episodes() - ep0) == [0, 1, 0, 1]
    q = env._sim.get(st.QPOS)
    assert np.allclose(q[1, :3], env._ref_qpos[0, :3], atol=1e-6) and not np.allclose(q[0, :3], env._ref_qpos[0, :3], atol=1e-6)
    assert env.device_reset_count() == int(env._sim.task_episodes().sum())
    env.close()


def test_reset_noise_can_be_limited_to_the_first_reset(emu):
    """fb_task_set_reset_noise: noisy initial states (decorrelation), exact start pose at later resets -- what bench.py does (SURVEY.md
    8(d) config 2 asks for noise on the INITIAL state; the reference's auto-resets return to the task's start pose)."""
    env = fly_envs.walk_imitation(terminal_com_dist=float('inf'), n_envs=4, lib_path=emu, device_task=True, reset_noise=0.05, seed=3)
    env.reset()
    q = env._sim.get(st.QPOS)[:, env._leg_act_qadr].astype(np.float64) - env.model.qpos0[env._leg_act_qadr]
    assert np.std(q) > 0.02
    env.set_reset_noise(0.0)
    env.request_reset([0, 2])
    env.step(np.zeros((4, 59), np.float32))
    q = env._sim.get(st.QPOS)[:, env._leg_act_qadr].astype(np.float64) - env.model.qpos0[env._leg_act_qadr]
    assert np.abs(q[[0, 2]]).max() < 1e-6 and np.std(q[[1, 3]]) > 0.01       # reset envs: exact start pose; the others moved on from their noisy one
    env.close()
