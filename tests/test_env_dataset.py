"""walk_imitation / flight_imitation with a reference *dataset* (SURVEY.md 8(f).3): per-env snippets in device slots, full-body
start pose and the DeepMimic reward.  Host-emulation build; the dataset is recorded from the stepper itself, so a walker
that replays the recorded actions must sit exactly on its reference (every Gaussian factor at its maximum)."""
import numpy as np
import pytest

import __graft_entry__ as ge
from flybody_b200 import fly_envs, rewards as rw, stepper as st
from flybody_b200.dm_env_shim import StepType
from flybody_b200.synthetic import rotate_vec_with_quat, reciprocal_quat

FUT = 64


@pytest.fixture(scope='module')
def emu():
    ge.build()
    return ge.EMU


def _record_dataset(emu, path, n_traj=3, T=FUT + 12):
    """roll the inference-mode env with random actions and store what a mocap dataset holds."""
    env = fly_envs.walk_imitation(terminal_com_dist=float('inf'), n_envs=n_traj, lib_path=emu, reset_noise=0.03, seed=3)
    m, sim = env.model, env._sim
    jn, sn = m.meta['jnt_names'], m.meta['site_names']
    joints = [n for n in m.meta['observable_joints'] if any(k in n for k in ('coxa', 'femur', 'tibia', 'tarsus'))]
    sites = [f'walker/claw_T{k}_{s}' for k in (1, 2, 3) for s in ('left', 'right')]
    jid = [jn.index(n) for n in joints]
    qadr, vadr, sid = m.jnt_qposadr[jid], m.jnt_dofadr[jid], [sn.index(n) for n in sites]
    rs = np.random.RandomState(0)
    actions = rs.uniform(-0.3, 0.3, (T, n_traj, 59)).astype(np.float32)
    env.reset()
    rec = {k: [] for k in ('root_qpos', 'qpos', 'root_qvel', 'qvel', 'root2site', 'joint_quat')}

    def grab():
        q, v = sim.get(st.QPOS).astype(np.float64), sim.get(st.QVEL).astype(np.float64)
        sx = sim.get(st.SITE_XPOS).astype(np.float64).reshape(n_traj, -1, 3)[:, sid]
        rq = q[:, env._root_q:env._root_q + 7]
        rec['root_qpos'].append(rq); rec['qpos'].append(q[:, qadr])
        rec['root_qvel'].append(v[:, env._root_v:env._root_v + 6]); rec['qvel'].append(v[:, vadr])
        rec['root2site'].append(rotate_vec_with_quat(sx - rq[:, None, :3], reciprocal_quat(rq[:, None, 3:7])))
        axes = np.stack([np.array(m.jnt_axis[j], np.float64) for j in jid])      # placeholder, replaced below
        return axes

    grab()
    for t in range(T - 1):
        env.step(actions[t])
        grab()
    d = {'timestep_seconds': np.float64(0.002), 'trajectory_lengths': np.full(n_traj, T, np.int64),
         'id2name/joints': np.array([n.split('/')[-1] for n in joints]), 'id2name/sites': np.array([n.split('/')[-1] for n in sites])}
    for k in range(n_traj):
        g = f'trajectories/{k}/'
        for key in ('root_qpos', 'qpos', 'root_qvel', 'qvel', 'root2site'):
            d[g + key] = np.stack([r[k] for r in rec[key]])
        d[g + 'joint_quat'] = np.tile([1.0, 0, 0, 0], (T, len(joints), 1))      # placeholder, filled by the second pass (needs the env's own feature code)
    np.savez(path, **d)
    env.close()
    return actions, d


def test_dataset_mode_tracks_own_snippet_and_rewards(emu, tmp_path):
    path = str(tmp_path / 'walk_ds.npz')
    actions, d = _record_dataset(emu, path)
    n_traj, T = 3, actions.shape[0]
    env = fly_envs.walk_imitation(ref_path=path, terminal_com_dist=float('inf'), n_envs=n_traj, lib_path=emu,
                                  random_state=np.random.RandomState(1))
    assert env._per_env_ref and not env._inference_mode
    # one snippet per env, in a known order
    for e in range(n_traj):
        env.task.set_next_trajectory_index(e)
        env._reset_envs(np.array([e]))
    assert env._slot_len == T and np.all(env._ref_len == T) and np.all(env._episode_steps == T - FUT - 1)
    env._sim.task_inputs(env._step_counter, np.ones(n_traj, np.uint8))
    rec = env._sim.read_task_obs(env._rec)
    obs0 = env._observation(rec)
    # root starts on its own snippet (x, y shifted to the origin by the loader), full-body pose from the snippet
    q = env._sim.get(st.QPOS)
    for e in range(n_traj):
        g = f'trajectories/{e}/'
        assert np.allclose(q[e, env._root_q + 2:env._root_q + 7], d[g + 'root_qpos'][0, 2:], atol=1e-6)
        assert np.allclose(q[e, env._mocap_qadr], d[g + 'qpos'][0], atol=1e-6)
        # future reference displacements come from the env's own slot
        rel = d[g + 'root_qpos'][:FUT + 1, :3] - d[g + 'root_qpos'][0, :3]
        ego = rotate_vec_with_quat(rel, reciprocal_quat(d[g + 'root_qpos'][0, 3:7]))
        assert np.allclose(obs0['walker/ref_displacement'][e], ego, atol=2e-5)
    # joint orientation features of the recorded poses: first pass through the env's feature code fills the dataset
    sl = env._obs_slices
    nj = len(env._mocap_qadr)
    f = rw.get_walker_features(rec[:, sl['_root_pose']].astype(np.float64), rec[:, sl['_mocap_qpos']].astype(np.float64),
                               rec[:, sl['_root_qvel']].astype(np.float64), rec[:, sl['_mocap_qvel']].astype(np.float64),
                               rec[:, sl['_mocap_sites']].astype(np.float64).reshape(n_traj, -1, 3),
                               rec[:, sl['_mocap_axes']].astype(np.float64).reshape(n_traj, nj, 3))
    assert np.allclose(np.linalg.norm(f['joint_quat'], axis=-1), 1.0, atol=1e-6)
    assert np.allclose(f['root2site'], np.stack([d[f'trajectories/{e}/root2site'][0] for e in range(n_traj)]), atol=2e-5)
    # wrong pose -> factors drop below their maxima; the maximum of the CoM factor is its weight 20
    fac = env._walk_reward_factors(rec, np.zeros(n_traj, np.int64))
    assert fac.shape == (n_traj, 10)
    assert np.allclose(fac[:, 0], 20.0, rtol=1e-6) and np.allclose(fac[:, 1], 1.0, atol=1e-6) and np.allclose(fac[:, 2], 1.0, atol=1e-5)
    assert np.all(fac[:, 3] < 1.0)            # joint_quat in the file is still a placeholder -> large angular distance
    assert np.allclose(fac[:, 4:], 1.0)       # wings at their spring reference


def test_dataset_mode_replay_keeps_maximal_reward(emu, tmp_path):
    path = str(tmp_path / 'walk_ds.npz')
    actions, d = _record_dataset(emu, path)
    n_traj, T = 3, actions.shape[0]
    # second pass: fill joint_quat by replaying the recorded states through the observation program
    env = fly_envs.walk_imitation(ref_path=path, terminal_com_dist=float('inf'), n_envs=n_traj, lib_path=emu)
    for e in range(n_traj):
        env.task.set_next_trajectory_index(e)
        env._reset_envs(np.array([e]))
    sl, nj = None, len(env._mocap_qadr)
    jq = np.zeros((T, n_traj, nj, 4))

    def features():
        env._sim.task_inputs(env._step_counter, np.zeros(n_traj, np.uint8))
        rec = env._sim.read_task_obs(env._rec)
        s = env._obs_slices
        return rw.joint_orientation_quat(rec[:, s['_mocap_axes']].astype(np.float64).reshape(n_traj, nj, 3),
                                         rec[:, s['_mocap_qpos']].astype(np.float64))
    # the dataset was recorded with reset noise on leg joints: the full-body start pose reproduces it
    jq[0] = features()
    env._needs_reset[:] = False
    for t in range(T - FUT - 1):
        env.step(actions[t])
        jq[t + 1] = features()
    jq[T - FUT:] = jq[T - FUT - 1]                            # rows past the episode end are never compared
    env.close()
    for e in range(n_traj):
        d[f'trajectories/{e}/joint_quat'] = jq[:, e]
    np.savez(path, **d)
    # third pass: the real thing -- replaying the actions keeps the four DeepMimic factors at their maxima (20, 1, 1, 1) at every step
    env = fly_envs.walk_imitation(ref_path=path, terminal_com_dist=0.3, n_envs=n_traj, lib_path=emu)
    for e in range(n_traj):
        env.task.set_next_trajectory_index(e)
        env._reset_envs(np.array([e]))
    env._needs_reset[:] = False
    for t in range(T - FUT - 2):
        ts = env.step(actions[t])
        assert np.all(ts.step_type == StepType.MID)
        fac = env._walk_reward_factors(env._rec, env._step_counter)
        assert np.allclose(fac[:, :4], [20.0, 1.0, 1.0, 1.0], rtol=1e-5), (t, fac[:, :4])     # on the reference: every Gaussian at its maximum
        assert np.all(fac[:, 4:] > 0.97) and np.all(fac[:, 4:] <= 1.0)                          # retracted wings sag a little under gravity
        assert np.allclose(ts.reward, np.prod(fac, axis=1))
        assert np.all(np.linalg.norm(ts.observation['walker/ref_displacement'][:, 0], axis=1) < 1e-4)
    ts = env.step(actions[T - FUT - 2])                       # step == episode_steps: good termination, discount 1
    assert np.all(ts.step_type == StepType.LAST) and np.all(ts.discount == 1.0) and np.all(np.isfinite(ts.reward))
    # a different policy falls off the reference: reward drops
    ts = env.step(np.zeros((n_traj, 59)))                      # auto-reset (FIRST), random snippets now
    assert np.all(ts.step_type == StepType.FIRST)
    for _ in range(5):
        ts = env.step(np.full((n_traj, 59), 0.4))
    assert np.all(ts.reward < 20.0) and np.all(ts.reward >= 0.0)


def test_flight_dataset_mode_per_env_trajectories(emu, tmp_path):
    from flybody_b200.synthetic import constant_speed_trajectory, com2root
    d = {'timestep_seconds': np.float64(2e-4)}
    speeds = (20.0, 35.0)
    for k, sp in enumerate(speeds):
        q, v = constant_speed_trajectory(n_steps=150, speed=sp, init_pos=(0.3 * k, 0, 1), body_rot_angle_y=-47.5, control_timestep=2e-4)
        d[f'trajectories/{k}/com_qpos'], d[f'trajectories/{k}/com_qvel'] = q, v
    path = str(tmp_path / 'flight_ds.npz'); np.savez(path, **d)
    env = fly_envs.flight_imitation(ref_path=path, randomize_start_step=False, n_envs=2, lib_path=emu)
    assert env._per_env_ref
    for e in range(2):
        env.task.set_next_trajectory_index(e)
        env._reset_envs(np.array([e]))
    env._needs_reset[:] = False
    assert env._slot_len == 150 and np.all(env._episode_steps == 150 - 6)
    v0 = env._sim.get(st.QVEL)
    assert np.allclose(v0[:, env._root_v], speeds, atol=1e-4)                  # root starts at the reference speed of its own trajectory
    rs = np.random.RandomState(0)
    for t in range(5):
        ts = env.step(rs.uniform(-0.2, 0.2, (2, 12)))
    assert np.all(np.isfinite(ts.reward)) and np.all(ts.reward > 0)
    # the 5-step look-ahead of each env follows its own speed: displacement between consecutive reference points
    disp = ts.observation['walker/ref_displacement']
    step_len = np.linalg.norm(disp[:, 2] - disp[:, 1], axis=1)
    assert np.allclose(step_len, np.array(speeds) * 2e-4, rtol=1e-3)
    # ghost sits on the env's own trajectory (x shifted to start at 0 by the loader)
    g = env._sim.get(st.QPOS)[:, env._ghost_q:env._ghost_q + 3]
    for e in range(2):
        com = d[f'trajectories/{e}/com_qpos'].copy(); com[:, :2] -= com[0, :2]
        root = com2root(com[:, :3], com[:, 3:7])
        assert np.allclose(g[e], root[5], atol=1e-5) or np.allclose(g[e], root[4], atol=1e-5)
