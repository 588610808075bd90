"""Walking-imitation reward and trajectory loaders against golden vectors produced by the reference's own
pure-Python functions (tests/golden/make_reward_goldens.py) and against a synthetic dataset in the HDF5 layout."""
import os

import numpy as np
import pytest

from flybody_b200 import rewards as rw
from flybody_b200 import trajectory_loaders as tl
from flybody_b200.synthetic import rotate_vec_with_quat, reciprocal_quat

G = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'reward_goldens.npz'))


def test_quat_z2vec_and_joint_orientation_match_reference():
    assert np.allclose(rw.quat_z2vec(G['axes']), G['quat_z2vec'], atol=1e-14)
    assert np.allclose(rw.joint_orientation_quat(G['axes'][3:], G['angles'][3:]), G['joint_orientation_quat'], atol=1e-14)


def test_egocentric_vectors_match_reference():
    ego = rotate_vec_with_quat(G['sites'] - G['root_pos'][:, None], reciprocal_quat(G['root_quat'])[:, None])
    assert np.allclose(ego, G['egocentric'], atol=1e-13)


def test_deep_mimic_factors_match_reference():
    snippet = {k[len('snippet_'):]: G[k] for k in G.files if k.startswith('snippet_')}
    wq, wv = G['walker_qpos'], G['walker_qvel']
    wf = rw.get_walker_features(wq[:, :7], wq[:, 7:], wv[:, :6], wv[:, 6:], G['walker_root2site'], G['walker_axes_ego'])
    rf = rw.get_reference_features(snippet, G['steps'])
    d = rw.compute_diffs(wf, rf)
    assert np.allclose(np.stack([d[k] for k in ('com', 'qvel', 'root2site', 'joint_quat')], 1), G['deep_mimic_diffs'], rtol=1e-12)
    f = rw.reward_factors_deep_mimic(wf, rf, weights=(20, 1, 1, 1))
    assert f.shape == (5, 4)
    assert np.allclose(f, G['deep_mimic_factors'], rtol=1e-11, atol=1e-300)


def _write_walking_dataset(path, rs, n_traj=12, nj=5, ns=3):
    d = {'timestep_seconds': np.float64(0.002), 'trajectory_lengths': np.zeros(n_traj, np.int64),
         'id2name/joints': np.array([f'j{i}' for i in range(nj)]), 'id2name/sites': np.array([f's{i}' for i in range(ns)])}
    for t in range(n_traj):
        T = 20 + 3 * t
        d['trajectory_lengths'][t] = T
        g = f'trajectories/{str(t).zfill(2)}/'
        d[g + 'root_qpos'] = rs.normal(size=(T, 7)); d[g + 'qpos'] = rs.normal(size=(T, nj))
        d[g + 'root_qvel'] = rs.normal(size=(T, 6)); d[g + 'qvel'] = rs.normal(size=(T, nj))
        d[g + 'root2site'] = rs.normal(size=(T, ns, 3)); d[g + 'joint_quat'] = rs.normal(size=(T, nj, 4))
    np.savez(path, **d)
    return d


def test_walking_loader_serves_snippets_like_the_reference(tmp_path):
    rs = np.random.RandomState(0)
    path = str(tmp_path / 'walk.npz')
    d = _write_walking_dataset(path, rs)
    ld = tl.HDF5WalkingTrajectoryLoader(path, random_state=np.random.RandomState(3))
    assert ld.num_trajectories == 12 and ld.timestep == 0.002
    assert ld.get_joint_names() == [f'j{i}' for i in range(5)] and ld.get_site_names() == ['s0', 's1', 's2']
    assert ld.trajectory_len(4) == 32
    s = ld.get_trajectory(traj_idx=4)
    g = 'trajectories/04/'
    assert s['qpos'].shape == (32, 12) and s['qvel'].shape == (32, 11)
    # x, y start above the origin; everything else is the stored data (trajectory_loaders.py:243-251)
    assert np.allclose(s['qpos'][0, :2], 0) and np.allclose(s['qpos'][:, 2:7], d[g + 'root_qpos'][:, 2:])
    assert np.allclose(s['qpos'][:, :2], d[g + 'root_qpos'][:, :2] - d[g + 'root_qpos'][0, :2])
    assert np.array_equal(s['qpos'][:, 7:], d[g + 'qpos']) and np.array_equal(s['root2site'], d[g + 'root2site'])
    s2 = ld.get_trajectory(traj_idx=4, start_step=5, end_step=9)
    assert s2['qpos'].shape[0] == 4 and np.array_equal(s2['joint_quat'], d[g + 'joint_quat'][5:9])
    # random choice follows the RandomState stream exactly as the reference would
    idx = np.random.RandomState(3).choice(np.arange(12))
    assert np.array_equal(ld.get_trajectory()['qpos'][:, 7:], d[f'trajectories/{str(idx).zfill(2)}/qpos'])
    sub = tl.HDF5WalkingTrajectoryLoader(path, traj_indices=[2, 7], random_state=np.random.RandomState(1))
    for _ in range(5):
        assert sub.get_trajectory()['qpos'].shape[0] in (26, 41)


def test_flight_loader_random_start(tmp_path):
    rs = np.random.RandomState(1)
    d = {'timestep_seconds': np.float64(2e-4)}
    for t in range(3):
        d[f'trajectories/{t}/com_qpos'] = rs.normal(size=(120, 7)); d[f'trajectories/{t}/com_qvel'] = rs.normal(size=(120, 6))
    path = str(tmp_path / 'flight.npz'); np.savez(path, **d)
    ld = tl.HDF5FlightTrajectoryLoader(path, randomize_start_step=True, random_state=np.random.RandomState(5))
    ref = np.random.RandomState(5)
    idx = ref.choice(np.arange(3)); start = ref.randint(120 - 50)
    q, v = ld.get_trajectory()
    assert q.shape[0] == 120 - start and np.allclose(q[0, :2], 0)
    assert np.array_equal(v, d[f'trajectories/{idx}/com_qvel'][start:])
    ld2 = tl.HDF5FlightTrajectoryLoader(path, randomize_start_step=False)
    q, v = ld2.get_trajectory(traj_idx=1, start_step=10, end_step=30)
    assert q.shape == (20, 7) and np.array_equal(q[:, 2:], d['trajectories/1/com_qpos'][10:30, 2:])


def test_hdf5_path_without_h5py_fails_loudly(tmp_path):
    try:
        import h5py  # noqa: F401
        pytest.skip('h5py present')
    except ImportError:
        pass
    with pytest.raises(ImportError, match='h5py'):
        tl.HDF5WalkingTrajectoryLoader(str(tmp_path / 'x.hdf5'))
