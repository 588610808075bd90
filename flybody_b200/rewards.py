"""Imitation reward of the walking task, batched over envs (reference `flybody/tasks/rewards.py:10-116`,
`flybody/quaternions.py:215-333`).  Every function takes a leading env dimension N where the reference handles one env;
with N = 1 the numbers are the reference's (tests/test_rewards.py checks this against golden values produced by the
reference's own pure-Python functions)."""
import numpy as np

from .synthetic import mult_quat, quat_dist_short_arc

# standard deviations of the four DeepMimic feature groups (reference `tasks/rewards.py:100-107`)
DEEP_MIMIC_STD = {'com': 0.078487, 'qvel': 53.7801, 'root2site': 0.0735, 'joint_quat': 1.2247}


def quat_z2vec(vec):
    """unit quaternion rotating the z axis onto `vec` [..., 3] (reference `quaternions.py:215-261`), including the
    edge cases vec = (0, 0, +-z) -> identity / half turn about x."""
    vec = np.array(vec, np.float64)
    edge = (vec[..., :2] == 0.0).all(-1)
    zsign = vec[..., 2].copy()
    vec[edge, 0] = 1.0                                        # placeholder, overwritten below
    vec = vec / np.linalg.norm(vec, axis=-1, keepdims=True)
    axis = np.stack([-vec[..., 1], vec[..., 0], np.zeros_like(vec[..., 0])], -1)
    axis /= np.linalg.norm(axis, axis=-1, keepdims=True)
    half = 0.5 * np.arccos(vec[..., 2:3])
    quat = np.concatenate([np.cos(half), np.sin(half) * axis], -1)
    quat[edge] = np.where(zsign[edge, None] < 0, np.array([0.0, 1, 0, 0]), np.array([1.0, 0, 0, 0]))
    return quat


def axis_angle_to_quat(axis, angle):
    """reference `quaternions.py:264-282`."""
    axis = axis / np.linalg.norm(axis, axis=-1, keepdims=True)
    half = 0.5 * np.asarray(angle)[..., None]
    return np.concatenate([np.cos(half), np.sin(half) * axis], -1)


def joint_orientation_quat(xaxis, qpos):
    """orientation of a hinge: rotate z onto the joint axis, then turn about the axis by the joint angle
    (reference `quaternions.py:310-333`)."""
    return mult_quat(axis_angle_to_quat(xaxis, qpos), quat_z2vec(xaxis))


def compute_diffs(walker_features, reference_features, n=2):
    """per-env sums of |difference|^n of every feature group; quaternion groups use the short-arc angle
    (reference `tasks/rewards.py:10-34`)."""
    diffs = {}
    for k, w in walker_features.items():
        r = reference_features[k]
        if 'quat' not in k:
            e = np.abs(w - r) ** n
        else:
            e = quat_dist_short_arc(w, r) ** n
        diffs[k] = e.reshape(e.shape[0], -1).sum(1)
    return diffs


def get_walker_features(root_qpos, mocap_qpos, root_qvel, mocap_qvel, root2site, axes_ego):
    """model pose features from what the observation program returns (reference `tasks/rewards.py:37-63`):
    root_qpos [N,7], mocap joint angles [N,nj] and velocities, egocentric site vectors [N,ns,3] (FB_OBS_SITES_EGO),
    joint axes in the root frame [N,nj,3] (FB_OBS_DOF_AXIS_EGO)."""
    joint_quat = joint_orientation_quat(axes_ego, mocap_qpos)
    return {'com': root_qpos[:, :3], 'qvel': np.concatenate([root_qvel, mocap_qvel], 1), 'root2site': root2site,
            'joint_quat': np.concatenate([root_qpos[:, None, 3:7], joint_quat], 1)}


def get_reference_features(snippet, step):
    """reference pose features at the per-env steps `step` [N] (reference `tasks/rewards.py:66-83`)."""
    qpos = snippet['qpos'][step]
    return {'com': qpos[:, :3], 'qvel': snippet['qvel'][step], 'root2site': snippet['root2site'][step],
            'joint_quat': np.concatenate([qpos[:, None, 3:7], snippet['joint_quat'][step]], 1)}


def reward_factors_deep_mimic(walker_features, reference_features, std=None, weights=(1, 1, 1, 1)):
    """[N, 4] un-normalised Gaussian factors for CoM position, velocities, end-effector vectors and joint orientations
    (reference `tasks/rewards.py:86-116`)."""
    std = DEEP_MIMIC_STD if std is None else std
    diffs = compute_diffs(walker_features, reference_features, n=2)
    f = np.stack([np.exp(-0.5 / std[k] ** 2 * diffs[k]) for k in walker_features], 1)
    return f * np.asarray(weights)
