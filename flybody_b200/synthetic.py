"""Constants, quaternion helpers and the synthetic reference trajectories shared by the task code
(reference `flybody/tasks/constants.py`, `flybody/quaternions.py`, `flybody/tasks/synthetic_trajectories.py`,
`flybody/tasks/task_utils.py`)."""
import numpy as np

_WALK_CONTROL_TIMESTEP = 2e-3      # reference tasks/constants.py:10-13
_WALK_PHYSICS_TIMESTEP = 2e-4
_TERMINAL_LINVEL = 50.0
_TERMINAL_ANGVEL = 200.0
_FLY_CONTROL_TIMESTEP = 2e-4       # tasks/constants.py:16-19
_FLY_PHYSICS_TIMESTEP = 5e-5
_TERMINAL_HEIGHT = 0.2
_TERMINAL_QACC = 1e14              # tasks/constants.py:21
_ACTION_CLASS_ORDER = ('adhesion', 'head', 'mouth', 'antennae', 'wings', 'abdomen', 'legs', 'user')  # fruitfly.py:25-32


# --- quaternion helpers on [..., 4] arrays (reference flybody/quaternions.py:13-76) -----------
def mult_quat(a, b):
    aw, ax, ay, az = np.moveaxis(a, -1, 0)
    bw, bx, by, bz = np.moveaxis(b, -1, 0)
    return np.stack([aw * bw - ax * bx - ay * by - az * bz, aw * bx + ax * bw + ay * bz - az * by,
                     aw * by - ax * bz + ay * bw + az * bx, aw * bz + ax * by - ay * bx + az * bw], -1)


def reciprocal_quat(q):
    return q * np.array([1.0, -1, -1, -1]) / np.sum(q * q, -1, keepdims=True)


def quat_dist_short_arc(q1, q2):
    """geodesic angle between unit quaternions, in [0, pi) (reference `quaternions.py:285-307`)."""
    q1 = q1 / np.linalg.norm(q1, axis=-1, keepdims=True)
    q2 = q2 / np.linalg.norm(q2, axis=-1, keepdims=True)
    return np.arccos(np.minimum(1.0, 2 * np.sum(q1 * q2, -1) ** 2 - 1))


def linear_tolerance(x, margin):
    """`dm_control.utils.rewards.tolerance(x, bounds=(0, 0), sigmoid='linear', margin, value_at_margin=0)`."""
    return np.clip(1.0 - np.abs(x) / margin, 0.0, 1.0)


def constant_speed_trajectory(n_steps, speed, yaw_speed=0.0, init_pos=(0, 0, 0.1278), init_heading=0.0,
                              body_rot_angle_y=0.0, body_rot_angle_x=0.0, control_timestep=0.002):
    """reference `tasks/synthetic_trajectories.py:10-70` (mju_quat2Vel restated: `quaternions.py:358-382`)."""
    qpos = np.zeros((n_steps, 7))
    qvel = np.zeros((n_steps, 6))
    qpos[0, :3] = init_pos
    qpos[:, 2] = init_pos[2]
    ya, xa = np.deg2rad(body_rot_angle_y), np.deg2rad(body_rot_angle_x)
    qpos[0, 3:] = [np.cos(ya / 2), 0.0, np.sin(ya / 2), 0.0]
    qpos[0, 3:] = mult_quat(np.array([np.cos(xa / 2), np.sin(xa / 2), 0.0, 0]), qpos[0, 3:])
    dq = np.array([np.cos(init_heading / 2), 0, 0, np.sin(init_heading / 2)])
    qpos[0, 3:] = mult_quat(dq, qpos[0, 3:])
    qvel[0, :2] = speed * np.array([np.cos(init_heading), np.sin(init_heading)])
    dtheta = yaw_speed * control_timestep
    dq = np.array([np.cos(dtheta / 2), 0, 0, np.sin(dtheta / 2)])
    axis = dq[1:]
    sin_a_2 = np.linalg.norm(axis)
    vel = np.zeros(3)
    if sin_a_2 > 0:
        speed_ang = 2 * np.arctan2(sin_a_2, dq[0])
        if speed_ang > np.pi:
            speed_ang -= 2 * np.pi
        vel = axis / sin_a_2 * speed_ang / 1.0
    qvel[:, 3:] = vel
    M = np.array([[np.cos(dtheta), -np.sin(dtheta)], [np.sin(dtheta), np.cos(dtheta)]])
    for i in range(1, n_steps):
        qvel[i, :2] = M @ qvel[i - 1, :2]
        qpos[i, :2] = qpos[i - 1, :2] + qvel[i, :2] * control_timestep
        qpos[i, 3:] = mult_quat(dq, qpos[i - 1, 3:])
    return qpos, qvel


def rotate_vec_with_quat(v, q):
    """v rotated by unit quaternion(s) q (reference `quaternions.py` rotate_vec_with_quat); broadcasts."""
    w, u = q[..., :1], q[..., 1:]
    t = 2.0 * np.cross(u, v)
    return v + w * t + np.cross(u, t)


_COM_OFFSET = np.array([-0.03697732, 0.00029205, -0.0142447])      # tasks/task_utils.py:237,259 (thorax frame)


def com2root(com, quat):
    """root-joint position from the CoM position (reference `tasks/task_utils.py:243-262`)."""
    return com + rotate_vec_with_quat(-_COM_OFFSET, quat)


def root2com(root_qpos):
    """inverse of com2root (reference `tasks/task_utils.py:223-240`)."""
    return root_qpos[..., :3] + rotate_vec_with_quat(_COM_OFFSET, root_qpos[..., 3:7])
