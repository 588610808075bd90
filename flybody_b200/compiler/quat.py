"""Small quaternion / rotation helpers for the model compiler (w, x, y, z order).

Conventions follow the reference's usage (Hamilton product, `flybody/quaternions.py:48-76`,
`flybody/fruitfly/fruitfly.py:34-60`)."""
import numpy as np


def qnorm(q):
    q = np.asarray(q, dtype=np.float64)
    n = np.linalg.norm(q)
    if n < 1e-15:
        return np.array([1.0, 0, 0, 0])
    return q / n


def qmul(a, b):
    aw, ax, ay, az = a
    bw, bx, by, bz = b
    return np.array([
        aw * bw - ax * bx - ay * by - az * bz,
        aw * bx + ax * bw + ay * bz - az * by,
        aw * by - ax * bz + ay * bw + az * bx,
        aw * bz + ax * by - ay * bx + az * bw,
    ])


def qconj(q):
    return np.array([q[0], -q[1], -q[2], -q[3]])


def neg_quat(q):
    """reference `fruitfly.py:34-38`: negates w only (same rotation as the conjugate)."""
    q = np.array(q, dtype=np.float64)
    q[0] *= -1
    return q


def q2mat(q):
    w, x, y, z = q
    return np.array([
        [w * w + x * x - y * y - z * z, 2 * (x * y - w * z), 2 * (x * z + w * y)],
        [2 * (x * y + w * z), w * w - x * x + y * y - z * z, 2 * (y * z - w * x)],
        [2 * (x * z - w * y), 2 * (y * z + w * x), w * w - x * x - y * y + z * z],
    ])


def qrot(q, v):
    return q2mat(q) @ np.asarray(v, dtype=np.float64)


def mat2q(R):
    """Rotation matrix -> unit quaternion (w>=0 branch chosen by largest diagonal term)."""
    R = np.asarray(R, dtype=np.float64)
    t = np.trace(R)
    if t > 0:
        s = np.sqrt(t + 1.0) * 2
        q = np.array([0.25 * s, (R[2, 1] - R[1, 2]) / s, (R[0, 2] - R[2, 0]) / s, (R[1, 0] - R[0, 1]) / s])
    elif R[0, 0] > R[1, 1] and R[0, 0] > R[2, 2]:
        s = np.sqrt(1.0 + R[0, 0] - R[1, 1] - R[2, 2]) * 2
        q = np.array([(R[2, 1] - R[1, 2]) / s, 0.25 * s, (R[0, 1] + R[1, 0]) / s, (R[0, 2] + R[2, 0]) / s])
    elif R[1, 1] > R[2, 2]:
        s = np.sqrt(1.0 + R[1, 1] - R[0, 0] - R[2, 2]) * 2
        q = np.array([(R[0, 2] - R[2, 0]) / s, (R[0, 1] + R[1, 0]) / s, 0.25 * s, (R[1, 2] + R[2, 1]) / s])
    else:
        s = np.sqrt(1.0 + R[2, 2] - R[0, 0] - R[1, 1]) * 2
        q = np.array([(R[1, 0] - R[0, 1]) / s, (R[0, 2] + R[2, 0]) / s, (R[1, 2] + R[2, 1]) / s, 0.25 * s])
    return qnorm(q)


def axisangle2q(axis, angle):
    axis = np.asarray(axis, dtype=np.float64)
    return np.hstack((np.cos(angle / 2), np.sin(angle / 2) * axis))


def z2quat(vec):
    """Quaternion rotating the +z axis onto `vec` (shortest arc)."""
    v = np.asarray(vec, dtype=np.float64)
    n = np.linalg.norm(v)
    if n < 1e-15:
        return np.array([1.0, 0, 0, 0])
    v = v / n
    z = np.array([0.0, 0, 1])
    ax = np.cross(z, v)
    s = np.linalg.norm(ax)
    if s < 1e-10:
        if v[2] > 0:
            return np.array([1.0, 0, 0, 0])
        return np.array([0.0, 1, 0, 0])
    ax = ax / s
    ang = np.arctan2(s, v[2])
    return axisangle2q(ax, ang)


def euler2q(e):
    """MJCF default eulerseq 'xyz' (intrinsic rotations about x, then y, then z)."""
    q = np.array([1.0, 0, 0, 0])
    for i, a in enumerate(e):
        ax = np.zeros(3)
        ax[i] = 1
        q = qmul(q, axisangle2q(ax, a))
    return q
