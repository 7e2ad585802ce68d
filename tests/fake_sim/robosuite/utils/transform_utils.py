"""robosuite.utils.transform_utils, the two functions the reference's 'val'-mode loss uses (models/losses.py:4,105),
restated from robosuite v1.0's published definitions: quaternions are (x, y, z, w); quat2axisangle returns
(axis, angle) in that version, which is what the reference's `_, angle = quat2axisangle(...)` unpacks."""
import math

import numpy as np


def quat_conjugate(q):
    return np.array([-q[0], -q[1], -q[2], q[3]], dtype=np.float32)


def quat_multiply(q1, q0):
    x0, y0, z0, w0 = q0
    x1, y1, z1, w1 = q1
    return np.array([
        x1 * w0 + y1 * z0 - z1 * y0 + w1 * x0,
        -x1 * z0 + y1 * w0 + z1 * x0 + w1 * y0,
        x1 * y0 - y1 * x0 + z1 * w0 + w1 * z0,
        -x1 * x0 - y1 * y0 - z1 * z0 + w1 * w0], dtype=np.float32)


def quat_inverse(q):
    return quat_conjugate(q) / np.dot(q, q)


def quat_distance(quaternion1, quaternion0):
    return quat_multiply(quaternion1, quat_inverse(quaternion0))


def quat2axisangle(quat):
    w = float(min(max(quat[3], -1.0), 1.0))
    den = math.sqrt(1.0 - w * w)
    if math.isclose(den, 0.0):
        return np.zeros(3), 0.0
    return np.asarray(quat[:3]) / den, 2.0 * math.acos(w)
