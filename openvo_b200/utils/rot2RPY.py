"""Display helper kept only for import-surface compatibility (ref: src/openVO/utils/rot2RPY.py:3-38; SURVEY.md §2 C5:
out of the hot path)."""
import numpy as np


def rot2RPY(T):
    """Both roll/pitch/yaw (X/Y/Z) factorizations of the rotation block of ``T``; each result is a (2, 1) array."""
    R = np.asarray(T)[:3, :3]
    roll, pitch, yaw = np.zeros((2, 1)), np.zeros((2, 1)), np.zeros((2, 1))
    c = float(np.hypot(R[0, 0], R[1, 0]))
    if abs(c) < 1e-4:  # gimbal lock
        pitch[:] = -R[2, 0] * (np.pi / 2)
        roll[:] = R[2, 0] * np.arctan2(-R[0, 1], R[1, 1])
    else:
        for i, cc in enumerate((c, -c)):
            pitch[i] = np.arctan2(-R[2, 0], cc)
            k = np.cos(pitch[i])
            roll[i] = np.arctan2(R[2, 1] / k, R[2, 2] / k)
            yaw[i] = np.arctan2(R[1, 0] / k, R[0, 0] / k)
    return roll, pitch, yaw
