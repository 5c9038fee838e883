import math

import numpy as np


def axangle2mat(axis, angle, is_normalized=False):
    """Rotation matrix of `angle` radians about `axis` (Rodrigues' formula)."""
    x, y, z = [float(v) for v in axis]
    if not is_normalized:
        n = math.sqrt(x * x + y * y + z * z)
        x, y, z = x / n, y / n, z / n
    c, s = math.cos(angle), math.sin(angle)
    C = 1.0 - c
    return np.array([[x * x * C + c, x * y * C - z * s, x * z * C + y * s],
                     [y * x * C + z * s, y * y * C + c, y * z * C - x * s],
                     [z * x * C - y * s, z * y * C + x * s, z * z * C + c]])
