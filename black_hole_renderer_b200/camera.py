"""Pinhole camera looking at the origin (host glue; reference: build_camera, render.py:93-127)."""
import math

import numpy as np


def build_camera(cam_pos, fov_deg, width, height):
    """Return (cam_pos, cam_right, cam_up, cam_forward, pixel_width, pixel_height), all float64.

    forward = -pos/|pos|, right = forward x z (x-axis when the camera sits on the z axis),
    up = right x forward; the image plane sits at distance 1 and spans 2 tan(fov/2) vertically.
    """
    pos = np.array(cam_pos, dtype=np.float64)
    forward = -pos / np.linalg.norm(pos)
    right = np.cross(forward, np.array([0.0, 0.0, 1.0]))
    length = np.linalg.norm(right)
    if length < 1e-6:
        right = np.array([1.0, 0.0, 0.0])
    else:
        right /= length
    up = np.cross(right, forward)
    up /= np.linalg.norm(up)
    plane_h = 2.0 * np.tan(np.radians(fov_deg) / 2)
    plane_w = plane_h * (width / height)
    return pos, right, up, forward, plane_w / width, plane_h / height


def build_camera_scalar(cam_pos, fov_deg, width, height):
    """`build_camera` in scalar float64 arithmetic (same operations in the same order, no numpy
    calls): ~3 us instead of ~100 us per frame of host time, which is on the latency path of every
    synchronous frame.  Returns plain tuples.  tests/test_host.py holds it to `build_camera`."""
    p0, p1, p2 = float(cam_pos[0]), float(cam_pos[1]), float(cam_pos[2])
    n = math.sqrt(p0 * p0 + p1 * p1 + p2 * p2)
    f0, f1, f2 = -p0 / n, -p1 / n, -p2 / n
    # forward x (0, 0, 1), written out like numpy.cross
    r0, r1, r2 = f1 * 1.0 - f2 * 0.0, f2 * 0.0 - f0 * 1.0, f0 * 0.0 - f1 * 0.0
    length = math.sqrt(r0 * r0 + r1 * r1 + r2 * r2)
    if length < 1e-6:
        r0, r1, r2 = 1.0, 0.0, 0.0
    else:
        r0, r1, r2 = r0 / length, r1 / length, r2 / length
    u0, u1, u2 = r1 * f2 - r2 * f1, r2 * f0 - r0 * f2, r0 * f1 - r1 * f0
    un = math.sqrt(u0 * u0 + u1 * u1 + u2 * u2)
    u0, u1, u2 = u0 / un, u1 / un, u2 / un
    plane_h = 2.0 * math.tan(math.radians(fov_deg) / 2)
    plane_w = plane_h * (width / height)
    return (p0, p1, p2), (r0, r1, r2), (u0, u1, u2), (f0, f1, f2), plane_w / width, plane_h / height, n
