"""Pinhole camera looking at the origin (host glue; reference: build_camera, render.py:93-127)."""
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
