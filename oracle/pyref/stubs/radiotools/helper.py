"""Minimal stand-in for radiotools.helper (absent in this image); used only at import time by the reference."""
import numpy as np


def spherical_to_cartesian(zenith, azimuth):
    sz = np.sin(zenith)
    return np.array([sz * np.cos(azimuth), sz * np.sin(azimuth), np.cos(zenith)])


def cartesian_to_spherical(x, y, z):
    r = np.sqrt(x * x + y * y + z * z)
    zenith = np.arccos(z / r)
    azimuth = np.arctan2(y, x)
    return zenith, azimuth


def get_normalized_angle(angle, degree=False, interval=np.deg2rad([0, 360])):
    import math
    if degree:
        interval = np.rad2deg(interval)
    delta = interval[1] - interval[0]
    return ((angle - interval[0]) % delta) + interval[0]
