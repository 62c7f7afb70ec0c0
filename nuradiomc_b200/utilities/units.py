"""Unit system of the reference (NuRadioReco/utilities/units.py:74,125,137): metre = ns = GHz = rad = 1."""
import math

m = meter = 1.0
cm = 1e-2 * m
km = 1e3 * m
ns = nanosecond = 1.0
s = second = 1e9 * ns
GHz = gigahertz = 1.0
MHz = megahertz = 1e-3 * GHz
Hz = hertz = 1e-9 * GHz
rad = radian = 1.0
deg = degree = math.pi / 180.0
speed_of_light = 0.299792458 * m / ns
