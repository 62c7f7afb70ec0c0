"""
TEST INFRASTRUCTURE -- numpy restatement of the reference's propagation effects, the step after the ray trace:
ray_tracing.apply_propagation_effects (NuRadioMC/SignalProp/analyticraytracing.py:2937-3033, in-ice branch, birefringence
off; the focusing factor is an input, oracle/focusing.py) and the Fresnel reflection coefficients (NuRadioReco/utilities/geometryUtilities.py:211-263).
Pinned by tests/golden/propagation_effects.npz (the reference's own functions, run by tests/golden/make_golden.py effects).
Only tests/ may import this module.
"""
import numpy as np


def fresnel_r_p(zenith, n_2, n_1):
    """geometryUtilities.py:211-236 (scimath sqrt: complex beyond total internal reflection)"""
    n = n_2 / n_1
    sq = np.emath.sqrt(n ** 2 - np.sin(zenith) ** 2)
    return np.conjugate((n ** 2 * np.cos(zenith) - sq) / (n ** 2 * np.cos(zenith) + sq))


def fresnel_r_s(zenith, n_2, n_1):
    """geometryUtilities.py:239-263"""
    n = n_2 / n_1
    sq = np.emath.sqrt(n ** 2 - np.sin(zenith) ** 2)
    return np.conjugate((np.cos(zenith) - sq) / (np.cos(zenith) + sq))


def apply_propagation_effects(spec, attenuation, reflection_angles, n_bottom_reflections, n_surface,
                              reflection_coefficient=None, reflection_phase_shift=None, focusing=None):
    """
    spec: (3, F) complex (eR, eTheta, ePhi); attenuation: (F,) or None; reflection_angles: per path segment, NaN = None.
    Returns (spec, r_theta, r_phi) as the reference leaves them (analyticraytracing.py:2963-3010).
    """
    spec = np.array(spec, complex)
    if attenuation is not None:
        spec *= attenuation                                              # :2964-2965
    r_theta = r_phi = 1.0 + 0j
    for a in np.atleast_1d(reflection_angles):                           # :2967-2999
        if a is None or np.isnan(a):
            continue
        rt, rp = fresnel_r_p(a, 1.0, n_surface), fresnel_r_s(a, 1.0, n_surface)
        spec[1] *= rt
        spec[2] *= rp
        r_theta, r_phi = rt, rp                                          # :2993-2994 overwrite: the field keeps the LAST reflection's
    k = int(n_bottom_reflections)
    if k > 0:                                                            # :3000-3010
        c = reflection_coefficient ** k * np.exp(1j * ((k * reflection_phase_shift) % (2 * np.pi)))
        spec[1] *= c
        spec[2] *= c
    if focusing is not None:                                             # :3012-3015
        spec[1:] *= focusing
    return spec, r_theta, r_phi
