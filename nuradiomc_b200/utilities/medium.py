"""
Ice models accepted by the analytic ray tracer: the exponential profile n(z) = n_ice - delta_n exp(z / z_0).

Mirrors the part of the reference the hot path uses: `IceModelSimple` (NuRadioMC/utilities/medium_base.py:206-277,
`add_reflective_bottom` :47-66) and the parameter sets of NuRadioMC/utilities/medium.py:57-154, `get_ice_model` :353-371.
Radiopropa-backed and birefringent media are out of scope.
"""
import numpy as np

from nuradiomc_b200.utilities import units


class IceModel:
    def __init__(self, z_air_boundary=0 * units.meter, z_bottom=None):
        self.z_air_boundary = z_air_boundary
        self.z_bottom = z_bottom
        self.reflection = None
        self.reflection_coefficient = None
        self.reflection_phase_shift = None

    def add_reflective_bottom(self, refl_z, refl_coef, refl_phase_shift):
        self.reflection = refl_z
        self.reflection_coefficient = refl_coef
        self.reflection_phase_shift = refl_phase_shift
        if not ((self.z_bottom is not None) and (self.z_bottom < self.reflection)):
            self.z_bottom = self.reflection - 1 * units.m


class IceModelSimple(IceModel):
    def __init__(self, n_ice, delta_n, z_0, z_shift=0 * units.meter, z_air_boundary=0 * units.meter, z_bottom=None):
        super().__init__(z_air_boundary, z_bottom)
        self.n_ice = n_ice
        self.delta_n = delta_n
        self.z_0 = z_0
        self.z_shift = z_shift

    def get_index_of_refraction(self, position):
        position = np.asarray(position, dtype=float)
        if position.ndim == 1:
            if (position[2] - self.z_air_boundary) <= 0:
                return self.n_ice - self.delta_n * np.exp((position[2] - self.z_shift) / self.z_0)
            return 1
        ior = self.n_ice - self.delta_n * np.exp((position[:, 2] - self.z_shift) / self.z_0)
        ior[position[:, 2] - self.z_air_boundary > 0] = 1.
        return ior


class southpole_simple(IceModelSimple):
    def __init__(self):
        super().__init__(z_bottom=-2820 * units.meter, n_ice=1.78, z_0=71. * units.meter, delta_n=0.426)


class southpole_2015(IceModelSimple):
    def __init__(self):
        super().__init__(z_bottom=-2820 * units.meter, n_ice=1.78, z_0=77. * units.meter, delta_n=0.423)


class ARAsim_southpole(IceModelSimple):
    def __init__(self):
        super().__init__(z_bottom=-2820 * units.meter, n_ice=1.78, z_0=75.75757575757576 * units.meter, delta_n=0.43)


class ARA_2022(IceModelSimple):
    def __init__(self):
        super().__init__(z_bottom=-2820 * units.meter, n_ice=1.78, z_0=49.5049505 * units.meter, delta_n=0.454)


class mooresbay_simple(IceModelSimple):
    def __init__(self):
        super().__init__(n_ice=1.78, z_0=34.5 * units.meter, delta_n=0.46)
        self.add_reflective_bottom(refl_z=-576 * units.m, refl_coef=0.82, refl_phase_shift=180 * units.deg)


class mooresbay_simple_2(IceModelSimple):
    def __init__(self):
        super().__init__(n_ice=1.78, z_0=37 * units.meter, delta_n=0.481)
        self.add_reflective_bottom(refl_z=-576 * units.m, refl_coef=0.82, refl_phase_shift=180 * units.deg)


class greenland_simple(IceModelSimple):
    def __init__(self):
        super().__init__(z_bottom=-3000 * units.meter, n_ice=1.78, z_0=37.25 * units.meter, delta_n=0.51)


class uniform_ice(IceModelSimple):
    """uniform ice (n = 1.78); rejected by the analytic ray tracer exactly as in the reference"""

    def __init__(self, z_bottom=None):
        super().__init__(z_bottom=z_bottom, n_ice=1.78, z_0=1 * units.meter, delta_n=0)


def get_ice_model(name):
    try:
        cls = globals()[name]
    except KeyError:
        raise NotImplementedError(f"The ice model '{name}' is not implemented.")
    return cls()
