"""
Interface every propagation module implements -- mirrors the reference's `ray_tracing_base`
(NuRadioMC/SignalProp/propagation_base_class.py:9-446): same constructor arguments, the same precedence
config > keyword > default for (n_freq, n_reflections, attenuation_model) (:86-133), the same derivation of the maximum
detector frequency (:64-80) and the same method set.
"""
import logging

import numpy as np

from nuradiomc_b200.utilities import units

logger = logging.getLogger('NuRadioMC.SignalProp.ray_tracing_base')


class ray_tracing_base:

    # (keyword argument, key in config['propagation'], default) of the three settings a configuration file may override (:86-133)
    _SETTINGS = (("n_frequencies_integration", "n_freq", 100), ("n_reflections", "n_reflections", 0), ("attenuation_model", "attenuation_model", "SP1"))

    def __init__(self, medium, attenuation_model=None, log_level=logging.NOTSET,
                 n_frequencies_integration=None, n_reflections=None, config=None,
                 detector=None, ray_tracing_2D_kwards={}, use_cpp=None):
        self.__logger = logging.getLogger('NuRadioMC.SignalProp.ray_tracing_base')
        self.__logger.setLevel(log_level)
        self._medium, self._config, self._detector = medium, config, detector
        self._set_arguments(n_frequencies_integration, n_reflections, attenuation_model)
        self._max_detector_frequency = self._nyquist_of(detector)
        self.reset_solutions()

    def _nyquist_of(self, detector):
        """largest Nyquist frequency over the stations, each station represented by its first channel (:64-80); None without detector"""
        if detector is None:
            return None
        best = None
        for sid in detector.get_station_ids():
            channels = detector.get_channel_ids(sid)
            rate = detector.get_sampling_frequency(sid, channels[0])
            if any(detector.get_sampling_frequency(sid, c) != rate for c in channels):
                self.__logger.warning(f"station {sid}: channels differ in sampling rate, the attenuation grid follows channel "
                                      f"{channels[0]} ({rate / units.GHz:.1f} GHz)")
            best = 0.5 * rate if best is None else max(best, 0.5 * rate)
        return best

    def _set_arguments(self, n_frequencies_integration, n_reflections, attenuation_model):
        """a value in config['propagation'] wins over the keyword argument (with a warning); defaults 100 / 0 / 'SP1' (:86-133)"""
        given = {"n_frequencies_integration": n_frequencies_integration, "n_reflections": n_reflections, "attenuation_model": attenuation_model}
        from_file = self._config['propagation'] if self._config is not None else {}
        for name, key, default in self._SETTINGS:
            value = given[name]
            if key in from_file:
                if value is not None:
                    self.__logger.warning(f"{name}: the configuration file ({from_file[key]}) overrides the argument ({value})")
                value = from_file[key]
            setattr(self, "_" + name, value or default)
        if self._n_reflections and getattr(self._medium, "reflection", None) is None:
            self.__logger.warning("bottom reflections requested for a medium without reflective layer: n_reflections set to 0")
            self._n_reflections = 0

    def reset_solutions(self):
        self._X1 = None
        self._X2 = None
        self._results = None

    def set_start_and_end_point(self, x1, x2):
        self.reset_solutions()
        self._X1 = np.array(x1, dtype=float)
        self._X2 = np.array(x2, dtype=float)
        if self._n_reflections:
            if self._X1[2] < self._medium.reflection or self._X2[2] < self._medium.reflection:
                msg = "start or stop point is below the reflective bottom layer at {:.1f}m".format(
                    self._medium.reflection / units.m)
                self.__logger.error(msg)
                raise AttributeError(msg)

    def use_optional_function(self, function_name, *args, **kwargs):
        if hasattr(self, function_name):
            getattr(self, function_name)(*args, **kwargs)

    def _undefined(self):
        self.__logger.error('function not defined')
        raise NotImplementedError

    def find_solutions(self):
        self._undefined()

    def has_solution(self):
        return len(self._results) > 0

    def get_number_of_solutions(self):
        return len(self._results)

    def get_results(self):
        return self._results

    def get_solution_type(self, iS):
        self._undefined()

    def get_path(self, iS, n_points=1000):
        self._undefined()

    def get_launch_vector(self, iS):
        self._undefined()

    def get_receive_vector(self, iS):
        self._undefined()

    def get_reflection_angle(self, iS):
        self._undefined()

    def get_path_length(self, iS, analytic=True):
        self._undefined()

    def get_travel_time(self, iS, analytic=True):
        self._undefined()

    def get_attenuation(self, iS, frequency, max_detector_freq=None):
        self._undefined()

    def apply_propagation_effects(self, efield, i_solution):
        self._undefined()

    def get_output_parameters(self):
        self._undefined()

    def get_raytracing_output(self, i_solution):
        self._undefined()

    def get_number_of_raytracing_solutions(self):
        """maximum number of solutions between two points (:424-429)"""
        return 2 + 4 * self._n_reflections

    def get_config(self):
        return self._config

    def set_config(self, config):
        self._config = config
