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

    def __init__(self, medium, attenuation_model=None, log_level=logging.NOTSET,
                 n_frequencies_integration=None, n_reflections=None, config=None,
                 detector=None, ray_tracing_2D_kwards={}, use_cpp=None):
        self.__logger = logging.getLogger('NuRadioMC.SignalProp.ray_tracing_base')
        self.__logger.setLevel(log_level)
        self._medium = medium
        self._config = config
        self._set_arguments(n_frequencies_integration, n_reflections, attenuation_model)

        self._detector = detector
        self._max_detector_frequency = None
        if self._detector is not None:
            # largest Nyquist frequency over the stations, taken from each station's first channel (:64-80)
            for station_id in self._detector.get_station_ids():
                channel_ids = self._detector.get_channel_ids(station_id)
                sampling_frequency = self._detector.get_sampling_frequency(station_id, channel_ids[0])
                for channel_id in channel_ids:
                    if self._detector.get_sampling_frequency(station_id, channel_id) != sampling_frequency:
                        self.__logger.warning(
                            f"Channels of station {station_id} have different sampling frequencies; using the one of "
                            f"channel {channel_ids[0]} ({sampling_frequency / units.GHz:.1f} GHz) for the attenuation grid.")
                if self._max_detector_frequency is None or sampling_frequency * .5 > self._max_detector_frequency:
                    self._max_detector_frequency = sampling_frequency * .5

        self._X1 = None
        self._X2 = None
        self._results = None

    def _set_arguments(self, n_frequencies_integration, n_reflections, attenuation_model):
        """config wins over keyword arguments (with a warning), defaults are 100 / 0 / 'SP1' (:86-133)"""
        self._n_frequencies_integration = None
        self._n_reflections = None
        self._attenuation_model = None
        if self._config is not None:
            prop = self._config['propagation']
            if 'n_freq' in prop:
                if n_frequencies_integration is not None:
                    self.__logger.warning(f"Overriding n_frequencies_integration from config file from "
                                          f"{n_frequencies_integration} to {prop['n_freq']}")
                self._n_frequencies_integration = prop['n_freq']
            if 'n_reflections' in prop:
                if n_reflections is not None:
                    self.__logger.warning(f"Overriding n_reflections from config file from {n_reflections} to "
                                          f"{prop['n_reflections']}")
                self._n_reflections = prop['n_reflections']
            if 'attenuation_model' in prop:
                if attenuation_model is not None:
                    self.__logger.warning(f"Overriding attenuation_model from config file from {attenuation_model} to "
                                          f"{prop['attenuation_model']}")
                self._attenuation_model = prop['attenuation_model']
        if self._n_frequencies_integration is None:
            self._n_frequencies_integration = n_frequencies_integration or 100
        if self._n_reflections is None:
            self._n_reflections = n_reflections or 0
        if self._attenuation_model is None:
            self._attenuation_model = attenuation_model or 'SP1'
        if self._n_reflections:
            if not hasattr(self._medium, "reflection") or self._medium.reflection is None:
                self.__logger.warning("Ray paths with bottom reflections requested but medium does not have any "
                                      "reflective layer, setting number of reflections to zero.")
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
