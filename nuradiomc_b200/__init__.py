"""
nuradiomc_b200 -- B200-native (sm_100a, FP64 CUDA) batched implementation of NuRadioMC's analytic ray tracer,
behind the reference's own propagation-module API.

    from nuradiomc_b200.SignalProp import propagation
    from nuradiomc_b200.utilities import medium
    prop = propagation.get_propagation_module('analytic')
    r = prop(medium.get_ice_model('southpole_2015'), attenuation_model='SP1')
    r.set_start_and_end_point(x1, x2); r.find_solutions(); ...          # reference's scalar API
    res = r.trace_batch(X1, X2, frequency=ff, max_detector_freq=fmax)   # new: all pairs in one device pass

The compute path is the CUDA library nuradiomc_b200/libnrmc_rt.so (C ABI: include/nrmc_rt.h).  There is no CPU
fallback: importing works without a GPU, running anything does not.
"""
__version__ = "0.1.0"
