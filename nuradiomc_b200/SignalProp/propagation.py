"""
Registry of propagation modules -- same names and return convention as the reference
(NuRadioMC/SignalProp/propagation.py:3-56).  Only the analytic ray tracer is in scope.
"""
solution_types = {
    1: 'direct',
    2: 'refracted',
    3: 'reflected'
}

solution_types_revert = {v: k for k, v in solution_types.items()}
available_modules = [
    'analytic',
    'radiopropa',
    'direct_ray'
]

reflection_case = {
    1: 'upwards launch vector',
    2: 'downward launch vector'
}


def get_propagation_module(name=None):
    """returns the python class of the respective propagation module (propagation.py:21-56)"""
    if name is None:
        from nuradiomc_b200.SignalProp.propagation_base_class import ray_tracing_base
        return ray_tracing_base
    elif name == available_modules[0]:
        from nuradiomc_b200.SignalProp.analyticraytracing import ray_tracing
        return ray_tracing
    elif name in available_modules:
        raise NotImplementedError(f"Module '{name}' is outside the scope of nuradiomc_b200 (analytic ray tracer only); "
                                  "use the reference implementation for it.")
    else:
        msg = "Module \'{}\' not implemented. Available modules: {}".format(name, str(available_modules))
        raise NotImplementedError(msg)
