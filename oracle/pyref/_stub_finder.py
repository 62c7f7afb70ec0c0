"""Import hook that returns permissive empty modules for heavy optional packages the reference imports at
module import time but never uses on the ray-tracing path (astropy, matplotlib, h5py, ...)."""
import importlib.abc
import importlib.machinery
import sys
import types

_PERMISSIVE = ("astropy", "matplotlib", "h5py", "peakutils", "toml", "pymongo", "tinydb", "dash")


class _Anything(types.ModuleType):
    __path__ = []

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        mod = _Anything(self.__name__ + "." + name)
        setattr(self, name, mod)
        return mod

    def __call__(self, *a, **k):
        return self


class _Finder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path, target=None):
        if fullname.split(".")[0] in _PERMISSIVE:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        return _Anything(spec.name)

    def exec_module(self, module):
        pass


def install():
    if not any(isinstance(f, _Finder) for f in sys.meta_path):
        sys.meta_path.append(_Finder())
