"""`src` -- drop-in import path of the reference repo (Neehan/WeatherModel keeps its code under src/).

Every `src.<x>` module resolves to the SAME module object as `weathermodel_b200.<x>`, so
`python -m src.pretraining.pretraining_main`, `from src.pretraining.models.weatherformer import
WeatherFormer` and unpickling of reference checkpoints (`torch.load(..., weights_only=False)` of whole
modules saved as src.pretraining.models.*) work against the B200 implementation unchanged.
"""
import importlib
import importlib.abc
import importlib.util
import sys

_REAL = "weathermodel_b200"


class _AliasLoader(importlib.abc.Loader):
    def __init__(self, real_name):
        self.real_name = real_name

    def create_module(self, spec):
        return importlib.import_module(self.real_name)

    def exec_module(self, module):  # already executed under its real name
        pass

    # `python -m src.<module>` (runpy) asks the loader for the code object / source of the module
    def _real_loader(self):
        return importlib.util.find_spec(self.real_name).loader

    def get_code(self, fullname):
        return self._real_loader().get_code(self.real_name)

    def get_source(self, fullname):
        return self._real_loader().get_source(self.real_name)

    def get_filename(self, fullname):
        return self._real_loader().get_filename(self.real_name)

    def is_package(self, fullname):
        return self._real_loader().is_package(self.real_name)


class _AliasFinder(importlib.abc.MetaPathFinder):
    def find_spec(self, fullname, path=None, target=None):
        if not fullname.startswith("src."):
            return None
        real = _REAL + fullname[3:]
        try:
            real_spec = importlib.util.find_spec(real)
        except (ImportError, ValueError):
            return None
        if real_spec is None:
            return None
        spec = importlib.util.spec_from_loader(fullname, _AliasLoader(real), is_package=real_spec.submodule_search_locations is not None)
        return spec


if not any(isinstance(f, _AliasFinder) for f in sys.meta_path):
    sys.meta_path.insert(0, _AliasFinder())
