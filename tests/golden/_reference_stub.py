"""Import hook used ONLY by the golden-vector generator (tests/golden/make_reference_golden.py), which runs in the
build container where /root/reference exists: every `tensorflow[.x.y]` import resolves to a MagicMock, so the
reference's pure Python / numpy logic (model settings, streaming post-processor, accuracy statistics, tpr/fpr) can
be imported and EXECUTED as is.  Nothing that needs TensorFlow arithmetic is called through it."""
import importlib.abc
import importlib.machinery
import sys
from unittest import mock

REFERENCE_ROOT = "/root/reference"


class _TFStubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    shim = None          # optional module whose public names become `tensorflow.<name>` (make_reference_augment_golden.py)

    def find_spec(self, fullname, path, target=None):
        if fullname == "tensorflow" or fullname.startswith("tensorflow."):
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        if spec.name == "tensorflow" and self.shim is not None:
            import types
            m = types.ModuleType("tensorflow")
            for k, v in vars(self.shim).items():
                if not k.startswith("_"):
                    setattr(m, k, v)
            m.__getattr__ = lambda name: mock.MagicMock(name="tensorflow." + name)      # anything else stays a mock
            m.__path__ = []
            m.__spec__ = spec
            m.__loader__ = self
            return m
        m = mock.MagicMock(name=spec.name)
        m.__path__ = []
        m.__name__ = spec.name
        m.__spec__ = spec
        m.__loader__ = self
        return m

    def exec_module(self, module):
        pass


def install(shim=None):
    """shim: a module of numpy-backed functions to serve as top-level `tensorflow` (everything it lacks, and every
    submodule, is still a MagicMock)."""
    _TFStubFinder.shim = shim
    if not any(isinstance(f, _TFStubFinder) for f in sys.meta_path):
        sys.meta_path.insert(0, _TFStubFinder())
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
