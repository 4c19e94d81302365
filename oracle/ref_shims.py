"""Import helpers for the UNMODIFIED upstream reference (TEST INFRASTRUCTURE ONLY).

Only usable in the build container where ``/root/reference`` is mounted; used by
``oracle/make_golden.py`` and by the container-only oracle-vs-reference tests.
Nothing is copied from the reference: its modules are imported in place, with
three stub modules for dependencies missing from this image.
"""
from __future__ import annotations

import os
import sys
import types

REFERENCE_ROOT = os.environ.get('RESDEPTH_REFERENCE_ROOT', '/root/reference')


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, 'lib', 'UNet.py'))


def _install_stub_modules():
    if 'easydict' not in sys.modules:
        try:
            import easydict  # noqa: F401
        except ImportError:
            m = types.ModuleType('easydict')

            class EasyDict(dict):
                def __init__(self, d=None, **kw):
                    super().__init__()
                    for k, v in dict(d or {}, **kw).items():
                        self[k] = v

                def __setitem__(self, k, v):
                    if isinstance(v, dict) and not isinstance(v, EasyDict):
                        v = EasyDict(v)
                    super().__setitem__(k, v)

                __setattr__ = __setitem__

                def __getattr__(self, k):
                    try:
                        return self[k]
                    except KeyError as e:
                        raise AttributeError(k) from e

            m.EasyDict = EasyDict
            sys.modules['easydict'] = m
    if 'torchsummary' not in sys.modules:
        try:
            import torchsummary  # noqa: F401
        except ImportError:
            m = types.ModuleType('torchsummary')
            m.summary = lambda *a, **k: None
            sys.modules['torchsummary'] = m
    if 'osgeo' not in sys.modules:
        try:
            import osgeo  # noqa: F401
        except ImportError:
            m = types.ModuleType('osgeo')
            g = types.ModuleType('osgeo.gdal')
            g.GA_ReadOnly = 0
            g.Dataset = type('Dataset', (), {})
            g.Open = lambda *a, **k: None
            m.gdal = g
            sys.modules['osgeo'] = m
            sys.modules['osgeo.gdal'] = g


def import_reference():
    """Returns the reference's ``lib`` package modules as a namespace:
    .UNet (module), .Trainer (module), .evaluation, .rasterutils, .data_normalization."""
    if not reference_available():
        raise RuntimeError(f'reference not mounted at {REFERENCE_ROOT}')
    _install_stub_modules()
    # our own package also has a sub-package called ``lib`` -- but it is only ever imported as
    # ``resdepth_b200.lib``, so the top-level name ``lib`` is free for the reference.
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import importlib
    ns = types.SimpleNamespace()
    ns.UNet = importlib.import_module('lib.UNet')
    ns.data_normalization = importlib.import_module('lib.data_normalization')
    ns.utils = importlib.import_module('lib.utils')          # must precede lib.Trainer (circular import)
    ns.Trainer = importlib.import_module('lib.Trainer')
    ns.evaluation = importlib.import_module('lib.evaluation')
    ns.rasterutils = importlib.import_module('lib.rasterutils')
    return ns
