"""Headless import shim for the UNMODIFIED reference (container only).

TEST INFRASTRUCTURE — never imported by the product path (platymatch_b200/).

`/root/reference/platymatch/__init__.py:7` imports `_dock_widget`, which pulls in
SimpleITK / PyQt5 / napari / qtpy / tifffile (`_dock_widget.py:1-9`, `utils/utils.py:3`).
None of those is installed here, so we register empty stub modules for them (with
`thread_worker` and `napari_hook_implementation` as identity decorators) and then import
the reference's numeric functions exactly as shipped.  Used only by
`oracle/make_golden.py` to dump golden vectors; `/root/reference` does not exist on the
GPU box, so nothing under tests/ or bench.py may import this module at run time.
"""
import sys
import types

REFERENCE_ROOT = "/root/reference"


class _Anything:
    """Attribute sink: any attribute / call / subclassing works and does nothing."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Anything()

    def __getattr__(self, name):
        return _Anything()


def _stub(name, **attrs):
    mod = types.ModuleType(name)
    mod.__dict__.update(attrs)
    mod.__getattr__ = lambda attr: _Anything  # PEP 562: any missing name -> a class
    sys.modules[name] = mod
    return mod


def install():
    ident = lambda f=None, **k: f if f is not None else (lambda g: g)
    for name in ["SimpleITK", "tifffile", "PyQt5", "PyQt5.QtCore", "PyQt5.QtWidgets", "qtpy",
                 "qtpy.QtWidgets", "qtpy.QtCore", "napari", "napari.qt", "napari.qt.threading",
                 "napari_plugin_engine", "skimage", "skimage.filters", "skimage.feature",
                 "skimage.morphology", "tqdm", "tqdm.contrib"]:
        if name not in sys.modules:
            _stub(name)
    sys.modules["napari.qt.threading"].thread_worker = ident
    sys.modules["napari_plugin_engine"].napari_hook_implementation = ident
    sys.modules["tqdm"].tqdm = lambda it, *a, **k: it
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)


def load():
    """Return a namespace with the reference's hot-path functions, unmodified."""
    install()
    from platymatch.utils.utils import get_centroid, get_mean_distance, get_error
    from platymatch.estimate_transform.shape_context import (
        get_unary, get_unary_distance, do_ransac, get_shape_context, get_bin_index, transform, get_Y)
    from platymatch.estimate_transform.find_transform import get_affine_transform, get_similar_transform
    from platymatch.estimate_transform.apply_transform import apply_affine_transform
    from platymatch.estimate_transform.perform_icp import perform_icp
    ns = types.SimpleNamespace(**{k: v for k, v in locals().items() if k != "ns"})
    return ns
