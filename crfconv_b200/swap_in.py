"""Zero-edit swap-in for the reference's training script.

``trainval.py`` and the dataset classes reach the hot path through three import surfaces:

    utils/__init__.py:9    import cpp_wrappers.cpp_subsampling.grid_subsampling as cpp_subsampling
    utils/__init__.py:10   import nearest_neighbors.lib.python.nearest_neighbors as nearest_neighbors
    models/__init__.py:2   from .point_conv_big import PointConvResNet as PointConvBig      (trainval.py:61 getattr(models, cfg.model_name))

``install()`` registers modules under exactly those dotted names in ``sys.modules`` BEFORE the reference packages are imported,
so their import statements bind to this package; ``patch_models(models)`` re-points ``models.PointConvBig`` (and the three
layer modules) on an already imported ``models`` package.  Nothing in the reference tree is edited or copied.
"""
from __future__ import annotations

import sys
import types

_KNN_CHAIN = ("nearest_neighbors", "nearest_neighbors.lib", "nearest_neighbors.lib.python", "nearest_neighbors.lib.python.nearest_neighbors")
_SUB_CHAIN = ("cpp_wrappers", "cpp_wrappers.cpp_subsampling", "cpp_wrappers.cpp_subsampling.grid_subsampling")


def _register_chain(names, leaf):
    parent = None
    for i, name in enumerate(names):
        mod = leaf if i == len(names) - 1 else sys.modules.get(name)
        if mod is None:
            mod = types.ModuleType(name)
            mod.__path__ = []                      # a package, so that dotted imports below it resolve through sys.modules
        sys.modules[name] = mod
        if parent is not None:
            setattr(parent, name.rsplit(".", 1)[1], mod)
        parent = mod


def install():
    """Make the reference's C-extension import paths resolve to the sm_100a implementations.  Idempotent."""
    from . import grid_subsampling, nearest_neighbors
    _register_chain(_KNN_CHAIN, nearest_neighbors)
    _register_chain(_SUB_CHAIN, grid_subsampling)
    return nearest_neighbors, grid_subsampling


def patch_models(models_pkg):
    """Re-point an imported reference ``models`` package at the drop-in layers: ``models.PointConvBig`` and the sub-modules
    ``models.point_conv_big`` / ``models.continuous_crf_conv_big`` / ``models.common`` (constructor and forward signatures,
    sub-module names and ``state_dict`` keys are identical, so ``Base.load`` reads the reference's checkpoints)."""
    from . import common, continuous_crf_conv_big, point_conv_big
    models_pkg.PointConvBig = point_conv_big.PointConvResNet
    for name, mod in (("point_conv_big", point_conv_big), ("continuous_crf_conv_big", continuous_crf_conv_big), ("common", common)):
        setattr(models_pkg, name, mod)
        sys.modules[models_pkg.__name__ + "." + name] = mod
    return models_pkg


def uninstall():
    for name in _KNN_CHAIN + _SUB_CHAIN:
        sys.modules.pop(name, None)
