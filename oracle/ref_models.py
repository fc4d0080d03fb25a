"""TEST INFRASTRUCTURE ONLY (oracle).  Imports the *unmodified* reference dense-API modules from /root/reference
(container only — that tree does not exist on the GPU box) so that oracle/layers.py can be pinned against them and
golden vectors generated (tests/golden/make_golden.py).

The reference's ``models/__init__.py`` imports the PyG family, which is broken as shipped (undefined ``DSPointConv``,
point_conv.py:85) and needs torch_geometric / torch_scatter / torch_cluster / torch_points3d, none of which is
installed.  So the three dense files are loaded individually under a private package name, with ``sys.modules`` stubs
for the absent third-party imports of models/common.py:4-6.  Only ``FastBatchNorm1d`` is actually *used* by the dense
path; its restatement lives in oracle/layers.py.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

REF = "/root/reference"


def available() -> bool:
    return os.path.exists(os.path.join(REF, "models", "continuous_crf_conv_big.py"))


def load():
    """Returns a namespace with the reference classes: MLP, PointConv, ResNetBBlock, Upsampling,
    ContinuousGaussianCRFConv, PointConvResNet."""
    if not available():
        raise RuntimeError("/root/reference is not mounted")
    from . import layers as _ol

    def stub(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules.setdefault(name, m)
        return sys.modules[name]

    unused = lambda *a, **k: (_ for _ in ()).throw(RuntimeError("stubbed third-party symbol"))  # noqa: E731
    stub("torch_geometric")
    stub("torch_geometric.nn", fps=unused, radius=unused, knn=unused)
    stub("torch_scatter", scatter=unused, scatter_max=unused)
    stub("torch_points3d")
    stub("torch_points3d.core")
    stub("torch_points3d.core.common_modules", FastBatchNorm1d=_ol.FastBatchNorm1d)

    pkg_name = "_crfconv_reference_models"
    if pkg_name not in sys.modules:
        pkg = types.ModuleType(pkg_name)
        pkg.__path__ = [os.path.join(REF, "models")]
        sys.modules[pkg_name] = pkg
        for mod in ("common", "continuous_crf_conv_big", "point_conv_big"):
            spec = importlib.util.spec_from_file_location(f"{pkg_name}.{mod}", os.path.join(REF, "models", mod + ".py"))
            m = importlib.util.module_from_spec(spec)
            sys.modules[f"{pkg_name}.{mod}"] = m
            spec.loader.exec_module(m)
    ns = types.SimpleNamespace()
    ns.MLP = sys.modules[f"{pkg_name}.common"].MLP
    pcb = sys.modules[f"{pkg_name}.point_conv_big"]
    ns.PointConv, ns.ResNetBBlock, ns.Upsampling, ns.PointConvResNet = pcb.PointConv, pcb.ResNetBBlock, pcb.Upsampling, pcb.PointConvResNet
    ns.ContinuousGaussianCRFConv = sys.modules[f"{pkg_name}.continuous_crf_conv_big"].ContinuousGaussianCRFConv
    return ns
