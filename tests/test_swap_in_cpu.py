"""N2: the reference's import statements must resolve to this package without editing the reference."""
import importlib
import sys
import types

import crfconv_b200.swap_in as swap_in


def test_reference_import_paths_resolve_to_the_dropins():
    try:
        nn_mod, sub_mod = swap_in.install()
        # utils/__init__.py:9-10 of the reference, verbatim
        import cpp_wrappers.cpp_subsampling.grid_subsampling as cpp_subsampling
        import nearest_neighbors.lib.python.nearest_neighbors as nearest_neighbors
        assert nearest_neighbors is nn_mod and cpp_subsampling is sub_mod
        assert callable(nearest_neighbors.knn) and callable(nearest_neighbors.knn_batch) and callable(cpp_subsampling.compute)
        assert importlib.import_module("nearest_neighbors.lib.python").nearest_neighbors is nn_mod
        swap_in.install()                                   # idempotent
        assert sys.modules["nearest_neighbors.lib.python.nearest_neighbors"] is nn_mod
    finally:
        swap_in.uninstall()
    assert "nearest_neighbors" not in sys.modules


def test_patch_models_repoints_the_big_network():
    fake = types.ModuleType("models")                       # stands in for the reference's models package
    fake.__path__ = []
    sys.modules["models"] = fake
    try:
        swap_in.patch_models(fake)
        from crfconv_b200.point_conv_big import PointConvResNet
        assert fake.PointConvBig is PointConvResNet         # trainval.py:61 getattr(models, cfg.model_name)
        from models.continuous_crf_conv_big import ContinuousGaussianCRFConv
        import crfconv_b200.continuous_crf_conv_big as ours
        assert ContinuousGaussianCRFConv is ours.ContinuousGaussianCRFConv
        net = fake.PointConvBig(in_channels=6, n_classes=13, use_crf=True, steps=1)     # same constructor keywords as trainval.py:61-64
        assert sum(p.numel() for p in net.parameters()) == 820_141
    finally:
        for k in [k for k in sys.modules if k == "models" or k.startswith("models.")]:
            sys.modules.pop(k)


def test_train_step_signature_is_importable_without_a_gpu():
    import crfconv_b200.train_dp as t
    assert callable(t.train_step) and callable(t.make_batch) and callable(t.main)
