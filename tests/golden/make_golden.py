"""Generates the committed golden fixtures (tests/golden/*.npz) by running the UNMODIFIED reference:

* the compiled reference C++ (oracle/_ref: nanoflann kNN wrapper, grid_subsampling core), and
* the reference PyTorch modules imported from /root/reference/models (oracle/ref_models.py).

Runs only in the build container (needs /root/reference):   python tests/golden/make_golden.py
The fixtures pin both the CPU oracle (tests/test_oracle_pinning.py, CPU) and the CUDA path (tests/test_*_gpu.py).
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import native as on          # noqa: E402
from oracle import ref_models, synthetic  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def knn_golden():
    d = {}
    pts = synthetic.room_cloud(2, 3000, seed=1)
    d["uni_pts"] = pts
    d["uni_k16"] = on.ref_knn_batch(pts, pts, 16, omp=True)
    d["uni_k32"] = on.ref_knn_batch(pts, pts, 32, omp=False)
    sub = np.ascontiguousarray(pts[:, ::4])
    d["uni_sub"] = sub
    d["uni_up_k1"] = on.ref_knn_batch(sub, pts, 1, omp=True)          # support = coarse, queries = fine
    q = synthetic.room_cloud(1, 777, seed=2, box=(10.0, 7.0, 4.0))[0] - 1.0   # queries partly outside the support bbox
    d["single_q"] = q
    d["single_k5"] = on.ref_knn(pts[0], q, 5, omp=True)
    lat = synthetic.lattice_cloud(10)
    d["lat_pts"] = lat
    d["lat_k16"] = on.ref_knn(lat, lat, 16)
    dup = synthetic.duplicated_cloud(2000, seed=3)
    d["dup_pts"] = dup
    d["dup_k16"] = on.ref_knn(dup, dup, 16)
    flat = synthetic.room_cloud(1, 1500, seed=4)[0]
    flat[:, 2] = 1.25                                                    # coplanar cloud
    d["flat_pts"] = flat
    d["flat_k16"] = on.ref_knn(flat, flat, 16)
    np.savez_compressed(os.path.join(OUT, "knn_golden.npz"), **d)
    print("knn_golden", {k: v.shape for k, v in d.items()})


def subsample_golden():
    rng = np.random.default_rng(5)
    d = {}
    pts = synthetic.room_cloud(1, 6000, seed=5)[0]
    feats = rng.integers(0, 256, (6000, 3)).astype(np.float32)
    cls = rng.integers(0, 13, (6000,)).astype(np.int32)
    cls2 = rng.integers(0, 4, (6000, 2)).astype(np.int32)
    d.update(pts=pts, feats=feats, cls=cls, cls2=cls2)
    for tag, dl in (("dl30", 0.30), ("dl08", 0.08)):
        p, f, c = on.ref_grid_subsample(pts, feats, cls, dl)
        d[f"{tag}_p"], d[f"{tag}_f"], d[f"{tag}_c"] = p, f, c
    d["ponly_p"] = on.ref_grid_subsample(pts, None, None, 0.2)
    p, c = on.ref_grid_subsample(pts, None, cls2, 0.5)
    d["c2_p"], d["c2_c"] = p, c
    p, f = on.ref_grid_subsample(pts - 3.7, feats, None, 0.25)         # negative coordinates
    d["neg_p"], d["neg_f"] = p, f
    np.savez_compressed(os.path.join(OUT, "subsample_golden.npz"), **d)
    print("subsample_golden", {k: v.shape for k, v in d.items()})


def _perturb(module, seed):
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in module.named_parameters():
            if name.endswith("batch_norm.weight"):
                p.copy_(1.0 + 0.2 * torch.randn(p.shape, generator=g))
            elif name.endswith("batch_norm.bias"):
                p.copy_(0.2 * torch.randn(p.shape, generator=g))
            elif name == "c" or name.endswith(".c"):
                p.copy_(torch.eye(p.shape[0]) + 0.1 * torch.randn(p.shape, generator=g))


def _run(module, inputs, grad_inputs, seed):
    """fwd+bwd in train mode with a random cotangent.  Returns dict of out, cotangent, grads, post-step buffers."""
    module.train()
    g = torch.Generator().manual_seed(seed)
    for t in grad_inputs:
        t.requires_grad_(True)
    out = module(*inputs)
    cot = torch.randn(out.shape, generator=g)
    (out * cot).sum().backward()
    d = {"out": out.detach().numpy(), "cot": cot.numpy()}
    for i, t in enumerate(grad_inputs):
        d[f"gin{i}"] = t.grad.numpy()
    for n, p in module.named_parameters():
        d["gparam." + n] = p.grad.numpy()
    for n, b in module.named_buffers():
        d["buf_after." + n] = b.detach().numpy().copy()
    return d


def layer_golden():
    ref = ref_models.load()
    knn = lambda s, q, k: on.ref_knn_batch(s, q, k, omp=True)   # noqa: E731
    d = {}
    # ---- CRF layer (B=2: the reference crashes at B=1, continuous_crf_conv_big.py:43)
    for tag, (Cu, Cp, steps, N) in {"crf_s1": (128, 64, 1, 512), "crf_s3": (64, 32, 3, 384)}.items():
        torch.manual_seed(10)
        m = ref.ContinuousGaussianCRFConv(Cu, Cp, Cp, steps=steps)
        _perturb(m, 11)
        inp = synthetic.crf_layer_inputs(2, N, 16, Cu, Cp, 4, seed=12, knn_batch_fn=knn)
        for n, v in m.state_dict().items():
            d[f"{tag}.sd.{n}"] = v.numpy().copy()
        d[f"{tag}.unary"], d[f"{tag}.pairwise"] = inp.unary.numpy().copy(), inp.pairwise.numpy().copy()
        d[f"{tag}.up_idx"], d[f"{tag}.neighbor_idx"] = inp.up_idx.numpy(), inp.neighbor_idx.numpy()
        r = _run(m, (inp.unary, inp.pairwise, inp.up_idx, inp.neighbor_idx), (inp.unary, inp.pairwise), 13)
        d.update({f"{tag}.{k}": v for k, v in r.items()})
    # ---- ResNetBBlock, plain (64→64) and strided (32→64)
    pos = synthetic.room_cloud(2, 512, seed=20)
    ms = synthetic.build_multiscale(pos, knn, num_scales=1, K=16, ratios=(4,), seed=21)[0]
    sub_pos = ms.pos[:, torch.randperm(512, generator=torch.Generator().manual_seed(21))[:128]]
    d["rb.pos"], d["rb.sub_pos"] = ms.pos.numpy(), sub_pos.numpy()
    d["rb.neighbor_idx"], d["rb.sub_idx"] = ms.neighbor_idx.numpy(), ms.sub_idx.numpy()
    torch.manual_seed(22)
    for tag, (cin, cout, strided) in {"rb_plain": (64, 64, False), "rb_strided": (32, 64, True), "rb_in6": (6, 32, False)}.items():
        m = ref.ResNetBBlock(cin, cout)
        _perturb(m, 23)
        x = torch.randn(2, 512, cin)
        for n, v in m.state_dict().items():
            d[f"{tag}.sd.{n}"] = v.numpy().copy()
        d[f"{tag}.x"] = x.numpy().copy()
        args = (x, (ms.pos, sub_pos), ms.sub_idx) if strided else (x, ms.pos, ms.neighbor_idx)
        r = _run(m, args, (x,), 24)
        d.update({f"{tag}.{k}": v for k, v in r.items()})
    # ---- Upsampling (use_crf=False decoder)
    m = ref.Upsampling(64, 32, 32)
    _perturb(m, 30)
    xd, xu = torch.randn(2, 128, 64), torch.randn(2, 512, 32)
    for n, v in m.state_dict().items():
        d[f"ups.sd.{n}"] = v.numpy().copy()
    d["ups.x_down"], d["ups.x_up"], d["ups.up_idx"] = xd.numpy().copy(), xu.numpy().copy(), ms.up_idx.numpy()
    r = _run(m, (xd, xu, ms.up_idx), (xd, xu), 31)
    d.update({f"ups.{k}": v for k, v in r.items()})
    np.savez_compressed(os.path.join(OUT, "layer_golden.npz"), **d)
    print("layer_golden", len(d), "arrays", sum(v.nbytes for v in d.values()) / 1e6, "MB")


def net_golden():
    """Full PointConvResNet(6, 13, use_crf=True, steps=1) fwd+bwd, B=2, N=4096 (levels 4096/1024/256/64/16).
    Weights come from torch.manual_seed (CPU RNG, reproducible) + _perturb; only inputs/outputs/selected grads stored."""
    ref = ref_models.load()
    knn = lambda s, q, k: on.ref_knn_batch(s, q, k, omp=True)   # noqa: E731
    pos = synthetic.room_cloud(2, 4096, seed=40)
    ms = synthetic.build_multiscale(pos, knn, num_scales=5, K=16, seed=41)
    torch.manual_seed(42)
    net = ref.PointConvResNet(6, 13, use_crf=True, steps=1)
    _perturb(net, 43)
    net.train()
    net.classifier[1].p = 0.0        # dropout off: its mask is an RNG stream, not arithmetic to be matched
    g = torch.Generator().manual_seed(44)
    x = torch.cat([torch.from_numpy(pos), torch.rand(2, 4096, 3, generator=g)], -1)
    y = torch.randint(0, 13, (2 * 4096,), generator=g)
    import types
    data = types.SimpleNamespace(x=x, multiscale=ms)
    logits = net(data)
    loss = torch.nn.functional.cross_entropy(logits, y)
    loss.backward()
    d = {"pos": pos, "x": x.numpy(), "y": y.numpy(), "logits": logits.detach().numpy(), "loss": np.float32(loss.item())}
    for n in ("conv1_1.lin_in.lin.weight", "conv3_2.point_conv.weight_nn.1.lin.weight", "conv5_2.lin_out.bn.batch_norm.weight",
              "deconv2.c", "deconv2.fusion_nn.lin.weight", "deconv4.unary_nn.0.lin.weight", "classifier.2.weight", "classifier.2.bias"):
        d["g." + n] = dict(net.named_parameters())[n].grad.numpy()
    np.savez_compressed(os.path.join(OUT, "net_golden.npz"), **d)
    print("net_golden loss", loss.item())


if __name__ == "__main__":
    which = sys.argv[1:] or ["knn", "subsample", "layer", "net"]
    torch.set_num_threads(8)
    for w in which:
        {"knn": knn_golden, "subsample": subsample_golden, "layer": layer_golden, "net": net_golden}[w]()
