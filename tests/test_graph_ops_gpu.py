"""GPU parity of the kNN-family extensions and the PyG-API graph builders / edge-list layers:
K > 32 kNN, knn_batch_distance_pick, FPS, radius search, edge softmax / aggregation kernels, GuideGaussianCRFConv, DiscreteCRFConv.
Index outputs are compared bit-exactly with the CPU oracle (oracle/oracle_native.cpp); layers within 1e-3 of oracle/layers.py."""
import numpy as np
import pytest
import torch

from oracle import layers as ol
from oracle import native as on
from oracle import synthetic

pytestmark = pytest.mark.gpu
TOL = 1e-3


def _rel(a, b, floor=0.0):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / max(float(b.abs().max()), floor, 1e-30))


@pytest.mark.parametrize("N,K", [(5000, 33), (20000, 48), (3000, 64), (40, 64), (100000, 128)])
def test_knn_large_k_bit_exact(N, K):
    from crfconv_b200 import nearest_neighbors as nn_
    pos = synthetic.room_cloud(2, N, seed=N + K)
    q = pos[:, : min(N, 2000)].copy()
    got = nn_.knn_batch(pos, q, K)                                   # host-pointer C ABI
    assert np.array_equal(got, on.knn_batch(pos, q, K))
    got_dev = nn_.knn_batch(torch.from_numpy(pos).cuda(), torch.from_numpy(q).cuda(), K).cpu().numpy()
    assert np.array_equal(got_dev, got)


def test_knn_large_k_on_a_lattice_keeps_the_index_tie_rule():
    from crfconv_b200 import nearest_neighbors as nn_
    pos = synthetic.lattice_cloud(10)[None]
    assert np.array_equal(nn_.knn_batch(pos, pos, 40), on.knn_batch(pos, pos, 40))


@pytest.mark.parametrize("B,N,Q,K,seed", [(2, 2000, 300, 16, 7), (3, 513, 600, 8, 123456789), (1, 40, 60, 32, 1)])
def test_distance_pick_matches_the_seeded_oracle(B, N, Q, K, seed):
    """Same mt19937 stream, same picks, same neighbours as the restated reference loop (knn_.cxx:138-203) — bit-exact."""
    from crfconv_b200 import nearest_neighbors as nn_
    pos = synthetic.room_cloud(B, N, seed=seed % 1000)
    oi, oq = on.knn_batch_distance_pick(pos, Q, K, seed)
    gi, gq = nn_.knn_batch_distance_pick(pos, Q, K, seed=seed)       # host-pointer C ABI, like the reference's Cython wrapper
    assert np.array_equal(gq, oq) and np.array_equal(gi, oi)
    di, dq = nn_.knn_batch_distance_pick(torch.from_numpy(pos).cuda(), Q, K, seed=seed)
    assert np.array_equal(di.cpu().numpy(), oi) and np.array_equal(dq.cpu().numpy(), oq)


def test_distance_pick_coverage_invariants_hold_like_the_compiled_reference():
    """The reference seeds with time(0): its picks are not reproducible, its invariants are — every query is a support point, its
    neighbours are its exact kNN, and a point is never picked twice before every point has been covered once."""
    from crfconv_b200 import nearest_neighbors as nn_
    pos = synthetic.room_cloud(2, 1500, seed=3)
    runs = [nn_.knn_batch_distance_pick(pos, 200, 16, seed=99)]
    if on.have_ref_knn():
        runs.append(on.ref_knn_batch_distance_pick(pos, 200, 16))
    for idx, q in runs:
        assert np.array_equal(idx, on.knn_batch(pos, q, 16))
        for b in range(2):
            picked = idx[b, :, 0]                                    # a support point is its own nearest neighbour
            assert np.array_equal(pos[b][picked], q[b])
            assert len(set(picked.tolist())) == len(picked)


@pytest.mark.parametrize("B,N,S", [(3, 4096, 1024), (1, 40960, 2048), (2, 100, 100)])
def test_fps_matches_oracle(B, N, S):
    from crfconv_b200 import graph_ops
    pos = synthetic.room_cloud(B, N, seed=N)
    got = graph_ops.furthest_point_sampling(torch.from_numpy(pos).cuda(), S).cpu().numpy()
    ptr = np.arange(B + 1) * N
    exp = on.fps(pos.reshape(-1, 3), ptr, S).reshape(B, S) - ptr[:-1, None]
    assert np.array_equal(got, exp)


def test_fps_ragged_batch_vector():
    from crfconv_b200 import graph_ops
    rng = np.random.default_rng(0)
    sizes = [700, 64, 1501]
    pos = rng.random((sum(sizes), 3)).astype(np.float32)
    batch = np.repeat(np.arange(3), sizes)
    got = graph_ops.fps(torch.from_numpy(pos).cuda(), torch.from_numpy(batch).cuda(), ratio=0.25).cpu().numpy()
    ptr = np.concatenate([[0], np.cumsum(sizes)])
    exp = on.fps(pos, ptr, [int(np.ceil(0.25 * s)) for s in sizes])
    assert np.array_equal(got, exp)


@pytest.mark.parametrize("sizes,r,K", [([3000, 3000], 0.35, 32), ([500, 1200, 77], 0.5, 9), ([4000], 0.2, 33), ([20000, 20000], 0.3, 64)])
def test_radius_matches_oracle(sizes, r, K):
    from crfconv_b200 import graph_ops
    rng = np.random.default_rng(len(sizes))
    pos = (rng.random((sum(sizes), 3)) * np.array([8.0, 6.0, 3.0])).astype(np.float32)
    batch = np.repeat(np.arange(len(sizes)), sizes)
    ptr = np.concatenate([[0], np.cumsum(sizes)])
    er, ec = on.radius(pos, ptr, pos, ptr, r, K)
    got = graph_ops.radius(torch.from_numpy(pos).cuda(), torch.from_numpy(pos).cuda(), r, torch.from_numpy(batch).cuda(),
                           torch.from_numpy(batch).cuda(), max_num_neighbors=K).cpu().numpy()
    assert np.array_equal(got[0], er) and np.array_equal(got[1], ec)
    g = graph_ops.radius_graph(torch.from_numpy(pos).cuda(), r, torch.from_numpy(batch).cuda(), loop=False, max_num_neighbors=K - 1).cpu().numpy()
    keep = er != ec
    assert np.array_equal(g[0], ec[keep]) and np.array_equal(g[1], er[keep])


def test_knn_graph_and_interpolate():
    from crfconv_b200 import graph_ops
    rng = np.random.default_rng(5)
    px, py = rng.random((900, 3)).astype(np.float32), rng.random((2500, 3)).astype(np.float32)
    bx, by = np.repeat([0, 1], [400, 500]), np.repeat([0, 1], [1000, 1500])
    x = torch.randn(900, 24, generator=torch.Generator().manual_seed(1))
    row, col = graph_ops.knn(torch.from_numpy(px).cuda(), torch.from_numpy(py).cuda(), 3, torch.from_numpy(bx).cuda(), torch.from_numpy(by).cuda())
    exp = np.concatenate([on.knn(px[:400], py[:1000], 3), on.knn(px[400:], py[1000:], 3) + 400]).reshape(-1)
    assert np.array_equal(col.cpu().numpy(), exp) and np.array_equal(row.cpu().numpy(), np.repeat(np.arange(2500), 3))
    xg = x.cuda().requires_grad_(True)
    out = graph_ops.knn_interpolate(xg, torch.from_numpy(px).cuda(), torch.from_numpy(py).cuda(), torch.from_numpy(bx).cuda(), torch.from_numpy(by).cuda(), k=3)
    out.square().sum().backward()
    xd = x.double().requires_grad_(True)
    d = torch.from_numpy(px).double()[exp] - torch.from_numpy(py).double().repeat_interleave(3, 0)
    w = 1.0 / torch.clamp((d * d).sum(-1, keepdim=True), min=1e-16)
    num = torch.zeros(2500, 24, dtype=torch.float64).index_add_(0, torch.arange(2500).repeat_interleave(3), xd[exp] * w)
    den = torch.zeros(2500, 1, dtype=torch.float64).index_add_(0, torch.arange(2500).repeat_interleave(3), w)
    ref = num / den
    ref.square().sum().backward()
    assert _rel(out, ref) < 1e-4 and _rel(xg.grad, xd.grad) < 1e-4
    g = graph_ops.knn_graph(torch.from_numpy(px).cuda(), 5, torch.from_numpy(bx).cuda(), loop=False).cpu().numpy()
    assert g.shape == (2, 900 * 5) and (g[0] != g[1]).all()


def _random_graph(N, kmax, seed):
    rng = np.random.default_rng(seed)
    deg = rng.integers(0, kmax + 1, N)
    row = np.repeat(np.arange(N), deg)
    col = rng.integers(0, N, row.size)
    return torch.from_numpy(np.stack([col, row]))                   # [source, target], grouped by target


@pytest.mark.parametrize("C", [16, 50, 256])
def test_edge_ops_vs_autograd(C):
    """edge softmax and weighted aggregation (forward + backward) on a ragged graph, incl. nodes without edges, any channel count."""
    from crfconv_b200 import graph_ops
    N = 700
    ei = _random_graph(N, 12, C)
    col, row = ei[0], ei[1]
    g = torch.Generator().manual_seed(C)
    y, x = torch.randn(N, C, generator=g) * 0.4, torch.randn(N, C, generator=g)
    cot = torch.randn(N, C, generator=g)
    yd, xd = y.double().requires_grad_(True), x.double().requires_grad_(True)
    s = ol._group_softmax(-((yd[row] - yd[col]) ** 2).sum(1, keepdim=True), row, N)
    out = torch.zeros(N, C, dtype=torch.float64).index_add_(0, row, s * xd[col])
    (out * cot.double()).sum().backward()
    eptr, colg, _ = graph_ops.csr_by_target(row.cuda(), col.cuda(), N)
    yg, xg = y.cuda().requires_grad_(True), x.cuda().requires_grad_(True)
    sg = graph_ops.EdgeSoftmax.apply(yg, eptr, colg)
    og = graph_ops.SpMM.apply(sg, xg, eptr, colg)
    (og * cot.cuda()).sum().backward()
    errs = {"s": _rel(sg, s[:, 0]), "out": _rel(og, out), "dy": _rel(yg.grad, yd.grad), "dx": _rel(xg.grad, xd.grad)}
    assert all(v < 1e-4 for v in errs.values()), errs


def _graph_inputs(sizes, r, K, seed):
    rng = np.random.default_rng(seed)
    pos = (rng.random((sum(sizes), 3))).astype(np.float32)
    batch = np.repeat(np.arange(len(sizes)), sizes)
    ptr = np.concatenate([[0], np.cumsum(sizes)])
    er, ec = on.radius(pos, ptr, pos, ptr, r, K + 1)
    keep = er != ec
    return torch.from_numpy(pos), torch.from_numpy(batch), torch.from_numpy(np.stack([ec[keep], er[keep]]))


@pytest.mark.parametrize("Cn,Ce,Co,steps", [(32, 16, 64, 1), (64, 64, 256, 2), (8, 6, 16, 3)])
def test_guide_crf_vs_oracle(Cn, Ce, Co, steps):
    """GuideGaussianCRFConv (continuous_crf_conv.py:9-69) incl. its internally built radius graph, forward + backward."""
    from crfconv_b200.continuous_crf_conv import GuideGaussianCRFConv
    pos, batch, ei = _graph_inputs([900, 1100], 0.12, 32, Co)
    torch.manual_seed(Co)
    mo = ol.GuideGaussianCRFConv(Cn, Ce, Co, radius=0.12, kernel_size=32, steps=steps).train()
    with torch.no_grad():
        mo.c.add_(0.05 * torch.randn(Co, Co))
    mp = GuideGaussianCRFConv(Cn, Ce, Co, radius=0.12, kernel_size=32, steps=steps)
    mp.load_state_dict(mo.state_dict())
    mp = mp.cuda().train()
    g = torch.Generator().manual_seed(1)
    x, y = torch.randn(2000, Cn, generator=g), torch.randn(2000, Ce, generator=g)
    cot = torch.randn(2000, Co, generator=g)
    x0, y0 = x.clone().requires_grad_(True), y.clone().requires_grad_(True)
    x1, y1 = x.cuda().requires_grad_(True), y.cuda().requires_grad_(True)
    o0 = mo(x0, y0, pos, batch, ei)
    (o0 * cot).sum().backward()
    o1 = mp(x1, y1, pos.cuda(), batch.cuda())                        # builds the radius graph itself
    (o1 * cot.cuda()).sum().backward()
    floor = 1e-3 * max(float(p.grad.abs().max()) for p in mo.parameters())
    errs = {"out": _rel(o1, o0), "dx": _rel(x1.grad, x0.grad), "dy": _rel(y1.grad, y0.grad)}
    po = dict(mo.named_parameters())
    errs.update({"g." + n: _rel(p.grad, po[n].grad, floor) for n, p in mp.named_parameters()})
    bo = dict(mo.named_buffers())
    errs.update({"b." + n: _rel(b.float(), bo[n].float()) for n, b in mp.named_buffers()})
    assert all(v < TOL for v in errs.values()), {k: v for k, v in errs.items() if v >= TOL}


def test_discrete_crf_vs_oracle():
    from crfconv_b200.discrete_crf_conv import DiscreteCRFConv
    pos, batch, ei = _graph_inputs([800, 700], 0.15, 16, 3)
    torch.manual_seed(3)
    mo = ol.DiscreteCRFConv(13, 32, hidden_channels=24, num_kernels=4, radius=0.15, kernel_size=16, steps=3)
    with torch.no_grad():
        mo.F.mul_(0.2)
        mo.C.add_(0.1 * torch.randn(13, 13))
    mp = DiscreteCRFConv(13, 32, hidden_channels=24, num_kernels=4, radius=0.15, kernel_size=16, steps=3)
    mp.load_state_dict(mo.state_dict())
    mp = mp.cuda()
    g = torch.Generator().manual_seed(2)
    p, f = torch.softmax(torch.randn(1500, 13, generator=g), -1), torch.randn(1500, 32, generator=g) * 0.3
    cot = torch.randn(1500, 13, generator=g)
    p0, f0 = p.clone().requires_grad_(True), f.clone().requires_grad_(True)
    p1, f1 = p.cuda().requires_grad_(True), f.cuda().requires_grad_(True)
    q0 = mo(pos, p0, f0, batch, ei)
    (q0 * cot).sum().backward()
    q1 = mp(pos.cuda(), p1, f1, batch.cuda())
    (q1 * cot.cuda()).sum().backward()
    errs = {"q": _rel(q1, q0), "dp": _rel(p1.grad, p0.grad), "df": _rel(f1.grad, f0.grad)}
    po = dict(mo.named_parameters())
    errs.update({"g." + n: _rel(q.grad, po[n].grad) for n, q in mp.named_parameters()})
    assert all(v < TOL for v in errs.values()), errs


def test_multiscale_builder_with_fps_matches_the_cpu_pipeline():
    """build_multiscale(sample_method='fps') (datasets/s3dis_dataset.py:434-437) against the oracle FPS + oracle kNN pipeline."""
    from crfconv_b200.multiscale import build_multiscale
    B, N = 2, 2048
    pos = synthetic.room_cloud(B, N, seed=9)
    ms = build_multiscale(torch.from_numpy(pos).cuda(), num_scales=3, ratio=(4, 4, 2), sample_method="fps")
    cur = pos
    for lvl, r in zip(ms, (4, 4, 2)):
        n = cur.shape[1]
        assert np.array_equal(lvl.neighbor_idx.cpu().numpy(), on.knn_batch(cur, cur, 16))
        ptr = np.arange(B + 1) * n
        choice = on.fps(cur.reshape(-1, 3), ptr, n // r).reshape(B, n // r) - ptr[:-1, None]
        sub = np.take_along_axis(cur, choice[..., None], 1)
        assert np.array_equal(lvl.up_idx.cpu().numpy(), on.knn_batch(sub, cur, 1))
        assert np.array_equal(lvl.sub_idx.cpu().numpy(), np.take_along_axis(on.knn_batch(cur, cur, 16), choice[..., None], 1))
        cur = sub
