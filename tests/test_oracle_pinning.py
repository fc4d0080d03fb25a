"""CPU: pins the oracle (oracle/) against the committed golden vectors, which were produced by the UNMODIFIED
reference (tests/golden/make_golden.py): compiled nanoflann wrapper, compiled grid_subsampling core, and the imported
reference PyTorch modules.  When oracle/_ref is present (container, or shipped prebuilt) it is cross-checked live too."""
import types

import numpy as np
import pytest
import torch

from oracle import layers as ol
from tests._util import grad_floor, rel_err
from oracle import native as on


# ---------------------------------------------------------------------------------------------------- kNN
def test_knn_oracle_bit_exact_on_tie_free_clouds(golden):
    g = golden("knn_golden")
    assert np.array_equal(on.knn_batch(g["uni_pts"], g["uni_pts"], 16), g["uni_k16"])
    assert np.array_equal(on.knn_batch(g["uni_pts"], g["uni_pts"], 32), g["uni_k32"])
    assert np.array_equal(on.knn_batch(g["uni_sub"], g["uni_pts"], 1), g["uni_up_k1"])
    assert np.array_equal(on.knn(g["uni_pts"][0], g["single_q"], 5), g["single_k5"])


@pytest.mark.parametrize("name", ["lat", "dup", "flat"])
def test_knn_oracle_distance_multiset_on_tie_heavy_clouds(golden, name):
    """nanoflann's order among equal distances is traversal-dependent (nanoflann.hpp:115-139); the contract that IS
    pinned is the ascending distance sequence, bit for bit."""
    g = golden("knn_golden")
    pts, ref_idx = g[name + "_pts"], g[name + "_k16"]
    idx, dist = on.knn(pts, pts, 16, return_dist=True)
    assert np.array_equal(dist, on.knn_distances(pts, pts, ref_idx))
    assert np.all(np.diff(dist, axis=1) >= 0)
    # canonical tie rule: equal distances come in ascending index order
    same = np.diff(dist, axis=1) == 0
    assert np.all(np.diff(idx, axis=1)[same] > 0)
    # rows whose K+1 smallest distances are all distinct must agree exactly with the reference
    idx17, d17 = on.knn(pts, pts, 17, return_dist=True)
    distinct = np.all(np.diff(d17, axis=1) > 0, axis=1)
    assert np.array_equal(idx[distinct], ref_idx[distinct])


def test_knn_oracle_k_larger_than_n():
    pts = np.random.default_rng(0).random((5, 3)).astype(np.float32)
    idx = on.knn(pts, pts, 8)
    assert np.all(idx[:, 5:] == 0) and np.all(np.sort(idx[:, :5], 1) == np.arange(5))


@pytest.mark.skipif(not on.have_ref_knn(), reason="compiled reference kNN not available")
def test_knn_oracle_vs_live_reference():
    rng = np.random.default_rng(7)
    pts = (rng.random((3, 4000, 3)) * [50, 50, 4]).astype(np.float32)
    assert np.array_equal(on.knn_batch(pts, pts, 16), on.ref_knn_batch(pts, pts, 16, omp=True))


# ------------------------------------------------------------------------------------------ grid subsampling
def test_subsample_oracle_matches_reference_rows_and_order(golden):
    g = golden("subsample_golden")
    for tag, dl in (("dl30", 0.30), ("dl08", 0.08)):
        p, f, c = on.grid_subsample(g["pts"], g["feats"], g["cls"], dl, order="reference")
        assert np.array_equal(p, g[tag + "_p"]) and np.array_equal(f, g[tag + "_f"]) and np.array_equal(c, g[tag + "_c"])
    assert np.array_equal(on.grid_subsample(g["pts"], None, None, 0.2, order="reference"), g["ponly_p"])
    p, c = on.grid_subsample(g["pts"], None, g["cls2"], 0.5, order="reference")
    assert np.array_equal(p, g["c2_p"]) and np.array_equal(c, g["c2_c"])
    p, f = on.grid_subsample(g["pts"] - 3.7, g["feats"], None, 0.25, order="reference")
    assert np.array_equal(p, g["neg_p"]) and np.array_equal(f, g["neg_f"])


def test_subsample_oracle_key_order_is_same_set(golden):
    g = golden("subsample_golden")
    p, f, c, keys = on.grid_subsample(g["pts"], g["feats"], g["cls"], 0.30, order="key", return_keys=True)
    assert np.all(np.diff(keys.astype(np.int64)) > 0)
    a = np.concatenate([p, f, c.astype(np.float32)], 1)
    b = np.concatenate([g["dl30_p"], g["dl30_f"], g["dl30_c"].astype(np.float32)], 1)
    assert np.array_equal(a[np.lexsort(a.T)], b[np.lexsort(b.T)])


@pytest.mark.skipif(not on.have_ref_subsample(), reason="compiled reference subsampling not available")
def test_subsample_oracle_vs_live_reference():
    rng = np.random.default_rng(3)
    pts = (rng.random((50000, 3)) * [8, 6, 3]).astype(np.float32)
    f = rng.integers(0, 256, (50000, 3)).astype(np.float32)
    c = rng.integers(0, 13, 50000).astype(np.int32)
    for dl in (0.04, 0.2, 1.0):
        r = on.ref_grid_subsample(pts, f, c, dl)
        o = on.grid_subsample(pts, f, c, dl, order="reference")
        assert all(np.array_equal(x, y) for x, y in zip(r, o))


# ------------------------------------------------------------------------------------------------- layers
def _load_sd(module, g, prefix):
    sd = {k[len(prefix) + 4:]: torch.from_numpy(v) for k, v in g.items() if k.startswith(prefix + ".sd.")}
    module.load_state_dict(sd)


def _check(module, g, tag, inputs, grad_inputs, tol=1e-4):
    module.train()
    for t in grad_inputs:
        t.requires_grad_(True)
    out = module(*inputs)
    (out * torch.from_numpy(g[tag + ".cot"])).sum().backward()

    floor = grad_floor(g, tag)

    def close(a, b, what):
        err = rel_err(a, b, floor if what.startswith("grad ") else 0.0)
        assert err < tol, f"{tag} {what}: rel err {err:.3e}"

    close(out.detach().numpy(), g[tag + ".out"], "out")
    for i, t in enumerate(grad_inputs):
        close(t.grad.numpy(), g[f"{tag}.gin{i}"], f"gin{i}")
    for n, p in module.named_parameters():
        close(p.grad.numpy(), g[f"{tag}.gparam.{n}"], "grad " + n)
    for n, b in module.named_buffers():
        close(b.detach().numpy(), g[f"{tag}.buf_after.{n}"], "buffer " + n)


@pytest.mark.parametrize("tag,Cu,Cp,steps", [("crf_s1", 128, 64, 1), ("crf_s3", 64, 32, 3)])
def test_crf_layer_oracle_vs_reference(golden, tag, Cu, Cp, steps):
    g = golden("layer_golden")
    m = ol.ContinuousGaussianCRFConv(Cu, Cp, Cp, steps=steps)
    _load_sd(m, g, tag)
    u, p = torch.from_numpy(g[tag + ".unary"]), torch.from_numpy(g[tag + ".pairwise"])
    _check(m, g, tag, (u, p, torch.from_numpy(g[tag + ".up_idx"]), torch.from_numpy(g[tag + ".neighbor_idx"])), (u, p))


@pytest.mark.parametrize("tag,cin,cout,strided", [("rb_plain", 64, 64, False), ("rb_strided", 32, 64, True), ("rb_in6", 6, 32, False)])
def test_resblock_oracle_vs_reference(golden, tag, cin, cout, strided):
    g = golden("layer_golden")
    m = ol.ResNetBBlock(cin, cout)
    _load_sd(m, g, tag)
    x = torch.from_numpy(g[tag + ".x"])
    pos, sub_pos = torch.from_numpy(g["rb.pos"]), torch.from_numpy(g["rb.sub_pos"])
    args = (x, (pos, sub_pos), torch.from_numpy(g["rb.sub_idx"])) if strided else (x, pos, torch.from_numpy(g["rb.neighbor_idx"]))
    _check(m, g, tag, args, (x,))


def test_upsampling_oracle_vs_reference(golden):
    g = golden("layer_golden")
    m = ol.Upsampling(64, 32, 32)
    _load_sd(m, g, "ups")
    xd, xu = torch.from_numpy(g["ups.x_down"]), torch.from_numpy(g["ups.x_up"])
    _check(m, g, "ups", (xd, xu, torch.from_numpy(g["ups.up_idx"])), (xd, xu))


def test_crf_layer_oracle_accepts_batch_of_one(golden):
    """The reference crashes at B=1 (.squeeze(), continuous_crf_conv_big.py:43); the restatement must not."""
    g = golden("layer_golden")
    m = ol.ContinuousGaussianCRFConv(128, 64, 64)
    _load_sd(m, g, "crf_s1")
    out = m(torch.from_numpy(g["crf_s1.unary"][:1]), torch.from_numpy(g["crf_s1.pairwise"][:1]),
            torch.from_numpy(g["crf_s1.up_idx"][:1]), torch.from_numpy(g["crf_s1.neighbor_idx"][:1]))
    assert out.shape == (1, 512, 64)


def test_distance_pick_oracle_against_the_compiled_reference():
    """knn_.cxx:138-271: the reference seeds mt19937 with time(0), so its picks cannot be replayed — what is pinned is (i) that the
    oracle's kNN of the reference's own picked queries equals the reference's neighbour lists, bit for bit, and (ii) that the seeded
    oracle loop satisfies the reference's invariants (no point picked twice before all are covered; queries are support points)."""
    import numpy as np
    from oracle import native as on
    from oracle import synthetic
    pos = synthetic.room_cloud(2, 800, seed=11)
    oi, oq = on.knn_batch_distance_pick(pos, 120, 12, seed=4)
    assert np.array_equal(oi, on.knn_batch(pos, oq, 12))
    for b in range(2):
        assert len(set(oi[b, :, 0].tolist())) == 120 and np.array_equal(pos[b][oi[b, :, 0]], oq[b])
    oi2, _ = on.knn_batch_distance_pick(pos, 120, 12, seed=4)
    assert np.array_equal(oi, oi2)                                    # deterministic in the seed
    if on.have_ref_knn():
        for omp in (False, True):
            ri, rq = on.ref_knn_batch_distance_pick(pos, 120, 12, omp=omp)
            assert np.array_equal(ri, on.knn_batch(pos, rq, 12))
            for b in range(2):
                assert len(set(ri[b, :, 0].tolist())) == 120 and np.array_equal(pos[b][ri[b, :, 0]], rq[b])


def test_large_k_oracle_against_the_compiled_reference():
    import numpy as np
    from oracle import native as on
    from oracle import synthetic
    if not on.have_ref_knn():
        return
    pos = synthetic.room_cloud(1, 3000, seed=12)
    assert np.array_equal(on.knn_batch(pos, pos[:, :200], 64), on.ref_knn_batch(pos, pos[:, :200].copy(), 64))
