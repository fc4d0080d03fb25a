"""GPU unit tests of the layer-specialised CRF kernels (csrc/crf_fused.cu) through the C ABI, each against a float64 torch
restatement of the same piece of models/continuous_crf_conv_big.py / models/common.py (autograd supplies the expected gradients).
Tolerance: 1e-4 relative (max-norm over the tensor) — an order of magnitude inside the layer-level 1e-3 bar; the 3xTF32
contractions measure ~1e-6."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
TOL = 1e-4


def _rel(a, b, floor=0.0):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    assert a.shape == b.shape, (a.shape, b.shape)
    return float((a - b).abs().max() / max(float(b.abs().max()), floor, 1e-30))


def _check(errs, tol=TOL):
    bad = {k: v for k, v in errs.items() if not v < tol}
    assert not bad, bad


def _bn_state(ops, C, H64, gamma, beta, eps=1e-5):
    """ops.BN filled with the batch statistics of H64 (float64 [M, C])."""
    st = ops.BN(C, H64.device, alloc_stats=False)
    mu = H64.mean(0)
    var = H64.var(0, unbiased=False)
    istd = 1.0 / torch.sqrt(var + eps)
    st.mean.copy_(mu.float()); st.invstd.copy_(istd.float())
    st.scale.copy_((gamma.double() * istd).float()); st.shift.copy_((beta.double() - mu * gamma.double() * istd).float())
    st.count, st.training = H64.shape[0], True
    return st


def _scratch(ops, dev):
    return torch.empty(ops.fused_part_floats(), device=dev), torch.zeros(ops.counter_ints(), dtype=torch.int32, device=dev)


@pytest.mark.parametrize("M,Cin,pro", [(5003, 64, False), (40960, 64, False), (2571, 128, False), (30001, 16, True), (7, 16, True)])
def test_lin16_fwd(M, Cin, pro):
    from crfconv_b200 import ops
    dev = torch.device("cuda")
    g = torch.Generator(device="cpu").manual_seed(M + Cin)
    X = (torch.randn(M, Cin, generator=g) * 1.3 + 0.2).to(dev)
    W = (torch.randn(16, Cin, generator=g) / Cin ** 0.5).to(dev)
    bnm = torch.nn.BatchNorm1d(16).to(dev)
    with torch.no_grad():
        bnm.weight.copy_(1 + 0.2 * torch.randn(16, generator=g)); bnm.bias.copy_(0.2 * torch.randn(16, generator=g))
        bnm.running_mean.copy_(torch.randn(16, generator=g)); bnm.running_var.copy_(1 + torch.rand(16, generator=g))
    rm0, rv0 = bnm.running_mean.clone().double(), bnm.running_var.clone().double()
    pre = None
    A = X.double()
    if pro:
        pre = ops.BN(16, dev, alloc_stats=False)
        pre.scale.copy_(1 + 0.3 * torch.randn(16, generator=g)); pre.shift.copy_(0.3 * torch.randn(16, generator=g))
        A = torch.nn.functional.leaky_relu(A * pre.scale.double() + pre.shift.double(), 0.1)
    st = ops.BN(16, dev, alloc_stats=False)
    part, cnt = _scratch(ops, dev)
    for rep in range(2):                       # second run: the counter was left at zero and the result is identical
        if rep:
            bnm.running_mean.copy_(rm0.float()); bnm.running_var.copy_(rv0.float())
        H = ops.lin16_fwd(X, W, st, bnm, part, cnt, pre=pre, pslope=0.1)
        He = A @ W.double().T
        mu, var = He.mean(0), He.var(0, unbiased=False)
        istd = 1 / torch.sqrt(var + bnm.eps)
        errs = {"H": _rel(H, He), "mean": _rel(st.mean, mu, floor=1e-3), "invstd": _rel(st.invstd, istd),
                "scale": _rel(st.scale, bnm.weight.double() * istd),
                "shift": _rel(st.shift, bnm.bias.double() - mu * bnm.weight.double() * istd, floor=1e-2),
                "running_mean": _rel(bnm.running_mean, 0.9 * rm0 + 0.1 * mu),
                "running_var": _rel(bnm.running_var, 0.9 * rv0 + 0.1 * (var * M / max(M - 1, 1)))}
        _check(errs)
        assert int(cnt.abs().sum().item()) == 0


@pytest.mark.parametrize("M,C1,C2,Cout", [(40960, 16, 0, 64), (20000, 64, 64, 64), (9000, 64, 0, 16), (600, 16, 0, 64)])
def test_linear_fwd_bn(M, C1, C2, Cout):
    """tcgen05 forward with the BatchNorm finalize in its tail (M >= 8192) and the two-launch path (small M)."""
    from crfconv_b200 import ops
    dev = torch.device("cuda")
    g = torch.Generator().manual_seed(M)
    X1 = torch.randn(M, C1, generator=g).to(dev)
    X2 = torch.randn(M, C2, generator=g).to(dev) if C2 else None
    W = (torch.randn(Cout, C1 + C2, generator=g) / (C1 + C2) ** 0.5).to(dev)
    bnm = torch.nn.BatchNorm1d(Cout).to(dev)
    st = ops.BN(Cout, dev)
    cnt = torch.zeros(ops.counter_ints(), dtype=torch.int32, device=dev)
    H = ops.linear_fwd_bn(X1, W, st, bnm, cnt, X2=X2)
    A = torch.cat([X1, X2], 1).double() if C2 else X1.double()
    He = A @ W.double().T
    mu, var = He.mean(0), He.var(0, unbiased=False)
    istd = 1 / torch.sqrt(var + bnm.eps)
    _check({"H": _rel(H, He), "mean": _rel(st.mean, mu, floor=1e-3), "invstd": _rel(st.invstd, istd), "scale": _rel(st.scale, istd),
            "shift": _rel(st.shift, -mu * istd, floor=1e-2), "running_mean": _rel(bnm.running_mean, 0.1 * mu, floor=1e-4),
            "running_var": _rel(bnm.running_var, 0.9 + 0.1 * var * M / (M - 1))})
    assert int(cnt.abs().sum().item()) == 0


@pytest.mark.parametrize("M", [4099, 245760])
def test_up16_fwd(M):
    from crfconv_b200 import ops
    dev = torch.device("cuda")
    g = torch.Generator().manual_seed(M)
    X = (torch.randn(M, 16, generator=g) + 0.3).to(dev)
    W = (torch.randn(64, 16, generator=g) / 4).to(dev)
    bnm = torch.nn.BatchNorm1d(64).to(dev)
    st = ops.BN(64, dev)
    cnt = torch.zeros(ops.counter_ints(), dtype=torch.int32, device=dev)
    H = ops.up16_fwd(X, W, st, bnm, cnt)
    He = X.double() @ W.double().T
    mu, var = He.mean(0), He.var(0, unbiased=False)
    istd = 1 / torch.sqrt(var + bnm.eps)
    _check({"H": _rel(H, He), "mean": _rel(st.mean, mu, floor=1e-3), "invstd": _rel(st.invstd, istd), "scale": _rel(st.scale, istd),
            "shift": _rel(st.shift, -mu * istd, floor=1e-2), "running_mean": _rel(bnm.running_mean, 0.1 * mu, floor=1e-4),
            "running_var": _rel(bnm.running_var, 0.9 + 0.1 * var * M / (M - 1))})
    assert int(cnt.abs().sum().item()) == 0


def test_bn_backward_prepare_fin():
    from crfconv_b200 import ops
    dev = torch.device("cuda")
    M, C = 33000, 64
    g = torch.Generator().manual_seed(1)
    H, dY = torch.randn(M, C, generator=g).to(dev), torch.randn(M, C, generator=g).to(dev)
    gamma, beta = (1 + 0.2 * torch.randn(C, generator=g)).to(dev), (0.2 * torch.randn(C, generator=g)).to(dev)
    st = _bn_state(ops, C, H.double(), gamma, beta)
    sums = torch.zeros(ops.STAT_SLOTS * 2 * C, device=dev)
    cnt = torch.zeros(ops.counter_ints(), dtype=torch.int32, device=dev)
    dg, db = torch.zeros(C, device=dev), torch.zeros(C, device=dev)
    ops.bn_backward_prepare_fin(dY, H, st, 0.1, dg, db, sums, cnt)
    pre = H.double() * st.scale.double() + st.shift.double()
    dV = torch.where(pre > 0, dY.double(), 0.1 * dY.double())
    Hh = (H.double() - st.mean.double()) * st.invstd.double()
    _check({"k1": _rel(st.k1, dV.mean(0)), "k2": _rel(st.k2, (dV * Hh).mean(0)), "dgamma": _rel(dg, (dV * Hh).sum(0)), "dbeta": _rel(db, dV.sum(0))})
    assert int(cnt.abs().sum().item()) == 0


def _bn_train64(H, gamma, beta, eps=1e-5):
    mu, var = H.mean(0), H.var(0, unbiased=False)
    return (H - mu) / torch.sqrt(var + eps) * gamma + beta


@pytest.mark.parametrize("M", [4099, 40960])
def test_mid16_bwd_and_in16(M):
    """Backward of  y = BN2(lrelu(BN1(X·W1ᵀ))·W2ᵀ)  through mid16_bwd → in16_dgrad / in16_wgrad, against autograd in float64."""
    from crfconv_b200 import ops
    dev = torch.device("cuda")
    for Cin in (64, 128):
        g = torch.Generator().manual_seed(M + Cin)
        X = torch.randn(M, Cin, generator=g).to(dev)
        W1 = (torch.randn(16, Cin, generator=g) / Cin ** 0.5).to(dev)
        W2 = (torch.randn(16, 16, generator=g) / 4).to(dev)
        g1, b1 = (1 + 0.2 * torch.randn(16, generator=g)).to(dev), (0.2 * torch.randn(16, generator=g)).to(dev)
        g2, b2 = (1 + 0.2 * torch.randn(16, generator=g)).to(dev), (0.2 * torch.randn(16, generator=g)).to(dev)
        cot = torch.randn(M, 16, generator=g).to(dev)
        dX0 = torch.randn(M, Cin, generator=g).to(dev)
        # float64 reference
        Xd = X.double().requires_grad_(True)
        W1d, W2d = W1.double().requires_grad_(True), W2.double().requires_grad_(True)
        g1d, b1d, g2d, b2d = (t.double().requires_grad_(True) for t in (g1, b1, g2, b2))
        H1 = Xd @ W1d.T
        V1 = _bn_train64(H1, g1d, b1d)
        V1.retain_grad()
        H2 = torch.nn.functional.leaky_relu(V1, 0.1) @ W2d.T
        y = _bn_train64(H2, g2d, b2d)
        (y * cot.double()).sum().backward()
        # product
        H1f, H2f = H1.detach().float(), H2.detach().float()
        bn1, bn2 = _bn_state(ops, 16, H1.detach(), g1, b1), _bn_state(ops, 16, H2.detach(), g2, b2)
        H2h = (H2.detach() - bn2.mean.double()) * bn2.invstd.double()
        bn2.k1.copy_(cot.double().mean(0).float()); bn2.k2.copy_((cot.double() * H2h).mean(0).float())
        n_small = 16 * 16 + 16 * Cin
        slots = torch.zeros(ops.GRAD_SLOTS * n_small, device=dev)
        part, cnt = _scratch(ops, dev)
        dg1, db1 = torch.zeros(16, device=dev), torch.zeros(16, device=dev)
        dV1 = ops.mid16_bwd(cot, H2f, bn2, H1f, bn1, 0.1, W2, slots, n_small, part, cnt, dg1, db1)
        H1h = (H1.detach() - bn1.mean.double()) * bn1.invstd.double()
        dV1e = V1.grad
        errs = {"dV1": _rel(dV1, dV1e), "k1": _rel(bn1.k1, dV1e.mean(0), floor=1e-4), "k2": _rel(bn1.k2, (dV1e * H1h).mean(0), floor=1e-4),
                "dgamma1": _rel(dg1, g1d.grad), "dbeta1": _rel(db1, b1d.grad)}
        dXp = dX0.clone()
        ops.in16_dgrad(dV1, H1f, bn1, W1, dXp, True)
        dXq = torch.empty_like(dX0)
        ops.in16_dgrad(dV1, H1f, bn1, W1, dXq, False)
        ops.in16_wgrad(dV1, H1f, bn1, X, slots[256:], n_small)
        dW = slots.view(ops.GRAD_SLOTS, n_small).double().sum(0)
        errs.update({"dX(+=)": _rel(dXp, dX0.double() + Xd.grad), "dX(=)": _rel(dXq, Xd.grad), "dW2": _rel(dW[:256].view(16, 16), W2d.grad),
                     "dW1": _rel(dW[256:].view(16, Cin), W1d.grad)})
        _check(errs)
        assert int(cnt.abs().sum().item()) == 0


@pytest.mark.parametrize("M", [3001, 40960])
def test_out16_bwd(M):
    """o = lrelu(BN(x·W3ᵀ)): one pass over dO gives dW3, dγ, dβ and (T, Q, a0) with dL/dx = T − a0 − x·Qᵀ."""
    from crfconv_b200 import ops
    dev = torch.device("cuda")
    g = torch.Generator().manual_seed(M)
    x = (torch.randn(M, 16, generator=g) + 0.3).to(dev)
    W3 = (torch.randn(64, 16, generator=g) / 4).to(dev)
    gm, bt = (1 + 0.2 * torch.randn(64, generator=g)).to(dev), (0.2 * torch.randn(64, generator=g)).to(dev)
    dO = torch.randn(M, 64, generator=g).to(dev)
    xd, Wd, gd, bd = x.double().requires_grad_(True), W3.double().requires_grad_(True), gm.double().requires_grad_(True), bt.double().requires_grad_(True)
    H3 = xd @ Wd.T
    o = torch.nn.functional.leaky_relu(_bn_train64(H3, gd, bd), 0.1)
    (o * dO.double()).sum().backward()
    bn3 = _bn_state(ops, 64, H3.detach(), gm, bt)
    part = torch.zeros(ops.out_bwd_part_floats(), device=dev)
    cnt = torch.zeros(ops.counter_ints(), dtype=torch.int32, device=dev)
    dg, db, dW3 = torch.zeros(64, device=dev), torch.zeros(64, device=dev), torch.zeros(64, 16, device=dev)
    Q, a0 = torch.empty(16, 16, device=dev), torch.empty(16, device=dev)
    T = ops.out16_bwd(dO, H3.detach().float(), bn3, 0.1, x, W3, part, cnt, dg, db, dW3, Q, a0)
    gx = T.double() - a0.double() - x.double() @ Q.double().T
    _check({"dx": _rel(gx, xd.grad), "dW3": _rel(dW3, Wd.grad), "dgamma": _rel(dg, gd.grad), "dbeta": _rel(db, bd.grad),
            "Q symmetric": _rel(Q, Q.T)})
    assert int(cnt.abs().sum().item()) == 0


@pytest.mark.parametrize("B,N,corr", [(2, 1500, False), (3, 4096, True), (1, 40960, True)])
def test_step_bwd_fused_matches_the_generic_kernels(B, N, corr):
    """Fused mean-field backward (in-kernel GC/GM, out_nn correction, y-layer BatchNorm sums) vs the generic step kernel + GEMMs."""
    from crfconv_b200 import ops
    from crfconv_b200.nearest_neighbors import knn_batch
    dev = torch.device("cuda")
    g = torch.Generator().manual_seed(N)
    M, F, K = B * N, 16, 16
    pos = torch.rand(B, N, 3, generator=g).to(dev)
    nbr = knn_batch(pos, pos, K)
    Hy = torch.randn(M, F, generator=g).to(dev) * 0.5
    z, gin, xT = (torch.randn(M, F, generator=g).to(dev) for _ in range(3))
    c = (torch.eye(F) + 0.1 * torch.randn(F, F, generator=g)).to(dev)
    Cm, Minv = ops.crf_compat_fwd(c)
    gamma = (1 + 0.2 * torch.randn(F, generator=g)).to(dev)
    bny = _bn_state(ops, F, Hy.double(), gamma, torch.zeros(F, device=dev))
    Qm = torch.randn(F, F, generator=g).to(dev) * 0.1
    Qm = (Qm + Qm.T).contiguous()
    a0 = torch.randn(F, generator=g).to(dev) * 0.1
    geff = (gin.double() - a0.double() - xT.double() @ Qm.double().T).float() if corr else gin
    # generic kernels
    Gz0, gp0, Gy0 = torch.empty(M, F, device=dev), torch.zeros(M, F, device=dev), torch.zeros(M, F, device=dev)
    m_o, v_o, h_o = (torch.empty(M, F, device=dev) for _ in range(3))
    ops.crf_step_bwd(Hy, bny.scale, z, z, nbr, Cm, Minv, geff, Gz0, gp0, Gy0, m_o, v_o, h_o, False, B, N, K)
    GC0, GM0 = m_o.double().T @ h_o.double(), v_o.double().T @ geff.double()
    # fused kernel
    n_small = 2 * F * F
    slots = torch.zeros(ops.GRAD_SLOTS * n_small, device=dev)
    Gz1, gp1, Gy1 = torch.empty(M, F, device=dev), torch.zeros(M, F, device=dev), torch.zeros(M, F, device=dev)
    ysum = torch.zeros(128, device=dev)
    cnt = torch.zeros(ops.counter_ints(), dtype=torch.int32, device=dev)
    dg, db = torch.zeros(F, device=dev), torch.zeros(F, device=dev)
    for variant in (2, 3):
        ops._lib.lib().crfconv_fused_tune(0, variant)
        slots.zero_(); gp1.zero_(); Gy1.zero_(); ysum.zero_(); dg.zero_()
        ops.crf_step_bwd_fused(Hy, bny, z, z, nbr, Cm, Minv, gin, xT if corr else None, Qm if corr else None, a0 if corr else None,
                               Gz1, False, gp1, Gy1, slots, slots[F * F:], n_small, ysum, B, N, K, True, cnt, gamma, dg, db)
        S = slots.view(ops.GRAD_SLOTS, n_small).double().sum(0)
        Hh = (Hy.double() - bny.mean.double()) * bny.invstd.double()
        s2 = (Gy0.double() * Hh).sum(0)
        _check({"Gz": _rel(Gz1, Gz0), "gprev": _rel(gp1, gp0), "Gy": _rel(Gy1, Gy0), "GC": _rel(S[:F * F].view(F, F), GC0),
                "GM": _rel(S[F * F:].view(F, F), GM0), "k2_y": _rel(bny.k2, s2 / M, floor=float(s2.abs().max()) / M * 1e-2),
                "dgamma_y": _rel(dg, s2, floor=float(s2.abs().max()) * 1e-2), "k1_y": float(bny.k1.abs().max())})
        assert int(cnt.abs().sum().item()) == 0
    ops._lib.lib().crfconv_fused_tune(0, 0)


def _pack_yx(Hy, z):
    M = Hy.shape[0]
    return torch.stack([Hy.view(M, 4, 4), z.view(M, 4, 4)], dim=2).reshape(M, 32).contiguous()


@pytest.mark.parametrize("B,N", [(2, 1501), (3, 4096), (1, 40960)])
def test_packed_mean_field_matches_the_two_table_kernels(B, N):
    """The packed {Hy | z} layout (one 128-byte line per gathered neighbour, 256-bit loads) computes what the two-table kernels do:
    producers (lin16_fwd's second output, upsample into the z half), forward step, fused backward step."""
    from crfconv_b200 import ops
    from crfconv_b200.nearest_neighbors import knn_batch
    dev = torch.device("cuda")
    g = torch.Generator().manual_seed(N + 1)
    M, F, K, Nc = B * N, 16, 16, max(N // 4, 1)
    Mc = B * Nc
    pos = torch.rand(B, N, 3, generator=g).to(dev)
    nbr = knn_batch(pos, pos, K)
    up = torch.randint(0, Nc, (B, N), generator=g).to(dev)
    H1 = torch.randn(M, F, generator=g).to(dev)
    Hu = torch.randn(Mc, F, generator=g).to(dev)
    W2 = (torch.randn(F, F, generator=g) / 4).to(dev)
    bn1 = _bn_state(ops, F, H1.double(), torch.ones(F, device=dev), torch.zeros(F, device=dev))
    bnu = _bn_state(ops, F, Hu.double(), (1 + 0.2 * torch.randn(F, generator=g)).to(dev), (0.1 * torch.randn(F, generator=g)).to(dev))
    part, cnt = _scratch(ops, dev)
    bnm = torch.nn.BatchNorm1d(F).to(dev)
    gamma = (1 + 0.2 * torch.randn(F, generator=g)).to(dev)
    with torch.no_grad():
        bnm.weight.copy_(gamma)
    bny = ops.BN(F, dev, alloc_stats=False)
    YX = torch.full((M, 32), float("nan"), device=dev)
    Hy = ops.lin16_fwd(H1, W2, bny, bnm, part, cnt, pre=bn1, pslope=0.1, packed_out=YX)
    z = ops.crf_upsample_fwd(Hu, bnu, up, B, N, Nc)
    ops.crf_upsample_fwd_packed(Hu, bnu, up, YX, B, N, Nc)
    assert torch.equal(YX, _pack_yx(Hy, z))                    # producers: bit-identical values, no float left unwritten
    c = (torch.eye(F) + 0.1 * torch.randn(F, F, generator=g)).to(dev)
    Cm, Minv = ops.crf_compat_fwd(c)
    x0 = ops.crf_step_fwd(Hy, bny.scale, z, z, nbr, Cm, Minv, B, N, K)
    x1 = ops.crf_step_fwd_packed(YX, bny.scale, nbr, Cm, Minv, B, N)
    _check({"x": _rel(x1, x0)}, tol=2e-6)
    gin, xT = (torch.randn(M, F, generator=g).to(dev) for _ in range(2))
    Qm = torch.randn(F, F, generator=g).to(dev) * 0.1
    Qm = (Qm + Qm.T).contiguous()
    a0 = torch.randn(F, generator=g).to(dev) * 0.1
    n_small = 2 * F * F
    res = []
    for packed in (False, True):
        slots = torch.zeros(ops.GRAD_SLOTS * n_small, device=dev)
        Gz, gp, Gy = torch.empty(M, F, device=dev), torch.zeros(M, F, device=dev), torch.zeros(M, F, device=dev)
        ysum = torch.zeros(128, device=dev)
        dg, db = torch.zeros(F, device=dev), torch.zeros(F, device=dev)
        bny.k1.zero_(); bny.k2.zero_()
        ops.crf_step_bwd_fused(YX if packed else Hy, bny, None if packed else z, None if packed else z, nbr, Cm, Minv, gin, xT, Qm, a0,
                               Gz, False, gp, Gy, slots, slots[F * F:], n_small, ysum, B, N, K, True, cnt, gamma, dg, db, packed=packed)
        res.append((Gz, gp, Gy, slots.view(ops.GRAD_SLOTS, n_small).double().sum(0), bny.k2.clone(), dg))
        assert int(cnt.abs().sum().item()) == 0
    names = ("Gz", "gprev", "Gy", "GC|GM", "k2_y", "dgamma_y")
    _check({n: _rel(a, b, floor=1e-6) for n, a, b in zip(names, res[1], res[0])}, tol=1e-5)   # same arithmetic; atomics reorder sums


def test_upsample_bwd_fused():
    from crfconv_b200 import ops
    dev = torch.device("cuda")
    B, N, Nc, F = 3, 5001, 1250, 16
    g = torch.Generator().manual_seed(5)
    Gz, G0 = torch.randn(B * N, F, generator=g).to(dev), torch.randn(B * N, F, generator=g).to(dev)
    up = torch.randint(0, Nc, (B, N), generator=g).to(dev)
    Hu = torch.randn(B * Nc, F, generator=g).to(dev)
    bnu = _bn_state(ops, F, Hu.double(), torch.ones(F, device=dev), torch.zeros(F, device=dev))
    Gu = torch.zeros(B * Nc, F, device=dev)
    part, cnt = _scratch(ops, dev)
    dg, db = torch.zeros(F, device=dev), torch.zeros(F, device=dev)
    ops.crf_upsample_bwd_fused(Gz, G0, up, Hu, bnu, Gu, B, N, Nc, part, cnt, dg, db)
    flat = (up + torch.arange(B, device=dev)[:, None] * Nc).reshape(-1)
    Ge = torch.zeros(B * Nc, F, device=dev, dtype=torch.float64).index_add_(0, flat, (Gz + G0).double())
    Hh = (Hu.double() - bnu.mean.double()) * bnu.invstd.double()
    _check({"Gu": _rel(Gu, Ge), "k1": _rel(bnu.k1, Ge.mean(0), floor=1e-3), "k2": _rel(bnu.k2, (Ge * Hh).mean(0), floor=1e-3),
            "dgamma": _rel(dg, (Ge * Hh).sum(0)), "dbeta": _rel(db, Ge.sum(0))})
    assert int(cnt.abs().sum().item()) == 0


@pytest.mark.parametrize("B,N,Cu,steps", [(2, 3000, 128, 1), (1, 9000, 64, 2)])
def test_fused_layer_matches_generic_layer(B, N, Cu, steps):
    """The whole layer on the layer-specialised kernels vs the same module on the generic kernels (same weights, same inputs)."""
    import crfconv_b200.continuous_crf_conv_big as cb
    from crfconv_b200.nearest_neighbors import knn_batch
    dev = torch.device("cuda")
    g = torch.Generator().manual_seed(N)
    pos = torch.rand(B, N, 3, generator=g).to(dev)
    nbr = knn_batch(pos, pos, 16)
    Nc = N // 4
    up = torch.randint(0, Nc, (B, N, 1), generator=g).to(dev)
    torch.manual_seed(0)
    m = cb.ContinuousGaussianCRFConv(Cu, 64, 64, steps=steps).to(dev).train()
    with torch.no_grad():
        m.c.copy_(torch.eye(16, device=dev) + 0.1 * torch.randn(16, 16, device=dev))
    unary, pairwise = torch.randn(B, Nc, Cu, generator=g).to(dev), torch.randn(B, N, 64, generator=g).to(dev)
    cot = torch.randn(B, N, 64, generator=g).to(dev)
    res = {}
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    for fused in (False, True):
        cb.USE_FUSED = fused
        m.load_state_dict(sd)
        m.zero_grad()
        u, p = unary.clone().requires_grad_(True), pairwise.clone().requires_grad_(True)
        out = m(u, p, up, nbr)
        (out * cot).sum().backward()
        res[fused] = {"out": out.detach(), "du": u.grad, "dp": p.grad, **{"g." + n: q.grad.clone() for n, q in m.named_parameters()},
                      **{"b." + n: b.clone().float() for n, b in m.named_buffers()}}
    cb.USE_FUSED = True
    from tests._util import rel_err_trimmed, rel_l2
    floor = 1e-3 * max(float(v.abs().max()) for k, v in res[False].items() if k.startswith("g."))
    errs = {}
    for k in res[False]:
        a, b = res[True][k].double().cpu().numpy(), res[False][k].double().cpu().numpy()
        if k in ("out", "du", "dp"):       # a LeakyReLU kink flip between two correct kernels moves single entries: trimmed max-norm + L2
            errs[k] = rel_err_trimmed(a, b)
            errs[k + "(l2)"] = rel_l2(a, b)
        else:
            errs[k] = _rel(res[True][k], res[False][k], floor if k.startswith("g.") else 0.0)
    # parameter gradients of the two paths differ by summation order only; the BatchNorm biases are sums with heavy cancellation
    _check(errs, tol=1e-3)


@pytest.mark.parametrize("B,Ns,Nq,K,d", [(2, 4096, 4096, 16, 8), (3, 5000, 1250, 16, 8), (1, 777, 777, 7, 8), (6, 40960, 40960, 16, 8),
                                          (2, 4096, 4096, 16, 16), (3, 5001, 1251, 16, 16), (1, 777, 777, 7, 16), (6, 40960, 10240, 16, 16)])
def test_pointconv_without_edge_tensors(B, Ns, Nq, K, d):
    """PointConv (hidden width 8 / 16) with the edge MLP recomputed in every pass and the weight gradients derived from sums
    (csrc/pointconv_fused.cu) against a float64 statement of models/point_conv_big.py:37-58 in torch (autograd), with the layer-by-layer
    kernels (which materialise the [E, 8] tensors) measured beside it: 1e-3 of each tensor's max for both; the fused path must not be
    further from the truth than 3x the layer-by-layer path wherever that matters (> 1e-4)."""
    import crfconv_b200.point_conv_big as pcb
    from crfconv_b200.nearest_neighbors import knn_batch
    dev = torch.device("cuda")
    g = torch.Generator().manual_seed(Ns + Nq + K)
    sup = (torch.rand(B, Ns, 3, generator=g) * torch.tensor([8.0, 6.0, 3.0])).to(dev)
    cen = sup if Nq == Ns else sup[:, torch.randperm(Ns, generator=g)[:Nq]].contiguous()
    idx = knn_batch(sup, cen, K)
    x0 = torch.randn(B, Ns, d, generator=g).to(dev)
    cot = torch.randn(B, Nq, d, generator=g).to(dev)
    torch.manual_seed(1)
    m = pcb.PointConv(d).to(dev).train()
    with torch.no_grad():
        for n, p in m.named_parameters():
            if "batch_norm.weight" in n:
                p.copy_(1 + 0.2 * torch.randn(p.shape, generator=g))
            elif "batch_norm.bias" in n:
                p.copy_(0.2 * torch.randn(p.shape, generator=g))
    # float64 truth
    P = {n: p.detach().double().requires_grad_(True) for n, p in m.named_parameters()}
    xt = x0.double().requires_grad_(True)
    bi = torch.arange(B, device=dev).view(B, 1, 1)
    rel = (cen.double()[:, :, None, :] - sup.double()[bi, idx]).reshape(-1, 3)

    def bn(h, gamma, beta):
        mu, var = h.mean(0), h.var(0, unbiased=False)
        return (h - mu) / torch.sqrt(var + 1e-5) * gamma + beta
    a1 = torch.nn.functional.leaky_relu(bn(rel @ P["weight_nn.0.lin.weight"].t(), P["weight_nn.0.bn.batch_norm.weight"], P["weight_nn.0.bn.batch_norm.bias"]), 0.1)
    w = bn(a1 @ P["weight_nn.1.lin.weight"].t(), P["weight_nn.1.bn.batch_norm.weight"], P["weight_nn.1.bn.batch_norm.bias"])
    ot = (w.view(B, Nq, K, d) * xt[bi, idx]).sum(2)
    (ot * cot.double()).sum().backward()
    truth = (ot.detach(), xt.grad, {n: p.grad for n, p in P.items()})
    res = []
    sd0 = {k: v.clone() for k, v in m.state_dict().items()}
    for fused in (False, True):
        m.load_state_dict(sd0)
        pcb.FUSED_EDGE_MLP = fused
        m.zero_grad(set_to_none=True)
        x = x0.clone().requires_grad_(True)
        out = m(x, sup if Nq == Ns else (sup, cen), idx)
        (out * cot).sum().backward()
        floor = 1e-3 * max(float(v.abs().max()) for v in truth[2].values())
        errs = {"out": _rel(out.detach(), truth[0]), "dx": _rel(x.grad, truth[1])}
        errs.update({"grad " + n: _rel(p.grad, truth[2][n], floor=floor) for n, p in m.named_parameters()})
        res.append((errs, {k: v.clone() for k, v in m.state_dict().items() if "running" in k}))
    pcb.FUSED_EDGE_MLP = True
    (e0, rs0), (e1, rs1) = res
    print("layer-by-layer:", {k: f"{v:.1e}" for k, v in e0.items()})
    print("fused         :", {k: f"{v:.1e}" for k, v in e1.items()})
    # 5,439 edges (the K = 7 case): one LeakyReLU unit that takes the other branch moves the first layer's gradients by ~1/edges
    tol = 1e-3 if B * Nq * K >= 100000 else 5e-3
    _check(e1, tol=tol)
    _check({k: v for k, v in e1.items() if v > 1e-4 and v > 3 * e0[k]}, tol=0.0) if B * Nq * K >= 100000 else None
    _check({"buf " + n: _rel(rs1[n], rs0[n], floor=1e-3) for n in rs0}, tol=2e-4)
