"""GPU unit parity of the Linear/BatchNorm C-ABI ops against an fp64 PyTorch statement of the same formulas, for both
kernel generations (generic 3xTF32 and persistent bf16x3) and every shape the CRF layer and the ResNet blocks use.
Tolerance relative to the tensor's max: forward 2e-5 (both generations contract the forward in 3xTF32), backward 2e-5 (generic,
3xTF32) / 2e-4 (fast, bf16x3) — well inside the 1e-3 layer budget.  The backward reference is built from the product's own
forward output H, so that a LeakyReLU branch decided differently at |pre-activation| ~ 1e-6 cannot masquerade as a backward error."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from crfconv_b200 import ops as o
    return o


def lrelu(v, s):
    return torch.where(v > 0, v, v * s)


def rel(a, b):
    return float((a.double() - b).abs().max() / b.abs().max().clamp_min(1e-30))


SHAPES = [  # (M, C1, C2, Cout)
    (9000, 128, 0, 16), (9000, 64, 0, 16), (9000, 16, 0, 16), (9000, 16, 0, 64), (9000, 64, 64, 64), (8200, 8, 0, 8),
    (8200, 32, 32, 32), (8200, 8, 0, 32), (8200, 64, 0, 8), (12345, 32, 0, 8), (8200, 64, 0, 64), (8200, 128, 0, 64), (8200, 6, 0, 32),
    (1000, 256, 256, 128), (960, 128, 0, 512), (3841, 512, 0, 256), (1280, 256, 256, 256), (320, 512, 0, 64), (1280, 256, 0, 64),      # deep levels: few rows, wide channels (linear_small.cu)
]


@pytest.mark.parametrize("fast", [0, 1])
@pytest.mark.parametrize("M,C1,C2,Cout", SHAPES)
def test_linear_fwd_and_bwd(ops, fast, M, C1, C2, Cout):
    from crfconv_b200 import _lib
    L = _lib.lib()
    g = torch.Generator(device="cuda").manual_seed(M + C1 + Cout)
    rn = lambda *s: torch.randn(*s, generator=g, device="cuda")   # noqa: E731
    X1, X2 = rn(M, C1), (rn(M, C2) if C2 else None)
    W = rn(Cout, C1 + C2) / (C1 + C2) ** 0.5
    sc1, sh1 = 1 + 0.2 * rn(C1), 0.2 * rn(C1)
    gamma, beta = 1 + 0.2 * rn(Cout), 0.2 * rn(Cout)
    dY = rn(M, Cout)
    tol = 2e-4 if fast else 2e-5
    ftol = 2e-5
    prev = L.crfconv_set_fast_path(fast)
    try:
        # ---------------- forward: H = [lrelu(X1*sc+sh, .1) | X2] Wᵀ with Σ/Σ² epilogue
        bn = ops.BN(Cout, X1.device)
        H = ops.linear_fwd(X1, W, scale1=sc1, shift1=sh1, slope1=0.1, X2=X2, stats=bn.stats)
        A = lrelu(X1.double() * sc1.double() + sh1.double(), 0.1)
        if C2:
            A = torch.cat([A, X2.double()], 1)
        Href = A @ W.double().t()
        assert rel(H, Href) < ftol
        st = bn.stats.view(-1, 2 * Cout).sum(0)                 # slotted partial sums (CRFCONV_STAT_SLOTS)
        assert rel(st[:Cout], Href.sum(0)) < 1e-4 and rel(st[Cout:], (Href ** 2).sum(0)) < 1e-4
        ops.bn_finalize_fwd(bn, M, gamma, beta, 1e-5, 0.1, True, None, None)
        mu, var = Href.mean(0), Href.var(0, unbiased=False)
        assert rel(bn.mean, mu) < 1e-4 and rel(bn.invstd, (var + 1e-5).rsqrt()) < 1e-4
        # ---------------- backward through lrelu(BN(H), .1), referenced to the product's own H
        Href = H.double()
        mu, var = Href.mean(0), Href.var(0, unbiased=False)
        Hh = (Href - mu) * (var + 1e-5).rsqrt()
        V = Hh * gamma.double() + beta.double()
        dV = torch.where(V > 0, dY.double(), dY.double() * 0.1)
        dgam, dbet = torch.zeros(Cout, device="cuda"), torch.zeros(Cout, device="cuda")
        ops.bn_backward_prepare(dY, H, bn, 0.1, dgam, dbet)
        assert rel(dbet, dV.sum(0)) < 1e-4 and rel(dgam, (dV * Hh).sum(0)) < 1e-4
        dH = gamma.double() * (var + 1e-5).rsqrt() * (dV - dV.mean(0) - Hh * (dV * Hh).mean(0))
        dX1, dX2 = torch.empty_like(X1), (torch.full_like(X2, 1.0) if C2 else None)
        dW = torch.zeros_like(W)
        ops.linear_bwd(dY, H, bn, 0.1, X1, W, scale1=sc1, shift1=sh1, slope1=0.1, X2=X2, dX1=dX1, dX2=dX2, acc2=True, dW=dW)
        dA = dH @ W.double()
        assert rel(dX1, dA[:, :C1]) < tol
        if C2:
            assert rel(dX2, dA[:, C1:] + 1.0) < tol          # acc2: accumulated onto the ones
        assert rel(dW, dH.t() @ A) < tol
        # ---------------- plain (no BN) wgrad-only and dgrad-only calls, as used for GC = mᵀh
        if C2 == 0:
            dW2 = torch.zeros_like(W)
            ops.linear_bwd(dY, None, None, 1.0, X1, W, dW=dW2)
            assert rel(dW2, dY.double().t() @ X1.double()) < tol
            dX = torch.empty_like(X1)
            ops.linear_bwd(dY, None, None, 1.0, X1, W, dX1=dX)
            assert rel(dX, dY.double() @ W.double()) < tol
    finally:
        L.crfconv_set_fast_path(prev)


@pytest.mark.parametrize("C1,Cout", [(16, 64), (64, 128)])
@pytest.mark.parametrize("fast", [0, 1])
def test_residual_activation_reference(ops, fast, C1, Cout):
    """act_ref path: the LeakyReLU branch is taken from a saved output (ResNetBBlock's lrelu(BN(h)+residual), point_conv_big.py:88)."""
    from crfconv_b200 import _lib
    L = _lib.lib()
    M = 8300
    g = torch.Generator(device="cuda").manual_seed(7)
    rn = lambda *s: torch.randn(*s, generator=g, device="cuda")   # noqa: E731
    X1, W, R, dY = rn(M, C1), rn(Cout, C1) / 4, rn(M, Cout), rn(M, Cout)
    gamma, beta = 1 + 0.2 * rn(Cout), 0.2 * rn(Cout)
    prev = L.crfconv_set_fast_path(fast)
    try:
        bn = ops.BN(Cout, X1.device)
        H = ops.linear_fwd(X1, W, stats=bn.stats)
        ops.bn_finalize_fwd(bn, M, gamma, beta, 1e-5, 0.1, True, None, None)
        out = ops.bn_act_fwd(H, bn, 0.01, R=R)
        Href = X1.double() @ W.double().t()
        mu, var = Href.mean(0), Href.var(0, unbiased=False)
        Hh = (Href - mu) * (var + 1e-5).rsqrt()
        oref = lrelu(Hh * gamma.double() + beta.double() + R.double(), 0.01)
        assert rel(out, oref) < 2e-5
        Href = H.double()
        mu, var = Href.mean(0), Href.var(0, unbiased=False)
        Hh = (Href - mu) * (var + 1e-5).rsqrt()
        dV = torch.where(out.double() > 0, dY.double(), dY.double() * 0.01)
        dgam, dbet = torch.zeros(Cout, device="cuda"), torch.zeros(Cout, device="cuda")
        ops.bn_backward_prepare(dY, H, bn, 0.01, dgam, dbet, act_ref=out)
        dH = gamma.double() * (var + 1e-5).rsqrt() * (dV - dV.mean(0) - Hh * (dV * Hh).mean(0))
        dX, dW = torch.empty_like(X1), torch.zeros_like(W)
        ops.linear_bwd(dY, H, bn, 0.01, X1, W, dX1=dX, dW=dW, act_ref=out)
        tol = 2e-4 if fast else 2e-5
        assert rel(dX, dH @ W.double()) < tol and rel(dW, dH.t() @ X1.double()) < tol
    finally:
        L.crfconv_set_fast_path(prev)


@pytest.mark.parametrize("M,C1,Cout,bn_on,bias,gather", [(245760, 128, 13, False, True, False), (50001, 32, 128, True, False, False),
                                                         (30000, 6, 8, True, False, False), (3840, 32, 16, True, False, True),
                                                         (977, 8, 32, False, True, True), (15, 64, 13, False, True, False),
                                                         (3840, 256, 128, True, False, True), (961, 512, 256, False, False, False)])
def test_wgrad_direct_kernel_shapes(ops, M, C1, Cout, bn_on, bias, gather):
    """The fragment-order weight-gradient kernel (linear_direct.cu) on the shapes the tiled kernels do not cover: Cout = 13 / 128,
    6 input channels, bias gradient, gathered rows (Upsampling.lin), ragged row counts; the last two cases are wide outputs over few
    rows (linear_small.cu wgrad_rows_kernel), one of them gathered."""
    g = torch.Generator(device="cuda").manual_seed(M + C1 + Cout)
    rn = lambda *s: torch.randn(*s, generator=g, device="cuda")   # noqa: E731
    B, rows_dst = 3, M // 3 if M % 3 == 0 else M
    if M % 3:
        B = 1
    rows_src = max(rows_dst // 4, 1)
    Xs = rn(B * rows_src if gather else M, C1)
    idx = torch.randint(0, rows_src, (M,), generator=g, device="cuda") if gather else None
    W = rn(Cout, C1) / C1 ** 0.5
    dY = rn(M, Cout)
    A = Xs.double()
    if gather:
        A = A[(torch.arange(M, device="cuda") // rows_dst) * rows_src + idx]
    H = bn = None
    dH = dY.double()
    if bn_on:
        H = (A @ W.double().t()).float()
        gamma, beta = 1 + 0.2 * rn(Cout), 0.2 * rn(Cout)
        bn = ops.BN(Cout, dY.device)
        st = torch.cat([H.double().sum(0), (H.double() ** 2).sum(0)]).float()
        bn.stats.zero_(); bn.stats.view(-1, 2 * Cout)[0].copy_(st)
        ops.bn_finalize_fwd(bn, M, gamma, beta, 1e-5, 0.1, True, None, None)
        Hd = H.double()
        mu, var = Hd.mean(0), Hd.var(0, unbiased=False)
        Hh = (Hd - mu) * (var + 1e-5).rsqrt()
        dV = torch.where(Hh * gamma.double() + beta.double() > 0, dY.double(), dY.double() * 0.1)
        ops.bn_backward_prepare(dY, H, bn, 0.1, torch.zeros(Cout, device="cuda"), torch.zeros(Cout, device="cuda"))
        dH = gamma.double() * (var + 1e-5).rsqrt() * (dV - dV.mean(0) - Hh * (dV * Hh).mean(0))
    dW = torch.zeros_like(W)
    db = torch.zeros(Cout, device="cuda") if bias else None
    ops.linear_bwd(dY, H, bn, 0.1 if bn_on else 1.0, Xs, W, idx1=idx, rows_dst=rows_dst, rows_src=rows_src, dW=dW, dbias=db)
    assert rel(dW, dH.t() @ A) < 2e-5
    if bias:
        assert rel(db, dH.sum(0)) < 2e-5


@pytest.mark.parametrize("M,K,N,pro", [(245760, 32, 128, False), (20001, 32, 128, True), (9000, 8, 64, True), (8200, 6, 32, False), (8193, 16, 128, False)])
def test_up_projection_forward(ops, M, K, N, pro):
    """Narrow input → wide output (linear_direct.cu upproj_kernel): H, Σ / Σ² statistics, optional BN + LeakyReLU prologue."""
    g = torch.Generator(device="cuda").manual_seed(M + K + N)
    rn = lambda *s: torch.randn(*s, generator=g, device="cuda")   # noqa: E731
    X, W = rn(M, K) + 0.3, rn(N, K) / K ** 0.5
    sc, sh = (1 + 0.2 * rn(K), 0.2 * rn(K)) if pro else (None, None)
    bn = ops.BN(N, X.device)
    H = ops.linear_fwd(X, W, scale1=sc, shift1=sh, slope1=0.1, stats=bn.stats)
    A = X.double()
    if pro:
        A = lrelu(A * sc.double() + sh.double(), 0.1)
    Hr = A @ W.double().t()
    assert rel(H, Hr) < 2e-5
    st = bn.stats.view(-1, 2 * N).double().sum(0)
    assert rel(st[:N], Hr.sum(0)) < 1e-4 and rel(st[N:], (Hr ** 2).sum(0)) < 1e-4


@pytest.mark.parametrize("M,Cout,Ktot", [(245760, 13, 128), (20001, 19, 128), (9000, 8, 64), (8200, 13, 32)])
def test_up_projection_input_gradient(ops, M, Cout, Ktot):
    """dX = dY[M, Cout]·W[Cout, Ktot] of a plain Linear with few outputs (the class head) — same kernel, weights indexed transposed."""
    g = torch.Generator(device="cuda").manual_seed(M + Cout)
    rn = lambda *s: torch.randn(*s, generator=g, device="cuda")   # noqa: E731
    dY, W, X = rn(M, Cout), rn(Cout, Ktot) / Ktot ** 0.5, rn(M, Ktot)
    dX = torch.full((M, Ktot), float("nan"), device="cuda")
    ops.linear_bwd(dY, None, None, 1.0, X, W, dX1=dX)
    assert rel(dX, dY.double() @ W.double()) < 2e-5
