"""Parity of the CRF layer vs the CPU oracle at several sizes, for the fast (bf16x3) and generic (3xTF32) contraction paths."""
import os, sys, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1:
    import torch
    from oracle import layers as ol, native as on, synthetic
    from crfconv_b200.continuous_crf_conv_big import ContinuousGaussianCRFConv
    N, steps = int(sys.argv[1]), int(sys.argv[2])
    B = 2
    knn = (lambda s, q, k: on.ref_knn_batch(s, q, k, omp=True)) if on.have_ref_knn() else on.knn_batch
    inp = synthetic.crf_layer_inputs(B, N, 16, 128, 64, 4, seed=N, knn_batch_fn=knn)
    torch.manual_seed(0)
    mo = ol.ContinuousGaussianCRFConv(128, 64, 64, steps=steps).train()
    with torch.no_grad():
        mo.c.add_(0.1 * torch.randn(16, 16))
    mp = ContinuousGaussianCRFConv(128, 64, 64, steps=steps); mp.load_state_dict(mo.state_dict()); mp = mp.cuda().train()
    u0, p0 = inp.unary.clone().requires_grad_(True), inp.pairwise.clone().requires_grad_(True)
    u1, p1 = inp.unary.cuda().requires_grad_(True), inp.pairwise.cuda().requires_grad_(True)
    cot = torch.randn(B, N, 64, generator=torch.Generator().manual_seed(1))
    (mo(u0, p0, inp.up_idx, inp.neighbor_idx) * cot).sum().backward()
    (mp(u1, p1, inp.up_idx.cuda(), inp.neighbor_idx.cuda()) * cot.cuda()).sum().backward()
    floor = 1e-3 * max(float(p.grad.abs().max()) for p in mo.parameters())
    rel = lambda a, b, f=0.0: float((a.cpu() - b).abs().max() / max(float(b.abs().max()), f))
    errs = {"d_unary": rel(u1.grad, u0.grad), "d_pair": rel(p1.grad, p0.grad)}
    po = dict(mo.named_parameters())
    for n, p in mp.named_parameters():
        errs[n] = rel(p.grad, po[n].grad, floor)
    worst = max(errs, key=errs.get)
    print(f"N={N} T={steps} mode={os.environ.get('CRFCONV_FORCE_GENERIC','0')}: worst {worst} {errs[worst]:.2e}")
else:
    for N, T in ((512, 3), (4096, 1), (4096, 3), (40960, 1), (40960, 3)):
        for gen in ("0", "1"):
            env = dict(os.environ, CRFCONV_FORCE_GENERIC=gen, CRFCONV_FAST_MIN_ROWS="0")
            subprocess.run([sys.executable, __file__, str(N), str(T)], env=env)
