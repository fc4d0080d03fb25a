import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import layers as ol, native as on, synthetic
from crfconv_b200.continuous_crf_conv_big import ContinuousGaussianCRFConv
N, steps = int(sys.argv[1]), int(sys.argv[2])
GENERIC = len(sys.argv) > 3 and sys.argv[3] == 'generic'      # third argument 'generic': the 3xTF32 generic kernels only (crfconv_set_fast_path(0))
if GENERIC:
    from crfconv_b200 import _lib
    _lib.lib().crfconv_set_fast_path(0)
B = 2
knn = (lambda s, q, k: on.ref_knn_batch(s, q, k, omp=True)) if on.have_ref_knn() else on.knn_batch
inp = synthetic.crf_layer_inputs(B, N, 16, 128, 64, 4, seed=N, knn_batch_fn=knn)
torch.manual_seed(0)
mo = ol.ContinuousGaussianCRFConv(128, 64, 64, steps=steps).train()
with torch.no_grad():
    mo.c.add_(0.1 * torch.randn(16, 16))
mp = ContinuousGaussianCRFConv(128, 64, 64, steps=steps); mp.load_state_dict(mo.state_dict()); mp = mp.cuda().train()
mo64 = ol.ContinuousGaussianCRFConv(128, 64, 64, steps=steps).double().train(); mo64.load_state_dict({k: v.double() if v.is_floating_point() else v for k, v in mo.state_dict().items()})
u0, p0 = inp.unary.clone().requires_grad_(True), inp.pairwise.clone().requires_grad_(True)
u1, p1 = inp.unary.cuda().requires_grad_(True), inp.pairwise.cuda().requires_grad_(True)
u2, p2 = inp.unary.double().requires_grad_(True), inp.pairwise.double().requires_grad_(True)
cot = torch.randn(B, N, 64, generator=torch.Generator().manual_seed(1))
o0 = mo(u0, p0, inp.up_idx, inp.neighbor_idx); (o0 * cot).sum().backward()
o1 = mp(u1, p1, inp.up_idx.cuda(), inp.neighbor_idx.cuda()); (o1 * cot.cuda()).sum().backward()
o2 = mo64(u2, p2, inp.up_idx, inp.neighbor_idx); (o2 * cot.double()).sum().backward()
rel = lambda a, b: float((a.detach().cpu().double() - b.detach().double()).abs().max() / float(b.detach().abs().max()))
print(f"N={N} T={steps} generic={int(GENERIC)}   [product vs fp64 oracle | fp32 oracle vs fp64 oracle]")
print(f"  out        {rel(o1, o2):.2e} | {rel(o0, o2):.2e}")
print(f"  d_unary    {rel(u1.grad, u2.grad):.2e} | {rel(u0.grad, u2.grad):.2e}")
print(f"  d_pair     {rel(p1.grad, p2.grad):.2e} | {rel(p0.grad, p2.grad):.2e}")
po, p64 = dict(mo.named_parameters()), dict(mo64.named_parameters())
for n, p in mp.named_parameters():
    print(f"  {n:40s} {rel(p.grad, p64[n].grad):.2e} | {rel(po[n].grad, p64[n].grad):.2e}   max|g|={float(p64[n].grad.abs().max()):.2e}")
