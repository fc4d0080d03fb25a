import os, sys, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import layers as ol, native as on, synthetic
from crfconv_b200.point_conv_big import PointConvResNet
from tests.golden.make_golden import _perturb
B, N = 2, int(sys.argv[1]) if len(sys.argv) > 1 else 4096
pos = synthetic.room_cloud(B, N, seed=40)
ms = synthetic.build_multiscale(pos, on.knn_batch, num_scales=5, K=16, seed=41)
torch.manual_seed(42)
net = PointConvResNet(6, 13, use_crf=True, steps=1); _perturb(net, 43); net.classifier[1].p = 0.0
onet = ol.PointConvResNet(6, 13, use_crf=True, steps=1); onet.load_state_dict(net.state_dict()); onet.classifier[1].p = 0.0
if os.environ.get("SMOOTH"):
    import torch.nn as nn
    from crfconv_b200 import point_conv_big as pcb
    for m in list(onet.modules()) + list(net.modules()):
        if isinstance(m, nn.LeakyReLU): m.negative_slope = 1.0
        if isinstance(m, (pcb.ResNetBBlock, ol.ResNetBBlock)): m.negative_slope = 1.0
onet = onet.double().train(); net = net.cuda().train()
g = torch.Generator().manual_seed(44)
x = torch.cat([torch.from_numpy(pos), torch.rand(B, N, 3, generator=g)], -1)
y = torch.randint(0, 13, (B * N,), generator=g)
acts = {}
def hook(name, store):
    def f(m, i, o): store[name] = o.detach()
    return f
so, sp = {}, {}
for n, m in onet.named_children(): m.register_forward_hook(hook(n, so))
for n, m in net.named_children(): m.register_forward_hook(hook(n, sp))
d0 = types.SimpleNamespace(x=x.double(), multiscale=[types.SimpleNamespace(pos=l.pos.double(), neighbor_idx=l.neighbor_idx, sub_idx=l.sub_idx, up_idx=l.up_idx) for l in ms])
d1 = types.SimpleNamespace(x=x.cuda(), multiscale=[types.SimpleNamespace(pos=l.pos.cuda(), neighbor_idx=l.neighbor_idx.cuda(), sub_idx=l.sub_idx.cuda(), up_idx=l.up_idx.cuda()) for l in ms])
lo = onet(d0); torch.nn.functional.cross_entropy(lo, y).backward()
lp = net(d1); torch.nn.functional.cross_entropy(lp, y.cuda()).backward()
print("forward activations (max-rel):")
for n in so:
    if n not in sp: continue
    a, b = sp[n].cpu().double(), so[n]
    print(f"  {n:12s} {float((a-b).abs().max()/b.abs().max()):.2e}")
print("param grads (L2 rel), from the loss backwards:")
po = dict(onet.named_parameters())
for n, p in reversed(list(net.named_parameters())):
    a, b = p.grad.cpu().double(), po[n].grad
    e = float((a-b).norm()/max(float(b.norm()), 1e-30))
    if e > 1e-4 or n.endswith("lin.weight"): print(f"  {n:50s} {e:.2e}  |g|={float(b.norm()):.2e}")
