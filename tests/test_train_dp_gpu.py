"""The data-parallel trainer (crfconv_b200/train_dp.py) on one GPU: a few optimisation steps of the full PointConvResNet on the
on-GPU multiscale pyramid must run, decrease the loss on a fixed batch and round-trip through Base.save / Base.load."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_train_steps_reduce_the_loss_and_checkpoint_round_trips(tmp_path):
    from crfconv_b200 import train_dp
    from crfconv_b200.distributed import FlatGradients
    from crfconv_b200.point_conv_big import PointConvResNet
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    model = PointConvResNet(in_channels=6, n_classes=8, use_crf=True, steps=1).to(dev).train()
    grads = FlatGradients(model)
    opt = torch.optim.SGD(model.parameters(), lr=0.05, momentum=0.9, weight_decay=1e-4)
    pos, feats, labels, gen = train_dp.synthetic_shard(2, 4096, 8, dev, seed=7)
    data = train_dp.make_batch(pos, feats, labels, generator=gen)
    losses = [float(train_dp.train_step(model, grads, opt, data)) for _ in range(6)]
    assert all(l == l for l in losses), losses                      # no NaN
    assert losses[-1] < losses[0], losses
    ck = tmp_path / "net.pth"
    model.save(str(ck))
    other = PointConvResNet(in_channels=6, n_classes=8, use_crf=True, steps=1).to(dev)
    other.load(str(ck))
    for (n1, p1), (n2, p2) in zip(model.state_dict().items(), other.state_dict().items()):
        assert n1 == n2 and torch.equal(p1, p2)


def test_graph_replay_matches_eager_gradients_for_the_full_network():
    """crfconv_b200.graphs.GraphedStep: the captured fwd+bwd of PointConvResNet must reproduce the eager gradients, and
    new inputs copied into the captured tensors must change the result."""
    import torch.nn.functional as F
    from crfconv_b200 import train_dp
    from crfconv_b200.distributed import FlatGradients
    from crfconv_b200.graphs import GraphedStep
    from crfconv_b200.point_conv_big import PointConvResNet
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    model = PointConvResNet(in_channels=6, n_classes=8, use_crf=True, steps=1).to(dev).train()
    model.classifier[1].p = 0.0                                       # dropout off: eager and replay must see the same function
    grads = FlatGradients(model)
    pos, feats, labels, gen = train_dp.synthetic_shard(2, 2048, 8, dev, seed=3)
    data = train_dp.make_batch(pos, feats, labels, generator=gen)

    def fwd_bwd():
        grads.zero()
        loss = F.cross_entropy(model(data), data.y.reshape(-1) - 1)
        loss.backward()
        return loss.detach()

    eager_loss = float(fwd_bwd())
    eager = grads.flat.clone()
    step = GraphedStep(fwd_bwd)
    loss = float(step.replay())
    torch.cuda.synchronize()
    assert abs(loss - eager_loss) < 1e-4 * abs(eager_loss)
    scale = float(eager.abs().max())
    assert float((grads.flat - eager).abs().max()) < 2e-3 * scale     # atomics order / kink flips only
    # a different batch through the SAME graph
    pos2, feats2, labels2, gen2 = train_dp.synthetic_shard(2, 2048, 8, dev, seed=4)
    data2 = train_dp.make_batch(pos2, feats2, labels2, generator=gen2)
    GraphedStep.copy_inputs(data, data2)
    loss2 = float(step.replay())
    ref2 = float(F.cross_entropy(model(data2), data2.y.reshape(-1) - 1))
    assert abs(loss2 - ref2) < 1e-3 * abs(ref2) and abs(loss2 - loss) > 1e-6


def test_direct_gradient_accumulation_equals_autograd_accumulation():
    """FlatGradients(direct=True): the backward kernels accumulate into the bound .grad views themselves; the gradients must equal
    the ones autograd accumulates (same kernels, same order — only the per-parameter fill / add launches disappear)."""
    import torch
    import torch.nn.functional as F
    from crfconv_b200 import train_dp
    from crfconv_b200.distributed import FlatGradients
    from crfconv_b200.point_conv_big import PointConvResNet
    dev = torch.device("cuda")
    torch.manual_seed(0)
    net = PointConvResNet(6, 13).to(dev).train()
    net.classifier[1].p = 0.0
    # kink-free (all LeakyReLU slopes 1): two runs of the same network differ in the last bit of their BatchNorm statistics (atomics),
    # and a LeakyReLU kink flip between the runs would swamp the comparison
    from crfconv_b200 import point_conv_big as pcb
    for m in net.modules():
        if isinstance(m, torch.nn.LeakyReLU):
            m.negative_slope = 1.0
        if isinstance(m, pcb.ResNetBBlock):
            m.negative_slope = 1.0
    pos, feat, lab, gen = train_dp.synthetic_shard(2, 2048, 13, dev, seed=5)
    data = train_dp.make_batch(pos, feat, lab, generator=gen)
    sd = {k: v.clone() for k, v in net.state_dict().items()}
    flats = []
    for direct in (False, True):
        net.load_state_dict(sd)
        for p in net.parameters():
            p.grad = None
            if hasattr(p, "_crf_direct_grad"):
                del p._crf_direct_grad
        fg = FlatGradients(net, direct=direct)
        for _ in range(2):                      # second pass: buffers are re-zeroed and reused
            fg.zero()
            F.cross_entropy(net(data), lab.reshape(-1) - 1).backward()
        flats.append(fg.flat.clone())
    a, b = flats
    assert float((a - b).abs().max()) <= 2e-4 * float(a.abs().max()), float((a - b).abs().max() / a.abs().max())
    assert float(b.abs().max()) > 0


@pytest.mark.parametrize("M,C,weighted,ignore", [(245760, 13, True, -1), (5001, 19, False, 3), (7, 8, True, -100), (4096, 40, True, 0)])
def test_cross_entropy_matches_torch(M, C, weighted, ignore):
    """csrc/loss.cu vs F.cross_entropy (fp64): loss, normaliser and gradient for 'mean' and 'sum', class weights, ignore_index."""
    import torch.nn.functional as Fn
    from crfconv_b200 import losses
    g = torch.Generator().manual_seed(M + C)
    x = (3 * torch.randn(M, C, generator=g)).cuda()
    t = torch.randint(0, C, (M,), generator=g).cuda()
    if ignore >= 0:
        t[::5] = ignore
    elif M > 10:
        t[::7] = ignore
    w = (0.5 + torch.rand(C, generator=g)).cuda() if weighted else None
    for red in ("mean", "sum"):
        xa = x.clone().requires_grad_(True)
        xb = x.double().clone().requires_grad_(True)
        la, norm = losses.cross_entropy(xa, t, weight=w, ignore_index=ignore, reduction=red, return_normaliser=True)
        lb = Fn.cross_entropy(xb, t, weight=w.double() if weighted else None, ignore_index=ignore, reduction=red)
        (2.5 * la).backward()
        (2.5 * lb).backward()
        assert abs(float(la) - float(lb)) <= 1e-5 * abs(float(lb))
        valid = t != ignore
        nref = (w[t.clamp(min=0)] * valid).sum().double() if weighted else valid.sum().double()
        assert abs(float(norm) - float(nref)) <= 1e-6 * float(nref)
        err = (xa.grad.double() - xb.grad).abs().max() / xb.grad.abs().max()
        assert float(err) < 1e-5, (red, float(err))
        if bool((~valid).any()):
            assert float(xa.grad[~valid].abs().max()) == 0.0
