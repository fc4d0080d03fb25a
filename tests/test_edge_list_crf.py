"""Edge-list (PyG-family) continuous-CRF layer, models/continuous_crf_conv.py:72-133: host logic on the CPU, parity on the GPU."""
import numpy as np
import pytest
import torch

from oracle import layers as ol
from oracle import native as on
from tests._util import rel_err, rel_err_trimmed, rel_l2

TOL = 1e-3      # relative fp32, as for the dense layers


def _knn_edges(pos, k):
    """knn_graph(pos, k) without self loops: for every target i its k nearest other points as sources."""
    idx = on.knn_batch(pos[None].numpy(), pos[None].numpy(), k + 1)[0][:, 1:]            # drop the self match
    i = torch.arange(pos.shape[0]).repeat_interleave(k)
    return torch.stack([i, torch.from_numpy(idx.astype(np.int64)).reshape(-1)])


def test_dense_neighbour_table_from_edge_lists():
    from crfconv_b200.continuous_crf_conv import dense_neighbours
    ei = torch.tensor([[0, 0, 1, 1, 2, 2], [1, 2, 0, 2, 0, 1]])
    nbr, padded = dense_neighbours(ei, 3)
    assert not padded and nbr.shape == (1, 3, 3) and nbr[0].tolist() == [[0, 1, 2], [1, 0, 2], [2, 0, 1]]
    perm = torch.tensor([4, 0, 2, 5, 1, 3])                                              # same graph, edges shuffled: grouped stably
    nbr, padded = dense_neighbours(ei[:, perm], 3)
    assert not padded and nbr[0].tolist() == [[0, 1, 2], [1, 0, 2], [2, 0, 1]]
    # ragged in-degrees (2, 1, 0): padded to 2 with the index of the extra node (3), which gets a row of its own
    nbr, padded = dense_neighbours(torch.tensor([[0, 1, 0], [1, 0, 2]]), 3)
    assert padded and nbr[0].tolist() == [[0, 1, 2], [1, 0, 3], [2, 3, 3], [3, 3, 3]]
    # same edge count as a regular graph but uneven: must not be mistaken for one
    nbr, padded = dense_neighbours(torch.tensor([[0, 0, 0, 1, 2, 2], [1, 2, 1, 0, 0, 1]]), 3)
    assert padded and nbr.shape == (1, 4, 4) and nbr[0, 0].tolist() == [0, 1, 2, 1] and nbr[0, 1].tolist() == [1, 0, 3, 3]


def test_state_dict_keys_match_the_reference_module():
    from crfconv_b200.continuous_crf_conv import ContinuousGaussianCRFConv
    ours = ContinuousGaussianCRFConv(64, 32, steps=2)
    ref = ol.EdgeListCRFConv(64, 32, steps=2)
    assert list(ours.state_dict().keys()) == list(ref.state_dict().keys())
    assert ours.hidden_channels == 8 and ours.out_channels == 32


@pytest.mark.gpu
@pytest.mark.parametrize("N,Cu,Co,k,steps,ragged", [(3000, 64, 32, 16, 1, False), (2048, 128, 64, 15, 2, False), (2500, 64, 32, 12, 2, True)])
def test_edge_list_layer_vs_oracle(N, Cu, Co, k, steps, ragged):
    from crfconv_b200.continuous_crf_conv import ContinuousGaussianCRFConv
    g = torch.Generator().manual_seed(N)
    pos = torch.rand(N, 3, generator=g)
    x, y = torch.randn(N, Cu, generator=g), torch.randn(N, Co, generator=g)
    ei = _knn_edges(pos, k)
    if ragged:                                      # radius-graph-like: drop a random 40 % of the edges (some nodes keep none) and shuffle
        keep = torch.rand(ei.shape[1], generator=g) > 0.4
        keep[: 3 * k] = False                       # nodes 0..2 end up with no incoming edge at all
        ei = ei[:, keep]
        ei = ei[:, torch.randperm(ei.shape[1], generator=g)]
    torch.manual_seed(1)
    mo = ol.EdgeListCRFConv(Cu, Co, steps=steps).train()
    with torch.no_grad():
        mo.c.add_(0.1 * torch.randn(mo.c.shape, generator=g))
        for n, p in mo.named_parameters():
            if n.endswith("1.weight"):
                p.copy_(1 + 0.2 * torch.randn(p.shape, generator=g))
            elif n.endswith("1.bias"):
                p.copy_(0.2 * torch.randn(p.shape, generator=g))
    mp = ContinuousGaussianCRFConv(Cu, Co, steps=steps)
    mp.load_state_dict(mo.state_dict())
    mp = mp.cuda().train()
    xc, yc = x.clone().requires_grad_(True), y.clone().requires_grad_(True)
    xg, yg = x.clone().cuda().requires_grad_(True), y.clone().cuda().requires_grad_(True)
    oo = mo(xc, yc, pos, ei)
    cot = torch.randn(oo.shape, generator=g)
    (oo * cot).sum().backward()
    og = mp(xg, yg, pos.cuda(), ei.cuda())
    (og * cot.cuda()).sum().backward()
    errs = {"out": rel_err_trimmed(og.detach().cpu().numpy(), oo.detach().numpy()), "out(l2)": rel_l2(og.detach().cpu().numpy(), oo.detach().numpy()),
            "dx": rel_err_trimmed(xg.grad.cpu().numpy(), xc.grad.numpy()), "dx(l2)": rel_l2(xg.grad.cpu().numpy(), xc.grad.numpy()),
            "dy": rel_err_trimmed(yg.grad.cpu().numpy(), yc.grad.numpy()), "dy(l2)": rel_l2(yg.grad.cpu().numpy(), yc.grad.numpy())}
    floor = 1e-3 * max(float(p.grad.abs().max()) for p in mo.parameters())
    po = dict(mo.named_parameters())
    loose = {}
    for n, p in mp.named_parameters():
        errs["grad(l2) " + n] = rel_l2(p.grad.cpu().numpy(), po[n].grad.numpy(), floor)
        loose["grad " + n] = rel_err(p.grad.cpu().numpy(), po[n].grad.numpy(), floor)
    for n, b in mp.named_buffers():
        errs["buf " + n] = rel_err(b.cpu().numpy(), dict(mo.named_buffers())[n].numpy())
    bad = {k_: v for k_, v in errs.items() if not v < TOL}
    bad.update({k_: v for k_, v in loose.items() if not v < 5 * TOL})
    assert not bad, bad


# ------------------------------------------------------------------ DepthwiseSeparablePointConv (point_conv.py:12-66)
def test_point_conv_state_dict_keys_and_ragged_rejection():
    from crfconv_b200.point_conv import DepthwiseSeparablePointConv, _regular_table
    ours, ref = DepthwiseSeparablePointConv(32, 64), ol.EdgeListPointConv(32, 64)
    assert list(ours.state_dict().keys()) == list(ref.state_dict().keys())
    assert "mlp4.0.bias" in ours.state_dict()
    same = DepthwiseSeparablePointConv(64, 64)
    assert not hasattr(same, "mlp4")
    t = _regular_table(torch.tensor([1, 2, 0, 2, 0, 1]), torch.tensor([0, 0, 1, 1, 2, 2]), 3)
    assert t.tolist() == [[[1, 2], [0, 2], [0, 1]]]
    with pytest.raises(NotImplementedError):
        _regular_table(torch.tensor([1, 2, 0]), torch.tensor([0, 0, 1]), 3)


def _compare(mo, mp, run_o, run_p, inputs_c, inputs_g, tol=TOL):
    oo = run_o()
    g = torch.Generator().manual_seed(5)
    cot = torch.randn(oo.shape, generator=g)
    (oo * cot).sum().backward()
    og = run_p()
    (og * cot.cuda()).sum().backward()
    errs = {"out": rel_err_trimmed(og.detach().cpu().numpy(), oo.detach().numpy()), "out(l2)": rel_l2(og.detach().cpu().numpy(), oo.detach().numpy())}
    for n, (c, gq) in enumerate(zip(inputs_c, inputs_g)):
        errs[f"din{n}"] = rel_err_trimmed(gq.grad.cpu().numpy(), c.grad.numpy())
        errs[f"din{n}(l2)"] = rel_l2(gq.grad.cpu().numpy(), c.grad.numpy())
    floor = 1e-3 * max(float(p.grad.abs().max()) for p in mo.parameters())
    po = dict(mo.named_parameters())
    loose = {}
    for n, p in mp.named_parameters():
        if n == "mlp4.0.bias":                                    # cancelled by the BatchNorm that follows: analytically zero — zeros here
            assert p.grad is not None and float(p.grad.abs().max()) == 0.0 and float(po[n].grad.abs().max()) < floor   # (rounding noise in the oracle)
            continue
        assert p.grad is not None, n
        errs["grad(l2) " + n] = rel_l2(p.grad.cpu().numpy(), po[n].grad.numpy(), floor)
        loose["grad " + n] = rel_err(p.grad.cpu().numpy(), po[n].grad.numpy(), floor)
    bo = dict(mo.named_buffers())
    for n, b in mp.named_buffers():
        errs["buf " + n] = rel_err(b.cpu().numpy(), bo[n].numpy(), 1e-6)
    bad = {k_: v for k_, v in errs.items() if not v < tol}
    bad.update({k_: v for k_, v in loose.items() if not v < 5 * tol})
    assert not bad, bad


@pytest.mark.gpu
@pytest.mark.parametrize("N,cin,cout,k", [(3000, 32, 64, 16), (2048, 64, 64, 12)])
def test_edge_list_point_conv_symmetric_vs_oracle(N, cin, cout, k):
    from crfconv_b200.point_conv import DepthwiseSeparablePointConv
    g = torch.Generator().manual_seed(N)
    pos = torch.rand(N, 3, generator=g)
    x = torch.randn(N, cin, generator=g)
    ei = _knn_edges(pos, k)                                      # row 0 = target, row 1 = source
    ei = torch.stack([ei[1], ei[0]])                              # PyG flow: row 0 = source j, row 1 = target i
    ei = torch.cat([ei, torch.arange(0, N, 7).repeat(2, 1)], dim=1)   # a few explicit self loops: removed and re-added by the layer
    ei = ei[:, torch.randperm(ei.shape[1], generator=g)]
    torch.manual_seed(2)
    mo = ol.EdgeListPointConv(cin, cout).train()
    mp = DepthwiseSeparablePointConv(cin, cout)
    mp.load_state_dict(mo.state_dict())
    mp = mp.cuda().train()
    xc, xg = x.clone().requires_grad_(True), x.clone().cuda().requires_grad_(True)
    _compare(mo, mp, lambda: mo(xc, pos, ei), lambda: mp(xg, pos.cuda(), ei.cuda()), [xc], [xg])


@pytest.mark.gpu
def test_edge_list_point_conv_bipartite_vs_oracle():
    from crfconv_b200.point_conv import DepthwiseSeparablePointConv
    g = torch.Generator().manual_seed(11)
    Ns, Nd, k, cin, cout = 4000, 1000, 16, 32, 64
    pos_s = torch.rand(Ns, 3, generator=g)
    pos_d = pos_s[torch.randperm(Ns, generator=g)[:Nd]].contiguous()
    x = torch.randn(Ns, cin, generator=g)
    idx = on.knn_batch(pos_s[None].numpy(), pos_d[None].numpy(), k)[0]                   # [Nd, k] sources of every target
    ei = torch.stack([torch.from_numpy(idx.astype(np.int64)).reshape(-1), torch.arange(Nd).repeat_interleave(k)])
    torch.manual_seed(3)
    mo = ol.EdgeListPointConv(cin, cout).train()
    mp = DepthwiseSeparablePointConv(cin, cout)
    mp.load_state_dict(mo.state_dict())
    mp = mp.cuda().train()
    xc, xg = x.clone().requires_grad_(True), x.clone().cuda().requires_grad_(True)
    _compare(mo, mp, lambda: mo(xc, (pos_s, pos_d), ei), lambda: mp(xg, (pos_s.cuda(), pos_d.cuda()), ei.cuda()), [xc], [xg])
