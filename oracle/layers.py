"""TEST INFRASTRUCTURE ONLY (oracle).  Plain-PyTorch (CPU, fp32 or fp64) restatement of the dense-API layers of the
reference hot path.  It is the checker for the CUDA layers and the ``port`` CPU baseline of bench.py; the product
package never imports it.

Restated from (all paths relative to /root/reference):
  * models/common.py:26-40                      MLP = Linear(bias = not bn) → FastBatchNorm1d → activation
  * torch_points3d.core.common_modules.FastBatchNorm1d  (third-party, NOT vendored in the reference and unpinned —
    restated from its published behaviour: wraps ``nn.BatchNorm1d(C, momentum=0.1)`` as attribute ``batch_norm``;
    [B,N,C] input ⇒ statistics over B·N per channel; [N,C] input ⇒ plain BatchNorm1d)
  * models/point_conv_big.py:8-58               PointConv            (depthwise-separable continuous conv)
  * models/point_conv_big.py:61-88              ResNetBBlock
  * models/point_conv_big.py:91-107             Upsampling
  * models/point_conv_big.py:110-167            PointConvResNet
  * models/continuous_crf_conv_big.py:7-78      ContinuousGaussianCRFConv (dense)

Parameter / buffer names are identical to the reference's so state_dicts are interchangeable; the forward passes are
written with advanced indexing instead of the reference's ``gather`` on a repeated int64 index, and the reference's
B=1 ``.squeeze()`` defect (continuous_crf_conv_big.py:43) is not reproduced.  Pinned against the imported reference
modules by tests/golden/make_golden.py (container only) → tests/golden/*.npz, checked in tests/test_oracle_pinning.py.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F


class FastBatchNorm1d(nn.Module):
    def __init__(self, num_features, momentum=0.1, **kwargs):
        super().__init__()
        self.batch_norm = nn.BatchNorm1d(num_features, momentum=momentum, **kwargs)

    def forward(self, x):
        if x.dim() == 2:
            return self.batch_norm(x)
        if x.dim() == 3:   # [B, N, C]: statistics over B*N
            return self.batch_norm(x.transpose(1, 2)).transpose(1, 2)
        raise ValueError("Non supported number of dimensions {}".format(x.dim()))


class MLP(nn.Module):                                   # common.py:26-40
    def __init__(self, in_channels, out_channels, bn=True, activation=None):
        super().__init__()
        self.lin = nn.Linear(in_channels, out_channels, bias=not bn)
        self.bn = FastBatchNorm1d(out_channels) if bn else None
        self.activation = activation

    def forward(self, x):
        x = self.lin(x)
        if self.bn is not None:
            x = self.bn(x)
        if self.activation is not None:
            x = self.activation(x)
        return x


def take_rows(x, idx):
    """x [B,N,F], idx [B,N',K] int64 → [B,N',K,F]   (point_conv_big.py:25-35, continuous_crf_conv_big.py:38-43)."""
    B = x.shape[0]
    b = torch.arange(B, device=x.device).view(B, 1, 1)
    return x[b, idx]


def _lrelu01():
    return nn.LeakyReLU(negative_slope=0.1)


class PointConv(nn.Module):                             # point_conv_big.py:8-58
    def __init__(self, d_model):
        super().__init__()
        self.weight_nn = nn.Sequential(MLP(3, d_model, activation=_lrelu01()), MLP(d_model, d_model, activation=None))

    def forward(self, x, pos, neighbor_idx):
        support, centres = (pos, pos) if torch.is_tensor(pos) else pos
        B, Nq, K = neighbor_idx.shape
        rel = centres.unsqueeze(2) - take_rows(support, neighbor_idx)              # centre − neighbour  (:40)
        w = self.weight_nn(rel.reshape(B, Nq * K, 3)).reshape(B, Nq, K, -1)       # BN statistics over all B·N'·K edges (:43)
        return (w * take_rows(x, neighbor_idx)).sum(dim=2)                         # (:56-57)


class ResNetBBlock(nn.Module):                          # point_conv_big.py:61-88
    negative_slope = 0.01                               # F.leaky_relu default (:88); a class attribute so tests can remove the kink

    def __init__(self, in_channels, out_channels):
        super().__init__()
        hidden = out_channels // 4
        self.lin_in = MLP(in_channels, hidden, activation=_lrelu01())
        self.lin_out = MLP(hidden, out_channels, activation=None)
        self.shortcut = MLP(in_channels, out_channels, activation=None) if in_channels != out_channels else nn.Identity()
        self.point_conv = PointConv(hidden)

    def forward(self, x, pos, neighbor_idx):
        residual = self.shortcut(x)
        if not torch.is_tensor(pos):                                               # strided block: max over neighbours (:74-77,81-82)
            residual = take_rows(residual, neighbor_idx).max(dim=2)[0]
        x = self.lin_out(self.point_conv(self.lin_in(x), pos, neighbor_idx))
        return F.leaky_relu(x + residual, self.negative_slope)                    # default slope 0.01 (:88)


class Upsampling(nn.Module):                            # point_conv_big.py:91-107
    def __init__(self, down_channels, up_channels, out_channels):
        super().__init__()
        self.lin = MLP(down_channels, up_channels, activation=_lrelu01())
        self.fusion = MLP(up_channels * 2, out_channels, activation=_lrelu01())

    def forward(self, x_down, x_up, up_idx, neighbor_idx=None):
        x_down = self.lin(take_rows(x_down, up_idx)[:, :, 0])
        return self.fusion(torch.cat([x_up, x_down], dim=-1))


class ContinuousGaussianCRFConv(nn.Module):             # continuous_crf_conv_big.py:7-78
    def __init__(self, unary_channels, pairwise_channels, out_channels=None, steps=1):
        super().__init__()
        self.unary_channels = unary_channels
        self.pairwise_channels = pairwise_channels
        self.out_channels = out_channels if out_channels is not None else pairwise_channels
        self.hidden_channels = self.out_channels // 4
        self.steps = steps
        h = self.hidden_channels
        self.unary_nn = nn.Sequential(MLP(unary_channels, h, activation=_lrelu01()), MLP(h, h, activation=None))
        self.pairwise_nn = nn.Sequential(MLP(pairwise_channels, h, activation=_lrelu01()), MLP(h, h, activation=None))
        self.out_nn = MLP(h, self.out_channels, activation=_lrelu01())
        self.fusion_nn = MLP(self.out_channels * 2, self.out_channels, activation=_lrelu01())
        self.c = nn.Parameter(torch.eye(h))                                        # nn.init.eye_ (:35-36)

    def forward(self, unary, pairwise, up_idx, neighbor_idx):
        nbr = neighbor_idx[:, :, 1:]                                               # drop column 0 (assumed self) (:45-47,57)
        u = self.unary_nn(unary)
        y = self.pairwise_nn(pairwise)
        z = take_rows(u, up_idx)[:, :, 0]                                          # nearest-coarse upsample (:60)
        d = (y.unsqueeze(2) - take_rows(y, nbr)).pow(2).sum(dim=-1, keepdim=True)  # (:49-54)
        s = (-d).softmax(dim=2)
        eye = torch.eye(self.hidden_channels, dtype=z.dtype, device=z.device)
        Cm = self.c.t() @ self.c                                                   # (:66)
        Minv = torch.linalg.inv(eye + Cm)
        x = z
        for _ in range(self.steps):                                                # (:68-72)
            m = (s * take_rows(x, nbr)).sum(dim=2)
            x = (z + m @ Cm) @ Minv
        x = self.out_nn(x)
        return self.fusion_nn(torch.cat([x, pairwise], dim=-1))                    # (:76)


class PointConvResNet(nn.Module):                       # point_conv_big.py:110-167
    def __init__(self, in_channels, n_classes, use_crf=True, steps=1):
        super().__init__()
        L = [32, 64, 128, 256, 512]
        self.C = n_classes
        prev = in_channels
        for lvl, ch in enumerate(L, start=1):
            setattr(self, f"conv{lvl}_1", ResNetBBlock(prev, ch))
            setattr(self, f"conv{lvl}_2", ResNetBBlock(ch, ch))
            prev = ch
        for lvl in (4, 3, 2, 1):
            down, up = L[lvl], L[lvl - 1]
            setattr(self, f"deconv{lvl}",
                    ContinuousGaussianCRFConv(down, up, up, steps=steps) if use_crf else Upsampling(down, up, up))
        self.classifier = nn.Sequential(MLP(L[0], L[0] * 4, activation=_lrelu01()), nn.Dropout(p=0.5),
                                        nn.Linear(L[0] * 4, n_classes))

    def forward(self, data):
        x, ms = data.x, data.multiscale
        skips = []
        for lvl in range(1, 6):
            if lvl == 1:
                x = self.conv1_1(x, ms[0].pos, ms[0].neighbor_idx)
            else:
                x = getattr(self, f"conv{lvl}_1")(x, (ms[lvl - 2].pos, ms[lvl - 1].pos), ms[lvl - 2].sub_idx)
            x = getattr(self, f"conv{lvl}_2")(x, ms[lvl - 1].pos, ms[lvl - 1].neighbor_idx)
            skips.append(x)
        for lvl in (4, 3, 2, 1):
            x = getattr(self, f"deconv{lvl}")(x, skips[lvl - 1], ms[lvl - 1].up_idx, ms[lvl - 1].neighbor_idx)
        return self.classifier(x).reshape(-1, self.C)



class EdgeListCRFConv(nn.Module):                         # continuous_crf_conv.py:72-133 (PyG family), restated without PyG
    """`torch_geometric.utils.softmax(src, index)` = exp(src − max_group) / (Σ_group + 1e-16); `scatter_add` = index_add."""

    def __init__(self, unary_channels, pairwise_channels, hidden_channels=None, out_channels=None, steps=1):
        super().__init__()
        self.out_channels = out_channels if out_channels is not None else pairwise_channels
        self.hidden_channels = hidden_channels if hidden_channels is not None else self.out_channels // 4
        self.steps = steps
        h, o = self.hidden_channels, self.out_channels
        self.unary_net = nn.Sequential(nn.Linear(unary_channels, h, bias=False), nn.BatchNorm1d(h))                    # :85-88
        self.pairwise_net = nn.Sequential(nn.Linear(pairwise_channels, h, bias=False), nn.BatchNorm1d(h))              # :89-92
        self.mlp = nn.Sequential(nn.Linear(h, o, bias=False), nn.BatchNorm1d(o), nn.LeakyReLU(inplace=True))           # :93-97
        self.fusion_net = nn.Sequential(nn.Linear(o * 2, o, bias=False), nn.BatchNorm1d(o), nn.LeakyReLU(inplace=True))  # :99-103
        self.c = nn.Parameter(torch.eye(h))                                                                            # :105,109-110

    def forward(self, x, y, pos, edge_index):
        N = pos.shape[0]
        i, j = edge_index
        x = self.unary_net(x)
        s = self.pairwise_net(y)
        s = -torch.sum((s[i] - s[j]) ** 2, dim=1, keepdim=True)                                                        # :117-118
        mx = torch.full((N, 1), float("-inf"), dtype=s.dtype).scatter_reduce(0, i[:, None], s, reduce="amax", include_self=True)
        ex = torch.exp(s - mx[i])
        den = torch.zeros((N, 1), dtype=s.dtype).index_add(0, i, ex)
        s = ex / (den[i] + 1e-16)
        z = x
        eye = torch.eye(self.hidden_channels, dtype=x.dtype)
        C = torch.mm(self.c.t(), self.c)
        for _ in range(self.steps):                                                                                    # :124-128
            x = torch.zeros_like(z).index_add(0, i, s * x[j])
            x = z + torch.mm(x, C)
            x = torch.mm(x, torch.linalg.inv(eye + C))
        x = self.mlp(x)
        return self.fusion_net(torch.cat([x, y], dim=-1))                                                              # :130-131



class EdgeListPointConv(nn.Module):                       # point_conv.py:12-66 (PyG family), restated without PyG
    """MessagePassing defaults: flow source_to_target (edge_index[0] = j, edge_index[1] = i), aggr = add."""

    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        h = out_channels // 4
        self.mlp1 = nn.Sequential(nn.Linear(3, h, bias=False), nn.BatchNorm1d(h), nn.LeakyReLU(inplace=True),
                                  nn.Linear(h, h, bias=False), nn.BatchNorm1d(h))                                      # :21-27
        self.mlp2 = nn.Sequential(nn.Linear(in_channels, h, bias=False), nn.BatchNorm1d(h), nn.LeakyReLU(inplace=True))   # :28-32
        self.mlp3 = nn.Sequential(nn.Linear(h, out_channels, bias=False), nn.BatchNorm1d(out_channels))                 # :33-36
        if in_channels != out_channels:
            self.mlp4 = nn.Sequential(nn.Linear(in_channels, out_channels), nn.BatchNorm1d(out_channels))              # :37-41

    def forward(self, x, pos, edge_index):
        src, dst = edge_index
        if torch.is_tensor(pos):                                                                                       # :46-48
            keep = src != dst
            loops = torch.arange(pos.size(0))
            src, dst = torch.cat([src[keep], loops]), torch.cat([dst[keep], loops])
            pos_src, pos_dst, n_dst = pos, pos, pos.size(0)
            residual = x
        else:                                                                                                          # :50-53
            pos_src, pos_dst = pos
            n_dst = pos_dst.size(0)
            residual = torch.full((n_dst, x.shape[1]), float("-inf"), dtype=x.dtype).scatter_reduce(
                0, dst[:, None].expand(-1, x.shape[1]), x[src], reduce="amax", include_self=True)
        if self.in_channels != self.out_channels:
            residual = self.mlp4(residual)
        x = self.mlp2(x)
        msg = self.mlp1(pos_dst[dst] - pos_src[src]) * x[src]                                                          # :61-65
        x = torch.zeros((n_dst, x.shape[1]), dtype=x.dtype).index_add(0, dst, msg)
        x = self.mlp3(x)
        return F.leaky_relu(x + residual)                                                                              # :58


def _group_softmax(src, index, num_nodes):                     # torch_geometric.utils.softmax
    mx = torch.full((num_nodes,) + src.shape[1:], -float("inf"), dtype=src.dtype).scatter_reduce(0, index.view(-1, *[1] * (src.dim() - 1)).expand_as(src), src, "amax")
    ex = (src - mx[index]).exp()
    den = torch.zeros((num_nodes,) + src.shape[1:], dtype=src.dtype).index_add_(0, index, ex)
    return ex / (den[index] + 1e-16)


class GuideGaussianCRFConv(nn.Module):                        # continuous_crf_conv.py:9-69 restated without PyG; the graph is an input
    def __init__(self, in_n_channels, in_e_channels, out_channels=None, radius=0.1, kernel_size=32, steps=1):
        super().__init__()
        self.out_channels = out_channels if out_channels is not None else in_e_channels
        self.radius, self.kernel_size, self.steps = radius, kernel_size, steps
        self.unary = nn.Sequential(nn.Linear(in_n_channels, self.out_channels, bias=False), nn.BatchNorm1d(self.out_channels))
        self.pairwise = nn.Sequential(nn.Linear(in_e_channels, self.out_channels, bias=False), nn.BatchNorm1d(self.out_channels),
                                      nn.LeakyReLU(inplace=True))
        self.c = nn.Parameter(torch.Tensor(self.out_channels, self.out_channels))
        nn.init.eye_(self.c)

    def forward(self, x, y, pos, batch=None, edge_index=None):
        N = pos.shape[0]
        col, row = edge_index[0], edge_index[1]                # :52  col, row = radius_graph(...)
        x = self.unary(x)
        y = self.pairwise(y)
        s = torch.sum((y[row] - y[col]) ** 2, dim=1, keepdim=True)
        s = _group_softmax(-s, row, N)
        z = x
        I = torch.eye(self.out_channels, dtype=x.dtype)
        C = torch.mm(self.c.t(), self.c)
        for _ in range(self.steps):
            x = s * x[col]
            x = torch.zeros((N, self.out_channels), dtype=x.dtype).index_add_(0, row, x)
            x = z + torch.mm(x, C)
            x = torch.mm(x, (I + C).inverse())
        return F.leaky_relu(x)


class DiscreteCRFConv(nn.Module):                             # discrete_crf_conv.py:11-63 restated without PyG; the graph is an input
    def __init__(self, n_channels, e_channels, hidden_channels=64, num_kernels=5, radius=0.2, kernel_size=32, steps=5):
        super().__init__()
        self.num_kernels, self.steps = num_kernels, steps
        self.F = nn.Parameter(torch.Tensor(num_kernels, e_channels, hidden_channels))
        self.W = nn.Parameter(torch.Tensor(num_kernels, 1))
        self.C = nn.Parameter(torch.Tensor(n_channels, n_channels))
        nn.init.uniform_(self.F)
        nn.init.constant_(self.W, 1 / num_kernels)
        nn.init.eye_(self.C)

    def forward(self, pos, p, f=None, batch=None, edge_index=None):
        N = pos.shape[0]
        col, row = edge_index[0], edge_index[1]
        u = -torch.log(p)
        f = f.unsqueeze(0).repeat(self.num_kernels, 1, 1)
        f = torch.bmm(f, self.F).permute((1, 0, 2))
        f = f[col] - f[row]
        w = torch.exp(-torch.sum(f ** 2, dim=-1))
        w = torch.mm(w, self.W)
        q = p
        for _ in range(self.steps):
            q = torch.zeros_like(p).index_add_(0, row, q[col] * w)
            q = torch.mm(q, self.C)
            q = torch.softmax(-u - q, dim=-1)
        return q
