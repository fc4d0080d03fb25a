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
