"""CPU, world_size = 2, gloo: host-side logic of the batch-sharded data parallelism — shard ranges, flat
gradient buffer, ONE all-reduce — checked against the single-process result.  A plain torch module stands in for the CUDA
layers (which need a GPU); BatchNorm-free so that sharded and unsharded gradients are identical by linearity."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from crfconv_b200.distributed import FlatGradients, shard_batch, shard_range


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _model():
    torch.manual_seed(0)
    return torch.nn.Sequential(torch.nn.Linear(6, 16), torch.nn.Tanh(), torch.nn.Linear(16, 3))


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        model = _model()
        fg = FlatGradients(model)
        g = torch.Generator().manual_seed(1)
        x, y = torch.randn(5, 40, 6, generator=g), torch.randn(5, 40, 3, generator=g)      # 5 clouds over 2 ranks: ragged shards
        xs, ys = shard_batch((x, y), rank, world)
        fg.zero()
        ((model(xs) - ys) ** 2).sum().backward()          # sum-loss ⇒ summed shard gradients == full-batch gradient
        before = fg.flat.clone()
        out = fg.all_reduce(average=False)
        assert out.data_ptr() == fg.flat.data_ptr()
        assert all(p.grad.data_ptr() >= fg.flat.data_ptr() for p in model.parameters())   # grads are views of the flat buffer
        # adopt mode (bind=False): autograd owns the gradient tensors; they are packed, reduced with ONE collective, copied back
        fa = FlatGradients(model, bind=False)
        fa.zero()
        assert all(p.grad is None for p in model.parameters())
        ((model(xs) - ys) ** 2).sum().backward()
        fa.all_reduce(average=False)
        adopted = torch.cat([p.grad.reshape(-1) for p in model.parameters()])
        assert torch.allclose(adopted, fg.flat, rtol=1e-5, atol=1e-5)
        # a parameter one rank's step never reached (grad None there) contributes zeros and receives the other ranks' sum
        fa.zero()
        if rank == 0:
            ((model(xs) - ys) ** 2).sum().backward()
            mine = [p.grad.clone() for p in model.parameters()]
        else:
            (model[0](xs) ** 2).sum().backward()                    # the second Linear is not reached on rank 1
            assert model[2].weight.grad is None
            mine = [p.grad.clone() if p.grad is not None else torch.zeros_like(p) for p in model.parameters()]
        fa.all_reduce(average=False)
        summed = [m.clone() for m in mine]
        for m in summed:
            dist.all_reduce(m)
        for p, m in zip(model.parameters(), summed):
            assert p.grad is not None and torch.allclose(p.grad, m, rtol=1e-6, atol=1e-6)
        q.put((rank, xs.shape[0], before, fg.flat.clone()))
    finally:
        dist.destroy_process_group()


def test_shard_ranges_cover_the_batch_exactly():
    for B in (1, 2, 5, 6, 16, 17):
        for W in (1, 2, 4, 8):
            r = [shard_range(B, k, W) for k in range(W)]
            assert r[0][0] == 0 and r[-1][1] == B and all(r[i][1] == r[i + 1][0] for i in range(W - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1


def _run_world(world):
    """Spawn `world` gloo ranks on a fresh port; one retry covers a lost port race / a slow first `import torch` in the children."""
    last = None
    for _ in range(2):
        port = _free_port()
        ctx = mp.get_context("spawn")
        q = ctx.Queue()
        procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
        for p in procs:
            p.start()
        try:
            res = sorted([q.get(timeout=150) for _ in range(world)], key=lambda t: t[0])
            for p in procs:
                p.join(timeout=60)
            if all(p.exitcode == 0 for p in procs):
                return res
            last = RuntimeError(f"exit codes {[p.exitcode for p in procs]}")
        except Exception as e:      # noqa: BLE001  (queue.Empty on a rendezvous failure)
            last = e
        for p in procs:
            if p.is_alive():
                p.terminate()
            p.join(timeout=10)
    raise last


@pytest.mark.timeout(400)
def test_flat_gradient_allreduce_matches_single_process():
    res = _run_world(2)
    assert [r[1] for r in res] == [3, 2]
    # single-process reference on the whole batch
    model = _model()
    g = torch.Generator().manual_seed(1)
    x, y = torch.randn(5, 40, 6, generator=g), torch.randn(5, 40, 3, generator=g)
    ((model(x) - y) ** 2).sum().backward()
    ref = torch.cat([p.grad.reshape(-1) for p in model.parameters()])
    for _, _, before, after in res:
        assert torch.allclose(after, ref, rtol=1e-5, atol=1e-5)
    assert torch.allclose(res[0][2] + res[1][2], ref, rtol=1e-5, atol=1e-5)
    assert not torch.allclose(res[0][2], ref, rtol=1e-3, atol=1e-3)        # a shard alone is NOT the full gradient


def _warmup_worker(rank, world, port, q):
    """bench.warm_up: rank 1's steps are slower than rank 0's, so their own clocks would stop the warm-up after different step counts;
    every step holds a collective, so the counts MUST agree (rank 0 decides, the decision is broadcast)."""
    import sys
    import time
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        class _NoCuda:                      # warm_up only calls torch.cuda.synchronize(); the rest is the real torch
            def __getattr__(self, name):
                return getattr(torch, name)

            class cuda:
                @staticmethod
                def synchronize():
                    pass
        calls = []

        def step(i):
            time.sleep(0.002 if rank == 0 else 0.0005)        # unsynchronised local work of different length ...
            t = torch.ones(1)
            dist.all_reduce(t)                                 # ... and the step's collective
            calls.append(i)
        n = bench.warm_up(_NoCuda(), dist, world, torch.device("cpu"), step, min_steps=3, min_seconds=0.15, chunk=4)
        t = torch.tensor([float(n)])
        dist.all_reduce(t)                                     # a mismatch would already have dead-locked above; this checks the count
        q.put((rank, n, len(calls), float(t.item())))
    finally:
        dist.destroy_process_group()


def test_bench_warm_up_runs_the_same_number_of_steps_on_every_rank():
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_warmup_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=150) for _ in range(2)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
    assert all(p.exitcode == 0 for p in procs)
    (_, n0, c0, s0), (_, n1, c1, s1) = res
    assert n0 == n1 == c0 == c1 and n0 >= 4 and n0 % 4 == 0 and s0 == s1 == 2.0 * n0
