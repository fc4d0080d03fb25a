#!/usr/bin/env python
"""bench.py — headline benchmark of the CRFConv hot path (BASELINE.json metric: CRFConv fwd+bwd points/s at N=40,960, k=16).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--clouds B] [--impl reference]

One *step* = forward + backward of ONE ``ContinuousGaussianCRFConv(128, 64, 64, steps=1)`` layer (SURVEY.md §8 config S1:
N=40,960 points, Nc=10,240 coarse points, K=16, Cu=128, Cp=Co=64, hidden F=16) over a batch of B synthetic S3DIS-room-shaped
clouds per GPU.  Clouds are independent ⇒ ranks hold different clouds (weak scaling); the only exchange is the NCCL
all-reduce of the layer's flat fp32 parameter gradient (13,440 floats), inside the timed step when N > 1.

value   : points/s with inputs resident in HBM (CUDA-event timed, max over ranks).
e2e     : same metric through the public module API with HOST (pinned) inputs: H2D of unary/pairwise/up_idx/neighbor_idx and
          a D2H read of the loss every step.
roofline: the dominant kernel (per-call CUDA-event timing of every C-ABI launch in a separate instrumented pass) and, as
          `roofline_step`, the whole layer against SURVEY.md §8(d)'s algorithmic bytes (79,298,560 B per cloud fwd+bwd).
cpu_baseline / --impl reference: the CPU port of the reference layer (oracle/layers.py, PyTorch CPU, all host threads) on a
          bounded sample of the same workload.  The reference itself is Python and is not present on the GPU box.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_POINTS, K_NBR, CU, CP, RATIO = 40960, 16, 128, 64, 4
ALGO_BYTES_PER_CLOUD = 4 * (2 * N_POINTS * CP + 3 * (N_POINTS // RATIO) * CU + 3 * N_POINTS * CP) + 16 * N_POINTS * (K_NBR + 1)  # 79,298,560
L2_BYTES = 126 * 2**20


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        time.sleep(0.05)
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i].lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ CPU arm
def cpu_layer_points_per_s(clouds=2, repeats=6, warm=1):
    """Reference layer restated on CPU (oracle/layers.py) fwd+bwd; kNN indices from the compiled reference when shipped."""
    import numpy as np
    import torch
    from oracle import layers as ol
    from oracle import native as on
    from oracle import synthetic
    torch.set_num_threads(os.cpu_count() or 1)
    knn = (lambda s, q, k: on.ref_knn_batch(s, q, k, omp=True)) if on.have_ref_knn() else on.knn_batch
    n = N_POINTS
    inp = synthetic.crf_layer_inputs(clouds, n, K_NBR, CU, CP, RATIO, seed=0, knn_batch_fn=knn)
    torch.manual_seed(0)
    layer = ol.ContinuousGaussianCRFConv(CU, CP, CP, steps=1).train()
    u, p = inp.unary.requires_grad_(True), inp.pairwise.requires_grad_(True)
    ts = []
    for i in range(warm + repeats):
        t0 = time.perf_counter()
        out = layer(u, p, inp.up_idx, inp.neighbor_idx)
        out.sum().backward()
        dt = time.perf_counter() - t0
        if i >= warm:
            ts.append(dt)
        layer.zero_grad(); u.grad = None; p.grad = None
    t = statistics.median(ts)
    return clouds * n / t, t, {"value": clouds * n / t, "unit": "points/s", "cores": torch.get_num_threads(), "kind": "port",
                               "sample": f"{clouds} clouds x {n} points, CRF layer fwd+bwd, median of {repeats} after {warm} warm-up "
                                         f"(oracle/layers.py on torch CPU, {torch.get_num_threads()} threads)"}


def run_reference(args, rank, world):
    if rank != 0:
        return
    steps = max(args.steps, 1)
    v, t, cb = cpu_layer_points_per_s(clouds=2, repeats=min(steps, 8), warm=min(max(args.warmup, 1), 2))
    line = {"impl": "reference", "metric": "CRFConv fwd+bwd points/s (N=40960,k=16)", "value": v, "unit": "points/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "single ContinuousGaussianCRFConv(128,64,64,steps=1) fwd+bwd, N=40960, Nc=10240, K=16; CPU arm: "
                                   "2 clouds per step (bounded sample of the same workload)"},
            "cpu_baseline": cb, "e2e": {"value": v, "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ GPU arm
def make_inputs(torch, B, dev, seed):
    from crfconv_b200 import nearest_neighbors as nn_
    g = torch.Generator(device="cpu").manual_seed(seed)
    pos = (torch.rand(B, N_POINTS, 3, generator=g) * torch.tensor([8.0, 6.0, 3.0])).to(dev)
    nbr = nn_.knn_batch(pos, pos, K_NBR)
    choice = torch.randperm(N_POINTS, generator=g)[: N_POINTS // RATIO].to(dev)
    up = nn_.knn_batch(pos[:, choice].contiguous(), pos, 1)
    unary = torch.randn(B, N_POINTS // RATIO, CU, generator=g).to(dev)
    pair = torch.randn(B, N_POINTS, CP, generator=g).to(dev)
    return {"pos": pos, "unary": unary, "pairwise": pair, "up_idx": up, "neighbor_idx": nbr}


def run_gpu(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: crfconv_b200 has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from crfconv_b200 import ops
    from crfconv_b200.continuous_crf_conv_big import ContinuousGaussianCRFConv
    from crfconv_b200 import nearest_neighbors as nn_

    B = args.clouds
    torch.manual_seed(1234)
    layer = ContinuousGaussianCRFConv(CU, CP, CP, steps=1).to(dev).train()
    with torch.no_grad():
        layer.c.add_(0.1 * torch.randn_like(layer.c))
    from crfconv_b200.distributed import FlatGradients
    # bind=False: autograd adopts the gradient tensors the fused layer returns — views of ONE flat allocation — so there is no
    # accumulate kernel per parameter, and the all-reduce runs in place on that allocation (one collective, no copies)
    fg = FlatGradients(layer, bind=False)

    in_bytes = B * (4 * ((N_POINTS // RATIO) * CU + N_POINTS * CP) + 8 * N_POINTS * (K_NBR + 1))
    nsets = max(2, -(-2 * L2_BYTES // max(in_bytes, 1)))     # rotate enough input sets that a step never finds its inputs in L2
    nsets = min(nsets, 8)
    sets = [make_inputs(torch, B, dev, seed=100 * rank + s) for s in range(nsets)]
    for s in sets:
        s["unary"].requires_grad_(True)
        s["pairwise"].requires_grad_(True)
    cot = torch.ones(B, N_POINTS, CP, device=dev)

    def eager_step(i, reduce=True):
        s = sets[i % nsets]
        fg.zero()
        out = layer(s["unary"], s["pairwise"], s["up_idx"], s["neighbor_idx"])
        out.backward(cot)
        if reduce:
            fg.all_reduce()                       # no-op at world == 1
        s["unary"].grad = None
        s["pairwise"].grad = None

    # The ~50 kernel launches of one fwd+bwd are captured ONCE per input set into a CUDA graph and replayed: with ≈1.4 ms of GPU
    # work per step the Python/ctypes launch path (≈1.8 ms per step) would otherwise be the bottleneck.  The gradient all-reduce
    # stays outside the graph.
    graphs, launches_per_step = [], 0
    if args.graph:
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for i in range(3):
                eager_step(i)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        for k in range(nsets):
            s = sets[k]
            g = torch.cuda.CUDAGraph()
            ops.COUNTERS["launches"] = 0
            with torch.cuda.graph(g):
                fg.zero()
                out = layer(s["unary"], s["pairwise"], s["up_idx"], s["neighbor_idx"])
                out.backward(cot)
            launches_per_step = ops.COUNTERS["launches"]
            s["unary"].grad = None
            s["pairwise"].grad = None
            graphs.append(g)

    def step(i):
        if graphs:
            graphs[i % nsets].replay()
            fg.all_reduce()
            ops.COUNTERS["launches"] += launches_per_step
        else:
            eager_step(i)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # clocks are sampled (20 ms period) from before the warm-up to the end of the timed region: same load throughout; the warm-up
    # runs for at least 0.5 s so that nvidia-smi is up and several samples land under load
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    t_w = time.perf_counter()
    i = 0
    while i < max(args.warmup, 3) or time.perf_counter() - t_w < 0.5:
        step(i)
        i += 1
        if i % 16 == 0:
            torch.cuda.synchronize()
    barrier()
    ops.COUNTERS["launches"] = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(args.steps):
        step(i)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = ops.COUNTERS["launches"]
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms_step = ms / args.steps
    value = world * B * N_POINTS / (ms_step * 1e-3)

    # ---- e2e: host (pinned) inputs, H2D every step, D2H of the loss
    host = [{k: v.detach().cpu().pin_memory() for k, v in s.items() if k != "pos"} for s in sets[:2]]
    h2d = sum(v.numel() * v.element_size() for v in host[0].values())
    loss_host = torch.zeros(1).pin_memory()

    # Two device-side staging buffers: the H2D copy of step i+1 runs on a copy stream while step i computes (the inputs of
    # every step still cross PCIe inside the timed region); fwd+bwd of each buffer is a captured CUDA graph when --graph.
    copy_stream = torch.cuda.Stream()
    stage = [{k: torch.empty_like(v, device=dev) for k, v in host[0].items()} for _ in range(2)]
    for st_ in stage:
        st_["unary"].requires_grad_(True)
        st_["pairwise"].requires_grad_(True)
    ready = [torch.cuda.Event() for _ in range(2)]
    done = [torch.cuda.Event() for _ in range(2)]
    loss_dev = [torch.zeros(1, device=dev) for _ in range(2)]

    def fwd_bwd(j):
        st_ = stage[j]
        fg.zero()
        out = layer(st_["unary"], st_["pairwise"], st_["up_idx"], st_["neighbor_idx"])
        out.backward(cot)
        loss_dev[j].copy_(out.detach().sum().reshape(1))
        st_["unary"].grad = None
        st_["pairwise"].grad = None

    e2e_graphs = []
    if args.graph:
        for j in range(2):
            for k, v in host[j].items():
                stage[j][k].data.copy_(v)
            torch.cuda.synchronize()
            gph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gph):
                fwd_bwd(j)
            e2e_graphs.append(gph)

    def upload(i):
        j = i % 2
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(done[j])                     # the previous user of this staging buffer has finished
            for k, v in host[j].items():
                stage[j][k].data.copy_(v, non_blocking=True)
            ready[j].record(copy_stream)

    def e2e_run(n):
        cur = torch.cuda.current_stream()
        for j in range(2):
            done[j].record(cur)
        upload(0)
        for i in range(n):
            j = i % 2
            if i + 1 < n:
                upload(i + 1)
            cur.wait_event(ready[j])
            if e2e_graphs:
                e2e_graphs[j].replay()
            else:
                fwd_bwd(j)
            fg.all_reduce()
            loss_host.copy_(loss_dev[j], non_blocking=True)
            done[j].record(cur)
            cur.synchronize()                                   # the caller reads the loss on the host every step

    e2e_run(3)
    barrier()
    e0.record()
    e2e_run(args.steps)
    e1.record()
    barrier()
    ms_e2e = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms_e2e], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_e2e = float(t.item())
    e2e_value = world * B * N_POINTS / (ms_e2e / args.steps * 1e-3)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- per-kernel instrumented pass (CUDA events around every C-ABI launch on the launching stream)
    peaks, peak_src = measured_peaks()
    prof = ops.profile_calls(lambda: eager_step(0, reduce=False), repeats=3)    # rank 0 only: no collectives from here on
    total_k = sum(v["ms"] for v in prof.values())
    top = max(prof.items(), key=lambda kv: kv[1]["ms"])
    kernels = {k: {"ms_per_step": round(v["ms"], 4), "calls_per_step": v["calls"], "share": round(v["ms"] / total_k, 4),
                   "algo_bytes_per_step": v["bytes"], "achieved_gbs": round(v["bytes"] / (v["ms"] * 1e-3) / 1e9, 1) if v["ms"] > 0 else None}
               for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])}
    top_ach = top[1]["bytes"] / (top[1]["ms"] * 1e-3) / 1e9
    traffic = None
    try:      # DRAM bytes (read + write) per launch of the same call, from the committed `ncu --set full` capture (profiles/)
        met = json.load(open(os.path.join(ROOT, "profiles", "ncu_r01_metrics.json")))
        traffic = met["calls"].get(top[0], {}).get("traffic_bytes")
    except Exception:
        pass
    roofline = {"kernel": top[0], "bound": "hbm", "achieved": round(top_ach, 1), "peak": peaks["hbm_gbs"], "unit": "GB/s",
                "frac": round(top_ach / peaks["hbm_gbs"], 4), "traffic": traffic, "peak_source": peak_src,
                "share_of_step": round(top[1]["ms"] / total_k, 4),
                "note": "achieved = algorithmic bytes of this kernel's calls in one step / their CUDA-event time"}
    step_ach = ALGO_BYTES_PER_CLOUD * B / (ms_step * 1e-3) / 1e9
    roofline_step = {"bound": "hbm", "achieved": round(step_ach, 1), "peak": peaks["hbm_gbs"], "unit": "GB/s",
                     "frac": round(step_ach / peaks["hbm_gbs"], 4), "frac_of_nominal_8TBs": round(step_ach / 8000.0, 4),
                     "algo_bytes_per_cloud": ALGO_BYTES_PER_CLOUD, "peak_source": peak_src}

    # ---- secondary metric of BASELINE.json: kNN queries/s (device-resident) on the same clouds
    pos = sets[0]["pos"]
    for _ in range(3):
        nn_.knn_batch(pos, pos, K_NBR)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(10):
        nn_.knn_batch(pos, pos, K_NBR)
    e1.record()
    torch.cuda.synchronize()
    knn_qps = 10 * B * N_POINTS / (e0.elapsed_time(e1) * 1e-3)

    # ---- secondary: the whole PointConvResNet (BASELINE configs[2]) fwd+bwd on the same clouds, replayed as one CUDA graph.
    # Reported next to the headline, never instead of it; any failure here leaves the bench line intact.
    network = None
    if args.graph and world == 1:
        try:
            import torch.nn.functional as Fn
            from crfconv_b200 import train_dp
            from crfconv_b200.graphs import GraphedStep
            from crfconv_b200.point_conv_big import PointConvResNet
            net = PointConvResNet(6, 13).to(dev).train()
            ngr = FlatGradients(net)
            npos, nfeat, nlab, ngen = train_dp.synthetic_shard(B, N_POINTS, 13, dev, seed=77)
            ndata = train_dp.make_batch(npos, nfeat, nlab, generator=ngen)

            def net_step():
                ngr.zero()
                loss = Fn.cross_entropy(net(ndata), ndata.y.reshape(-1) - 1)
                loss.backward()
                return loss.detach()

            gs = GraphedStep(net_step)
            for _ in range(2):
                gs.replay()
            torch.cuda.synchronize()
            e0.record()
            for _ in range(5):
                gs.replay()
            e1.record()
            torch.cuda.synchronize()
            net_ms = e0.elapsed_time(e1) / 5
            network = {"metric": "PointConvResNet fwd+bwd points/s", "value": B * N_POINTS / (net_ms * 1e-3), "unit": "points/s", "ms_per_step": net_ms,
                       "config": f"PointConvResNet(6, 13, use_crf=True), B={B}, N=40960, 5-level pyramid prebuilt on the device, one CUDA graph"}
            del gs, net, ngr, ndata
        except Exception as exc:                                     # noqa: BLE001
            network = {"error": f"{type(exc).__name__}: {exc}"[:200]}

    _, _, cb = cpu_layer_points_per_s(clouds=2, repeats=6, warm=1)
    line = {"metric": "CRFConv fwd+bwd points/s (N=40960,k=16)", "value": value, "unit": "points/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"single ContinuousGaussianCRFConv(128,64,64,steps=1) fwd+bwd, N=40960, Nc=10240, K=16, "
                                   f"{B} clouds per GPU per step (SURVEY.md §8 S1 / BASELINE configs[0] shape on the GPU)",
                       "clouds_per_gpu": B, "precision": "3xTF32 tensor-core contractions, fp32 elsewhere" if ops.PRECISION == 0 else "TF32",
                       "launch": "one CUDA graph per input set, replayed" if args.graph else "eager (per-kernel launches from Python)",
                       "l2": f"rotating {nsets} input sets of {in_bytes / 2**20:.0f} MiB each (> {L2_BYTES / 2**20:.0f} MiB L2)",
                       "parallelism": f"dp{world}: clouds sharded, one NCCL all-reduce of the flat gradient" if world > 1 else "single GPU"},
            "e2e": {"value": e2e_value, "unit": "points/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "roofline_step": roofline_step, "kernels": kernels,
            "knn": {"metric": "kNN queries/s", "value": knn_qps, "unit": "queries/s", "config": f"B={B}, N=Q=40960, K=16, device-resident"},
            "network": network,
            "cpu_baseline": cb}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--clouds", type=int, default=6, help="clouds per GPU per step")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-graph", dest="graph", action="store_false", help="launch every kernel from Python instead of replaying CUDA graphs")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world == 1 and args.gpus > 1:
        # launched without torchrun: re-launch under torch.distributed.run, one rank per GPU
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}", "--master-addr", "127.0.0.1",
               "--master-port", str(29500 + os.getpid() % 2000), os.path.abspath(__file__), "--gpus", str(args.gpus), "--steps", str(args.steps),
               "--warmup", str(args.warmup), "--clouds", str(args.clouds)] + ([] if args.graph else ["--no-graph"])
        sys.exit(subprocess.call(cmd))
    run_gpu(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
