#!/usr/bin/env python
"""bench.py — images/s through the attentional-pooling head on synthetic conv5 features.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A *step* is one pass of the hot path over one batch: BASELINE.json configs[1] — rank-1
attentional pooling, 32 feature maps of 14x14x2048 fp32 per GPU, K=393 — ONE kernel launch; at N>1
the single all-gather of the [32,393] logits is fused into that kernel (NVLink peer stores + step
flags, `--collective nccl` for a NCCL all_gather instead).  Consecutive steps are independent batches:
four are kept in flight (`--streams 4`, each lane with its own workspace) and back-to-back launches on
a lane overlap through programmatic dependent launch.  Weak scaling: every rank owns 32 images per step.  Prints ONE JSON line on rank 0 (contract in the task statement):
  value        whole-job images/s with X resident in HBM (CUDA events, max over ranks)
  e2e          same metric through the C-ABI host-buffer entry (apa_attn_pool_fwd_host): pinned
               host X -> device, head, logits -> host, inside the timed region
  roofline     fused kernel alone: algorithmic bytes / CUDA-event duration vs MEASURED_PEAKS.json
  cpu_baseline the oracle's PyTorch-CPU fp32 direct-form port timed on this box's host cores
L2 hygiene: the inputs rotate over a ring of X buffers larger than 2x L2, so every step reads
X from HBM.  `--impl reference` times the CPU port (TF-1.1 itself cannot be installed).
"""
from __future__ import annotations

import argparse
import math
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

B_PER_GPU, HW, C, K = 32, 196, 2048, 393
FLAGS = {"SINGLE_LAYER_ATT": True}
X_BYTES = B_PER_GPU * HW * C * 4
W_BYTES = (C + 1) * (K + 1) * 4
OUT_BYTES = B_PER_GPU * K * 4
ALGO_BYTES = X_BYTES + W_BYTES + OUT_BYTES  # 54.66 MB per launch
L2_BYTES = 126 * 1024 * 1024
METRIC = "imgs/sec attn-pool 14x14x2048->393 logits"


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(object):
    """nvidia-smi clocks/throttle reasons sampled while the GPU is under the benchmark's load."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None
        self.path = "/tmp/apa_clocks_%d_%d.csv" % (os.getpid(), gpu_index)

    def start(self):
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.close()
        sm, mx, reasons = [], [], set()
        try:
            with open(self.path) as f:
                for line in f:
                    p = [x.strip() for x in line.split(",")]
                    if len(p) < 9:
                        continue
                    try:
                        sm.append(float(p[1]))
                        mx.append(float(p[2]))
                    except ValueError:
                        continue
                    for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                          "sw_power_cap"), p[5:9]):
                        if val.lower().startswith("active"):
                            reasons.add(name)
            os.remove(self.path)
        except Exception:
            pass
        if sm:
            out["sm_mhz"] = statistics.median(sm)
            out["sm_max_mhz"] = max(mx)
            out["samples"] = len(sm)
        out["reasons"] = sorted(reasons)
        return out


def _best_thread_count(fn):
    """The CPU arm gets its best shot: sweep the intra-op thread count (all cores is often NOT the fastest
    for this memory-bound graph on a many-core host) and keep the fastest; leaves torch set to it."""
    import torch
    ncpu = os.cpu_count() or 1
    cands = sorted(set(c for c in (8, 16, 32, 64, ncpu) if c <= ncpu))
    best, best_t = ncpu, float("inf")
    with torch.no_grad():
        for c in cands:
            torch.set_num_threads(c)
            fn()
            t0 = time.perf_counter()
            fn()
            dt = time.perf_counter() - t0
            if dt < best_t:
                best, best_t = c, dt
    torch.set_num_threads(best)
    return best


def cpu_port_rate(seconds, images_per_call=B_PER_GPU):
    """images/s of the op-for-op PyTorch-CPU fp32 direct-form port (oracle (ii)), all host threads."""
    import torch
    from oracle import attn_pool_oracle as orc
    fl = dict(orc.DEFAULT_FLAGS, **FLAGS)
    params = {k: torch.from_numpy(v) for k, v in orc.make_params(K, C, flags=fl, seed=42).items()}
    X = torch.from_numpy(orc.make_features(images_per_call, HW, C, seed=42))
    cores = _best_thread_count(lambda: orc.attn_pool_head_torch_cpu(X, params, fl, K))
    with torch.no_grad():
        orc.attn_pool_head_torch_cpu(X, params, fl, K)  # warm-up
        n, t0 = 0, time.perf_counter()
        while True:
            orc.attn_pool_head_torch_cpu(X, params, fl, K)
            n += 1
            dt = time.perf_counter() - t0
            if dt >= seconds:
                break
    return n * images_per_call / dt, cores, n, dt


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU path for this head.  TF 1.1 / Python 2.7 cannot be
    installed here (no network, no wheel), so this times the oracle's op-for-op PyTorch-CPU port."""
    if rank != 0:
        return
    import torch
    from oracle import attn_pool_oracle as orc
    fl = dict(orc.DEFAULT_FLAGS, **FLAGS)
    params = {k: torch.from_numpy(v) for k, v in orc.make_params(K, C, flags=fl, seed=42).items()}
    # bound the whole run to ~2 minutes: shrink the per-step sample if K is large
    probe = torch.from_numpy(orc.make_features(4, HW, C, seed=42))
    cores = _best_thread_count(lambda: orc.attn_pool_head_torch_cpu(
        torch.from_numpy(orc.make_features(B_PER_GPU, HW, C, seed=42)), params, fl, K))
    with torch.no_grad():
        orc.attn_pool_head_torch_cpu(probe, params, fl, K)
        t0 = time.perf_counter()
        orc.attn_pool_head_torch_cpu(probe, params, fl, K)
        t_img = (time.perf_counter() - t0) / 4
    budget = 120.0
    n_img = int(max(1, min(B_PER_GPU, budget / (max(1, args.steps + args.warmup) * t_img))))
    X = torch.from_numpy(orc.make_features(n_img, HW, C, seed=42))
    with torch.no_grad():
        for _ in range(args.warmup):
            orc.attn_pool_head_torch_cpu(X, params, fl, K)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            orc.attn_pool_head_torch_cpu(X, params, fl, K)
        dt = time.perf_counter() - t0
    value = args.steps * n_img / dt
    sample = "%d images/step (of the %d-image batch), %d steps, PyTorch-CPU fp32 direct form" % (
        n_img, B_PER_GPU, args.steps)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "images/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "configs[1]: rank-1 attn pool, synthetic 14x14x2048 feats, K=393, batch=32",
                   "note": "TF-1.1 CPU path is not installable here; oracle port timed instead"},
        "cpu_baseline": {"value": value, "unit": "images/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=50)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cluster", type=int, default=0, help="force CTAs per image (0 = auto)")
    ap.add_argument("--fused", type=int, default=-1, help="1 fused kernel, 0 two kernels, -1 auto")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--streams", type=int, default=4, help="independent batches in flight (1..4)")
    ap.add_argument("--no-pdl", action="store_true", help="disable programmatic dependent launch")
    ap.add_argument("--trigger-lead", type=int, default=0, help="PDL trigger lead in tiles (0 = library default, -1 = none)")
    ap.add_argument("--no-coop", action="store_true", help="never use the cooperative launch attribute")
    ap.add_argument("--collective", default="peer", choices=["peer", "nccl"],
                    help="N>1: all-gather fused into the kernel as NVLink peer stores, or a NCCL all_gather per step")
    ap.add_argument("--general", action="store_true", help="force the grid-barrier kernel (not the cluster-local one)")
    ap.add_argument("--split", action="store_true", help="force the two-kernel STREAM + CONTRACT pipeline")
    ap.add_argument("--shape", type=int, default=-1, choices=[-1, 0, 1, 2],
                    help="cluster-local kernel shape (-1 = library default; 0: 4 img/cluster, 2 CTAs/SM; "
                         "1: 4 img/cluster, 3 CTAs/SM; 2: 8 img/cluster, 3 CTAs/SM)")
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    ap.add_argument("--e2e-steps", type=int, default=30)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from attentionalpoolingaction_b200.custom_ops import attn_pool_factory as ops
    from attentionalpoolingaction_b200 import dist as apa_dist
    from oracle import attn_pool_oracle as orc  # inputs generator + cpu_baseline leg only

    assert torch.cuda.is_available(), "bench.py needs a GPU (there is no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    assert world == args.gpus, "--gpus %d but WORLD_SIZE=%d" % (args.gpus, world)

    fl = dict(orc.DEFAULT_FLAGS, **FLAGS)
    params = orc.make_params(K, C, flags=fl, seed=42)
    bits, rnk = ops.flags_from_cfg_net(fl)
    tuning = ops.make_tuning(args.cluster, args.fused, no_pdl=int(args.no_pdl), no_cooperative=int(args.no_coop))
    tuning.pdl_trigger_lead = args.trigger_lead
    tuning.form = 1 if args.general else (7 if args.split else (0 if args.shape < 0 else 4 + args.shape))
    nstreams = max(1, min(4, args.streams))
    # one op (= one workspace) per in-flight batch: consecutive steps are independent batches, so two
    # of them overlap on the GPU (one streams X from HBM while the other is in its latency-bound tail)
    opl = [ops.AttnPoolOp({k: torch.from_numpy(v) for k, v in params.items()}, K, HW, channels=C, rank=rnk,
                          flags=bits, tuning=tuning, device=dev) for _ in range(nstreams)]
    op = opl[0]
    plan = ops.attn_pool_plan(op.desc(B_PER_GPU), tuning)

    # ring of inputs > 2x L2 so that every step streams X from HBM
    nbuf = max(2, -(-2 * L2_BYTES // X_BYTES) + 1)
    g = torch.Generator(device=dev).manual_seed(42 + rank)
    ring = [torch.relu(torch.randn((B_PER_GPU, HW, C), device=dev, generator=g)) for _ in range(nbuf)]
    logits_l = [torch.empty((B_PER_GPU, K), device=dev) for _ in range(nstreams)]
    logits = logits_l[0]
    gathered_l = [torch.empty((world * B_PER_GPU, K), device=dev) if world > 1 else None for _ in range(nstreams)]
    collective = "none"
    peer_l = None
    if world > 1:
        collective = args.collective
        if collective == "peer":
            try:
                peer_l = [apa_dist.PeerGather(B_PER_GPU, K, dev) for _ in range(nstreams)]
            except Exception as e:  # symmetric memory not available on this box: NCCL all_gather
                sys.stderr.write("peer gather unavailable (%s); using NCCL all_gather\n" % e)
                collective = "nccl"

    def step(i, with_collective=True):
        j = i % nstreams
        if world > 1 and with_collective and collective == "peer":
            opl[j].forward_gather(ring[i % nbuf], peer_l[j].struct)   # kernel stores into every peer's buffer
            return
        opl[j](ring[i % nbuf], logits=logits_l[j])
        if world > 1 and with_collective:
            apa_dist.all_gather_logits(logits_l[j], gathered_l[j])

    # parity spot-check of what is being timed (first buffer, rank-local) -- outside timed regions
    step(0)
    torch.cuda.synchronize()
    ref, _ = orc.attn_pool_head(ring[0][:4].cpu().numpy(), params, fl, K, want_pose=False)
    mine = (peer_l[0].out[rank * B_PER_GPU:rank * B_PER_GPU + 4] if (world > 1 and collective == "peer")
            else logits[:4])
    parity = orc.rel_err(mine.cpu().numpy(), ref)
    assert parity <= 1e-4, "bench parity check failed: %.3e" % parity

    side = torch.cuda.Stream(device=dev)
    lanes = [side] + [torch.cuda.Stream(device=dev) for _ in range(nstreams - 1)]

    def enqueue(cnt, with_collective):
        """`cnt` steps on the current stream, step i on lane i % nstreams (fork/join around the extra lane)."""
        cur = torch.cuda.current_stream(dev)
        for ln in lanes[1:]:
            ln.wait_stream(cur)
        for i in range(cnt):
            ln = cur if i % nstreams == 0 else lanes[i % nstreams]
            with torch.cuda.stream(ln):
                step(i, with_collective)
        for ln in lanes[1:]:
            cur.wait_stream(ln)

    def make_runner(nsteps, with_collective):
        """Returns run() that enqueues exactly `nsteps` steps; CUDA graphs of one ring pass when possible."""
        if args.no_graph or nsteps == 0:
            return (lambda: enqueue(nsteps, with_collective)), "stream"
        try:
            ring_steps = math.lcm(nbuf, nstreams)
            full, rem = divmod(nsteps, ring_steps)
            graphs = []
            with torch.cuda.stream(side):
                for cnt in ((ring_steps,) if full else ()) + ((rem,) if rem else ()):
                    enqueue(cnt, with_collective)  # warm the exact launches before capture
                    torch.cuda.synchronize()
                    gr = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(gr, stream=side):
                        enqueue(cnt, with_collective)
                    graphs.append(gr)
            torch.cuda.synchronize()
            g_full = graphs[0] if full else None
            g_rem = graphs[-1] if rem else None

            def run():
                for _ in range(full):
                    g_full.replay()
                if g_rem is not None:
                    g_rem.replay()
            return run, "cuda-graph(%d steps/replay, %d batches in flight)" % (ring_steps, nstreams)
        except Exception as e:  # capture refused (e.g. cooperative launch): plain stream launches
            sys.stderr.write("graph capture failed (%s); using stream launches\n" % e)
            torch.cuda.synchronize()
            return (lambda: enqueue(nsteps, with_collective)), "stream"

    def timed(run):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dist.barrier()
            ms = float(t.item())
        return ms

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()

    run_w, _ = make_runner(args.warmup, True)
    run_k, launch_mode = make_runner(args.steps, True)
    run_w()
    torch.cuda.synchronize()
    ms_total = timed(run_k)                       # <- the K timed steps
    ms_per_step = ms_total / args.steps
    value = world * B_PER_GPU * args.steps / (ms_total * 1e-3)

    # roofline: the fused kernel alone (no collective), long enough for clock sampling (~1 s)
    n_roof = int(min(200000, max(2000, 1.0e3 / max(ms_per_step, 1e-3))))
    run_r, _ = make_runner(n_roof, False)
    run_r()
    torch.cuda.synchronize()
    ms_roof = timed(run_r) / n_roof
    peak, peak_src = _peaks()
    achieved = ALGO_BYTES / (ms_roof * 1e-3) / 1e9

    # e2e: host buffers through apa_attn_pool_fwd_host (H2D + kernel + D2H + sync per step)
    xh = [torch.relu(torch.randn((B_PER_GPU, HW, C), generator=torch.Generator().manual_seed(rank))).pin_memory()
          for _ in range(2)]
    lh = torch.empty((B_PER_GPU, K)).pin_memory()
    for i in range(3):
        op.forward_host(xh[i % 2], lh)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(args.e2e_steps):
        op.forward_host(xh[i % 2], lh)
    t_e2e = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([t_e2e], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_e2e = float(t.item())
    e2e_value = world * B_PER_GPU * args.e2e_steps / t_e2e

    clocks = sampler.stop() if rank == 0 else None

    if rank == 0:
        cpu_v, cores, ncalls, cpu_dt = cpu_port_rate(args.cpu_seconds)
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            with open(tp) as f:
                traffic = json.load(f).get("dram_bytes_per_launch")
        line = {
            "metric": METRIC, "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {
                "workload": "configs[1]: rank-1 attn pool, synthetic 14x14x2048 feats, K=393, batch=32 per GPU",
                "per_gpu_batch": B_PER_GPU, "global_batch": world * B_PER_GPU,
                "parallelism": ("batch-sharded x%d, one all-gather of logits per step (%s)" % (
                                    world, "fused into the kernel: NVLink peer stores + step flags"
                                    if collective == "peer" else "NCCL all_gather_into_tensor")) if world > 1
                               else "single GPU",
                "l2": "inputs rotate over %d buffers x %.1f MB (%.0f MB > 2x 126 MB L2): every step reads HBM"
                      % (nbuf, X_BYTES / 1e6, nbuf * X_BYTES / 1e6),
                "launch": launch_mode, "plan": dict(plan), "batches_in_flight": nstreams, "parity_rel_err_vs_fp64_oracle": parity,
                "clock_sampling": "nvidia-smi every 100 ms across warm-up, the K timed steps, the "
                                  "roofline loop (%d launches) and the e2e loop" % n_roof,
            },
            "e2e": {"value": e2e_value, "unit": "images/s", "h2d_bytes_per_step": X_BYTES,
                    "d2h_bytes_per_step": OUT_BYTES, "steps": args.e2e_steps,
                    "api": "apa_attn_pool_fwd_host (pinned host X -> HBM, head, logits -> host, sync)"},
            "gpu_launches": args.steps * int(plan["launches"]),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "kernel": "apa_pool_local_kernel" if plan.get("cluster_local") else "apa_pool_fused_kernel", "us_per_launch": ms_roof * 1e3,
                         "duration": "CUDA-event time of the loop / launches (%d independent launches in flight)"
                                     % nstreams,
                         "algorithmic_bytes_per_launch": ALGO_BYTES, "launches_timed": n_roof},
            "cpu_baseline": {"value": cpu_v, "unit": "images/s", "cores": cores, "kind": "port",
                             "sample": "%d calls x %d images in %.1f s, PyTorch-CPU fp32 direct form"
                                       % (ncalls, B_PER_GPU, cpu_dt)},
            "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sys.stdout.flush()
    sys.stderr.flush()
    # Captured CUDA graphs hold NCCL kernels; tearing the communicator down under them can hang.  Everything
    # is printed and synchronised: leave without running destructors.
    os._exit(0)


if __name__ == "__main__":
    main()
