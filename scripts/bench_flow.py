#!/usr/bin/env python
"""A/B of the single-map head's forms on one shape: us/step and TB/s of algorithmic bytes for
(form, lanes, inputs_stable, max_ctas), graph-captured, inputs cycled through > 2x L2; plus the latency of ONE cold
launch (events around a single call after a device sync).
usage: bench_flow.py [B=32] [K=393] [F=1] [HW=196]"""
import os, sys, math
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from attentionalpoolingaction_b200.custom_ops import attn_pool_factory as ops
from oracle import attn_pool_oracle as orc

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
K = int(sys.argv[2]) if len(sys.argv) > 2 else 393
F = int(sys.argv[3]) if len(sys.argv) > 3 else 1
HW = int(sys.argv[4]) if len(sys.argv) > 4 else 196
C = 2048
dev = torch.device("cuda:0")
fl = dict(orc.DEFAULT_FLAGS, SINGLE_LAYER_ATT=True)
params = orc.make_params(K, C, flags=fl, seed=42)
bits, rank = ops.flags_from_cfg_net(fl)
P = {k: torch.from_numpy(v) for k, v in params.items()}
xbytes = B * HW * C * 4
algo = xbytes + (C + 1) * (K + 1) * 4 + (B // F) * K * 4
nbuf = max(2, int(2 * 126 * 2**20 // xbytes) + 1)
g = torch.Generator(device=dev).manual_seed(3)
xs = [torch.relu(torch.randn((B, HW, C), device=dev, generator=g)) for _ in range(nbuf)]
ref, _ = orc.attn_pool_head(xs[0][:2 * F].cpu().numpy(), params, fl, K, frames_per_video=F, want_pose=False)

CONFIGS = []
for form, name in ((ops.APA_FORM_FLOW, "flow"), (ops.APA_FORM_CLUSTER, "cluster")):
    for lanes in (1, 2, 4, 6):
        for stable in (0, 1):
            CONFIGS.append((name, form, lanes, stable, 0))
CONFIGS += [("flow", ops.APA_FORM_FLOW, 4, 1, 49), ("flow", ops.APA_FORM_FLOW, 6, 1, 37), ("flow", ops.APA_FORM_FLOW, 8, 1, 25)]
LEADS = (0,)


def run_cfg(name, form, lanes, stable, max_ctas, lead=0):
    tun = ops.make_tuning(form=form, inputs_stable=stable, stream_ctas=max_ctas, pdl_trigger_lead=lead)
    opl = [ops.AttnPoolOp(P, K, HW, channels=C, frames=F, rank=rank, flags=bits, tuning=tun, device=dev)
           for _ in range(lanes)]
    plan = ops.attn_pool_plan(opl[0].desc(B), tun)
    outs = [torch.empty((B // F, K), device=dev) for _ in range(lanes)]
    opl[0](xs[0], logits=outs[0])
    torch.cuda.synchronize()
    err = orc.rel_err(outs[0][:2].cpu().numpy(), ref)
    assert err <= 1e-4, (name, err)
    side = torch.cuda.Stream()
    extra = [torch.cuda.Stream() for _ in range(lanes - 1)]
    n = math.lcm(nbuf, lanes) * 2

    def enqueue():
        cur = torch.cuda.current_stream()
        for s in extra:
            s.wait_stream(cur)
        for i in range(n):
            with torch.cuda.stream(cur if i % lanes == 0 else extra[i % lanes - 1]):
                opl[i % lanes](xs[i % nbuf], logits=outs[i % lanes])
        for s in extra:
            cur.wait_stream(s)

    with torch.cuda.stream(side):
        enqueue()
        torch.cuda.synchronize()
        gr = torch.cuda.CUDAGraph()
        try:
            with torch.cuda.graph(gr, stream=side):
                enqueue()
            run = gr.replay
            mode = "graph"
        except Exception as e:
            torch.cuda.synchronize()
            run, mode = enqueue, "stream"
        for _ in range(3):
            run()
        torch.cuda.synchronize()
        reps = max(5, int(4000 / n))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            run()
        e1.record()
        torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / (reps * n)
    # short window like the driver's: 20 steps, one graph replay, 4 repeats
    # single cold launch latency
    lat = []
    for i in range(12):
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        opl[0](xs[(i + 3) % nbuf], logits=outs[0])
        b.record()
        torch.cuda.synchronize()
        lat.append(a.elapsed_time(b) * 1e3)
    lat.sort()
    print("%-9s lanes=%d stable=%d max_ctas=%3d lead=%2d [%s, kind %d, grid %d]: %7.2f us/step  %5.2f TB/s  frac %.3f | one launch: %6.1f us (median of 12)"
          % (name, lanes, stable, max_ctas, lead, mode, plan["form"], plan["grid"], us, algo / us / 1e6,
             algo / us / 1e6 / 6.5748, lat[len(lat) // 2]), flush=True)


print("B=%d HW=%d K=%d F=%d: %.2f MB algorithmic per step, roofline %.2f us at 6574.8 GB/s" % (B, HW, K, F, algo / 1e6, algo / 6574.8e3))
for c in CONFIGS:
    for lead in (LEADS if (c[1] == ops.APA_FORM_FLOW and c[3] == 1) else (0,)):
        try:
            run_cfg(*c, lead=lead)
        except Exception as e:
            print("%-9s lanes=%d stable=%d max_ctas=%d: FAILED %s" % (c[0], c[2], c[3], c[4], e), flush=True)
