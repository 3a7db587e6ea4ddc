#!/usr/bin/env python
"""In-kernel timeline of the flow form (ApaTuning.debug_timestamps): one cold launch alone, then the last NOPS
launches of a long queue on `lanes` streams (each op has its own workspace, so each keeps its own stamps).
usage: flow_timeline.py [lanes=1] [B=32]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from attentionalpoolingaction_b200.custom_ops import attn_pool_factory as ops
from oracle import attn_pool_oracle as orc

S = int(sys.argv[1]) if len(sys.argv) > 1 else 1
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
HW, C, K = 196, 2048, 393
NOPS = 8
fl = dict(orc.DEFAULT_FLAGS, SINGLE_LAYER_ATT=True)
params = orc.make_params(K, C, flags=fl, seed=42)
bits, rank = ops.flags_from_cfg_net(fl)
tun = ops.make_tuning(form=ops.APA_FORM_FLOW, inputs_stable=1, debug_timestamps=1)
P = {k: torch.from_numpy(v) for k, v in params.items()}
opl = [ops.AttnPoolOp(P, K, HW, rank=rank, flags=bits, tuning=tun) for _ in range(NOPS)]
xs = [torch.relu(torch.randn((B, HW, C), device="cuda")) for _ in range(8)]
outs = [torch.empty((B, K), device="cuda") for _ in range(NOPS)]
plan = ops.attn_pool_plan(opl[0].desc(B), tun)
G = plan["grid"]
U = plan["contract_grid"]


def stamps(op):
    d, ws, _ = op._workspace(B, (False, False, False))
    return ws[4096:4096 + 320 * 64 * 8].view(torch.int64).reshape(320, 64).cpu().numpy()


def report(ts, t0, tag):
    st = ts[:min(G, 160)]
    ct = ts[160:160 + min(U, 160)]
    s0, s1 = (st[:, 62] - t0) / 1e3, (st[:, 63] - t0) / 1e3
    c0, c1 = (ct[:, 62] - t0) / 1e3, (ct[:, 63] - t0) / 1e3
    clk = 1.965e3  # cycles per us (nominal)
    waited = (ct[:, 1] - ct[:, 0]) / clk
    gl = (ct[:, 2] - ct[:, 0]) / clk
    arr = (ct[:, 3] - ct[:, 0]) / clk
    end = (ct[:, 5] - ct[:, 0]) / clk
    sm = st[:, 61]
    cnt = np.bincount(sm.astype(np.int64), minlength=148)
    hist = [int((cnt == i).sum()) for i in range(4)]
    print("%s STREAM start %6.2f..%6.2f end %6.2f..%6.2f (mean len %5.2f) | CONTRACT start %6.2f..%6.2f end %6.2f..%6.2f "
          "per-CTA since start: waited %5.2f gloaded %5.2f arrived %5.2f end %5.2f | SMs with 0/1/2/3 STREAM CTAs %s"
          % (tag, s0.min(), s0.max(), s1.min(), s1.max(), (s1 - s0).mean(), c0.min(), c0.max(), c1.min(), c1.max(),
             waited.mean(), gl.mean(), arr.mean(), end.mean(), hist), flush=True)
    return s0.min()


# ---- one cold launch alone
for i in range(3):
    opl[0](xs[i], logits=outs[0])
    torch.cuda.synchronize()
ts = stamps(opl[0])
t0 = min(ts[:G, 62].min(), ts[160:160 + min(U, 160), 62].min())
report(ts, t0, "alone   ")

# ---- steady state
streams = [torch.cuda.Stream() for _ in range(S)]
WARM = 6 * NOPS


def enqueue():
    for i in range(WARM + NOPS):
        j = i % NOPS
        with torch.cuda.stream(streams[i % S]):
            opl[j](xs[i % 8], logits=outs[j])


enqueue(); torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
side = torch.cuda.Stream()
with torch.cuda.stream(side):
    with torch.cuda.graph(g, stream=side):
        cur = torch.cuda.current_stream()
        for s in streams:
            s.wait_stream(cur)
        enqueue()
        for s in streams:
            cur.wait_stream(s)
g.replay(); torch.cuda.synchronize()
g.replay(); torch.cuda.synchronize()
all_ts = [stamps(op) for op in opl]
order = [(WARM + k) % NOPS for k in range(NOPS)]
t0 = min(all_ts[order[0]][:G, 62].min(), all_ts[order[0]][160:160 + min(U, 160), 62].min())
print("lanes %d, last %d launches, us since the first of them started" % (S, NOPS))
starts = []
for k, j in enumerate(order):
    starts.append(report(all_ts[j], t0, "launch %d" % k))
print("start-to-start spacing (us):", [round(b - a, 2) for a, b in zip(starts, starts[1:])])
