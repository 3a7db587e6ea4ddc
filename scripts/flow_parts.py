#!/usr/bin/env python
"""Where the flow form's time goes: the chain with its tail cut off (ApaTuning.debug_cut_tail = 1: STREAM alone, 2: +
REDUCE, 3: + CONTRACT; results are then wrong by construction) against the full chain, by lanes, trigger lead and
STREAM CTAs per call."""
import os, sys, math
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from attentionalpoolingaction_b200.custom_ops import attn_pool_factory as ops
from oracle import attn_pool_oracle as orc

B, HW, C, K = 32, 196, 2048, 393
dev = torch.device("cuda:0")
fl = dict(orc.DEFAULT_FLAGS, SINGLE_LAYER_ATT=True)
params = orc.make_params(K, C, flags=fl, seed=42)
bits, rank = ops.flags_from_cfg_net(fl)
P = {k: torch.from_numpy(v) for k, v in params.items()}
xbytes = B * HW * C * 4
nbuf = 7
xs = [torch.relu(torch.randn((B, HW, C), device=dev)) for _ in range(nbuf)]


def run(tag, fused, lanes, lead, max_ctas=0, shape=0):
    tun = ops.make_tuning(form=ops.APA_FORM_FLOW, inputs_stable=1, debug_cut_tail=fused, pdl_trigger_lead=lead, stream_ctas=max_ctas, stream_shape=shape)
    opl = [ops.AttnPoolOp(P, K, HW, rank=rank, flags=bits, tuning=tun, device=dev) for _ in range(lanes)]
    outs = [torch.empty((B, K), device=dev) for _ in range(lanes)]
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
        with torch.cuda.graph(gr, stream=side):
            enqueue()
        for _ in range(3):
            gr.replay()
        torch.cuda.synchronize()
        reps = max(5, int(3000 / n))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            gr.replay()
        e1.record()
        torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / (reps * n)
    print("%-22s lanes=%d lead=%2d G=%3d: %6.2f us/step  %5.2f TB/s of X" % (tag, lanes, lead, max_ctas or 148, us, xbytes / us / 1e6), flush=True)


import itertools
for lanes in (1, 2, 4):
    for cut, tag in ((1, "STREAM alone"), (2, "STREAM + REDUCE"), (3, "STREAM+REDUCE+CONTRACT"), (0, "full chain")):
        run(tag, cut, lanes, 0)
for lanes, g in itertools.product((4, 6, 8), (25, 37, 49, 74, 148)):
    run("full chain", 0, lanes, 0, g)
