#!/usr/bin/env python
"""Time every form of the single-map head (grid-barrier kernel, cluster-local shapes, STREAM+CONTRACT pipeline)
on a few shapes: graph-captured back-to-back calls on one stream, inputs cycled through > L2 worth of buffers.
Prints us/call per form; parity of every form is checked against the oracle on the first buffer."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from attentionalpoolingaction_b200.custom_ops import attn_pool_factory as ops
from oracle import attn_pool_oracle as orc

SHAPES = [  # (name, B, HW, C, K, F, flags)
    ("cfg2 b32", 32, 196, 2048, 393, 1, dict(SINGLE_LAYER_ATT=True)),
    ("b30 (not % 4)", 30, 196, 2048, 393, 1, dict(SINGLE_LAYER_ATT=True)),
    ("hico 10-crop b20 k600", 20, 196, 2048, 600, 10, dict(SINGLE_LAYER_ATT=True)),
    ("b8", 8, 196, 2048, 393, 1, dict(SINGLE_LAYER_ATT=True)),
    ("b128", 128, 196, 2048, 393, 1, dict(SINGLE_LAYER_ATT=True)),
    ("b32 softmax", 32, 196, 2048, 393, 1, dict(SINGLE_LAYER_ATT=True, SOFTMAX_ATT=True)),
    ("b32 rank3 relu", 32, 196, 2048, 393, 1, dict(SINGLE_LAYER_ATT=True, RANK=3, RELU_ATT=True)),
    ("b32 c1024", 32, 196, 1024, 393, 1, dict(SINGLE_LAYER_ATT=True)),
]
FORMS = [("auto", 0), ("grid-barrier", 1), ("local shape0", 4), ("local shape1", 5), ("local shape2", 6), ("pipeline", 7)]


def main():
    dev = torch.device("cuda:0")
    for name, B, HW, C, K, F, flags in SHAPES:
        fl = dict(orc.DEFAULT_FLAGS, **flags)
        params = orc.make_params(K, C, flags=fl, seed=1)
        bits, rank = ops.flags_from_cfg_net(fl)
        P = {k: torch.from_numpy(v) for k, v in params.items()}
        nbuf = max(2, int(2 * 126e6 // (B * HW * C * 4)) + 1)
        g = torch.Generator(device=dev).manual_seed(3)
        xs = [torch.relu(torch.randn((B, HW, C), device=dev, generator=g)) for _ in range(nbuf)]
        ref, _ = orc.attn_pool_head(xs[0][:F * 2].cpu().numpy(), params, fl, K, frames_per_video=F, want_pose=False)
        row = []
        for fname, code in FORMS:
            tun = ops.make_tuning()
            tun.form = code
            op = ops.AttnPoolOp(P, K, HW, channels=C, frames=F, rank=rank, flags=bits, tuning=tun, device=dev)
            plan = ops.attn_pool_plan(op.desc(B), tun)
            kind = {0: "grid-barrier", 1: "local", 2: "pipeline"}[plan["cluster_local"]]
            if code in (4, 5, 6) and kind != "local":
                row.append("%s: n/a" % fname)
                continue
            out = torch.empty((B // F, K), device=dev)
            op(xs[0], logits=out)
            torch.cuda.synchronize()
            err = orc.rel_err(out[:2].cpu().numpy(), ref)
            assert err <= 1e-4, (name, fname, err)
            side = torch.cuda.Stream()
            n = 2 * nbuf
            with torch.cuda.stream(side):
                for i in range(n):
                    op(xs[i % nbuf], logits=out)
                torch.cuda.synchronize()
                gr = torch.cuda.CUDAGraph()
                try:
                    with torch.cuda.graph(gr, stream=side):
                        for i in range(n):
                            op(xs[i % nbuf], logits=out)
                    run = gr.replay
                except Exception:   # cooperative launches cannot be captured
                    torch.cuda.synchronize()
                    run = lambda: [op(xs[i % nbuf], logits=out) for i in range(n)]
                for _ in range(5):
                    run()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                reps = 40
                e0.record()
                for _ in range(reps):
                    run()
                e1.record()
                torch.cuda.synchronize()
            us = e0.elapsed_time(e1) * 1e3 / (reps * n)
            row.append("%s[%s]: %.2f us" % (fname, kind, us))
        print("%-24s %s" % (name, " | ".join(row)), flush=True)


main()
