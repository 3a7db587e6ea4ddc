#!/usr/bin/env python
"""Dump the fused kernel's in-kernel timeline (ApaTuning.debug_timestamps) for the bench workload."""
import os, sys, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from attentionalpoolingaction_b200.custom_ops import attn_pool_factory as ops
from oracle import attn_pool_oracle as orc

def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    cluster = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    fused = int(sys.argv[3]) if len(sys.argv) > 3 else 1
    pieces = int(sys.argv[4]) if len(sys.argv) > 4 else 0
    HW, C, K = 196, 2048, 393
    fl = dict(orc.DEFAULT_FLAGS, SINGLE_LAYER_ATT=True)
    params = orc.make_params(K, C, flags=fl, seed=42)
    bits, rank = ops.flags_from_cfg_net(fl)
    tun = ops.make_tuning(cluster, fused, debug_timestamps=1, no_pdl=pieces, max_ctas=(int(sys.argv[6]) if len(sys.argv) > 6 else 0))
    tun.pdl_trigger_lead = int(sys.argv[7]) if len(sys.argv) > 7 else 0
    tun.form = int(sys.argv[8]) if len(sys.argv) > 8 else 0
    op = ops.AttnPoolOp({k: torch.from_numpy(v) for k, v in params.items()}, K, HW, rank=rank, flags=bits, tuning=tun)
    xs = [torch.relu(torch.randn((B, HW, C), device="cuda")) for _ in range(8)]
    for x in xs:
        op(x)
    torch.cuda.synchronize()
    warm = int(sys.argv[5]) if len(sys.argv) > 5 else 0
    op(xs[0])   # cold in L2 after 7 other buffers
    if warm:      # steady state: many back-to-back launches (programmatic dependent launch overlaps them);
        for i in range(1, 1 + warm):   # the timeline read below is the LAST launch's
            op(xs[i % 8])
    torch.cuda.synchronize()
    d, ws, _ = op._workspace(B, (False, False, False))
    ts = ws[2048:2048 + 320 * 64 * 8].view(torch.int64).reshape(320, 64).cpu().numpy()
    plan = ops.attn_pool_plan(d, tun)
    n = plan["grid"]
    ts = ts[:n]
    ghz = 1.965
    rel = lambda a: (a - ts[:, 0:1]) / ghz / 1e3   # us since CTA start
    g0 = ts[:, 62].min()
    print("plan", dict(plan))
    print("kernel span (globaltimer) %.2f us" % ((ts[:, 63].max() - g0) / 1e3))
    print("CTA start skew (globaltimer) max %.2f us" % ((ts[:, 62].max() - g0) / 1e3))
    st = np.sort((ts[:, 62] - g0) / 1e3)
    print("CTA start times (us), every 16th:", np.round(st[::16], 2).tolist())
    names = ["start", "cluster_sync", "phase1_end", "dsmem_done(gown_ready)", "grid_barrier(out_ready)", "end", "grid_arrive(gathered)", "p2_operands(p2 computed)"]
    ev = rel(ts[:, 0:8])
    for i, nm in enumerate(names):
        print("%-28s mean %.2f  min %.2f  max %.2f us" % (nm, ev[:, i].mean(), ev[:, i].min(), ev[:, i].max()))
    tiles = rel(ts[:, 8:32]); issue = rel(ts[:, 32:56])
    for cta in (0, n // 2):
        nt = int((ts[cta, 8:32] > 0).sum())
        print("CTA %d tiles ready (us):" % cta, np.round(tiles[cta, :nt], 2).tolist())
        nt = int((ts[cta, 32:56] > 0).sum())
        print("CTA %d copies issued (us):" % cta, np.round(issue[cta, :nt], 2).tolist())

main()
