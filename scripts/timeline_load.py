#!/usr/bin/env python
"""Timeline of the cluster-local head kernel UNDER LOAD: the last 12 launches of a long queue each write their
in-kernel timestamps (ApaTuning.debug_timestamps) to their own workspace, so the steady-state overlap of
consecutive launches (start/end on %globaltimer, phase marks on clock64) can be read back.
usage: timeline_load.py [streams=3] [shape=2] [lead=0]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from attentionalpoolingaction_b200.custom_ops import attn_pool_factory as ops
from oracle import attn_pool_oracle as orc


def main():
    S = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    shape = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    lead = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    B, HW, C, K = 32, 196, 2048, 393
    NOPS = 12
    fl = dict(orc.DEFAULT_FLAGS, SINGLE_LAYER_ATT=True)
    params = orc.make_params(K, C, flags=fl, seed=42)
    bits, rank = ops.flags_from_cfg_net(fl)
    tun = ops.make_tuning(debug_timestamps=1)
    tun.pdl_trigger_lead = lead
    tun.form = 7 if shape == 7 else 4 + shape
    P = {k: torch.from_numpy(v) for k, v in params.items()}
    opl = [ops.AttnPoolOp(P, K, HW, rank=rank, flags=bits, tuning=tun) for _ in range(NOPS)]
    xs = [torch.relu(torch.randn((B, HW, C), device="cuda")) for _ in range(8)]
    outs = [torch.empty((B, K), device="cuda") for _ in range(NOPS)]
    for i, op in enumerate(opl):
        op(xs[i % 8], logits=outs[i])
    torch.cuda.synchronize()
    streams = [torch.cuda.Stream() for _ in range(S)]
    WARM = 10 * NOPS

    def enqueue():
        for i in range(WARM + NOPS):
            j = i % NOPS
            with torch.cuda.stream(streams[i % S]):
                opl[j](xs[i % 8], logits=outs[j])
    # capture as a graph so that launch overhead does not pace the queue
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
    rows, crows = [], []
    for j, op in enumerate(opl):
        d, ws, _ = op._workspace(B, (False, False, False))
        plan = ops.attn_pool_plan(d, tun)
        ts = ws[2048:2048 + 320 * 64 * 8].view(torch.int64).reshape(320, 64).cpu().numpy()
        rows.append(ts[:min(160, plan["grid"])])
        if plan["cluster_local"] == 2:
            crows.append(ts[160:160 + min(160, plan["contract_grid"])])
    if crows:
        t0 = min(r[:, 62].min() for r in rows)
        order = np.argsort([r[:, 62].min() for r in rows])
        print("streams %d STREAM+CONTRACT pipeline lead %d; us since the first of the last %d launches started" % (S, lead, NOPS))
        print("launch | STREAM start(min..max) end(min..max)  per-CTA: first tile, stream end, end | CONTRACT start(min..max) end(min..max) per-CTA: waited, g loaded, contracted, end")
        for j in order:
            a, c = rows[j], crows[j]
            ra = (a[:, 0:64] - a[:, 0:1]) / 1.965 / 1e3
            rc = (c[:, 0:8] - c[:, 0:1]) / 1.965 / 1e3
            print("%4d | %6.2f..%6.2f  %6.2f..%6.2f   %5.2f %5.2f %5.2f | %6.2f..%6.2f  %6.2f..%6.2f   %5.2f %5.2f %5.2f %5.2f" % (
                j, (a[:, 62].min() - t0) / 1e3, (a[:, 62].max() - t0) / 1e3, (a[:, 63].min() - t0) / 1e3, (a[:, 63].max() - t0) / 1e3,
                ra[:, 8].mean(), ra[:, 2].mean(), ra[:, 5].mean(),
                (c[:, 62].min() - t0) / 1e3, (c[:, 62].max() - t0) / 1e3, (c[:, 63].min() - t0) / 1e3, (c[:, 63].max() - t0) / 1e3,
                rc[:, 1].mean(), rc[:, 2].mean(), rc[:, 3].mean(), rc[:, 5].mean()))
        ends = sorted((crows[j][:, 63].max() - t0) / 1e3 for j in order)
        print("end-to-end spacing (us):", np.round(np.diff(ends), 2).tolist())
        mid = rows[order[NOPS // 2]]
        tiles = (mid[:, 8:32] - mid[:, 0:1]) / 1.965 / 1e3
        nt = int((mid[0, 8:32] > 0).sum())
        print("middle launch, CTA 0 tiles ready (us):", np.round(tiles[0, :nt], 2).tolist())
        smid = mid[:, 61]
        cnt = np.bincount(smid.astype(np.int64), minlength=148)
        print("middle launch: STREAM CTAs on %d distinct SMs; CTAs per SM histogram (0,1,2,3+):" % int((cnt > 0).sum()),
              [int((cnt == 0).sum()), int((cnt == 1).sum()), int((cnt == 2).sum()), int((cnt >= 3).sum())])
        return
    t0 = min(r[:, 62].min() for r in rows)
    ghz = 1.965
    order = np.argsort([r[:, 62].min() for r in rows])
    print("streams %d shape %d lead %d; times in us since the first of the last %d launches started" % (S, shape, lead, NOPS))
    print("launch  start(min..max)    end(min..max)     life   | per-CTA means since CTA start: sync  p1_end  x1_done  sum_done  p2_done  x2_done  end")
    prev_end = None
    for j in order:
        ts = rows[j]
        st = (ts[:, 62] - t0) / 1e3
        en = (ts[:, 63] - t0) / 1e3
        rel = (ts[:, 0:8] - ts[:, 0:1]) / ghz / 1e3
        # slots: 0 start, 1 cluster sync, 2 phase1 end, 3 exchange-1 done, 6 quarter sum done, 7 phase 2 done, 4 exchange-2 done, 5 end
        m = lambda k: rel[:, k].mean()
        print("%4d   %6.2f..%6.2f   %6.2f..%6.2f   %6.2f  | %5.2f  %5.2f  %5.2f  %5.2f  %5.2f  %5.2f  %5.2f"
              % (j, st.min(), st.max(), en.min(), en.max(), en.max() - st.min(), m(1), m(2), m(3), m(6), m(7), m(4), m(5)))
    ends = sorted((rows[j][:, 63].max() - t0) / 1e3 for j in order)
    print("end-to-end spacing (us):", np.round(np.diff(ends), 2).tolist())
    # first-tile-ready and tile cadence of CTA 0 of a middle launch
    mid = rows[order[NOPS // 2]]
    tiles = (mid[:, 8:32] - mid[:, 0:1]) / ghz / 1e3
    nt = int((mid[0, 8:32] > 0).sum())
    print("middle launch, CTA 0 tiles ready (us):", np.round(tiles[0, :nt], 2).tolist())
    print("middle launch, CTA 64 (or last) tiles ready (us):", np.round(tiles[min(64, len(tiles) - 1), :nt], 2).tolist())


main()
