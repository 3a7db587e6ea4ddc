#!/usr/bin/env python
"""Time the dense / tensor-bound configurations (BASELINE configs 3 and 5, PER_CLASS) on one GPU."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, torch
from attentionalpoolingaction_b200.custom_ops import attn_pool_factory as ops
from oracle import attn_pool_oracle as orc

def timeit(fn, n=10):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

def case(name, B, HW, K, flags, flops_per_img, check=4):
    fl = dict(orc.DEFAULT_FLAGS, **flags)
    params = orc.make_params(K, 2048, flags=fl, seed=1)
    bits, rank = ops.flags_from_cfg_net(fl)
    op = ops.AttnPoolOp({k: torch.from_numpy(v) for k, v in params.items()}, K, HW, rank=rank, flags=bits)
    X = torch.relu(torch.randn((B, HW, 2048), device="cuda"))
    ms = timeit(lambda: op(X))
    ref, _ = orc.attn_pool_head(X[:check].cpu().numpy(), params, fl, K, want_pose=False)
    err = orc.rel_err(op(X)[0][:check].cpu().numpy(), ref)
    print("%-34s B=%4d  %8.3f ms  %10.0f img/s  %7.1f TFLOP/s (algorithmic, fp32-equivalent)  rel err %.2e"
          % (name, B, ms, B / ms * 1e3, flops_per_img * B / ms / 1e9, err), flush=True)

def case_frames(name, B, HW, K, F):
    fl = dict(orc.DEFAULT_FLAGS, SINGLE_LAYER_ATT=True)
    params = orc.make_params(K, 2048, flags=fl, seed=1)
    bits, rank = ops.flags_from_cfg_net(fl)
    op = ops.AttnPoolOp({k: torch.from_numpy(v) for k, v in params.items()}, K, HW, frames=F, rank=rank, flags=bits)
    xs = [torch.relu(torch.randn((B, HW, 2048), device="cuda")) for _ in range(3)]   # 3 x 514 MB > L2
    it = [0]
    def run():
        it[0] += 1
        op(xs[it[0] % 3])
    ms = timeit(run, 30)
    ref, _ = orc.attn_pool_head(xs[0][:F].cpu().numpy(), params, fl, K, frames_per_video=F, want_pose=False)
    err = orc.rel_err(op(xs[0])[0][:1].cpu().numpy(), ref)
    gbs = B * HW * 2048 * 4 / ms / 1e6
    print("%-34s B=%4d  %8.3f ms  %10.0f crop-maps/s  %7.1f GB/s of X (%.2f of 6574.8)  rel err %.2e"
          % (name, B, ms, B / ms * 1e3, gbs, gbs / 6574.8, err), flush=True)


HW, C = 196, 2048
case_frames("config4/GPU: HICO K=600 10-crop", 320, HW, 600, 10)
case("config3: pose-attn rank2 (003)", 128, HW, 393, dict(RANK=2), 2 * HW * C * 768)
case("per-class attention K=393", 32, HW, 393, dict(SINGLE_LAYER_ATT=True, PER_CLASS=True), 2 * HW * C * 2 * 393)
case("topdown end-point requested", 32, HW, 393, dict(SINGLE_LAYER_ATT=True, PER_CLASS=True, RANK=2), 2 * HW * C * 2 * 393 * 2)
Xg = torch.relu(torch.randn((8, 784, 2048), device="cuda"))
ms = timeit(lambda: ops.gram(Xg), 3)
print("config5: gram 28x28x2048 B=8            %8.3f ms  %7.1f TFLOP/s" % (ms, 2 * 784 * 2048 * 2048 * 8 / ms / 1e9))
