#!/usr/bin/env python
"""Time the training path (apa_attn_pool_fwd_train + apa_attn_pool_bwd) on one GPU at BASELINE config 2's shape.
Bytes-based roofline: the forward reads X once, the backward reads it once and writes dX once (csrc/apa_train.cu,
the one-pass kernels that serve rank 1 with relu / plain attention); 3 |X| per step is the floor."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from attentionalpoolingaction_b200.custom_ops import attn_pool_factory as ops
from oracle import attn_pool_oracle as orc

B, HW, C, K = (int(sys.argv[1]) if len(sys.argv) > 1 else 32), 196, 2048, 393
keep = float(sys.argv[2]) if len(sys.argv) > 2 else 0.2
fl = dict(orc.DEFAULT_FLAGS, SINGLE_LAYER_ATT=True)
params = orc.make_params(K, C, flags=fl, seed=1)
bits, rnk = ops.flags_from_cfg_net(fl)
dev = torch.device("cuda:0")
op = ops.AttnPoolTrainOp({k: torch.from_numpy(v) for k, v in params.items()}, K, HW, rank=rnk, flags=bits, device=dev)
nbuf = max(3, (3 * 126 * 2**20) // (B * HW * C * 4) + 1)
xs = [torch.relu(torch.randn((B, HW, C), device=dev)) for _ in range(nbuf)]
dl = torch.randn((B, K), device=dev)
dX = torch.empty_like(xs[0])
train = ops.make_train(keep, 5)
grads = None

def timeit(fn, n):
    for i in range(5): fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n): fn(i)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3

def fwd(i): op.forward(xs[i % nbuf], train)
def step(i):
    global grads
    op.forward(xs[i % nbuf], train)
    _, grads = op.backward(dl, want_dX=dX, grads=grads)
def step_frozen(i):
    global grads
    op.forward(xs[i % nbuf], train)
    _, grads = op.backward(dl, want_dX=False, grads=grads)

if len(sys.argv) > 3 and sys.argv[3] == "profile":   # under ncu: a few whole steps, nothing else
    for i in range(6): step(i)
    torch.cuda.synchronize(); sys.exit(0)
xb = B * HW * C * 4
t_f, t_s, t_z = timeit(fwd, 200), timeit(step, 200), timeit(step_frozen, 200)
peak = 6574.8
try: peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception: pass
out = {"workload": "training step of the head, B=%d 14x14x2048 -> %d, dropout keep %.2f" % (B, K, keep),
       "fwd_train_us": t_f, "fwd_plus_bwd_us": t_s, "fwd_plus_bwd_frozen_backbone_us": t_z,
       "images_per_s": B / t_s * 1e6,
       "roofline": {"bound": "hbm", "bytes_per_step": 3 * xb, "achieved_gbs": 3 * xb / t_s / 1e3, "peak_gbs": peak,
                    "frac": 3 * xb / t_s / 1e3 / peak,
                    "fwd_only": {"bytes": xb, "achieved_gbs": xb / t_f / 1e3, "frac": xb / t_f / 1e3 / peak},
                    "note": "algorithmic bytes: |X| read forward, |X| read + |X| written backward"}}
print(json.dumps(out))
