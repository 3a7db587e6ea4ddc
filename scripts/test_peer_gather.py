#!/usr/bin/env python
"""torchrun --nproc-per-node N scripts/test_peer_gather.py : the fused peer-store all-gather on N GPUs vs NCCL."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, torch, torch.distributed as dist
from attentionalpoolingaction_b200 import dist as apa_dist
from attentionalpoolingaction_b200.custom_ops import attn_pool_factory as ops
from oracle import attn_pool_oracle as orc

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
B, HW, C, K = 32, 196, 2048, 393
fl = dict(orc.DEFAULT_FLAGS, SINGLE_LAYER_ATT=True)
params = orc.make_params(K, C, flags=fl, seed=3)
bits, rnk = ops.flags_from_cfg_net(fl)
op = ops.AttnPoolOp({k: torch.from_numpy(v) for k, v in params.items()}, K, HW, rank=rnk, flags=bits, device=dev)
pg = apa_dist.PeerGather(B, K, dev)
ok = True
for it in range(6):
    X = torch.relu(torch.randn((B, HW, C), device=dev, generator=torch.Generator(device=dev).manual_seed(100 * it + rank)))
    local = op(X)[0].clone()
    ref = apa_dist.all_gather_logits(local)
    op.forward_gather(X, pg)
    torch.cuda.synchronize()
    same = torch.equal(pg.out, ref)
    ok = ok and same
    if rank == 0:
        print("iter %d gathered == nccl all_gather: %s (max abs diff %.3e)" % (it, same, (pg.out - ref).abs().max().item()), flush=True)
t = torch.tensor([1 if ok else 0], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MIN)
if rank == 0:
    print("PEER_GATHER_OK" if t.item() == 1 else "PEER_GATHER_MISMATCH", flush=True)
dist.barrier(); torch.cuda.synchronize(); sys.stdout.flush(); os._exit(0 if t.item() == 1 else 1)
