#!/usr/bin/env python
"""torchrun --nproc-per-node N scripts/test_peer_gather.py : the fused peer-store all-gather on N GPUs against
ncclAllGather, bit for bit -- strict (depth 1), pipelined (ring of 3, consumer 2 steps behind, explicit wait), both
forms of the head, ranks deliberately skewed, and under CUDA-graph replay."""
import os, sys, time
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
P = {k: torch.from_numpy(v) for k, v in params.items()}
xs = [torch.relu(torch.randn((B, HW, C), device=dev, generator=torch.Generator(device=dev).manual_seed(100 * it + rank)))
      for it in range(8)]
ok = True


def nccl_refs(form):
    """What ncclAllGather makes of the same form's local logits (the two forms differ in the last bits)."""
    plain = ops.AttnPoolOp(P, K, HW, rank=rnk, flags=bits, tuning=ops.make_tuning(form=form), device=dev)
    r = [apa_dist.all_gather_logits(plain(x)[0].clone()) for x in xs]
    torch.cuda.synchronize()
    return r


def check(name, cond):
    global ok
    ok = ok and bool(cond)
    if rank == 0:
        print("%-60s %s" % (name, "OK" if cond else "MISMATCH"), flush=True)


for form, fname in ((ops.APA_FORM_FLOW, "flow"), (ops.APA_FORM_CLUSTER, "cluster")):
    refs = nccl_refs(form)
    # ---- strict: after every call the gathered buffer of that call is complete
    op = ops.AttnPoolOp(P, K, HW, rank=rnk, flags=bits, tuning=ops.make_tuning(form=form), device=dev)
    pg = apa_dist.PeerGather(B, K, dev, depth=1)
    good = True
    for it in range(6):
        if it == 3 and rank == world - 1:
            time.sleep(0.05)                     # a straggling rank
        step = op.forward_gather(xs[it], pg)
        torch.cuda.synchronize()
        good = good and torch.equal(pg.buffer(step), refs[it])
    check("%s: strict, depth 1, 6 calls == nccl all_gather" % fname, good)
    # ---- pipelined: ring of 3, the call only guarantees step - 2; the newest ones after an explicit wait
    op = ops.AttnPoolOp(P, K, HW, rank=rnk, flags=bits, device=dev,
                        tuning=ops.make_tuning(form=form, inputs_stable=1, gather_wait_lag=2))
    pg = apa_dist.PeerGather(B, K, dev, depth=3)
    good = True
    for it in range(8):
        if it == 4 and rank == 0:
            time.sleep(0.05)
        step = op.forward_gather(xs[it], pg)
        if step > 2:
            torch.cuda.synchronize()             # the consumer of step - 2 (stream order would do)
            good = good and torch.equal(pg.buffer(step - 2), refs[it - 2])
        dist.barrier()                           # nobody runs more than a call ahead of a consumer that is still reading
    pg.wait(0)
    torch.cuda.synchronize()
    good = good and torch.equal(pg.buffer(8), refs[7]) and torch.equal(pg.buffer(7), refs[6])
    check("%s: ring of 3, consumer 2 steps behind, explicit wait" % fname, good)

refs = nccl_refs(ops.APA_FORM_FLOW)
# ---- CUDA-graph replay of 6 pipelined calls, three times: the step counter lives on the device
op = ops.AttnPoolOp(P, K, HW, rank=rnk, flags=bits, device=dev,
                    tuning=ops.make_tuning(inputs_stable=1, gather_wait_lag=2))
pg = apa_dist.PeerGather(B, K, dev, depth=3)
side = torch.cuda.Stream()
with torch.cuda.stream(side):
    for it in range(6):
        op.forward_gather(xs[it], pg)
    pg.wait(0)
    torch.cuda.synchronize()
    dist.barrier()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr, stream=side):
        for it in range(6):
            op.forward_gather(xs[it], pg)
        pg.wait(0)
    for _ in range(3):
        gr.replay()
    torch.cuda.synchronize()
dist.barrier()
n = pg.sync_step()
check("graph replay: %d calls counted on the device, last two buffers right" % n,
      n == 24 and torch.equal(pg.buffer(n), refs[5]) and torch.equal(pg.buffer(n - 1), refs[4]))

t = torch.tensor([1 if ok else 0], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MIN)
if rank == 0:
    print("PEER_GATHER_OK" if t.item() == 1 else "PEER_GATHER_MISMATCH", flush=True)
dist.barrier(); torch.cuda.synchronize()
code = 0 if t.item() == 1 else 1
del pg, op
dist.destroy_process_group()
sys.exit(code)
