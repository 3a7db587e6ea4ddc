#!/usr/bin/env python
"""Small shapes of every kernel family, for compute-sanitizer (scripts/gpu_sanitize.sh): each form of the single-map
head, the fused all-gather on one rank, the dense path on the tcgen05 GEMM, the Gram, the training path, the eval tail.  Each case is checked against
the oracle so a sanitizer-clean run is also a correct one."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from attentionalpoolingaction_b200 import dist as apa_dist
from attentionalpoolingaction_b200.custom_ops import attn_pool_factory as ops
from oracle import attn_pool_oracle as orc

which = sys.argv[1] if len(sys.argv) > 1 else "all"
S = dict(SINGLE_LAYER_ATT=True)


def head(name, B, HW, C, K, flags, F=1, form=0, gather=False, want=(False, False, False), reps=2, shape=0):
    fl = dict(orc.DEFAULT_FLAGS, **flags)
    params = orc.make_params(K, C, 64, 5, flags=fl, seed=2)
    X = orc.make_features(B, HW, C, seed=2)
    ref, _ = orc.attn_pool_head(X, params, fl, K, frames_per_video=F, want_pose=False)
    bits, rank = ops.flags_from_cfg_net(fl)
    t = ops.make_tuning(form=form, inputs_stable=1, stream_shape=shape)
    op = ops.AttnPoolOp({k: torch.from_numpy(v) for k, v in params.items()}, K, HW, channels=C, frames=F, rank=rank,
                        flags=bits, num_pose_keypoints=5, tuning=t, device="cuda:0")
    Xd = torch.from_numpy(X).cuda()
    if gather:
        pg = apa_dist.PeerGather(B // F, K, torch.device("cuda:0"), depth=2)
        for _ in range(reps + 1):
            op.forward_gather(Xd, pg)
        torch.cuda.synchronize()
        got = pg.buffer(reps + 1).cpu().numpy()
    else:
        for _ in range(reps):           # back-to-back launches: the PDL overlap is part of what is checked
            got = op(Xd, want_attn=want[0], want_topdown=want[1], want_pose=want[2])[0]
        torch.cuda.synchronize()
        got = got.cpu().numpy()
    err = orc.rel_err(got, ref)
    print("%-36s rel err %.2e %s" % (name, err, "OK" if err <= 1e-4 else "MISMATCH"), flush=True)
    assert err <= 1e-4


if which in ("all", "flow"):
    head("flow plain b8", 8, 49, 2048, 37, S, form=ops.APA_FORM_FLOW)
    head("flow softmax b5 odd", 5, 25, 2048, 37, dict(S, SOFTMAX_ATT=True), form=ops.APA_FORM_FLOW)
    head("flow rank2 relu frames2 c1024", 8, 16, 1024, 70, dict(S, RANK=2, RELU_ATT=True), F=2, form=ops.APA_FORM_FLOW)
    head("flow gather (world 1, depth 2)", 8, 49, 2048, 37, S, form=ops.APA_FORM_FLOW, gather=True)
    head("flow attention end point", 4, 25, 2048, 37, dict(S, SOFTMAX_ATT=True), form=ops.APA_FORM_FLOW,
         want=(True, False, False))
    head("flow stream shape 1", 8, 49, 2048, 37, S, form=ops.APA_FORM_FLOW, shape=1)
    head("flow stream shape 2", 8, 49, 2048, 37, dict(S, RELU_ATT=True), form=ops.APA_FORM_FLOW, shape=2)
if which in ("all", "cluster"):
    head("cluster form b8", 8, 49, 2048, 37, S, form=ops.APA_FORM_CLUSTER)
    head("cluster form softmax", 8, 25, 2048, 37, dict(S, SOFTMAX_ATT=True), form=ops.APA_FORM_CLUSTER)
    head("cluster form rank2", 4, 25, 2048, 37, dict(S, RANK=2), form=ops.APA_FORM_CLUSTER)
    head("cluster form gather (world 1)", 8, 49, 2048, 37, S, form=ops.APA_FORM_CLUSTER, gather=True)
if which in ("all", "dense"):
    head("config 003 (tcgen05 pose GEMM + a0)", 8, 64, 2048, 37, dict(RANK=2), form=ops.APA_FORM_FLOW)
    head("per-class dense path", 4, 64, 2048, 40, dict(S, PER_CLASS=True))
    head("top-down end point", 2, 64, 2048, 40, S, want=(True, True, False))
    Xg = torch.relu(torch.randn((2, 64, 256), device="cuda"))
    G = ops.gram(Xg)
    torch.cuda.synchronize()
    err = orc.rel_err(G.cpu().numpy(), orc.gram(Xg.cpu().numpy()))
    print("%-36s rel err %.2e" % ("gram 2 x 64 x 256", err), flush=True)
    assert err <= 1e-4
if which in ("all", "train"):
    from oracle import attn_pool_backward as bwd, eval_tail as et

    def train(name, B, HW, C, K, flags, F=1, keep=0.2):
        fl = dict(orc.DEFAULT_FLAGS, **flags)
        params = orc.make_params(K, C, flags=fl, seed=3)
        bits, rank = ops.flags_from_cfg_net(fl)
        rng = np.random.RandomState(1)
        X = np.maximum(rng.randn(B, HW, C), 0).astype(np.float32)
        dl = rng.randn(B // F, K).astype(np.float32)
        op = ops.AttnPoolTrainOp({k: torch.from_numpy(v) for k, v in params.items()}, K, HW, channels=C, frames=F,
                                 rank=rank, flags=bits, device="cuda:0")
        tr = ops.make_train(keep, 9)
        for _ in range(2):
            op.forward(torch.from_numpy(X).cuda(), tr)
            dX, grads = op.backward(torch.from_numpy(dl).cuda())
        torch.cuda.synchronize()
        mask = bwd.dropout_mask(9, X.shape, keep)
        ref_dX, ref_g = bwd.backward_autograd(X, params, fl, K, dl, frames_per_video=F, dropout_mask=mask)
        err = max([orc.rel_err(dX.cpu().numpy(), ref_dX)] +
                  [float(np.abs(grads[n].cpu().numpy().reshape(g.shape) - g).max() / max(np.abs(g).max(), 1e-3))
                   for n, g in ref_g.items()])
        print("%-36s rel err %.2e %s" % (name, err, "OK" if err <= 1e-4 else "MISMATCH"), flush=True)
        assert err <= 1e-4

    train("train one-pass relu b6", 6, 49, 2048, 37, S)
    train("train one-pass plain c512 frames2", 4, 30, 512, 33, dict(S, RELU_ATT=False), F=2, keep=0.5)
    train("train two-pass softmax", 3, 25, 2048, 37, dict(S, SOFTMAX_ATT=True))
    train("train two-pass rank 3", 3, 25, 1024, 37, dict(S, RANK=3))
    x = torch.randn((19, 37), device="cuda")
    lab = torch.randint(0, 37, (19,), device="cuda", dtype=torch.int32)
    cnt = torch.zeros(2, dtype=torch.int64, device="cuda")
    pred, probs = ops.eval_tail(x, lab, cnt)
    mAP, aps = ops.mean_ap(x, lab)
    torch.cuda.synchronize()
    ref_map, ref_aps = et.compute_map(x.cpu().numpy(), lab.cpu().numpy())
    err = float(np.abs(aps.cpu().numpy() - ref_aps).max())
    print("%-36s rel err %.2e" % ("eval tail + mAP 19 x 37", err), flush=True)
    assert err <= 1e-12 and np.array_equal(pred.cpu().numpy(), et.predictions(x.cpu().numpy()))
    from oracle import losses
    loss, dl = ops.softmax_xent_loss(x, lab, 0.5)
    pl, dp = ops.pose_l2_loss(torch.randn((3, 25, 5), device="cuda"), torch.zeros((3, 25, 5), device="cuda"),
                              torch.ones((3, 5), dtype=torch.bool, device="cuda"))
    torch.cuda.synchronize()
    ref, _ = losses.softmax_xent(x.cpu().numpy(), lab.cpu().numpy(), 0.5)
    print("%-36s rel err %.2e" % ("losses", abs(float(loss) - ref) / abs(ref)), flush=True)
    assert abs(float(loss) - ref) <= 1e-5 * abs(ref)
print("SANITIZE_CASES_DONE", flush=True)
