"""The CUDA head (through the C ABI) against the golden vectors produced by the REFERENCE'S OWN
nets_factory.py (tests/golden/, see tests/golden/make_golden.py and oracle/ref_graph.py)."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
with open(os.path.join(GOLD, "index.json")) as f:
    INDEX = json.load(f)
CASES = sorted(k for k, v in INDEX.items() if "raises" not in v)
TOL = 1e-4


@pytest.mark.parametrize("name", CASES)
def test_cuda_head_matches_reference_graph_fixture(name):
    import torch
    from attentionalpoolingaction_b200.custom_ops import attn_pool_factory as ops
    from oracle import attn_pool_oracle as orc
    meta = INDEX[name]
    z = np.load(os.path.join(GOLD, name + ".npz"))
    variables = {k[4:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("var:")}
    eps = {k[3:]: z[k] for k in z.files if k.startswith("ep:")}
    X = z["X"]
    H, W = meta["hw"]
    bits, rank = ops.flags_from_cfg_net(meta["flags"])
    op = ops.AttnPoolOp(variables, meta["num_classes"], H * W, channels=meta["channels"], frames=meta["frames"],
                        rank=rank, flags=bits, num_pose_keypoints=meta["num_pose_keypoints"], device="cuda:0")
    logits, got = op(torch.from_numpy(X).cuda(), want_attn=True, want_topdown=True,
                     want_pose="PoseLogits" in eps)
    torch.cuda.synchronize()
    orc.assert_parity(logits.cpu().numpy(), z["logits"], TOL, name + " logits")
    for key in ("PosePrelogitsBasedAttention", "TopDownAttention", "PoseLogits"):
        if key in eps:
            ref = eps[key]
            g = got[key].cpu().numpy().reshape(ref.shape)   # the fixture keeps H and W separate
            orc.assert_parity(g, ref, TOL, name + " " + key)
    # without any end-point requested the fast path must give the same logits
    logits2, _ = op(torch.from_numpy(X).cuda())
    torch.cuda.synchronize()
    orc.assert_parity(logits2.cpu().numpy(), z["logits"], TOL, name + " logits (no end-points)")


FORM_CASES = sorted(k for k in CASES if k.startswith("c2048_"))


@pytest.mark.parametrize("name", FORM_CASES)
def test_every_form_of_the_head_sees_a_reference_fixture(name):
    """C = 2048 fixtures from the reference graph through EVERY form of the single-map head (ApaTuning.form), with
    the plan asserted: the headline kernels are pinned by a reference-generated vector directly, not transitively."""
    import torch
    from attentionalpoolingaction_b200.custom_ops import attn_pool_factory as ops
    from oracle import attn_pool_oracle as orc
    meta = INDEX[name]
    z = np.load(os.path.join(GOLD, name + ".npz"))
    variables = {k[4:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("var:")}
    X = torch.from_numpy(z["X"]).cuda()
    H, W = meta["hw"]
    bits, rank = ops.flags_from_cfg_net(meta["flags"])
    forms = [(ops.APA_FORM_FLOW, 0, 0), (ops.APA_FORM_FLOW, 1, 0), (ops.APA_FORM_FLOW, 2, 0)]
    if meta["frames"] == 1:
        forms += [(ops.APA_FORM_CLUSTER, 0, 1), (ops.APA_FORM_CLUSTER, 0, 3)]
    seen = set()
    for form, sshape, cshape in forms:
        t = ops.make_tuning(form=form, stream_shape=sshape, cluster_shape=cshape)
        op = ops.AttnPoolOp(variables, meta["num_classes"], H * W, channels=meta["channels"], frames=meta["frames"],
                            rank=rank, flags=bits, num_pose_keypoints=meta["num_pose_keypoints"], tuning=t,
                            device="cuda:0")
        plan = ops.attn_pool_plan(op.desc(X.shape[0]), t)
        assert plan["path"] == 0 and plan["form"] == form, (form, dict(plan))
        seen.add(plan["form"])
        logits, _ = op(X)
        torch.cuda.synchronize()
        orc.assert_parity(logits.cpu().numpy(), z["logits"], TOL, "%s form %d" % (name, form))
    assert seen == {f for f, _, _ in forms}


def test_reference_build_error_is_an_abi_error():
    from attentionalpoolingaction_b200.custom_ops import attn_pool_factory as ops
    meta = INDEX["softmax_rank2_build_error"]
    bits, rank = ops.flags_from_cfg_net(meta["flags"])
    assert ops.attn_pool_check(ops.make_desc(2, 9, 32, 7, rank=rank, flags=bits)) == ops.APA_ERR_SOFTMAX_RANK
