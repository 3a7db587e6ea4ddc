"""Parity of the CUDA head (through the C ABI) against the CPU oracle.  All tests need a B200.

Tolerance (north_star: <= 1e-4 rel on logits): max|d| / max|ref| <= 1e-4 AND
allclose(rtol=1e-4, atol=1e-4*max|ref|), against the fp64 oracle.  The fp32 SIMT/FMA kernels here
land around 1e-6; the 1e-4 bar is the contract.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TOL = 1e-4


def _setup(B, HW, C, K, flags, seed=0, F=1, pose_channels=768, kp=16):
    import torch
    from oracle import attn_pool_oracle as orc
    fl = dict(orc.DEFAULT_FLAGS, **flags)
    params = orc.make_params(K, C, pose_channels, kp, flags=fl, seed=seed)
    X = orc.make_features(B, HW, C, seed=seed)
    return fl, params, X


def _run(B, HW, C, K, flags, seed=0, F=1, want=(False, False, False), tuning=None, pose_channels=768, kp=16):
    import torch
    from attentionalpoolingaction_b200.custom_ops import attn_pool_factory as ops
    from oracle import attn_pool_oracle as orc
    fl, params, X = _setup(B, HW, C, K, flags, seed, F, pose_channels, kp)
    ref, ref_ep = orc.attn_pool_head(X, params, fl, K, frames_per_video=F, want_pose=want[2])
    bits, rank = ops.flags_from_cfg_net(fl)
    op = ops.AttnPoolOp({k: torch.from_numpy(v) for k, v in params.items()}, K, HW, channels=C, frames=F,
                        rank=rank, flags=bits, num_pose_keypoints=kp, tuning=tuning, device="cuda:0")
    logits, ep = op(torch.from_numpy(X).cuda(), want_attn=want[0], want_topdown=want[1], want_pose=want[2])
    torch.cuda.synchronize()
    return logits.cpu().numpy(), {k: v.cpu().numpy() for k, v in ep.items()}, ref, ref_ep


SINGLE = dict(SINGLE_LAYER_ATT=True)

FUSED_CASES = [
    # (id, B, HW, C, K, flags, F)
    ("cfg002_b4", 4, 196, 2048, 393, SINGLE, 1),
    ("cfg2_full_b32", 32, 196, 2048, 393, SINGLE, 1),
    ("hw225_b3", 3, 225, 2048, 393, SINGLE, 1),
    ("relu", 5, 196, 2048, 393, dict(SINGLE, RELU_ATT=True), 1),
    ("softmax", 5, 196, 2048, 393, dict(SINGLE, SOFTMAX_ATT=True), 1),
    ("softmax_relu", 2, 49, 2048, 101, dict(SINGLE, SOFTMAX_ATT=True, RELU_ATT=True), 1),
    ("rank2", 4, 196, 2048, 393, dict(SINGLE, RANK=2), 1),
    ("rank3_relu", 3, 196, 2048, 393, dict(SINGLE, RANK=3, RELU_ATT=True), 1),
    ("rank4", 2, 100, 2048, 51, dict(SINGLE, RANK=4), 1),
    ("hico_k600_10crop", 20, 196, 2048, 600, SINGLE, 10),
    ("frames3_rank2", 6, 49, 2048, 21, dict(SINGLE, RANK=2), 3),
    ("c1024", 4, 196, 1024, 393, SINGLE, 1),
    ("b1", 1, 196, 2048, 393, SINGLE, 1),
    ("k1", 3, 196, 2048, 1, SINGLE, 1),
    ("hw1", 3, 1, 2048, 17, SINGLE, 1),
    ("hw7_softmax", 3, 7, 2048, 17, dict(SINGLE, SOFTMAX_ATT=True), 1),
    ("multiwave_b200", 200, 49, 2048, 40, SINGLE, 1),
    ("multiwave_b77_softmax", 77, 64, 2048, 33, dict(SINGLE, SOFTMAX_ATT=True), 1),
]


@pytest.mark.parametrize("case", FUSED_CASES, ids=[c[0] for c in FUSED_CASES])
def test_fused_path_parity(case):
    """The single-map fast path in its default form (flow), every flag set and the edge shapes."""
    from oracle import attn_pool_oracle as orc
    _, B, HW, C, K, flags, F = case
    got, _, ref, _ = _run(B, HW, C, K, flags, F=F)
    orc.assert_parity(got, ref, TOL)


CLUSTER_CASES = [
    # the cluster form (APA_FORM_CLUSTER): B % 4 == 0, F == 1, C == 2048 (id, B, HW, K, flags)
    ("b32_k393", 32, 196, 393, SINGLE),
    ("b4_k600_nc3", 4, 196, 600, SINGLE),
    ("b8_softmax", 8, 196, 393, dict(SINGLE, SOFTMAX_ATT=True)),
    ("b8_rank3_relu", 8, 225, 393, dict(SINGLE, RANK=3, RELU_ATT=True)),
    ("b12_k1", 12, 49, 1, SINGLE),
    ("b4_hw3_ragged", 4, 3, 37, dict(SINGLE, SOFTMAX_ATT=True)),
    ("b100_multigroup", 100, 49, 129, SINGLE),
    ("b48_k768_rank2", 48, 36, 768, dict(SINGLE, RANK=2)),
    ("b8_pose_source", 8, 196, 393, dict(RANK=2)),
]


@pytest.mark.parametrize("case", CLUSTER_CASES, ids=[c[0] for c in CLUSTER_CASES])
def test_cluster_form(case):
    """The 16-CTA-cluster kernel (the tail stays in shared memory) against the oracle and the flow form, all of its
    shapes, bit-for-bit run-to-run, with and without the caller's inputs_stable promise."""
    import torch
    from attentionalpoolingaction_b200.custom_ops import attn_pool_factory as ops
    from oracle import attn_pool_oracle as orc
    _, B, HW, K, flags = case
    fl = dict(orc.DEFAULT_FLAGS, **flags)
    bits, rank = ops.flags_from_cfg_net(fl)
    t = ops.make_tuning(form=ops.APA_FORM_CLUSTER)
    plan = ops.attn_pool_plan(ops.make_desc(B, HW, 2048, K, rank=rank, flags=bits), t)
    assert plan["form"] == ops.APA_FORM_CLUSTER and plan["launches"] == (1 if fl["SINGLE_LAYER_ATT"] else 2)
    got, _, ref, _ = _run(B, HW, 2048, K, flags, seed=31, tuning=t)
    orc.assert_parity(got, ref, TOL)
    got2, _, _, _ = _run(B, HW, 2048, K, flags, seed=31)        # the default (flow) form agrees to fp32 rounding
    orc.assert_parity(got2, ref, TOL)
    assert np.abs(got - got2).max() <= 1e-5 * max(1e-30, np.abs(ref).max())
    for shape in (1, 2, 3):
        for stable in (0, 1):
            t = ops.make_tuning(form=ops.APA_FORM_CLUSTER, cluster_shape=shape, inputs_stable=stable)
            got3, _, _, _ = _run(B, HW, 2048, K, flags, seed=31, tuning=t)
            orc.assert_parity(got3, ref, TOL)
            got4, _, _, _ = _run(B, HW, 2048, K, flags, seed=31, tuning=t)
            assert np.array_equal(got3, got4)


def test_cluster_form_falls_back_to_flow_where_it_does_not_apply():
    from attentionalpoolingaction_b200.custom_ops import attn_pool_factory as ops
    from oracle import attn_pool_oracle as orc
    t = ops.make_tuning(form=ops.APA_FORM_CLUSTER)
    for B, C, F in ((9, 2048, 1), (8, 1024, 1), (8, 2048, 2)):
        assert ops.attn_pool_plan(ops.make_desc(B, 196, C, 393, frames=F), t)["form"] == ops.APA_FORM_FLOW
        got, _, ref, _ = _run(B, 196, C, 393, SINGLE, F=F, seed=4, tuning=t)
        orc.assert_parity(got, ref, TOL)


@pytest.mark.parametrize("no_pdl", [0, 1])
def test_back_to_back_launches(no_pdl):
    """Back-to-back calls on one stream overlap (programmatic dependent launch): the next launch streams
    X while the previous one is in its tail and both share one workspace.  Results must not depend on it."""
    import torch
    from attentionalpoolingaction_b200.custom_ops import attn_pool_factory as ops
    from oracle import attn_pool_oracle as orc
    for flags in (SINGLE, dict(SINGLE, SOFTMAX_ATT=True), dict(SINGLE, RANK=2, RELU_ATT=True)):
        fl, params, X = _setup(32, 196, 2048, 393, flags, seed=12)
        bits, rank = ops.flags_from_cfg_net(fl)
        op = ops.AttnPoolOp({k: torch.from_numpy(v) for k, v in params.items()}, 393, 196, rank=rank, flags=bits,
                            tuning=ops.make_tuning(no_pdl=no_pdl), device="cuda:0")
        xs = [torch.from_numpy(np.roll(X, i, axis=0).copy()).cuda() for i in range(4)]
        outs = [torch.empty((32, 393), device="cuda") for _ in range(4)]
        for rep in range(10):                       # 40 launches queued without any host sync
            for i in range(4):
                op(xs[i], logits=outs[i])
        torch.cuda.synchronize()
        ref, _ = orc.attn_pool_head(X, params, fl, 393, want_pose=False)
        for i in range(4):
            orc.assert_parity(outs[i].cpu().numpy(), np.roll(ref, i, axis=0), TOL)


@pytest.mark.parametrize("B,form,depth", [(32, 0, 1), (32, 0, 3), (9, 0, 2), (32, 2, 1), (8, 2, 2)])
def test_fused_gather_single_rank(B, form, depth):
    """apa_attn_pool_fwd_gather with world == 1: the peer-store epilogue, the ring of gathered buffers and the step
    flags on one GPU (both forms)."""
    import torch
    from attentionalpoolingaction_b200 import dist as apa_dist
    from attentionalpoolingaction_b200.custom_ops import attn_pool_factory as ops
    from oracle import attn_pool_oracle as orc
    fl, params, X = _setup(B, 196, 2048, 393, SINGLE, seed=17)
    bits, rank = ops.flags_from_cfg_net(fl)
    op = ops.AttnPoolOp({k: torch.from_numpy(v) for k, v in params.items()}, 393, 196, rank=rank, flags=bits,
                        tuning=ops.make_tuning(form=form), device="cuda:0")
    pg = apa_dist.PeerGather(B, 393, torch.device("cuda:0"), depth=depth)
    xs = [torch.from_numpy(np.roll(X, i, axis=0).copy()).cuda() for i in range(5)]
    steps = [op.forward_gather(xs[i], pg) for i in range(5)]
    torch.cuda.synchronize()
    assert steps == [1, 2, 3, 4, 5]
    ref, _ = orc.attn_pool_head(X, params, fl, 393, want_pose=False)
    for i in range(5 - depth, 5):                # the last `depth` calls are still in the ring
        orc.assert_parity(pg.buffer(steps[i]).cpu().numpy(), np.roll(ref, i, axis=0), TOL)
    assert int(pg.flags[0].item()) == 5        # one data flag per launch
    assert pg.sync_step() == 5                 # counted on the device


def test_two_batches_in_flight():
    """Two ops (two workspaces) on two streams concurrently: results identical to running them alone."""
    import torch
    from attentionalpoolingaction_b200.custom_ops import attn_pool_factory as ops
    from oracle import attn_pool_oracle as orc
    fl, params, X = _setup(32, 196, 2048, 393, SINGLE, seed=8)
    bits, rank = ops.flags_from_cfg_net(fl)
    P = {k: torch.from_numpy(v) for k, v in params.items()}
    opl = [ops.AttnPoolOp(P, 393, 196, rank=rank, flags=bits, device="cuda:0") for _ in range(2)]
    Xd = [torch.from_numpy(X).cuda(), torch.from_numpy(X[::-1].copy()).cuda()]
    alone = [opl[i](Xd[i])[0].clone() for i in range(2)]
    torch.cuda.synchronize()
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    outs = [[], []]
    for it in range(20):
        for i in range(2):
            with torch.cuda.stream(streams[i]):
                outs[i].append(opl[i](Xd[i])[0].clone())
    torch.cuda.synchronize()
    for i in range(2):
        for o in outs[i]:
            assert torch.equal(o, alone[i])
    ref, _ = orc.attn_pool_head(X, params, fl, 393, want_pose=False)
    orc.assert_parity(alone[0].cpu().numpy(), ref, TOL)


def test_config003_pose_attention_rank2():
    """BASELINE config 3: attention from relu(X Wp + bp) (768 ch), RANK=2 chain (quirk Q1)."""
    from oracle import attn_pool_oracle as orc
    got, ep, ref, ref_ep = _run(3, 196, 2048, 393, dict(RANK=2), want=(True, False, True))
    orc.assert_parity(got, ref, TOL)
    orc.assert_parity(ep["PoseLogits"], ref_ep["PoseLogits"], TOL, "PoseLogits")
    orc.assert_parity(ep["PosePrelogitsBasedAttention"], ref_ep["PosePrelogitsBasedAttention"], TOL, "attention")


def test_config003_fused_a0_epilogue():
    """Config 003 without the PoseLogits end-point: the 2048->768 tensor-core GEMM never writes its output,
    its epilogue emits relu(P).w_0 partial dots that the streaming kernel sums (plus the attention end-point)."""
    from oracle import attn_pool_oracle as orc
    for flags in (dict(), dict(RANK=2, RELU_ATT=True), dict(SOFTMAX_ATT=True)):
        got, ep, ref, ref_ep = _run(5, 196, 2048, 393, flags, seed=21, want=(True, False, False))
        orc.assert_parity(got, ref, TOL)
        orc.assert_parity(ep["PosePrelogitsBasedAttention"], ref_ep["PosePrelogitsBasedAttention"], TOL, "attention")


DENSE_CASES = [
    ("per_class", 2, 196, 2048, 393, dict(SINGLE, PER_CLASS=True), 1),
    ("per_class_softmax", 2, 49, 2048, 57, dict(SINGLE, PER_CLASS=True, SOFTMAX_ATT=True), 1),
    ("per_class_relu_rank2", 2, 49, 2048, 37, dict(SINGLE, PER_CLASS=True, RELU_ATT=True, RANK=2), 2),
    ("per_class_pose_src", 2, 49, 2048, 37, dict(PER_CLASS=True), 1),
    ("c512_falls_to_dense", 3, 49, 512, 37, SINGLE, 1),
]


@pytest.mark.parametrize("case", DENSE_CASES, ids=[c[0] for c in DENSE_CASES])
def test_dense_path_parity(case):
    from oracle import attn_pool_oracle as orc
    _, B, HW, C, K, flags, F = case
    got, _, ref, _ = _run(B, HW, C, K, flags, F=F)
    orc.assert_parity(got, ref, TOL)


@pytest.mark.parametrize("flags", [SINGLE, dict(SINGLE, SOFTMAX_ATT=True), dict(SINGLE, RANK=3, RELU_ATT=True),
                                   dict(SINGLE, PER_CLASS=True, RANK=2)],
                         ids=["plain", "softmax", "rank3relu", "perclass_rank2"])
def test_end_points(flags):
    """The three observable end-points (nets_factory.py:160, :287, :309) with the reference's shapes."""
    from oracle import attn_pool_oracle as orc
    K = 29 if flags.get("PER_CLASS") else 393
    got, ep, ref, ref_ep = _run(3, 49, 2048, K, flags, want=(True, True, True))
    orc.assert_parity(got, ref, TOL)
    for name in ("PosePrelogitsBasedAttention", "TopDownAttention", "PoseLogits"):
        assert ep[name].shape == ref_ep[name].shape, name
        orc.assert_parity(ep[name], ref_ep[name], TOL, name)
    if flags.get("SOFTMAX_ATT"):
        np.testing.assert_allclose(ep["PosePrelogitsBasedAttention"].sum(axis=1), 1.0, rtol=1e-5)


def test_bitwise_deterministic():
    import torch
    from attentionalpoolingaction_b200.custom_ops import attn_pool_factory as ops
    from oracle import attn_pool_oracle as orc
    fl, params, X = _setup(32, 196, 2048, 393, SINGLE, seed=5)
    bits, rank = ops.flags_from_cfg_net(fl)
    op = ops.AttnPoolOp({k: torch.from_numpy(v) for k, v in params.items()}, 393, 196, rank=rank, flags=bits,
                        device="cuda:0")
    Xd = torch.from_numpy(X).cuda()
    a = op(Xd)[0].clone()
    for _ in range(5):
        b = op(Xd)[0]
        assert torch.equal(a, b)


def test_properties_at_full_size():
    """Size-independent properties at BASELINE config-4 per-GPU size (320 crop maps, K=600, 10-crop):
    batch-permutation equivariance, crop-mean linearity, linearity in the top-down weights."""
    import torch
    from attentionalpoolingaction_b200.custom_ops import attn_pool_factory as ops
    from oracle import attn_pool_oracle as orc
    B, HW, C, K, F = 320, 196, 2048, 600, 10
    fl = dict(orc.DEFAULT_FLAGS, SINGLE_LAYER_ATT=True)
    params = orc.make_params(K, C, flags=fl, seed=11)
    g = torch.Generator(device="cuda").manual_seed(11)
    X = torch.relu(torch.randn((B, HW, C), device="cuda", generator=g))
    P = {k: torch.from_numpy(v) for k, v in params.items()}
    bits, rank = ops.flags_from_cfg_net(fl)
    op1 = ops.AttnPoolOp(P, K, HW, frames=1, rank=rank, flags=bits, device="cuda:0")
    opF = ops.AttnPoolOp(P, K, HW, frames=F, rank=rank, flags=bits, device="cuda:0")
    per_crop = op1(X)[0]
    pooled = opF(X)[0]
    scale = per_crop.abs().max().item()
    assert torch.equal(op1(X)[0], per_crop)      # the same call again: bit-identical
    # crop mean == mean of per-crop logits
    assert (pooled - per_crop.reshape(B // F, F, K).mean(1)).abs().max().item() <= 1e-5 * scale
    # permuting the batch permutes the logits -- to fp32 rounding: the rows of the whole batch are cut into equal
    # runs, so where an image is cut (and with it the order of its partial sums) depends on its position
    perm = torch.randperm(B, device="cuda", generator=g)
    assert (op1(X[perm].contiguous())[0] - per_crop[perm]).abs().max().item() <= 1e-5 * scale
    # linear in the top-down weights and biases
    P2 = dict(P)
    tw, tb = "PosePrelogitsBasedAttention/Conv/weights", "PosePrelogitsBasedAttention/Conv/biases"
    P2[tw] = P[tw] * 2.0
    P2[tb] = P[tb] * 2.0
    op2 = ops.AttnPoolOp(P2, K, HW, frames=1, rank=rank, flags=bits, device="cuda:0")
    assert (op2(X)[0] - 2.0 * per_crop).abs().max().item() <= 1e-5 * scale
    # and the oracle agrees on a slice of it
    ref, _ = orc.attn_pool_head(X[:20].cpu().numpy(), params, fl, K, frames_per_video=F, want_pose=False)
    orc.assert_parity(pooled[:2].cpu().numpy(), ref, TOL)


def test_host_buffer_entry_point():
    import torch
    from attentionalpoolingaction_b200.custom_ops import attn_pool_factory as ops
    from oracle import attn_pool_oracle as orc
    fl, params, X = _setup(8, 196, 2048, 393, SINGLE, seed=9)
    ref, _ = orc.attn_pool_head(X, params, fl, 393, want_pose=False)
    bits, rank = ops.flags_from_cfg_net(fl)
    op = ops.AttnPoolOp({k: torch.from_numpy(v) for k, v in params.items()}, 393, 196, rank=rank, flags=bits,
                        device="cuda:0")
    out = op.forward_host(torch.from_numpy(X).pin_memory())
    orc.assert_parity(out.numpy(), ref, TOL)


def test_network_fn_shim():
    """get_network_fn(...) -> network_fn(images) -> (logits, end_points), 5-D input folds frames."""
    import torch
    from attentionalpoolingaction_b200 import config, nets_factory
    from oracle import attn_pool_oracle as orc
    cfg = config.cfg_from_dict({"NET": {"USE_POSE_PRELOGITS_BASED_ATTENTION": True,
                                        "USE_POSE_PRELOGITS_BASED_ATTENTION_SINGLE_LAYER_ATT": True}})
    fl = dict(orc.DEFAULT_FLAGS, SINGLE_LAYER_ATT=True)
    params = orc.make_params(393, 2048, flags=fl, seed=2)
    X = orc.make_features(6, 225, 2048, seed=2)
    fn = nets_factory.get_network_fn("resnet_v1_101", 393, 16, cfg,
                                     params={k: torch.from_numpy(v) for k, v in params.items()})
    images = torch.from_numpy(X).cuda().reshape(2, 3, 15, 15, 2048)
    logits, ep = fn(images, end_point_names=("PosePrelogitsBasedAttention",))
    ref, ref_ep = orc.attn_pool_head(X, params, fl, 393, frames_per_video=3, want_pose=False)
    orc.assert_parity(logits.cpu().numpy(), ref, TOL)
    assert tuple(ep["PosePrelogitsBasedAttention"].shape) == (6, 15, 15, 1)
    orc.assert_parity(ep["PosePrelogitsBasedAttention"].reshape(6, 225, 1).cpu().numpy(),
                      ref_ep["PosePrelogitsBasedAttention"], TOL)
    # the functional form runs the same op
    pt = {k: torch.from_numpy(v) for k, v in params.items()}
    logits2, ep2 = nets_factory.attn_pool_head(torch.from_numpy(X).cuda().reshape(6, 15, 15, 2048), pt, fl, 393,
                                               frames_per_video=3)
    orc.assert_parity(logits2.cpu().numpy(), ref, TOL)
    assert list(ep2) == ["Logits"]
    # the reference's per-frame end-points (nets_factory.py:352,357): un-pooled logits on request
    logits3, ep3 = fn(images, end_point_names=("logits_beforePool",))
    assert tuple(ep3["logits_beforePool"].shape) == (6, 393) and ep3["Logits"] is ep3["logits_beforePool"]
    orc.assert_parity(ep3["logits_beforePool"].cpu().numpy(), ref_ep["logits_beforePool"], TOL)
    orc.assert_parity(logits3.cpu().numpy(), ref, TOL)


def test_gram_matches_oracle():
    import torch
    from attentionalpoolingaction_b200.custom_ops import attn_pool_factory as ops
    from oracle import attn_pool_oracle as orc
    for (B, HW, C) in ((2, 49, 256), (3, 196, 512), (1, 784, 2048), (2, 9, 32)):
        X = orc.make_features(B, HW, C, seed=4)
        G = ops.gram(torch.from_numpy(X).cuda()).cpu().numpy()
        orc.assert_parity(G, orc.gram(X), TOL, "gram %s" % ((B, HW, C),))
        assert np.allclose(G, np.transpose(G, (0, 2, 1)), rtol=1e-4, atol=1e-4 * np.abs(G).max())


def test_errors_are_loud():
    import torch
    from attentionalpoolingaction_b200.custom_ops import attn_pool_factory as ops
    from oracle import attn_pool_oracle as orc
    fl = dict(orc.DEFAULT_FLAGS, SINGLE_LAYER_ATT=True, SOFTMAX_ATT=True, RANK=2)
    params = orc.make_params(10, 2048, flags=fl)
    bits, rank = ops.flags_from_cfg_net(fl)
    op = ops.AttnPoolOp({k: torch.from_numpy(v) for k, v in params.items()}, 10, 49, rank=rank, flags=bits,
                        device="cuda:0")
    with pytest.raises(ops.ApaError) as e:
        op(torch.zeros((2, 49, 2048), device="cuda"))
    assert e.value.status == ops.APA_ERR_SOFTMAX_RANK
    with pytest.raises(orc.ReferenceGraphError):
        orc.attn_pool_head(np.zeros((2, 49, 2048), np.float32), params, fl, 10)
