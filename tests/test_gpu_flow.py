"""The flow form of the single-map head (csrc/apa_pool_flow.cu: row-balanced STREAM kernel + REDUCE / CONTRACT /
FINISH tail kernels, per-call buffers kept twice) against the fp64 oracle through the C ABI.  Same tolerance as
tests/test_gpu_parity.py."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TOL = 1e-4
SINGLE = dict(SINGLE_LAYER_ATT=True)


def _flow_tuning(**kw):
    from attentionalpoolingaction_b200.custom_ops import attn_pool_factory as ops
    return ops.make_tuning(form=ops.APA_FORM_FLOW, **kw)


def _run(B, HW, C, K, flags, seed=0, F=1, want=(False, False, False), tuning=None):
    import torch
    from attentionalpoolingaction_b200.custom_ops import attn_pool_factory as ops
    from oracle import attn_pool_oracle as orc
    fl = dict(orc.DEFAULT_FLAGS, **flags)
    params = orc.make_params(K, C, 768, 16, flags=fl, seed=seed)
    X = orc.make_features(B, HW, C, seed=seed)
    ref, ref_ep = orc.attn_pool_head(X, params, fl, K, frames_per_video=F, want_pose=want[2])
    bits, rank = ops.flags_from_cfg_net(fl)
    tuning = tuning if tuning is not None else _flow_tuning()
    plan = ops.attn_pool_plan(ops.make_desc(B, HW, C, K, frames=F, rank=rank, flags=bits), tuning, *want)
    assert plan["form"] == ops.APA_FORM_FLOW and plan["path"] == 0, plan
    op = ops.AttnPoolOp({k: torch.from_numpy(v) for k, v in params.items()}, K, HW, channels=C, frames=F,
                        rank=rank, flags=bits, tuning=tuning, device="cuda:0")
    logits, ep = op(torch.from_numpy(X).cuda(), want_attn=want[0], want_topdown=want[1], want_pose=want[2])
    torch.cuda.synchronize()
    return logits.cpu().numpy(), {k: v.cpu().numpy() for k, v in ep.items()}, ref, ref_ep


FLOW_CASES = [
    # (id, B, HW, C, K, flags, F)
    ("cfg2_b32", 32, 196, 2048, 393, SINGLE, 1),
    ("b1_eval_batch", 1, 196, 2048, 393, SINGLE, 1),
    ("b1_hw225", 1, 225, 2048, 393, SINGLE, 1),
    ("b3_hw225", 3, 225, 2048, 393, SINGLE, 1),
    ("b9_odd", 9, 196, 2048, 393, SINGLE, 1),
    ("b33_odd", 33, 196, 2048, 393, SINGLE, 1),
    ("b128_four_video_tiles", 128, 196, 2048, 393, SINGLE, 1),
    ("relu", 5, 196, 2048, 393, dict(SINGLE, RELU_ATT=True), 1),
    ("softmax", 5, 196, 2048, 393, dict(SINGLE, SOFTMAX_ATT=True), 1),
    ("softmax_b32", 32, 196, 2048, 393, dict(SINGLE, SOFTMAX_ATT=True), 1),
    ("softmax_relu", 2, 49, 2048, 101, dict(SINGLE, SOFTMAX_ATT=True, RELU_ATT=True), 1),
    ("rank2", 4, 196, 2048, 393, dict(SINGLE, RANK=2), 1),
    ("rank3_relu", 3, 196, 2048, 393, dict(SINGLE, RANK=3, RELU_ATT=True), 1),
    ("rank4", 2, 100, 2048, 51, dict(SINGLE, RANK=4), 1),
    ("hico_k600_10crop", 20, 196, 2048, 600, SINGLE, 10),
    ("hico_k600_10crop_b320", 320, 196, 2048, 600, SINGLE, 10),
    ("frames3_rank2", 6, 49, 2048, 21, dict(SINGLE, RANK=2), 3),
    ("frames2_softmax", 8, 36, 2048, 77, dict(SINGLE, SOFTMAX_ATT=True), 2),
    ("c1024", 4, 196, 1024, 393, SINGLE, 1),
    ("c1024_softmax_b40", 40, 64, 1024, 130, dict(SINGLE, SOFTMAX_ATT=True), 1),
    ("k1", 3, 196, 2048, 1, SINGLE, 1),
    ("k64_k65", 3, 30, 2048, 65, SINGLE, 1),
    ("hw1", 3, 1, 2048, 17, SINGLE, 1),
    ("hw7_softmax", 3, 7, 2048, 17, dict(SINGLE, SOFTMAX_ATT=True), 1),
    ("many_small_images_b200", 200, 49, 2048, 40, SINGLE, 1),
    ("b77_hw64_softmax", 77, 64, 2048, 33, dict(SINGLE, SOFTMAX_ATT=True), 1),
    ("pose_source_rank2", 8, 196, 2048, 393, dict(RANK=2), 1),
]


@pytest.mark.parametrize("case", FLOW_CASES, ids=[c[0] for c in FLOW_CASES])
def test_flow_parity(case):
    from oracle import attn_pool_oracle as orc
    _, B, HW, C, K, flags, F = case
    got, _, ref, _ = _run(B, HW, C, K, flags, F=F, seed=11)
    orc.assert_parity(got, ref, TOL)
    got2, _, _, _ = _run(B, HW, C, K, flags, F=F, seed=11)
    assert np.array_equal(got, got2)          # fixed reduction orders: bit-reproducible run to run


@pytest.mark.parametrize("g", [1, 7, 37, 148])
def test_flow_any_number_of_stream_ctas(g):
    """ApaTuning.stream_ctas = number of STREAM CTAs: any cut of the row space, and every shape of the STREAM kernel,
    gives the same logits to fp32 rounding."""
    from oracle import attn_pool_oracle as orc
    for flags in (SINGLE, dict(SINGLE, SOFTMAX_ATT=True), dict(SINGLE, RANK=2, RELU_ATT=True)):
        for shape in (0, 1, 2):
            got, _, ref, _ = _run(9, 196, 2048, 393, flags, seed=3, tuning=_flow_tuning(stream_ctas=g, stream_shape=shape))
            orc.assert_parity(got, ref, TOL)


@pytest.mark.parametrize("lead", [0, -1, 3])
@pytest.mark.parametrize("stable", [0, 1])
@pytest.mark.parametrize("pdl", [0, 1])
def test_flow_back_to_back_same_workspace(stable, pdl, lead):
    """40 launches queued on one stream, one workspace (the counters only ever grow), changing inputs; with
    inputs_stable=1 consecutive launches overlap through programmatic dependent launch."""
    import torch
    from attentionalpoolingaction_b200.custom_ops import attn_pool_factory as ops
    from oracle import attn_pool_oracle as orc
    fl = dict(orc.DEFAULT_FLAGS, **SINGLE)
    params = orc.make_params(393, 2048, flags=fl, seed=5)
    X = orc.make_features(32, 196, 2048, seed=5)
    bits, rank = ops.flags_from_cfg_net(fl)
    op = ops.AttnPoolOp({k: torch.from_numpy(v) for k, v in params.items()}, 393, 196, rank=rank, flags=bits,
                        tuning=_flow_tuning(inputs_stable=stable, no_pdl=1 - pdl, pdl_trigger_lead=lead), device="cuda:0")
    Xd = torch.from_numpy(X).cuda()
    xs = [torch.roll(Xd, i, 0).contiguous() for i in range(4)]
    outs = [torch.empty((32, 393), device="cuda") for _ in range(4)]
    torch.cuda.synchronize()
    for rep in range(10):
        for i in range(4):
            op(xs[i], logits=outs[i])
    torch.cuda.synchronize()
    ref, _ = orc.attn_pool_head(X, params, fl, 393, want_pose=False)
    for i in range(4):
        orc.assert_parity(outs[i].cpu().numpy(), np.roll(ref, i, axis=0), TOL)
    # a different shape on the SAME op: different grid, same counters
    X2 = orc.make_features(5, 196, 2048, seed=6)
    got = op(torch.from_numpy(X2).cuda())[0]
    got = op(torch.from_numpy(X2).cuda())[0]
    torch.cuda.synchronize()
    ref2, _ = orc.attn_pool_head(X2, params, fl, 393, want_pose=False)
    orc.assert_parity(got.cpu().numpy(), ref2, TOL)


def test_flow_graph_replay_and_lanes():
    """CUDA-graph capture of 8 launches on two forked lanes (two ops), replayed 5 times."""
    import torch
    from attentionalpoolingaction_b200.custom_ops import attn_pool_factory as ops
    from oracle import attn_pool_oracle as orc
    fl = dict(orc.DEFAULT_FLAGS, **SINGLE)
    params = orc.make_params(393, 2048, flags=fl, seed=8)
    X = orc.make_features(32, 196, 2048, seed=8)
    bits, rank = ops.flags_from_cfg_net(fl)
    P = {k: torch.from_numpy(v) for k, v in params.items()}
    opl = [ops.AttnPoolOp(P, 393, 196, rank=rank, flags=bits, tuning=_flow_tuning(inputs_stable=1), device="cuda:0")
           for _ in range(2)]
    Xd = torch.from_numpy(X).cuda()
    outs = [torch.empty((32, 393), device="cuda") for _ in range(2)]
    side, lane = torch.cuda.Stream(), torch.cuda.Stream()

    def enqueue():
        cur = torch.cuda.current_stream()
        lane.wait_stream(cur)
        for i in range(8):
            with torch.cuda.stream(cur if i % 2 == 0 else lane):
                opl[i % 2](Xd, logits=outs[i % 2])
        cur.wait_stream(lane)

    with torch.cuda.stream(side):
        enqueue()
        torch.cuda.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr, stream=side):
            enqueue()
        for _ in range(5):
            outs[0].zero_(); outs[1].zero_()
            gr.replay()
        torch.cuda.synchronize()
    ref, _ = orc.attn_pool_head(X, params, fl, 393, want_pose=False)
    for o in outs:
        orc.assert_parity(o.cpu().numpy(), ref, TOL)


def test_flow_attention_end_point_and_config003():
    from oracle import attn_pool_oracle as orc
    for flags in (SINGLE, dict(SINGLE, SOFTMAX_ATT=True), dict(SINGLE, RANK=2, RELU_ATT=True), dict(RANK=2)):
        got, ep, ref, ref_ep = _run(5, 196, 2048, 393, flags, seed=21, want=(True, False, False))
        orc.assert_parity(got, ref, TOL)
        orc.assert_parity(ep["PosePrelogitsBasedAttention"], ref_ep["PosePrelogitsBasedAttention"], TOL, "attention")
