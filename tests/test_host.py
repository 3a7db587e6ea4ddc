"""CPU: host-side mirror of the reference interface — config keys, flag mapping, network_fn argument
checks, shard arithmetic.  No kernels are launched."""
import pytest


def test_cfg_defaults_and_strict_merge():
    from attentionalpoolingaction_b200 import config
    cfg = config.default_cfg()
    assert cfg.NET.USE_POSE_PRELOGITS_BASED_ATTENTION is False
    assert cfg.NET.USE_POSE_PRELOGITS_BASED_ATTENTION_RANK == 1 and cfg.NET.DROPOUT == -1.0
    cfg = config.cfg_from_dict({"NET": {"USE_POSE_PRELOGITS_BASED_ATTENTION": True,
                                        "USE_POSE_PRELOGITS_BASED_ATTENTION_RANK": 2}})
    assert cfg.NET.USE_POSE_PRELOGITS_BASED_ATTENTION_RANK == 2
    with pytest.raises(KeyError):
        config.cfg_from_dict({"NET": {"NOT_A_KEY": 1}})
    with pytest.raises(ValueError):
        config.cfg_from_dict({"NET": {"USE_POSE_PRELOGITS_BASED_ATTENTION_RANK": "two"}})


def test_cfg_from_reference_style_yaml(tmp_path):
    from attentionalpoolingaction_b200 import config
    from attentionalpoolingaction_b200.custom_ops import attn_pool_factory as ops
    y = tmp_path / "002.yaml"     # the NET block of experiments/002_MPII_ResNet_withAttention.yaml:5-7
    y.write_text("GPUS: '0,1,2,3'\nMODEL_NAME: 'resnet_v1_101'\nNET:\n"
                 "  USE_POSE_PRELOGITS_BASED_ATTENTION: True\n"
                 "  USE_POSE_PRELOGITS_BASED_ATTENTION_SINGLE_LAYER_ATT: True\nTRAIN:\n  ITER_SIZE: 2\n")
    cfg = config.cfg_from_file(str(y))
    bits, rank = ops.flags_from_cfg_net(cfg.NET)
    assert bits == ops.APA_FLAG_SINGLE_LAYER_ATT and rank == 1
    y3 = tmp_path / "003.yaml"    # experiments/003_*.yaml:6-7: attention from the pose pre-logits
    y3.write_text("NET:\n  USE_POSE_PRELOGITS_BASED_ATTENTION: True\n")
    bits, rank = ops.flags_from_cfg_net(config.cfg_from_file(str(y3)).NET)
    assert bits == 0 and rank == 1


def test_flag_mapping_accepts_short_and_full_names():
    from attentionalpoolingaction_b200.custom_ops import attn_pool_factory as ops
    bits, rank = ops.flags_from_cfg_net({"SINGLE_LAYER_ATT": True, "RELU_ATT": True, "RANK": 3})
    assert bits == ops.APA_FLAG_SINGLE_LAYER_ATT | ops.APA_FLAG_RELU_ATT and rank == 3
    bits, _ = ops.flags_from_cfg_net({"USE_POSE_PRELOGITS_BASED_ATTENTION_PER_CLASS": True,
                                      "USE_POSE_PRELOGITS_BASED_ATTENTION_SOFTMAX_ATT": True})
    assert bits == ops.APA_FLAG_PER_CLASS | ops.APA_FLAG_SOFTMAX_ATT


def test_variable_scope_names_match_reference():
    from attentionalpoolingaction_b200.custom_ops import attn_pool_factory as ops
    assert ops.attn_scope(0) == "PosePrelogitsBasedAttention/Conv2d_PrePose_Attn"
    assert ops.attn_scope(2) == "PosePrelogitsBasedAttention/Conv2d_PrePose_Attn2"
    assert ops.topdown_scope(0) == "PosePrelogitsBasedAttention/Conv"
    assert ops.topdown_scope(1) == "PosePrelogitsBasedAttention/Conv_1"


def test_get_network_fn_argument_checks():
    from attentionalpoolingaction_b200 import config, nets_factory
    cfg = config.default_cfg()
    with pytest.raises(ValueError):
        nets_factory.get_network_fn("no_such_net", 393, 16, cfg)
    with pytest.raises(ValueError):
        nets_factory.get_network_fn("resnet_v1_101", 393, 16, cfg, params={})  # head not enabled
    cfg.NET.USE_POSE_PRELOGITS_BASED_ATTENTION = True
    with pytest.raises(ValueError):
        nets_factory.get_network_fn("resnet_v1_101", 393, 16, cfg)            # no variables
    fn = nets_factory.get_network_fn("resnet_v1_101", 393, 16, cfg, is_training=True, params={})
    assert callable(fn) and fn.default_image_size == 224   # training graph: built lazily at the first call
    cfg.NET.USE_TEMPORAL_ATT = True
    with pytest.raises(NotImplementedError):
        nets_factory.get_network_fn("resnet_v1_101", 393, 16, cfg, params={})


def test_head_call_logic_with_a_stand_in_op(monkeypatch):
    """`network_fn` / `attn_pool_head` host logic (frame folding, one op per shape, end-point axes, caching per
    `params` object) with the CUDA op replaced by a recorder: no GPU involved, nothing numeric checked."""
    import torch
    from attentionalpoolingaction_b200 import config, nets_factory
    from attentionalpoolingaction_b200.custom_ops import attn_pool_factory as ops
    built = []

    class FakeOp(object):
        def __init__(self, params, num_classes, hw, channels=2048, frames=1, rank=1, flags=0,
                     num_pose_keypoints=16, tuning=None, device=None):
            self.args = dict(K=num_classes, hw=hw, C=channels, F=frames, rank=rank, flags=flags)
            built.append(self)

        def __call__(self, X, want_attn=False, want_topdown=False, want_pose=False, logits=None):
            B, HW, C = X.shape
            assert HW == self.args["hw"] and C == self.args["C"]
            R = self.args["rank"]
            tail = (R,) if R > 1 else ()
            eps = {"Logits": torch.zeros(B // self.args["F"], self.args["K"])}
            if want_attn:
                eps["PosePrelogitsBasedAttention"] = torch.zeros((B, HW, 1) + tail)
            if want_topdown:
                eps["TopDownAttention"] = torch.zeros((B, HW, self.args["K"]) + tail)
            if want_pose:
                eps["PoseLogits"] = torch.zeros(B, HW, 16)
            return eps["Logits"], eps

    monkeypatch.setattr(ops, "AttnPoolOp", FakeOp)
    cfg = config.default_cfg()
    cfg.NET.USE_POSE_PRELOGITS_BASED_ATTENTION = True
    cfg.NET.USE_POSE_PRELOGITS_BASED_ATTENTION_SINGLE_LAYER_ATT = True
    cfg.NET.USE_POSE_PRELOGITS_BASED_ATTENTION_RANK = 2
    params = {"dummy": 0}
    fn = nets_factory.get_network_fn("resnet_v1_101", 393, 16, cfg, params=params)
    logits, eps = fn(torch.zeros(6, 7, 7, 2048), end_point_names=("PosePrelogitsBasedAttention", "TopDownAttention"))
    assert logits.shape == (6, 393)
    assert eps["PosePrelogitsBasedAttention"].shape == (6, 7, 7, 1, 2)      # nets_factory.py:271-274,287
    assert eps["TopDownAttention"].shape == (6, 7, 7, 393, 2)               # :305-309
    logits, _ = fn(torch.zeros(2, 3, 7, 7, 2048))                           # [B, F, H, W, C] (:119-125)
    assert logits.shape == (2, 393) and built[-1].args["F"] == 3
    fn(torch.zeros(4, 7, 7, 2048))
    assert len(built) == 2                                                  # same (HW, C, frames): op re-used

    flags = {"SINGLE_LAYER_ATT": True, "RANK": 1}
    n0 = len(built)
    logits, eps = nets_factory.attn_pool_head(torch.zeros(8, 49, 2048), params, flags, 600)
    assert logits.shape == (8, 600) and list(eps) == ["Logits"]
    nets_factory.attn_pool_head(torch.zeros(8, 49, 2048), params, flags, 600)
    assert len(built) == n0 + 1                                             # cached per params object
    logits, eps = nets_factory.attn_pool_head(torch.zeros(20, 7, 7, 2048), params, flags, 600, frames_per_video=10,
                                              end_point_names=("PoseLogits",))
    assert logits.shape == (2, 600) and eps["PoseLogits"].shape == (20, 7, 7, 16)
    nets_factory.attn_pool_head(torch.zeros(8, 49, 2048), dict(params), flags, 600)
    assert len(built) == n0 + 3                                             # another params object: another op
    with pytest.raises(NotImplementedError):
        nets_factory.attn_pool_head(torch.zeros(8, 49, 2048), params, flags, 600, is_training=True)
    with pytest.raises(ValueError):
        nets_factory.attn_pool_head(torch.zeros(8, 49, 2048), None, flags, 600)
    # per-frame end-points (nets_factory.py:352,357): 'Logits' / 'logits_beforePool' are the UN-pooled logits
    logits, eps = nets_factory.attn_pool_head(torch.zeros(20, 7, 7, 2048), params, flags, 600, frames_per_video=10,
                                              end_point_names=("logits_beforePool",))
    assert logits.shape == (2, 600) and eps["logits_beforePool"].shape == (20, 600) and eps["Logits"].shape == (20, 600)
    assert built[-1].args["F"] == 1                                         # ... served by a frames=1 op
    # a different tuning is a different cache entry; invalidation drops the device copies of one params object
    n1 = len(built)
    nets_factory.attn_pool_head(torch.zeros(8, 49, 2048), params, flags, 600, tuning=ops.make_tuning(stream_ctas=37))
    assert len(built) == n1 + 1
    nets_factory.invalidate_head_cache(params)
    nets_factory.attn_pool_head(torch.zeros(8, 49, 2048), params, flags, 600)
    assert len(built) == n1 + 2
    for i in range(12):                                                     # bounded: at most 8 params objects kept
        nets_factory.attn_pool_head(torch.zeros(8, 49, 2048), {"dummy": i}, flags, 600)
    assert len(nets_factory._HEAD_STATE) <= 8


def test_shard_bounds_cover_batch_exactly():
    from attentionalpoolingaction_b200 import dist
    for n in (1, 7, 8, 32, 256, 257):
        for world in (1, 2, 4, 8):
            spans = [dist.shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        dist.shard_bounds(8, 3, 2)


def test_shard_features_keeps_videos_whole():
    import torch
    from attentionalpoolingaction_b200 import dist
    X = torch.arange(30).reshape(30, 1)
    parts = [dist.shard_features(X, 10, r, 2) for r in range(2)]
    assert [p.shape[0] for p in parts] == [20, 10]
    with pytest.raises(ValueError):
        dist.shard_features(X, 7, 0, 2)


def test_training_network_fn_host_logic(monkeypatch):
    """get_network_fn(is_training=True) with the CUDA op replaced by a recorder: keep probability from cfg.NET.DROPOUT
    (nets_factory.py:143-146), a new dropout seed per call, the autograd node hands every variable its gradient."""
    import torch
    from attentionalpoolingaction_b200 import config, nets_factory
    from attentionalpoolingaction_b200.custom_ops import attn_pool_factory as ops
    seen = []

    class FakeTrainOp(object):
        def __init__(self, params, num_classes, hw, channels=2048, frames=1, rank=1, flags=0, device=None, **kw):
            self.K, self.F, self.rank, self.C = num_classes, frames, rank, channels

        def forward(self, X, train=None, logits=None):
            seen.append((float(train.dropout_keep), int(train.dropout_seed), tuple(X.shape)))
            self.X = X
            return torch.zeros(X.shape[0] // self.F, self.K)

        def backward(self, dlogits, want_dX=True, grads=None):
            g = {}
            for r in range(self.rank):
                g[ops.attn_scope(r) + "/weights"] = torch.full((self.C, 1) if r == 0 else (1, 1), 1.0 + r)
                g[ops.attn_scope(r) + "/biases"] = torch.full((1,), 2.0 + r)
                g[ops.topdown_scope(r) + "/weights"] = torch.full((self.C, self.K), 3.0 + r)
                g[ops.topdown_scope(r) + "/biases"] = torch.full((self.K,), 4.0 + r)
            return (torch.full_like(self.X, 5.0) if want_dX else None), g

    monkeypatch.setattr(ops, "AttnPoolTrainOp", FakeTrainOp)
    monkeypatch.setattr(torch.Tensor, "is_cuda", property(lambda self: True))   # the shim insists on CUDA variables
    cfg = config.default_cfg()
    cfg.NET.USE_POSE_PRELOGITS_BASED_ATTENTION = True
    cfg.NET.USE_POSE_PRELOGITS_BASED_ATTENTION_SINGLE_LAYER_ATT = True
    cfg.NET.USE_POSE_PRELOGITS_BASED_ATTENTION_RANK = 2
    C, K = 8, 5
    params = {}
    for r in range(2):
        params[ops.attn_scope(r) + "/weights"] = torch.zeros((C, 1) if r == 0 else (1, 1), requires_grad=True)
        params[ops.attn_scope(r) + "/biases"] = torch.zeros(1, requires_grad=True)
        params[ops.topdown_scope(r) + "/weights"] = torch.zeros(C, K, requires_grad=True)
        params[ops.topdown_scope(r) + "/biases"] = torch.zeros(K, requires_grad=True)
    fn = nets_factory.get_network_fn("resnet_v1_101", K, 16, cfg, is_training=True, params=params)
    X = torch.zeros(4, 3, 3, C, requires_grad=True)
    logits, eps = fn(X)
    logits.sum().backward()
    assert seen[-1] == (pytest.approx(0.2), 1, (4, 9, C))                   # DROPOUT = -1 -> keep 0.2; first seed 1
    assert X.grad.shape == X.shape and float(X.grad[0, 0, 0, 0]) == 5.0
    assert float(params[ops.topdown_scope(1) + "/weights"].grad[0, 0]) == 4.0
    assert float(params[ops.attn_scope(1) + "/biases"].grad[0]) == 3.0
    fn(X)
    assert seen[-1][1] == 2                                                  # a new mask every call
    fn(X, dropout_seed=77)
    assert seen[-1][1] == 77
    cfg.NET.DROPOUT = 0.3                                                    # keep = 1 - DROPOUT (:145-146)
    fn2 = nets_factory.get_network_fn("resnet_v1_101", K, 16, cfg, is_training=True, params=params)
    fn2(X.detach())                                                          # frozen backbone: no dX requested
    assert seen[-1][0] == pytest.approx(0.7)
    with pytest.raises(NotImplementedError):
        fn(X, end_point_names=("PoseLogits",))
