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
    with pytest.raises(NotImplementedError):
        nets_factory.get_network_fn("resnet_v1_101", 393, 16, cfg, is_training=True, params={})
    cfg.NET.USE_TEMPORAL_ATT = True
    with pytest.raises(NotImplementedError):
        nets_factory.get_network_fn("resnet_v1_101", 393, 16, cfg, params={})


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
