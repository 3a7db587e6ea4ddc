"""`get_network_fn` shim with the reference's signature and return contract
(models/slim/nets/nets_factory.py:94-118, :375-380), for callers that hand over conv5 features.

    network_fn = get_network_fn('resnet_v1_101', num_classes, num_pose_keypoints, cfg, params=variables)
    logits, end_points = network_fn(last_conv)      # eval.py:181 / train.py:399 call shape

`images` here is the backbone's last conv map (end_points['resnet_v1_101/block4'],
nets_factory.py:63-67,137): [B,H,W,C] or, for videos, [B,F,H,W,C] which is folded to
[B*F,H,W,C] and whose logits are averaged over F (nets_factory.py:119-125, :354-374).
The backbone itself is out of scope.  `params` maps the reference's TF
variable names to arrays (what the checkpoint restore would have put in the variable store).
"""
from __future__ import annotations

import functools

from .custom_ops import attn_pool_factory as ops

# nets_factory.py:63-67
last_conv_map = {
    "inception_v3": "Mixed_7c",
    "inception_v2_tsn": "InceptionV2_TSN/inception_5c",
    "resnet_v1_101": "resnet_v1_101/block4",
    "vgg_16": "vgg_16/conv5",
}
last_conv_channels = {"resnet_v1_101": 2048, "inception_v2_tsn": 1024, "inception_v3": 2048, "vgg_16": 512}


def get_network_fn(name, num_classes, num_pose_keypoints, cfg, weight_decay=0.0, is_training=False,
                   params=None, tuning=None):
    """Returns `network_fn(images) -> (logits, end_points)` for the attentional-pooling head.

    Raises ValueError for unknown `name` (nets_factory.py:113-114), and for head variants that
    are out of scope (anything but USE_POSE_PRELOGITS_BASED_ATTENTION).
    """
    if name not in last_conv_map:
        raise ValueError("Name of network unknown %s" % name)
    net = cfg.NET
    if not net.USE_POSE_PRELOGITS_BASED_ATTENTION:
        raise ValueError("only the USE_POSE_PRELOGITS_BASED_ATTENTION head is built here (DESIGN.md §7)")
    if getattr(net, "USE_POSE_PRELOGITS_BASED_ATTENTION_WITH_POSE_FEAT", False):
        raise NotImplementedError("WITH_POSE_FEAT changes C to 2064 (nets_factory.py:289-295): out of scope")
    if getattr(net, "USE_TEMPORAL_ATT", False):
        raise NotImplementedError("USE_TEMPORAL_ATT (nets_factory.py:361-372) is out of scope")
    if params is None:
        raise ValueError("params (TF variable name -> array) is required: the head has trained variables")
    flags, rank = ops.flags_from_cfg_net(net)
    state = {}
    if is_training:
        # nets_factory.py:143-146: the dropout arg-scope of the head
        drop = float(getattr(net, "DROPOUT", -1.0))
        keep = 0.2 if drop < 0 else 1.0 - drop
        counter = [0]

        def network_fn(images, end_point_names=(), dropout_seed=None):
            """Training graph (train.py:399-420): logits carry a grad_fn; `loss.backward()` fills `.grad` of X and
            of every variable in `params` on the path (they must be float32 CUDA tensors with requires_grad)."""
            if end_point_names:
                raise NotImplementedError("end-points other than Logits are not materialised in training mode")
            if dropout_seed is None:
                counter[0] += 1
                dropout_seed = counter[0]
            logits = _run_head_train(state, images, params, flags, rank, num_classes, keep, dropout_seed)
            return logits, {"Logits": logits}

        network_fn.default_image_size = 224
        return network_fn

    def network_fn(images, end_point_names=()):
        return _run_head(state, images, params, flags, rank, num_classes, num_pose_keypoints, end_point_names, tuning)

    network_fn.default_image_size = 224
    return network_fn


def _run_head(state, images, params, flags, rank, num_classes, num_pose_keypoints, end_point_names, tuning,
              frames_per_video=1):
    """Shared by `network_fn` and `attn_pool_head`: one AttnPoolOp per (HW, C, frames) in `state`."""
    if images.dim() == 5:                       # nets_factory.py:121-125
        frames_per_video = images.shape[1]
        images = images.reshape((-1,) + tuple(images.shape[2:]))
    if images.dim() == 3:                       # [B*F, HW, C]: no spatial axes to restore
        B, HW, C = images.shape
        H = W = None
    else:
        B, H, W, C = images.shape
        HW = H * W
    key = (HW, C, frames_per_video)
    op = state.get(key)
    if op is None:
        op = ops.AttnPoolOp(params, num_classes, HW, channels=C, frames=frames_per_video, rank=rank,
                            flags=flags, num_pose_keypoints=num_pose_keypoints, tuning=tuning,
                            device=images.device)
        state[key] = op
    want = set(end_point_names)
    per_frame = frames_per_video > 1 and (want & {"Logits", "logits_beforePool"})
    if per_frame:
        # nets_factory.py:352,357: with several frames per video the reference's end_points['Logits'] and
        # ['logits_beforePool'] are the UN-pooled [B*F, K] logits and only the returned tensor is the frame mean
        # (:354-374).  The fast path pools inside the kernel, so these two end-points cost a frames=1 call and the
        # mean is taken from its result (same numbers up to fp32 rounding).
        key1 = (HW, C, 1)
        op1 = state.get(key1)
        if op1 is None:
            op1 = ops.AttnPoolOp(params, num_classes, HW, channels=C, frames=1, rank=rank, flags=flags,
                                 num_pose_keypoints=num_pose_keypoints, tuning=tuning, device=images.device)
            state[key1] = op1
        op = op1
    logits, eps = op(images.reshape(B, HW, C),
                     want_attn="PosePrelogitsBasedAttention" in want,
                     want_topdown="TopDownAttention" in want,
                     want_pose="PoseLogits" in want)
    if per_frame:
        eps["Logits"] = logits
        eps["logits_beforePool"] = logits
        logits = logits.reshape(B // frames_per_video, frames_per_video, -1).mean(dim=1)
    if H is not None:   # restore the spatial axes the reference's end-points carry
        for k in ("PosePrelogitsBasedAttention", "TopDownAttention", "PoseLogits"):
            if k in eps:
                t = eps[k]
                eps[k] = t.reshape((B, H, W) + tuple(t.shape[2:]))
    return logits, eps


def _run_head_train(state, images, params, flags, rank, num_classes, keep, seed, frames_per_video=1):
    """Training-mode call: one AttnPoolTrainOp per (HW, C, frames); autograd sees one node for the whole head."""
    import torch
    if images.dim() == 5:
        frames_per_video = images.shape[1]
        images = images.reshape((-1,) + tuple(images.shape[2:]))
    X = images if images.dim() == 3 else images.reshape(images.shape[0], -1, images.shape[-1])
    _, HW, C = X.shape
    key = ("train", HW, C, frames_per_video)
    op = state.get(key)
    if op is None:
        op = ops.AttnPoolTrainOp(params, num_classes, HW, channels=C, frames=frames_per_video, rank=rank,
                                 flags=flags, device=X.device)
        state[key] = op
    names = []
    for r in range(rank):
        names += [ops.attn_scope(r) + "/weights", ops.attn_scope(r) + "/biases",
                  ops.topdown_scope(r) + "/weights", ops.topdown_scope(r) + "/biases"]
    variables = [params[n] for n in names]
    for n, t in zip(names, variables):
        if not (torch.is_tensor(t) and t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
            raise TypeError("training mode reads the variables in place: %r must be a contiguous float32 CUDA tensor" % n)
    train = ops.make_train(keep, seed)

    class _Head(torch.autograd.Function):
        @staticmethod
        def forward(ctx, x, *vs):
            ctx.x_shape = x.shape
            ctx.need_dx = x.requires_grad
            return op.forward(x, train)

        @staticmethod
        def backward(ctx, dlogits):
            dX, grads = op.backward(dlogits, want_dX=ctx.need_dx)
            out = [dX.reshape(ctx.x_shape) if dX is not None else None]
            for n, v in zip(names, variables):
                out.append(grads[n].reshape(v.shape))
            return tuple(out)

    return _Head.apply(X.contiguous(), *variables)


_HEAD_STATE = {}   # id(params) -> (params, {shape key -> AttnPoolOp}); holds the dict so the id stays unique
_HEAD_STATE_MAX = 8


def invalidate_head_cache(params=None):
    """Drop the device copies `attn_pool_head` keeps: of one `params` object (after its arrays were reloaded or
    updated in place) or of all of them."""
    if params is None:
        _HEAD_STATE.clear()
    else:
        _HEAD_STATE.pop(id(params), None)


def attn_pool_head(X, params, cfg_net_flags, num_classes, is_training=False, frames_per_video=1,
                   num_pose_keypoints=16, end_point_names=(), tuning=None):
    """Functional form of the head: conv5 features in, `(logits, end_points)` out.

    X: float32 CUDA tensor `[B*F, H, W, C]` or `[B*F, HW, C]`; `params`: TF variable name -> array/tensor
    (nets_factory.py:260-269, :300-303, :151-159); `cfg_net_flags`: `cfg.NET`-like object or dict with the
    `USE_POSE_PRELOGITS_BASED_ATTENTION_*` flags (src/config.py:182-209).  The device-resident copy of the
    variables and the workspace are cached per (`params` object, flags, tuning) — at most 8 `params` objects; call
    `invalidate_head_cache(params)` after updating the arrays of a `params` dict in place.
    """
    if params is None:
        raise ValueError("params (TF variable name -> array) is required: the head has trained variables")
    flags, rank = ops.flags_from_cfg_net(cfg_net_flags)
    if is_training:
        raise NotImplementedError("use get_network_fn(..., is_training=True): the training graph needs the dropout "
                                  "seed and autograd plumbing it sets up")
    entry = _HEAD_STATE.get(id(params))
    if entry is None or entry[0] is not params:
        while len(_HEAD_STATE) >= _HEAD_STATE_MAX:       # bounded: the oldest entry goes first
            _HEAD_STATE.pop(next(iter(_HEAD_STATE)))
        entry = (params, {})
        _HEAD_STATE[id(params)] = entry
    tkey = bytes(tuning) if tuning is not None else None   # ApaTuning is a ctypes struct: its bytes are its value
    state = entry[1].setdefault((flags, rank, int(num_classes), int(num_pose_keypoints), tkey), {})
    return _run_head(state, X, params, flags, rank, num_classes, num_pose_keypoints, end_point_names, tuning,
                     frames_per_video=frames_per_video)
