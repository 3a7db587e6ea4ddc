"""Backward of the attentional-pooling head — CPU oracle.  TEST INFRASTRUCTURE ONLY.

SURVEY.md section 8(f) rank 1: `train.py:399-420,509-513` differentiates the graph that
`nets_factory.py:242-352` builds; TF does it by autodiff, so there is no reference source to follow
line by line.  Two independent statements are kept here so that a CUDA backward has something to
be held against:

* `backward_autograd` — float64 PyTorch autograd through the op-for-op forward port
  (`attn_pool_oracle.attn_pool_head_torch_cpu`, which mirrors `nets_factory.py:242-374`), with the
  training-time dropout of `nets_factory.py:296` (`slim.dropout(last_conv)` feeds the top-down convs
  only; the attention branch sees the un-dropped features) as an explicit, pre-scaled mask.
* `backward_reassociated` — NumPy closed form for the single-map case (M == 1) written in the
  algebra of the fast path: the only per-position quantities are the attention
  value and two dot products with per-image vectors, so a kernel needs two more passes over X
  (one to form dA = X.u + t and the attention-branch deltas, one to write dX) and never the
  [HW, K] top-down maps.

PARITY PIN STATUS: unpinned against the reference (no TF here, and the reference holds no gradient
fixtures); the two statements are checked against each other and against finite differences in
`tests/test_oracle_backward.py`.  No product code uses this file; the CUDA backward
(`apa_attn_pool_bwd`, csrc/apa_train.cu) is held against `backward_autograd` in `tests/test_gpu_backward.py`.

`dropout_mask` restates the library's counter-based mask (include/apa.h: apa_dropout_mask) in NumPy so that the
oracle can be given exactly the mask the kernels apply; TF's own mask comes from its Philox stream and cannot be
reproduced bit for bit by anyone, only its distribution (Bernoulli(keep), kept values scaled by 1/keep).
"""
from __future__ import annotations

import numpy as np

from . import attn_pool_oracle as orc


_M64 = (1 << 64) - 1


def _splitmix64(z):
    """Vectorised splitmix64 on uint64 arrays (wrap-around arithmetic)."""
    with np.errstate(over="ignore"):
        z = (z + np.uint64(0x9E3779B97F4A7C15)).astype(np.uint64)
        z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)).astype(np.uint64)
        z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)).astype(np.uint64)
        return z ^ (z >> np.uint64(31))


def _fmix32(h):
    """murmur3's 32-bit finaliser on uint32 arrays."""
    with np.errstate(over="ignore"):
        h = h ^ (h >> np.uint32(16))
        h = (h * np.uint32(0x85EBCA6B)).astype(np.uint32)
        h = h ^ (h >> np.uint32(13))
        h = (h * np.uint32(0xC2B2AE35)).astype(np.uint32)
        return h ^ (h >> np.uint32(16))


def dropout_mask(seed, shape, keep):
    """include/apa.h apa_dropout_mask restated: key = splitmix64(seed); for group g = i // 4,
    a = fmix32(g ^ low32(key)), b = fmix32(g ^ high32(key)); the group's four elements take the 16-bit fields
    a & 0xFFFF, a >> 16, b & 0xFFFF, b >> 16 and are kept (value 1/keep) iff the field < round(keep * 65536)."""
    n = int(np.prod(shape))
    if n % 4:
        raise ValueError("element count must be a multiple of 4")
    if keep >= 1.0:
        return np.ones(shape, dtype=np.float32)
    key = int(_splitmix64(np.array([int(seed) & _M64], dtype=np.uint64))[0])
    g = np.arange(n // 4, dtype=np.uint32)
    a = _fmix32(g ^ np.uint32(key & 0xFFFFFFFF))
    b = _fmix32(g ^ np.uint32(key >> 32))
    fields = np.stack([a & np.uint32(0xFFFF), a >> np.uint32(16), b & np.uint32(0xFFFF), b >> np.uint32(16)],
                      axis=1).reshape(-1)
    thr = np.uint32(int(float(np.float32(keep)) * 65536.0 + 0.5))
    scale = np.float32(1.0) / np.float32(keep)
    return np.where(fields < thr, scale, np.float32(0)).astype(np.float32).reshape(shape)


def forward_train_torch(X, params, flags, num_classes, frames_per_video=1, dropout_mask=None):
    """Training-mode forward in torch (any dtype): as `attn_pool_head_torch_cpu`, but the top-down convs
    read `X * dropout_mask` (mask already divided by keep_prob; nets_factory.py:296, config.py:231)."""
    import torch
    f = dict(orc.DEFAULT_FLAGS, **(flags or {}))
    R = int(f["RANK"])
    last_conv = X
    if not f["SINGLE_LAYER_ATT"]:
        net = torch.relu(torch.matmul(last_conv, params["PoseLogits/ExtraConv2d_1x1/weights"]) +
                         params["PoseLogits/ExtraConv2d_1x1/biases"])              # :151-157
    else:
        net = last_conv                                                            # :247-248
    all_att = []
    for r in range(R):                                                             # :259-270 (chain, Q1)
        s = "PosePrelogitsBasedAttention/" + orc.attn_scope_name(r)
        net = torch.matmul(net, params[s + "/weights"]) + params[s + "/biases"]
        all_att.append(net)
    attention = torch.stack(all_att, dim=-1) if R > 1 else all_att[0]
    if f["SOFTMAX_ATT"]:                                                           # :276-284
        if attention.dim() != 3:
            raise orc.ReferenceGraphError("SOFTMAX_ATT with RANK>1 (nets_factory.py:278)")
        attention = torch.softmax(attention.transpose(1, 2), dim=-1).transpose(1, 2)
    if f["RELU_ATT"]:                                                              # :285-286
        attention = torch.relu(attention)
    dropped = last_conv if dropout_mask is None else last_conv * dropout_mask      # :296
    all_td = []
    for r in range(R):                                                             # :299-304
        t = "PosePrelogitsBasedAttention/" + orc.topdown_scope_name(r)
        all_td.append(torch.matmul(dropped, params[t + "/weights"]) + params[t + "/biases"])
    topdown = torch.stack(all_td, dim=-1) if R > 1 else all_td[0]
    logits = (attention * topdown).mean(dim=1)                                     # :322-325
    if R > 1:
        logits = logits.sum(dim=-1)                                                # :326-328
    if frames_per_video > 1:                                                       # :354-374
        logits = logits.reshape(-1, frames_per_video, logits.shape[-1]).mean(dim=1)
    return logits


def _used_names(flags):
    f = dict(orc.DEFAULT_FLAGS, **(flags or {}))
    names = []
    if not f["SINGLE_LAYER_ATT"]:
        names += ["PoseLogits/ExtraConv2d_1x1/weights", "PoseLogits/ExtraConv2d_1x1/biases"]
    for r in range(int(f["RANK"])):
        a = "PosePrelogitsBasedAttention/" + orc.attn_scope_name(r)
        t = "PosePrelogitsBasedAttention/" + orc.topdown_scope_name(r)
        names += [a + "/weights", a + "/biases", t + "/weights", t + "/biases"]
    return names


def backward_autograd(X, params, flags, num_classes, dlogits, frames_per_video=1, dropout_mask=None):
    """d(sum(logits * dlogits)) / d(X, every variable on the path), float64.  Returns (dX, {name: grad})."""
    import torch
    names = _used_names(flags)
    Xt = torch.tensor(np.asarray(X, dtype=np.float64), requires_grad=True)
    P = {k: torch.tensor(np.asarray(v, dtype=np.float64), requires_grad=(k in names)) for k, v in params.items()}
    mask = None if dropout_mask is None else torch.tensor(np.asarray(dropout_mask, dtype=np.float64))
    logits = forward_train_torch(Xt, P, flags, num_classes, frames_per_video, mask)
    loss = (logits * torch.tensor(np.asarray(dlogits, dtype=np.float64))).sum()
    grads = torch.autograd.grad(loss, [Xt] + [P[n] for n in names])
    return grads[0].numpy(), {n: g.numpy() for n, g in zip(names, grads[1:])}


def backward_reassociated(X, params, flags, num_classes, dlogits, frames_per_video=1, dropout_mask=None):
    """Same gradients for the single-map head (PER_CLASS off), in the fast path's algebra.  float64 NumPy.

    Forward, Xd = X * mask:
        a_0 = S.w_0 + c_0 (S = X or P = relu(X Wp + bp));  a_r = w_r a_{r-1} + c_r;  A_r = act(a_r)
        g_r[b] = sum_hw A_r Xd          s_r[b] = sum_hw A_r
        logits[v] = 1/(HW F) sum_{frames of v} sum_r (g_r.B_r + d_r s_r)
    """
    f = dict(orc.DEFAULT_FLAGS, **(flags or {}))
    if f["PER_CLASS"]:
        raise NotImplementedError("closed form is for the single-map head")
    R, F = int(f["RANK"]), int(frames_per_video)
    X = np.asarray(X, dtype=np.float64)
    P = {k: np.asarray(v, dtype=np.float64) for k, v in params.items()}
    Bn, HW, C = X.shape
    Xd = X if dropout_mask is None else X * np.asarray(dropout_mask, dtype=np.float64)
    G = np.repeat(np.asarray(dlogits, dtype=np.float64), F, axis=0) / (HW * F)        # per frame, scale folded in

    # ---- forward quantities that the backward re-uses
    single = bool(f["SINGLE_LAYER_ATT"])
    if single:
        S = X
    else:
        Wp, bp = P["PoseLogits/ExtraConv2d_1x1/weights"], P["PoseLogits/ExtraConv2d_1x1/biases"]
        pre = X @ Wp + bp
        S = np.maximum(pre, 0)
    aw = [P["PosePrelogitsBasedAttention/%s/weights" % orc.attn_scope_name(r)] for r in range(R)]
    ab = [P["PosePrelogitsBasedAttention/%s/biases" % orc.attn_scope_name(r)] for r in range(R)]
    tw = [P["PosePrelogitsBasedAttention/%s/weights" % orc.topdown_scope_name(r)] for r in range(R)]
    tb = [P["PosePrelogitsBasedAttention/%s/biases" % orc.topdown_scope_name(r)] for r in range(R)]
    a = [S @ aw[0][:, 0] + ab[0][0]]                                                   # [B, HW]
    for r in range(1, R):
        a.append(aw[r][0, 0] * a[r - 1] + ab[r][0])
    if f["SOFTMAX_ATT"]:
        if R > 1:
            raise orc.ReferenceGraphError("SOFTMAX_ATT with RANK>1 (nets_factory.py:278)")
        e = np.exp(a[0] - a[0].max(axis=1, keepdims=True))
        A = [e / e.sum(axis=1, keepdims=True)]
    else:
        A = list(a)
    if f["RELU_ATT"]:
        A = [np.maximum(v, 0) for v in A]

    grads = {}
    dX = np.zeros_like(X)
    dA = []
    for r in range(R):
        g = np.einsum("bh,bhc->bc", A[r], Xd)                                          # [B, C]
        s = A[r].sum(axis=1)                                                           # [B]
        t = "PosePrelogitsBasedAttention/" + orc.topdown_scope_name(r)
        grads[t + "/weights"] = g.T @ G                                                # [C, K]
        grads[t + "/biases"] = s @ G                                                   # [K]
        u = G @ tw[r].T                                                                # [B, C]  = B_r G[b]
        tt = G @ tb[r]                                                                 # [B]     = d_r . G[b]
        dA.append(np.einsum("bhc,bc->bh", Xd, u) + tt[:, None])                        # pass 1 over X
        dXd = A[r][:, :, None] * u[:, None, :]                                         # pass 2 over X
        dX += dXd if dropout_mask is None else dXd * np.asarray(dropout_mask, dtype=np.float64)

    # ---- through relu / softmax, then down the rank chain
    da = []
    for r in range(R):
        d = dA[r]
        if f["RELU_ATT"]:
            src = A[r] if f["SOFTMAX_ATT"] else a[r]   # relu after softmax acts on positive values: identity
            d = d * (src > 0)
        da.append(d)
    if f["SOFTMAX_ATT"]:
        soft = A[0]                                    # relu(softmax) == softmax
        da[0] = soft * (da[0] - (soft * da[0]).sum(axis=1, keepdims=True))
    delta = [None] * R
    for r in range(R - 1, -1, -1):
        delta[r] = da[r] + (aw[r + 1][0, 0] * delta[r + 1] if r + 1 < R else 0.0)
    for r in range(1, R):
        s_ = "PosePrelogitsBasedAttention/" + orc.attn_scope_name(r)
        grads[s_ + "/weights"] = np.array([[np.sum(delta[r] * a[r - 1])]])
        grads[s_ + "/biases"] = np.array([np.sum(delta[r])])
    s0 = "PosePrelogitsBasedAttention/" + orc.attn_scope_name(0)
    grads[s0 + "/weights"] = np.einsum("bh,bhc->c", delta[0], S)[:, None]              # [Cs, 1]
    grads[s0 + "/biases"] = np.array([np.sum(delta[0])])
    dS = delta[0][:, :, None] * aw[0][:, 0][None, None, :]
    if single:
        dX += dS
    else:
        dpre = dS * (pre > 0)
        grads["PoseLogits/ExtraConv2d_1x1/weights"] = np.einsum("bhc,bhp->cp", X, dpre)
        grads["PoseLogits/ExtraConv2d_1x1/biases"] = dpre.sum(axis=(0, 1))
        dX += dpre @ Wp.T
    return dX, grads
