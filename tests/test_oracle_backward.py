"""Backward oracle of the head: the re-associated closed form a CUDA backward
would follow vs float64 autograd through the op-for-op forward port, and autograd vs finite differences.
CPU only; no CUDA backward exists yet."""
import numpy as np
import pytest

from oracle import attn_pool_backward as bwd
from oracle import attn_pool_oracle as orc

CASES = [
    ("cfg002", dict(SINGLE_LAYER_ATT=True), 1, False),
    ("relu", dict(SINGLE_LAYER_ATT=True, RELU_ATT=True), 1, False),
    ("softmax", dict(SINGLE_LAYER_ATT=True, SOFTMAX_ATT=True), 1, False),
    ("softmax_relu", dict(SINGLE_LAYER_ATT=True, SOFTMAX_ATT=True, RELU_ATT=True), 1, False),
    ("rank3_chain", dict(SINGLE_LAYER_ATT=True, RANK=3), 1, False),
    ("rank2_relu_frames", dict(SINGLE_LAYER_ATT=True, RANK=2, RELU_ATT=True), 3, False),
    ("cfg003_pose_source_rank2", dict(RANK=2), 1, False),
    ("cfg003_softmax", dict(SOFTMAX_ATT=True), 2, False),
    ("train_dropout", dict(SINGLE_LAYER_ATT=True), 1, True),
    ("train_dropout_cfg003", dict(RANK=2, RELU_ATT=True), 2, True),
]


def _setup(flags, F, dropout, B=6, HW=10, C=24, K=7, pose=9):
    fl = dict(orc.DEFAULT_FLAGS, **flags)
    params = orc.make_params(K, C, pose, 4, flags=fl, seed=3)
    params = {k: v.astype(np.float64) * 10 for k, v in params.items()}   # O(1) attention values
    rng = np.random.RandomState(11)
    X = np.maximum(rng.randn(B, HW, C), 0)
    dlogits = rng.randn(B // F, K)
    mask = (rng.rand(B, HW, C) < 0.2) / 0.2 if dropout else None        # keep_prob 0.2 (config.py:231)
    return fl, params, X, dlogits, mask


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_reassociated_backward_equals_autograd(case):
    _, flags, F, dropout = case
    fl, params, X, dlogits, mask = _setup(flags, F, dropout)
    dX_a, g_a = bwd.backward_autograd(X, params, fl, 7, dlogits, F, mask)
    dX_r, g_r = bwd.backward_reassociated(X, params, fl, 7, dlogits, F, mask)
    # absolute floor: the softmax bias gradient is exactly 0 (shift invariance), the closed form leaves 1e-16
    scale = max([np.abs(dX_a).max()] + [np.abs(v).max() for v in g_a.values()])
    np.testing.assert_allclose(dX_r, dX_a, rtol=1e-9, atol=1e-12 * scale)
    assert set(g_r) == set(g_a)
    for name in g_a:
        assert g_r[name].shape == g_a[name].shape, name
        np.testing.assert_allclose(g_r[name], g_a[name], rtol=1e-9, atol=1e-12 * scale, err_msg=name)


def test_autograd_matches_finite_differences():
    """The truth itself: central differences of sum(logits * dlogits) along random directions."""
    import torch
    fl, params, X, dlogits, mask = _setup(dict(RANK=2, RELU_ATT=True), 2, True)
    X = X + 0.05          # keep relu arguments away from the kink
    dX, grads = bwd.backward_autograd(X, params, fl, 7, dlogits, 2, mask)

    def loss(Xv, Pv):
        Pt = {k: torch.tensor(v) for k, v in Pv.items()}
        out = bwd.forward_train_torch(torch.tensor(Xv), Pt, fl, 7, 2, torch.tensor(mask))
        return float((out * torch.tensor(dlogits)).sum())

    rng = np.random.RandomState(5)
    eps = 1e-6
    d = rng.randn(*X.shape)
    fd = (loss(X + eps * d, params) - loss(X - eps * d, params)) / (2 * eps)
    assert abs(fd - np.sum(dX * d)) <= 1e-6 * max(1.0, abs(fd))
    for name, g in grads.items():
        d = rng.randn(*g.shape)
        hi = dict(params); lo = dict(params)
        hi[name] = params[name] + eps * d
        lo[name] = params[name] - eps * d
        fd = (loss(X, hi) - loss(X, lo)) / (2 * eps)
        assert abs(fd - np.sum(g * d)) <= 1e-6 * max(1.0, abs(fd)), name


def test_forward_train_without_mask_is_the_eval_forward():
    import torch
    fl, params, X, _, _ = _setup(dict(RANK=2), 2, False)
    Pt = {k: torch.tensor(v) for k, v in params.items()}
    got = bwd.forward_train_torch(torch.tensor(X), Pt, fl, 7, 2).numpy()
    ref, _ = orc.attn_pool_head(X, params, fl, 7, frames_per_video=2, want_pose=False)
    np.testing.assert_allclose(got, ref, rtol=1e-12, atol=1e-13)


def test_dropout_mask_restatement_is_a_fair_bernoulli():
    """oracle.dropout_mask (the NumPy restatement of include/apa.h apa_dropout_mask): keep rate, scale, independence
    between the four fields of a group, between neighbouring groups and between seeds."""
    from oracle import attn_pool_backward as bwd
    m = bwd.dropout_mask(77, (8, 196, 2048), 0.2)
    k = m > 0
    assert set(np.unique(m)) == {np.float32(0), np.float32(5)}
    assert abs(k.mean() - 0.2) < 2e-3 and abs(m.mean() - 1.0) < 1e-2
    g = k.reshape(-1, 4)
    for i in range(4):
        for j in range(i + 1, 4):
            assert abs((g[:, i] & g[:, j]).mean() - 0.04) < 2e-3
    assert abs((k[:, :, :-4] & k[:, :, 4:]).mean() - 0.04) < 2e-3          # neighbouring groups
    assert abs((k & (bwd.dropout_mask(78, (8, 196, 2048), 0.2) > 0)).mean() - 0.04) < 2e-3   # another step's seed
    assert np.array_equal(bwd.dropout_mask(77, (8, 196, 2048), 1.0), np.ones((8, 196, 2048), np.float32))
