"""CUDA training path (csrc/apa_train.cu: apa_attn_pool_fwd_train / apa_attn_pool_bwd / apa_dropout_mask) against the
float64 autograd oracle (oracle/attn_pool_backward.py: backward_autograd over the op-for-op port of
nets_factory.py:242-374 with the dropout of :296).  Tolerance: 1e-4 relative to the largest gradient entry
(fp32 kernels vs a float64 truth), the same bar as the forward parity tests."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _rel(a, b, floor=1e-30):
    """max |a - b| relative to max |b| (or to `floor`, for gradients that are zero by construction: the sum of a
    softmax's input gradients, for one)."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), floor))


CASES = [
    # name, B(images), F, HW, C, K, flags, keep
    ("config2 flags, keep 0.2", 8, 1, 196, 2048, 393, dict(SINGLE_LAYER_ATT=True), 0.2),
    ("config2 flags, no dropout", 4, 1, 196, 2048, 393, dict(SINGLE_LAYER_ATT=True), 1.0),
    ("plain attention (RELU_ATT off)", 4, 1, 196, 2048, 393, dict(SINGLE_LAYER_ATT=True, RELU_ATT=False), 0.2),
    ("softmax attention", 4, 1, 196, 2048, 393, dict(SINGLE_LAYER_ATT=True, SOFTMAX_ATT=True), 0.2),
    ("rank 3 chain", 4, 1, 196, 2048, 393, dict(SINGLE_LAYER_ATT=True, RANK=3), 0.2),
    ("rank 2 chain, plain", 3, 1, 49, 256, 51, dict(SINGLE_LAYER_ATT=True, RANK=2, RELU_ATT=False), 0.5),
    ("videos of 5 frames, K=600", 10, 5, 196, 2048, 600, dict(SINGLE_LAYER_ATT=True), 0.2),
    ("ragged: 3 images, HW=225, C=1024", 3, 1, 225, 1024, 393, dict(SINGLE_LAYER_ATT=True), 0.2),
    ("one image, 9 videos tile edge", 9, 1, 30, 512, 33, dict(SINGLE_LAYER_ATT=True), 0.7),
]


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_forward_train_and_backward_match_autograd(case):
    import torch
    from attentionalpoolingaction_b200.custom_ops import attn_pool_factory as ops
    from oracle import attn_pool_oracle as orc, attn_pool_backward as bwd
    name, B, F, HW, C, K, fl, keep = case
    fl = dict(orc.DEFAULT_FLAGS, **fl)
    params = orc.make_params(K, C, flags=fl, seed=11)
    bits, rnk = ops.flags_from_cfg_net(fl)
    dev = torch.device("cuda:0")
    rng = np.random.RandomState(5)
    X = np.maximum(rng.randn(B, HW, C), 0).astype(np.float32)
    dl = rng.randn(B // F, K).astype(np.float32)
    op = ops.AttnPoolTrainOp({k: torch.from_numpy(v) for k, v in params.items()}, K, HW, channels=C, frames=F,
                             rank=rnk, flags=bits, device=dev)
    train = ops.make_train(keep, dropout_seed=1234)
    Xd = torch.from_numpy(X).to(dev)
    # the mask the kernels apply == its NumPy restatement, bit for bit
    mask = bwd.dropout_mask(1234, X.shape, keep)
    assert np.array_equal(ops.dropout_mask(train, X.shape, dev).cpu().numpy(), mask)
    logits = op.forward(Xd, train)
    dX, grads = op.backward(torch.from_numpy(dl).to(dev))
    torch.cuda.synchronize()
    ref_logits = bwd.forward_train_torch(torch.from_numpy(X.astype(np.float64)),
                                         {k: torch.from_numpy(v.astype(np.float64)) for k, v in params.items()},
                                         fl, K, F, torch.from_numpy(mask.astype(np.float64))).numpy()
    assert _rel(logits.cpu().numpy(), ref_logits) <= 1e-4
    ref_dX, ref_g = bwd.backward_autograd(X, params, fl, K, dl, frames_per_video=F, dropout_mask=mask)
    assert _rel(dX.cpu().numpy(), ref_dX) <= 1e-4, "dX"
    assert set(grads) == set(ref_g)
    gscale = max(float(np.abs(v).max()) for v in ref_g.values())
    for n, g in grads.items():
        assert tuple(g.shape) == tuple(np.asarray(params[n]).reshape(g.shape).shape)
        assert _rel(g.cpu().numpy().reshape(ref_g[n].shape), ref_g[n], floor=1e-2 * gscale) <= 1e-4, n
    # bit-reproducible: all reductions have a fixed order
    logits2 = op.forward(Xd, train).clone()
    dX2, grads2 = op.backward(torch.from_numpy(dl).to(dev))
    assert torch.equal(logits2, logits) and torch.equal(dX2, dX)
    assert all(torch.equal(grads2[n], grads[n]) for n in grads)


def test_keep_one_equals_eval_forward():
    """train == NULL (or keep 1) is the eval graph: same logits as apa_attn_pool_fwd up to fp32 rounding."""
    import torch
    from attentionalpoolingaction_b200.custom_ops import attn_pool_factory as ops
    from oracle import attn_pool_oracle as orc
    fl = dict(orc.DEFAULT_FLAGS, SINGLE_LAYER_ATT=True)
    params = orc.make_params(393, 2048, flags=fl, seed=2)
    bits, rnk = ops.flags_from_cfg_net(fl)
    P = {k: torch.from_numpy(v) for k, v in params.items()}
    X = torch.relu(torch.randn((32, 196, 2048), device="cuda", generator=torch.Generator(device="cuda").manual_seed(1)))
    a = ops.AttnPoolTrainOp(P, 393, 196, rank=rnk, flags=bits).forward(X, None)
    b, _ = ops.AttnPoolOp(P, 393, 196, rank=rnk, flags=bits)(X)
    assert orc.rel_err(a.cpu().numpy(), b.cpu().numpy()) <= 1e-5


def test_frozen_backbone_and_unsupported():
    import torch
    from attentionalpoolingaction_b200.custom_ops import attn_pool_factory as ops
    from oracle import attn_pool_oracle as orc
    fl = dict(orc.DEFAULT_FLAGS, SINGLE_LAYER_ATT=True)
    params = orc.make_params(51, 256, flags=fl, seed=2)
    bits, rnk = ops.flags_from_cfg_net(fl)
    P = {k: torch.from_numpy(v) for k, v in params.items()}
    X = torch.relu(torch.randn((2, 49, 256), device="cuda"))
    op = ops.AttnPoolTrainOp(P, 51, 49, channels=256, rank=rnk, flags=bits)
    train = ops.make_train(0.2, 7)
    op.forward(X, train)
    dl = torch.randn((2, 51), device="cuda")
    dX, g1 = op.backward(dl, want_dX=False)       # dX == NULL: the variables' gradients only
    assert dX is None
    dX2, g2 = op.backward(dl)
    assert all(torch.equal(g1[n], g2[n]) for n in g1) and dX2 is not None
    # PER_CLASS and the 768-channel pose branch are not served by the training path
    for extra in (dict(PER_CLASS=True), dict(SINGLE_LAYER_ATT=False)):
        f2 = dict(orc.DEFAULT_FLAGS, SINGLE_LAYER_ATT=True); f2.update(extra)
        p2 = orc.make_params(51, 256, flags=f2, seed=2)
        b2, r2 = ops.flags_from_cfg_net(f2)
        o2 = ops.AttnPoolTrainOp({k: torch.from_numpy(v) for k, v in p2.items()}, 51, 49, channels=256, rank=r2, flags=b2)
        with pytest.raises(ops.ApaError) as e:
            o2.forward(X, train)
        assert e.value.status == ops.APA_ERR_UNSUPPORTED
    with pytest.raises(ops.ApaError):
        op.forward(X, ops.make_train(0.0, 1))     # keep must be in (0, 1]


def test_network_fn_in_training_mode_backpropagates():
    """get_network_fn(is_training=True) (train.py:399-420): `loss.backward()` reaches X and the head's variables,
    and equals the float64 autograd of the reference graph with the same mask."""
    import torch
    from attentionalpoolingaction_b200 import config, nets_factory
    from oracle import attn_pool_oracle as orc, attn_pool_backward as bwd
    cfg = config.default_cfg()
    cfg.NET.USE_POSE_PRELOGITS_BASED_ATTENTION = True
    cfg.NET.USE_POSE_PRELOGITS_BASED_ATTENTION_SINGLE_LAYER_ATT = True
    fl = dict(orc.DEFAULT_FLAGS, SINGLE_LAYER_ATT=True)
    K, HW, C, B = 393, 196, 2048, 4
    np_params = orc.make_params(K, C, flags=fl, seed=4)
    params = {k: torch.from_numpy(v).cuda().requires_grad_(True) for k, v in np_params.items()}
    fn = nets_factory.get_network_fn("resnet_v1_101", K, 16, cfg, is_training=True, params=params)
    X = torch.relu(torch.randn((B, 14, 14, C), device="cuda")).requires_grad_(True)
    labels = torch.randint(0, K, (B,), device="cuda")
    logits, eps = fn(X, dropout_seed=99)
    loss = torch.nn.functional.cross_entropy(logits, labels)       # the softmax loss of train.py:404-411
    loss.backward()
    assert eps["Logits"] is logits and X.grad is not None and X.grad.shape == X.shape
    mask = bwd.dropout_mask(99, (B, HW, C), 0.2)                    # cfg.NET.DROPOUT = -1 -> keep 0.2
    Xn = X.detach().cpu().numpy().reshape(B, HW, C)
    lt = logits.detach().double().cpu().requires_grad_(True)
    torch.nn.functional.cross_entropy(lt, labels.cpu()).backward()
    ref_dX, ref_g = bwd.backward_autograd(Xn, np_params, fl, K, lt.grad.numpy(), dropout_mask=mask)
    assert _rel(X.grad.cpu().numpy().reshape(B, HW, C), ref_dX) <= 1e-4
    for n, g in ref_g.items():
        assert _rel(params[n].grad.cpu().numpy().reshape(g.shape), g) <= 1e-4, n
