"""Training losses behind the head: CPU — the NumPy restatement (oracle/losses.py) against
float64 PyTorch autograd of the expressions of loss.py:29-79; GPU — apa_softmax_xent_loss / apa_pose_l2_loss through
the C ABI against the restatement (1e-5 relative: fp32 exp/log vs float64), bit-reproducible."""
import numpy as np
import pytest


def test_oracle_losses_match_autograd():
    import torch
    from oracle import losses
    rng = np.random.RandomState(0)
    x = rng.randn(7, 13) * 3
    lab = rng.randint(0, 13, size=7)
    xt = torch.tensor(x, dtype=torch.float64, requires_grad=True)
    ref = 0.3 * torch.nn.functional.cross_entropy(xt, torch.tensor(lab), reduction="sum") / 7
    ref.backward()
    loss, grad = losses.softmax_xent(x, lab, 0.3)
    assert abs(loss - float(ref.detach())) <= 1e-12 and np.abs(grad - xt.grad.numpy()).max() <= 1e-12
    p, h = rng.randn(3, 5, 5, 4), np.maximum(rng.randn(3, 5, 5, 4), 0)
    valid = rng.rand(3, 4) > 0.4
    pt = torch.tensor(p, dtype=torch.float64, requires_grad=True)
    # loss.py:55-64,69-70 written out: per channel, 0.5 * sum / (B*HW), where(valid), mean over the batch, sum, weight
    lv = 0.5 * ((torch.tensor(h) - pt) ** 2).sum(dim=(1, 2)) / (3 * 25)    # reduce_sum(ones([B, H, W])) = 75
    ref = 2.5 * torch.where(torch.tensor(valid), lv, torch.zeros_like(lv)).mean(dim=0).sum()
    ref.backward()
    loss, grad = losses.pose_l2(p, h, valid, 2.5)
    assert abs(loss - float(ref.detach())) <= 1e-12 and np.abs(grad - pt.grad.numpy()).max() <= 1e-12


@pytest.mark.gpu
@pytest.mark.parametrize("rows,K", [(1, 393), (32, 393), (33, 600), (5, 7)])
def test_gpu_softmax_xent_matches_oracle(rows, K):
    import torch
    from attentionalpoolingaction_b200.custom_ops import attn_pool_factory as ops
    from oracle import losses
    rng = np.random.RandomState(rows + K)
    x = (rng.randn(rows, K) * 4).astype(np.float32)
    lab = rng.randint(0, K, size=rows).astype(np.int32)
    loss, grad = ops.softmax_xent_loss(torch.from_numpy(x).cuda(), torch.from_numpy(lab).cuda(), weight=0.7)
    ref_loss, ref_grad = losses.softmax_xent(x, lab, 0.7)
    assert abs(float(loss) - ref_loss) <= 1e-5 * max(1.0, abs(ref_loss))
    assert np.abs(grad.cpu().numpy() - ref_grad).max() <= 1e-5 * np.abs(ref_grad).max()
    loss2, grad2 = ops.softmax_xent_loss(torch.from_numpy(x).cuda(), torch.from_numpy(lab).cuda(), weight=0.7)
    assert torch.equal(loss, loss2) and torch.equal(grad, grad2)
    loss3, none = ops.softmax_xent_loss(torch.from_numpy(x).cuda(), torch.from_numpy(lab).cuda(), weight=0.7, want_grad=False)
    assert none is None and torch.equal(loss, loss3)


@pytest.mark.gpu
@pytest.mark.parametrize("B,H,Kp", [(1, 14, 16), (8, 14, 16), (3, 15, 5)])
def test_gpu_pose_l2_matches_oracle(B, H, Kp):
    import torch
    from attentionalpoolingaction_b200.custom_ops import attn_pool_factory as ops
    from oracle import losses
    rng = np.random.RandomState(B * 31 + Kp)
    p = rng.randn(B, H, H, Kp).astype(np.float32)
    h = np.maximum(rng.randn(B, H, H, Kp), 0).astype(np.float32)
    valid = rng.rand(B, Kp) > 0.3
    loss, grad = ops.pose_l2_loss(torch.from_numpy(p).cuda(), torch.from_numpy(h).cuda(), torch.from_numpy(valid).cuda(), 1.5)
    ref_loss, ref_grad = losses.pose_l2(p, h, valid, 1.5)
    assert abs(float(loss) - ref_loss) <= 1e-5 * max(1e-6, abs(ref_loss))
    assert grad.shape == (B, H, H, Kp)
    assert np.abs(grad.cpu().numpy() - ref_grad).max() <= 1e-5 * max(np.abs(ref_grad).max(), 1e-12)


@pytest.mark.gpu
def test_train_step_chain_head_loss_backward():
    """logits -> apa_softmax_xent_loss -> dlogits -> apa_attn_pool_bwd equals float64 autograd of the whole chain."""
    import torch
    from attentionalpoolingaction_b200.custom_ops import attn_pool_factory as ops
    from oracle import attn_pool_oracle as orc, attn_pool_backward as bwd, losses
    fl = dict(orc.DEFAULT_FLAGS, SINGLE_LAYER_ATT=True)
    K, HW, C, B = 393, 196, 2048, 4
    params = orc.make_params(K, C, flags=fl, seed=6)
    bits, rnk = ops.flags_from_cfg_net(fl)
    rng = np.random.RandomState(2)
    X = np.maximum(rng.randn(B, HW, C), 0).astype(np.float32)
    lab = rng.randint(0, K, size=B).astype(np.int32)
    op = ops.AttnPoolTrainOp({k: torch.from_numpy(v) for k, v in params.items()}, K, HW, rank=rnk, flags=bits)
    train = ops.make_train(0.2, 42)
    logits = op.forward(torch.from_numpy(X).cuda(), train)
    loss, dl = ops.softmax_xent_loss(logits, torch.from_numpy(lab).cuda())
    dX, grads = op.backward(dl)
    mask = bwd.dropout_mask(42, X.shape, 0.2)
    ref_logits = bwd.forward_train_torch(torch.from_numpy(X.astype(np.float64)),
                                         {k: torch.from_numpy(v.astype(np.float64)) for k, v in params.items()},
                                         fl, K, 1, torch.from_numpy(mask.astype(np.float64))).numpy()
    ref_loss, ref_dl = losses.softmax_xent(ref_logits, lab)
    ref_dX, ref_g = bwd.backward_autograd(X, params, fl, K, ref_dl, dropout_mask=mask)
    assert abs(float(loss) - ref_loss) <= 1e-5 * abs(ref_loss)
    assert np.abs(dX.cpu().numpy() - ref_dX).max() <= 1e-4 * np.abs(ref_dX).max()
    for n, g in ref_g.items():
        assert np.abs(grads[n].cpu().numpy().reshape(g.shape) - g).max() <= 1e-4 * max(np.abs(g).max(), 1e-6), n


def test_pose_loss_restatement_equals_the_channel_loop_of_the_reference():
    """loss.py:31-69 walked channel by channel the way the reference writes it (unstack on the last axis, zip,
    0.5 * reduce_sum(square(.)) / reduce_sum(ones), tf.where(valid, ., 0), reduce_mean, sum of the channel losses),
    against the vectorised restatement in oracle/losses.py."""
    from oracle import losses
    rng = np.random.RandomState(4)
    B, H, W, Kp = 5, 6, 7, 9
    logits_pose = rng.randn(B, H, W, Kp)
    labels_pose = np.maximum(rng.randn(B, H, W, Kp), 0)
    labels_pose_valid = rng.rand(B, Kp) > 0.5
    loss_elements = []
    for c in range(Kp):                                              # :35 zip(channels_valid, channels_logits, channels_labels)
        v, lbl, lgt = labels_pose_valid[:, c], logits_pose[..., c], labels_pose[..., c]   # (names bound as the reference binds them)
        mask = np.ones_like(lgt)                                     # :55
        loss_val = 0.5 * np.square(lbl - lgt).sum(axis=(1, 2)) / mask.sum()   # :56-57 -- NB: mask.sum() counts the BATCH too
        loss_elements.append(np.mean(np.where(v, loss_val, 0.0)))    # :60-64
    tot = 1.7 * float(np.sum(loss_elements))                         # :69-70
    ours, _ = losses.pose_l2(logits_pose, labels_pose, labels_pose_valid, 1.7)
    # tf.reduce_sum(mask) at :57 sums ones of shape [B, H, W]: the divisor is B*H*W, not H*W -- the restatement (and the
    # kernel) must divide by it as well
    assert abs(ours - tot) <= 1e-12, (ours, tot)
