"""The two training losses behind the head — CPU oracle.  TEST INFRASTRUCTURE ONLY.

NumPy restatement (float64) of
  loss.py:72-79   tf.losses.softmax_cross_entropy(onehot(labels_action), logits_action, weights=action_loss_wt):
                  TF 1.1's `compute_weighted_loss` with a scalar weight = sum_i(w * l_i) / number of rows (w != 0)
  loss.py:29-70   the masked L2 heat-map loss, un-sampled branch (cfg.TRAIN.LOSS_FN_POSE_SAMPLED False, config.py):
                  per channel 0.5 * sum_hw (lbl - lgt)^2 / reduce_sum(mask), mask = ones([B, H, W]) -- the divisor is
                  B*HW, the batch included (:55-57; the zip at :35 binds the names the other way round, which a squared
                  difference does not see), `tf.where(valid, loss_val, 0)` mean over the batch (:60-64), sum over
                  channels (:69), times pose_loss_wt (:70).  `tests/test_losses.py` walks the same lines channel by
                  channel.
and of their gradients with respect to the logits (TF derives them by autodiff; `tests/test_losses.py` checks these
closed forms against float64 PyTorch autograd of the same expressions).

PARITY PIN STATUS: unpinned against the reference (TensorFlow cannot run here and the reference holds no loss fixtures).
"""
from __future__ import annotations

import numpy as np


def softmax_xent(logits, labels, weight=1.0):
    """(loss, d loss / d logits)."""
    x = np.asarray(logits, dtype=np.float64)
    lab = np.asarray(labels).astype(np.int64)
    n = x.shape[0]
    m = x.max(axis=1, keepdims=True)
    lse = np.log(np.exp(x - m).sum(axis=1)) + m[:, 0]
    loss = weight * np.sum(lse - x[np.arange(n), lab]) / n
    p = np.exp(x - lse[:, None])
    p[np.arange(n), lab] -= 1.0
    return float(loss), weight / n * p


def pose_l2(pose_logits, pose_labels, valid, weight=1.0):
    """(loss, d loss / d pose_logits); pose_* [B, HW, Kp] (or [B, H, W, Kp]), valid [B, Kp]."""
    x = np.asarray(pose_logits, dtype=np.float64)
    y = np.asarray(pose_labels, dtype=np.float64)
    shape = x.shape
    x, y = x.reshape(shape[0], -1, shape[-1]), y.reshape(shape[0], -1, shape[-1])
    v = np.asarray(valid).astype(bool)
    B, HW, Kp = x.shape
    loss_val = 0.5 * np.square(y - x).sum(axis=1) / (B * HW)            # [B, Kp]   (:55-57: reduce_sum(ones([B,H,W])))
    per_channel = np.where(v, loss_val, 0.0).mean(axis=0)               # [Kp]      (:60-64)
    loss = weight * per_channel.sum()                                   # (:69-70)
    grad = weight * (x - y) * v[:, None, :] / (HW * B * B)
    return float(loss), grad.reshape(shape)
