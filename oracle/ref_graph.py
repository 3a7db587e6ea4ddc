"""Run the REFERENCE'S OWN head code on NumPy.  TEST INFRASTRUCTURE ONLY (container-only).

TensorFlow 1.1 cannot be installed here, but `models/slim/nets/nets_factory.py` is plain Python
that compiles under Python 3.  This module installs a small NumPy-backed stand-in for the ~20
TF / TF-Slim calls that file makes (`slim.conv2d` 1x1, `slim.dropout`, `slim.arg_scope`,
`tf.variable_scope`, `tf.nn.softmax`, `tf.transpose`, `tf.reshape`, `tf.stack`, `tf.split`,
`tf.reduce_mean`, `tf.reduce_sum`, `tf.squeeze`, ...), stubs the backbone so that
`end_points['resnet_v1_101/block4']` is the feature tensor we feed, imports the reference file
UNMODIFIED from /root/reference, and calls its `get_network_fn(...)(images)`.

What this pins: every structural decision of the reference graph — op order, which tensor feeds
which conv (the rank chain, quirk Q1), default slim scope names ('Conv', 'Conv_1'), stacking
axes, the softmax transpose/reshape dance (and its failure for RANK>1, Q3), mean-not-sum, biases,
frame folding and pooling.  What it cannot pin: the floating-point behaviour of TF's own kernels
(the primitives below are restated from their documented semantics, evaluated in float64).

Nothing under tests/ -m gpu, smoke() or bench.py imports this file: /root/reference does not exist
on the GPU box.  `tests/golden/make_golden.py` uses it once to write the committed fixtures.
"""
from __future__ import annotations

import contextlib
import importlib.util
import os
import sys
import types

import numpy as np

REFERENCE_ROOT = os.environ.get("APA_REFERENCE_ROOT", "/root/reference")
NETS_FACTORY = os.path.join(REFERENCE_ROOT, "models", "slim", "nets", "nets_factory.py")


# ----------------------------------------------------------------------------- tensors
class _Shape(object):
    def __init__(self, shp):
        self._s = list(int(x) for x in shp)
        self.ndims = len(self._s)

    def as_list(self):
        return list(self._s)

    def __getitem__(self, i):
        return self._s[i]

    def __len__(self):
        return len(self._s)


class T(object):
    """A 'tf.Tensor' that is just an eagerly evaluated float64 ndarray."""

    def __init__(self, a):
        self.a = np.asarray(a, dtype=np.float64)

    def get_shape(self):
        return _Shape(self.a.shape)

    def set_shape(self, shp):
        pass

    def __mul__(self, o):
        return T(self.a * _a(o))

    __rmul__ = __mul__

    def __add__(self, o):
        return T(self.a + _a(o))

    __radd__ = __add__

    def __getitem__(self, idx):
        return T(self.a[idx])


def _a(x):
    return x.a if isinstance(x, T) else np.asarray(x, dtype=np.float64)


# ----------------------------------------------------------------------------- the stand-in
class _State(object):
    def __init__(self, variables):
        self.variables = variables          # TF variable name -> ndarray ([1,1,Cin,Cout] or [Cin,Cout])
        self.scope = []                     # variable-scope stack
        self.used_names = {}                # scope path -> set of op-scope names already taken
        self.arg_scopes = []                # stack of {fn: kwargs}
        self.touched = []                   # variable names in creation order


def _build_tf(state):
    tf = types.ModuleType("tensorflow")

    @contextlib.contextmanager
    def variable_scope(name, *a, **k):
        state.scope.append(name)
        try:
            yield
        finally:
            state.scope.pop()

    @contextlib.contextmanager
    def name_scope(name, *a, **k):
        yield

    def _unique(default):
        path = "/".join(state.scope)
        taken = state.used_names.setdefault(path, set())
        name, i = default, 0
        while name in taken:              # slim: 'Conv', 'Conv_1', 'Conv_2', ...
            i += 1
            name = "%s_%d" % (default, i)
        taken.add(name)
        return name

    def _scoped_kwargs(fn, kwargs):
        merged = {}
        for layer in state.arg_scopes:
            merged.update(layer.get(fn, {}))
        merged.update(kwargs)
        return merged

    def relu(x, *a, **k):
        return T(np.maximum(_a(x), 0))

    def softmax(x, *a, **k):             # tf.nn.softmax: over the last axis
        v = _a(x)
        e = np.exp(v - v.max(axis=-1, keepdims=True))
        return T(e / e.sum(axis=-1, keepdims=True))

    def conv2d(inputs, num_outputs, kernel_size, **kwargs):
        kw = _scoped_kwargs(conv2d, kwargs)
        assert list(kernel_size) == [1, 1], "only 1x1 convs occur on this path"
        scope = kw.get("scope")
        op_name = scope if scope is not None else _unique("Conv")
        if scope is not None:
            state.used_names.setdefault("/".join(state.scope), set()).add(scope)
        prefix = "/".join(state.scope + [op_name])
        x = _a(inputs)
        w = np.asarray(state.variables[prefix + "/weights"], dtype=np.float64)
        state.touched.append(prefix + "/weights")
        if w.ndim == 4:
            assert w.shape[:2] == (1, 1)
            w = w[0, 0]
        assert w.shape == (x.shape[-1], num_outputs), (prefix, w.shape, x.shape, num_outputs)
        y = np.tensordot(x, w, axes=([x.ndim - 1], [0]))       # Conv2D 1x1, stride 1, SAME
        if kw.get("normalizer_fn", None) is None:              # slim adds biases iff no normalizer
            b = np.asarray(state.variables[prefix + "/biases"], dtype=np.float64)
            state.touched.append(prefix + "/biases")
            y = y + b                                           # BiasAdd
        else:
            raise AssertionError("batch-norm conv on the head path: not expected")
        act = kw.get("activation_fn", relu)                     # slim.conv2d default activation is relu
        out = T(y)
        return act(out) if act is not None else out

    def dropout(x, *a, **kwargs):
        kw = _scoped_kwargs(dropout, kwargs)
        assert kw.get("is_training", True) is False, "oracle runs the eval graph (dropout == identity)"
        return x

    @contextlib.contextmanager
    def arg_scope(ops_or_scope, **kwargs):
        if isinstance(ops_or_scope, (list, tuple)):
            layer = {fn: dict(kwargs) for fn in ops_or_scope}
        else:
            layer = {}                                          # the backbone's arg_scope: every head conv
        state.arg_scopes.append(layer)                          # overrides what it would set (normalizer_fn,
        try:                                                    # activation_fn, initialisers) explicitly
            yield
        finally:
            state.arg_scopes.pop()

    def transpose(x, perm):
        v = _a(x)
        if len(perm) != v.ndim:                                 # TF raises at graph build (quirk Q3)
            raise ValueError("Dimension must be %d but is %d for 'transpose'" % (v.ndim, len(perm)))
        return T(np.transpose(v, perm))

    def reshape(x, shape):
        return T(np.reshape(_a(x), [int(s) for s in shape]))

    def stack(values, axis=0):
        return T(np.stack([_a(v) for v in values], axis=axis))

    def unstack(x, axis=0):
        v = _a(x)
        return [T(np.take(v, i, axis=axis)) for i in range(v.shape[axis])]

    def split(value, num_or_size_splits, axis=0):
        return [T(p) for p in np.split(_a(value), num_or_size_splits, axis=axis)]

    def concat(values, axis):
        return T(np.concatenate([_a(v) for v in values], axis=axis))

    def reduce_mean(x, axis=None, keep_dims=False):
        ax = tuple(axis) if isinstance(axis, (list, tuple)) else axis
        return T(_a(x).mean(axis=ax, keepdims=keep_dims))

    def reduce_sum(x, axis=None, keep_dims=False):
        ax = tuple(axis) if isinstance(axis, (list, tuple)) else axis
        return T(_a(x).sum(axis=ax, keepdims=keep_dims))

    def squeeze(x, axis=None, name=None):
        return T(np.squeeze(_a(x), axis=tuple(axis) if axis is not None else None))

    def expand_dims(x, axis):
        return T(np.expand_dims(_a(x), axis))

    tf.variable_scope = variable_scope
    tf.name_scope = name_scope
    tf.transpose, tf.reshape, tf.stack, tf.unstack, tf.split, tf.concat = transpose, reshape, stack, unstack, split, concat
    tf.reduce_mean, tf.reduce_sum, tf.squeeze, tf.expand_dims = reduce_mean, reduce_sum, squeeze, expand_dims
    tf.random_normal_initializer = lambda *a, **k: ("random_normal", a, k)
    tf.zeros_initializer = lambda *a, **k: ("zeros",)
    tf.constant_initializer = lambda *a, **k: ("constant", a)
    tf.nn = types.SimpleNamespace(relu=relu, softmax=softmax)
    tf.logging = types.SimpleNamespace(info=lambda *a, **k: None)
    slim = types.SimpleNamespace(conv2d=conv2d, dropout=dropout, arg_scope=arg_scope)
    tf.contrib = types.SimpleNamespace(slim=slim)
    return tf


class _AnyModule(types.ModuleType):
    """Stands in for the backbone modules nets_factory.py imports: any attribute is a dummy callable."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return lambda *a, **k: None


def _backbone(images, num_classes, is_training=False, train_top_bn=False, **kwargs):
    """resnet_v1_101 stand-in: the 'image' IS the conv5 map (post-ReLU features)."""
    return T(np.zeros((images.get_shape().as_list()[0], num_classes))), {"resnet_v1_101/block4": images}


def make_cfg(flags):
    """cfg.NET as the head reads it (src/config.py:160-231 names; everything the head branches on)."""
    P = "USE_POSE_PRELOGITS_BASED_ATTENTION"
    net = types.SimpleNamespace()
    net.DROPOUT = -1.0
    net.TRAIN_TOP_BN = False
    net.LAST_CONV_MAP_FOR_POSE = types.SimpleNamespace(resnet_v1_101="resnet_v1_101/block4")
    net.USE_POSE_ATTENTION_LOGITS = False
    net.USE_POSE_LOGITS_DIRECTLY = False
    net.USE_POSE_LOGITS_DIRECTLY_v2 = False
    net.USE_COMPACT_BILINEAR_POOLING = False
    setattr(net, P, True)
    setattr(net, P + "_SINGLE_LAYER_ATT", bool(flags.get("SINGLE_LAYER_ATT", False)))
    setattr(net, P + "_SOFTMAX_ATT", bool(flags.get("SOFTMAX_ATT", False)))
    setattr(net, P + "_RELU_ATT", bool(flags.get("RELU_ATT", False)))
    setattr(net, P + "_PER_CLASS", bool(flags.get("PER_CLASS", False)))
    setattr(net, P + "_RANK", int(flags.get("RANK", 1)))
    setattr(net, P + "_WITH_POSE_FEAT", False)
    setattr(net, P + "_WITH_POSE_FEAT_2LAYER", False)
    net.USE_TEMPORAL_ATT = False
    return types.SimpleNamespace(NET=net)


def run_reference_head(X, variables, flags, num_classes, num_pose_keypoints=16, frames_per_video=1, hw_shape=None):
    """Execute /root/reference/models/slim/nets/nets_factory.py::get_network_fn(...)(images).

    X: [B*F, HW, C]; it is presented to the reference as [B, F, H, W, C] (or [B, H, W, C] when F == 1).
    Returns (logits ndarray, end_points {name: ndarray}, variable names the graph created, in order).
    """
    if not os.path.exists(NETS_FACTORY):
        raise FileNotFoundError(NETS_FACTORY)
    X = np.asarray(X, dtype=np.float64)
    BF, HW, C = X.shape
    if hw_shape is None:
        h = int(round(HW ** 0.5))
        hw_shape = (h, HW // h) if h * (HW // h) == HW else (1, HW)
    H, W = hw_shape
    assert H * W == HW
    state = _State(variables)
    tf = _build_tf(state)
    saved = {}
    fakes = {"tensorflow": tf, "compact_bilinear_pooling": _AnyModule("compact_bilinear_pooling"),
             "nets": _AnyModule("nets")}
    for sub in ("alexnet", "cifarnet", "inception", "lenet", "overfeat", "resnet_v1", "resnet_v2", "vgg"):
        m = _AnyModule("nets." + sub)
        fakes["nets." + sub] = m
        setattr(fakes["nets"], sub, m)
    fakes["nets.resnet_v1"].__dict__["resnet_v1_101"] = _backbone
    for k, v in fakes.items():
        saved[k] = sys.modules.get(k)
        sys.modules[k] = v
    path_before = list(sys.path)
    try:
        spec = importlib.util.spec_from_file_location("_apa_reference_nets_factory", NETS_FACTORY)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)      # the reference file, unmodified
        network_fn = mod.get_network_fn("resnet_v1_101", num_classes, num_pose_keypoints, make_cfg(flags),
                                        weight_decay=0.0, is_training=False)
        if frames_per_video > 1:
            images = T(X.reshape(BF // frames_per_video, frames_per_video, H, W, C))
        else:
            images = T(X.reshape(BF, H, W, C))
        logits, end_points = network_fn(images)
    finally:
        sys.path[:] = path_before
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    eps = {k: v.a for k, v in end_points.items() if isinstance(v, T)}
    return logits.a, eps, list(state.touched)
