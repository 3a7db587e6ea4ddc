"""Loader + thin Python wrappers for the B200 attentional-pooling library.

Mirrors the reference's op packaging (src/custom_ops/custom_ops_factory.py:7-28): the shared
library sits next to this file, is loaded by absolute path at import time, and each entry
point is wrapped by a same-named Python function.  The reference loads with
`tf.load_op_library`; here the boundary is a plain C ABI (include/apa.h) loaded with ctypes.
PyTorch tensors are only containers for device memory and the CUDA stream.

There is NO CPU fallback: importing without the built library raises, and calling without
an sm_100 device returns the library's error.
"""
from __future__ import annotations

import ctypes
import os
from collections import OrderedDict

cur_path = os.path.realpath(__file__)
ROOT_PATH = os.path.dirname(cur_path)
LIB_NAME = "libapa_attn_pool.so"
LIB_PATH = os.path.join(ROOT_PATH, LIB_NAME)

APA_ABI_VERSION = 2
APA_MAX_RANK = 4
APA_FLAG_SINGLE_LAYER_ATT = 1 << 0
APA_FLAG_SOFTMAX_ATT = 1 << 1
APA_FLAG_RELU_ATT = 1 << 2
APA_FLAG_PER_CLASS = 1 << 3

APA_OK = 0
APA_ERR_NULL = -1
APA_ERR_SHAPE = -2
APA_ERR_RANK = -3
APA_ERR_SOFTMAX_RANK = -4
APA_ERR_WORKSPACE = -5
APA_ERR_UNSUPPORTED = -6
APA_ERR_ALIGN = -7
APA_ERR_DEVICE = -8

# every symbol include/apa.h declares (tests/test_abi.py checks the header against this list)
ABI_SYMBOLS = (
    "apa_abi_version",
    "apa_strerror",
    "apa_attn_pool_check",
    "apa_attn_pool_workspace",
    "apa_attn_pool_workspace_init",
    "apa_attn_pool_fwd",
    "apa_attn_pool_fwd_host",
    "apa_attn_pool_fwd_gather",
    "apa_peer_gather_wait",
    "apa_attn_pool_train_saved_bytes",
    "apa_attn_pool_train_workspace",
    "apa_attn_pool_fwd_train",
    "apa_attn_pool_bwd",
    "apa_dropout_mask",
    "apa_softmax_xent_loss",
    "apa_pose_l2_loss",
    "apa_eval_tail",
    "apa_mean_ap",
    "apa_gram_fwd",
    "apa_attn_pool_plan",
)


class ApaDesc(ctypes.Structure):
    _fields_ = [
        ("abi_version", ctypes.c_int32),
        ("batch", ctypes.c_int32),
        ("hw", ctypes.c_int32),
        ("channels", ctypes.c_int32),
        ("num_classes", ctypes.c_int32),
        ("frames", ctypes.c_int32),
        ("rank", ctypes.c_int32),
        ("flags", ctypes.c_int32),
        ("pose_channels", ctypes.c_int32),
        ("num_pose_keypoints", ctypes.c_int32),
        ("reserved", ctypes.c_int32 * 6),
    ]


_FP = ctypes.c_void_p  # const float* (device)


class ApaWeights(ctypes.Structure):
    _fields_ = [
        ("attn_w", _FP * APA_MAX_RANK),
        ("attn_b", _FP * APA_MAX_RANK),
        ("topdown_w", _FP * APA_MAX_RANK),
        ("topdown_b", _FP * APA_MAX_RANK),
        ("pose_w", _FP),
        ("pose_b", _FP),
        ("pose_logits_w", _FP),
        ("pose_logits_b", _FP),
    ]


class ApaTuning(ctypes.Structure):
    _fields_ = [
        ("form", ctypes.c_int32),
        ("stream_ctas", ctypes.c_int32),
        ("stream_shape", ctypes.c_int32),
        ("cluster_shape", ctypes.c_int32),
        ("no_pdl", ctypes.c_int32),
        ("inputs_stable", ctypes.c_int32),
        ("pdl_trigger_lead", ctypes.c_int32),
        ("gather_wait_lag", ctypes.c_int32),
        ("debug_timestamps", ctypes.c_int32),
        ("debug_cut_tail", ctypes.c_int32),
        ("tail_smem_kb", ctypes.c_int32),
        ("reserved", ctypes.c_int32 * 5),
    ]


# ApaTuning.form (include/apa.h APA_FORM_*)
APA_FORM_AUTO = 0
APA_FORM_FLOW = 1
APA_FORM_CLUSTER = 2


APA_MAX_PEERS = 8


class ApaPeerGather(ctypes.Structure):
    _fields_ = [
        ("world", ctypes.c_int32),
        ("rank", ctypes.c_int32),
        ("depth", ctypes.c_int32),
        ("reserved0", ctypes.c_int32),
        ("peer_logits", ctypes.c_void_p * APA_MAX_PEERS),
        ("peer_flags", ctypes.c_void_p * APA_MAX_PEERS),
        ("counters", ctypes.c_void_p),
    ]


class ApaTrain(ctypes.Structure):
    _fields_ = [
        ("dropout_keep", ctypes.c_float),
        ("reserved0", ctypes.c_int32),
        ("dropout_seed", ctypes.c_uint64),
    ]


class ApaGrads(ctypes.Structure):
    _fields_ = [
        ("dX", ctypes.c_void_p),
        ("attn_w", ctypes.c_void_p * APA_MAX_RANK),
        ("attn_b", ctypes.c_void_p * APA_MAX_RANK),
        ("topdown_w", ctypes.c_void_p * APA_MAX_RANK),
        ("topdown_b", ctypes.c_void_p * APA_MAX_RANK),
    ]


class ApaError(RuntimeError):
    def __init__(self, status, where=""):
        self.status = int(status)
        msg = _lib.apa_strerror(self.status).decode() if _lib is not None else "apa error"
        super().__init__("%s%s (status %d)" % (where + ": " if where else "", msg, self.status))


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "%s not found next to %s -- build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C attentionalpoolingaction_b200/csrc`). There is no CPU fallback." % (LIB_NAME, cur_path))
    lib = ctypes.CDLL(LIB_PATH)
    c = ctypes
    lib.apa_abi_version.restype = c.c_int
    lib.apa_abi_version.argtypes = []
    lib.apa_strerror.restype = c.c_char_p
    lib.apa_strerror.argtypes = [c.c_int]
    lib.apa_attn_pool_check.restype = c.c_int
    lib.apa_attn_pool_check.argtypes = [c.POINTER(ApaDesc)]
    lib.apa_attn_pool_workspace.restype = c.c_size_t
    lib.apa_attn_pool_workspace.argtypes = [c.POINTER(ApaDesc), c.c_int, c.c_int, c.c_int]
    lib.apa_attn_pool_workspace_init.restype = c.c_int
    lib.apa_attn_pool_workspace_init.argtypes = [c.c_void_p, c.c_size_t, c.c_void_p]
    lib.apa_attn_pool_fwd.restype = c.c_int
    lib.apa_attn_pool_fwd.argtypes = [c.POINTER(ApaDesc), c.c_void_p, c.POINTER(ApaWeights), c.c_void_p,
                                      c.c_void_p, c.c_void_p, c.c_void_p, c.c_void_p, c.c_size_t,
                                      c.POINTER(ApaTuning), c.c_void_p]
    lib.apa_attn_pool_fwd_host.restype = c.c_int
    lib.apa_attn_pool_fwd_host.argtypes = [c.POINTER(ApaDesc), c.c_void_p, c.POINTER(ApaWeights), c.c_void_p,
                                           c.c_void_p, c.c_void_p, c.c_void_p, c.c_size_t,
                                           c.POINTER(ApaTuning), c.c_void_p]
    lib.apa_attn_pool_fwd_gather.restype = c.c_int
    lib.apa_attn_pool_fwd_gather.argtypes = [c.POINTER(ApaDesc), c.c_void_p, c.POINTER(ApaWeights),
                                             c.POINTER(ApaPeerGather), c.c_void_p, c.c_size_t,
                                             c.POINTER(ApaTuning), c.c_void_p]
    lib.apa_peer_gather_wait.restype = c.c_int
    lib.apa_peer_gather_wait.argtypes = [c.POINTER(ApaPeerGather), c.c_int32, c.c_void_p]
    lib.apa_attn_pool_train_saved_bytes.restype = c.c_size_t
    lib.apa_attn_pool_train_saved_bytes.argtypes = [c.POINTER(ApaDesc)]
    lib.apa_attn_pool_train_workspace.restype = c.c_size_t
    lib.apa_attn_pool_train_workspace.argtypes = [c.POINTER(ApaDesc)]
    lib.apa_attn_pool_fwd_train.restype = c.c_int
    lib.apa_attn_pool_fwd_train.argtypes = [c.POINTER(ApaDesc), c.c_void_p, c.POINTER(ApaWeights), c.POINTER(ApaTrain),
                                            c.c_void_p, c.c_void_p, c.c_size_t, c.c_void_p, c.c_size_t, c.c_void_p]
    lib.apa_attn_pool_bwd.restype = c.c_int
    lib.apa_attn_pool_bwd.argtypes = [c.POINTER(ApaDesc), c.c_void_p, c.POINTER(ApaWeights), c.POINTER(ApaTrain),
                                      c.c_void_p, c.c_void_p, c.c_size_t, c.POINTER(ApaGrads), c.c_void_p,
                                      c.c_size_t, c.c_void_p]
    lib.apa_dropout_mask.restype = c.c_int
    lib.apa_dropout_mask.argtypes = [c.POINTER(ApaTrain), c.c_size_t, c.c_void_p, c.c_void_p]
    lib.apa_softmax_xent_loss.restype = c.c_int
    lib.apa_softmax_xent_loss.argtypes = [c.c_int32, c.c_int32, c.c_void_p, c.c_void_p, c.c_float, c.c_void_p, c.c_void_p,
                                          c.c_void_p, c.c_void_p]
    lib.apa_pose_l2_loss.restype = c.c_int
    lib.apa_pose_l2_loss.argtypes = [c.c_int32, c.c_int32, c.c_int32, c.c_void_p, c.c_void_p, c.c_void_p, c.c_float,
                                     c.c_void_p, c.c_void_p, c.c_void_p, c.c_void_p]
    lib.apa_eval_tail.restype = c.c_int
    lib.apa_eval_tail.argtypes = [c.c_int32, c.c_int32, c.c_void_p, c.c_int32, c.c_void_p, c.c_void_p, c.c_void_p,
                                  c.c_void_p, c.c_void_p]
    lib.apa_mean_ap.restype = c.c_int
    lib.apa_mean_ap.argtypes = [c.c_int32, c.c_int32, c.c_void_p, c.c_void_p, c.c_void_p, c.c_void_p, c.c_void_p]
    lib.apa_gram_fwd.restype = c.c_int
    lib.apa_gram_fwd.argtypes = [c.c_int32, c.c_int32, c.c_int32, c.c_void_p, c.c_void_p, c.c_void_p]
    lib.apa_attn_pool_plan.restype = c.c_int
    lib.apa_attn_pool_plan.argtypes = [c.POINTER(ApaDesc), c.POINTER(ApaTuning), c.c_int, c.c_int, c.c_int,
                                       c.POINTER(c.c_int32 * 8)]
    if lib.apa_abi_version() != APA_ABI_VERSION:
        raise ImportError("ABI mismatch: library %d, wrapper %d" % (lib.apa_abi_version(), APA_ABI_VERSION))
    return lib


_lib = None
_lib = _load()
attn_pool_lib = _lib  # like `pose_to_heatmap_fn = tf.load_op_library(...)` in the reference

# TF variable names of the head (nets_factory.py:147-160, :243-303)
POSE_W = "PoseLogits/ExtraConv2d_1x1/weights"
POSE_B = "PoseLogits/ExtraConv2d_1x1/biases"
POSE_LOGITS_W = "PoseLogits/Conv2d_1c_1x1/weights"
POSE_LOGITS_B = "PoseLogits/Conv2d_1c_1x1/biases"


def attn_scope(rank_id):
    return "PosePrelogitsBasedAttention/Conv2d_PrePose_Attn" + (str(rank_id) if rank_id >= 1 else "")


def topdown_scope(rank_id):
    return "PosePrelogitsBasedAttention/Conv" + ("_%d" % rank_id if rank_id >= 1 else "")


def flags_from_cfg_net(net):
    """cfg.NET.* (src/config.py:182-209) -> (APA flag bits, rank).  `net` is attribute- or dict-like."""
    def get(key, default):
        full = "USE_POSE_PRELOGITS_BASED_ATTENTION_" + key
        if isinstance(net, dict):
            return net.get(full, net.get(key, default))
        return getattr(net, full, getattr(net, key, default))
    flags = 0
    if get("SINGLE_LAYER_ATT", False):
        flags |= APA_FLAG_SINGLE_LAYER_ATT
    if get("SOFTMAX_ATT", False):
        flags |= APA_FLAG_SOFTMAX_ATT
    if get("RELU_ATT", False):
        flags |= APA_FLAG_RELU_ATT
    if get("PER_CLASS", False):
        flags |= APA_FLAG_PER_CLASS
    return flags, int(get("RANK", 1))


def make_desc(batch, hw, channels, num_classes, frames=1, rank=1, flags=APA_FLAG_SINGLE_LAYER_ATT,
              pose_channels=768, num_pose_keypoints=16):
    d = ApaDesc()
    d.abi_version = APA_ABI_VERSION
    d.batch, d.hw, d.channels, d.num_classes = int(batch), int(hw), int(channels), int(num_classes)
    d.frames, d.rank, d.flags = int(frames), int(rank), int(flags)
    d.pose_channels, d.num_pose_keypoints = int(pose_channels), int(num_pose_keypoints)
    return d


def make_tuning(form=APA_FORM_AUTO, stream_ctas=0, stream_shape=0, cluster_shape=0, no_pdl=0, inputs_stable=0,
                pdl_trigger_lead=0, gather_wait_lag=0, debug_timestamps=0, debug_cut_tail=0, tail_smem_kb=0):
    """ApaTuning (include/apa.h).  `inputs_stable=1` is a promise by the caller: X and the weights were completely
    written before the previous kernel in the stream was launched (lets back-to-back calls overlap)."""
    t = ApaTuning()
    t.form, t.stream_ctas, t.stream_shape, t.cluster_shape = int(form), int(stream_ctas), int(stream_shape), int(cluster_shape)
    t.no_pdl, t.inputs_stable, t.pdl_trigger_lead = int(no_pdl), int(inputs_stable), int(pdl_trigger_lead)
    t.gather_wait_lag, t.debug_timestamps, t.debug_cut_tail = int(gather_wait_lag), int(debug_timestamps), int(debug_cut_tail)
    t.tail_smem_kb = int(tail_smem_kb)
    return t


def attn_pool_check(desc):
    return int(_lib.apa_attn_pool_check(ctypes.byref(desc)))


def attn_pool_workspace(desc, want_attn=False, want_topdown=False, want_pose=False):
    return int(_lib.apa_attn_pool_workspace(ctypes.byref(desc), int(want_attn), int(want_topdown),
                                            int(want_pose)))


def attn_pool_plan(desc, tuning=None, want_attn=False, want_topdown=False, want_pose=False):
    out = (ctypes.c_int32 * 8)()
    st = _lib.apa_attn_pool_plan(ctypes.byref(desc), ctypes.byref(tuning) if tuning is not None else None,
                                 int(want_attn), int(want_topdown), int(want_pose), ctypes.byref(out))
    if st != APA_OK:
        raise ApaError(st, "apa_attn_pool_plan")
    # form: APA_FORM_FLOW / APA_FORM_CLUSTER on the single-map fast path (path 0), 0 on the dense path;
    # grid = CTAs of the streaming kernel, contract_grid = CTAs of the flow form's CONTRACT kernel
    keys = ("path", "form", "grid", "launches", "smem_bytes", "contract_grid")
    return OrderedDict(zip(keys, list(out)[:6]))


def strerror(status):
    return _lib.apa_strerror(int(status)).decode()


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


class AttnPoolOp(object):
    """One configured head: descriptor + device-resident variables + workspace.

    `params` maps the reference's TF variable names to float32 CUDA tensors shaped
    [Cin, Cout] / [Cout] (HWIO with the leading 1,1 dropped; 4-D HWIO tensors are accepted
    and viewed).  The op keeps references so the pointers stay valid.
    """

    def __init__(self, params, num_classes, hw, channels=2048, batch=1, frames=1, rank=1,
                 flags=APA_FLAG_SINGLE_LAYER_ATT, num_pose_keypoints=16, tuning=None, device=None):
        import torch
        self._torch = torch
        self.device = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())
        self.num_classes, self.hw, self.channels = int(num_classes), int(hw), int(channels)
        self.frames, self.rank, self.flags = int(frames), int(rank), int(flags)
        self.num_pose_keypoints = int(num_pose_keypoints)
        self.tuning = tuning
        self._keep = []
        self.weights = ApaWeights()
        single = bool(flags & APA_FLAG_SINGLE_LAYER_ATT)

        def dev(name, required=True):
            if name not in params:
                if required:
                    raise KeyError("missing variable %r" % name)
                return None
            t = params[name]
            if not torch.is_tensor(t):
                t = torch.as_tensor(t)
            t = t.to(device=self.device, dtype=torch.float32)
            if t.dim() == 4:  # HWIO [1,1,Cin,Cout]
                t = t.reshape(t.shape[2], t.shape[3])
            t = t.contiguous()
            self._keep.append(t)
            return t

        self.pose_channels = 768
        pw = dev(POSE_W, required=not single)
        if pw is not None:
            self.pose_channels = int(pw.shape[1])
            self.weights.pose_w = _ptr(pw)
            self.weights.pose_b = _ptr(dev(POSE_B))
            plw = dev(POSE_LOGITS_W, required=False)
            if plw is not None:
                self.num_pose_keypoints = int(plw.shape[1])
                self.weights.pose_logits_w = _ptr(plw)
                self.weights.pose_logits_b = _ptr(dev(POSE_LOGITS_B))
        for r in range(self.rank):
            self.weights.attn_w[r] = _ptr(dev(attn_scope(r) + "/weights")).value
            self.weights.attn_b[r] = _ptr(dev(attn_scope(r) + "/biases")).value
            self.weights.topdown_w[r] = _ptr(dev(topdown_scope(r) + "/weights")).value
            self.weights.topdown_b[r] = _ptr(dev(topdown_scope(r) + "/biases")).value
        self._has_pose_logits = bool(self.weights.pose_logits_w)
        self._ws = {}  # (batch, want flags) -> (desc, workspace tensor)

    def desc(self, batch):
        return make_desc(batch, self.hw, self.channels, self.num_classes, self.frames, self.rank, self.flags,
                         self.pose_channels, self.num_pose_keypoints)

    def _workspace(self, batch, want):
        key = (int(batch),) + tuple(bool(w) for w in want)
        hit = self._ws.get(key)
        if hit is not None:
            return hit
        torch = self._torch
        d = self.desc(batch)
        st = attn_pool_check(d)
        if st != APA_OK:
            raise ApaError(st, "apa_attn_pool_check")
        nbytes = attn_pool_workspace(d, *want)
        ws = torch.zeros(nbytes + 256, dtype=torch.uint8, device=self.device)
        off = (-ws.data_ptr()) % 256
        ws_al = ws[off:off + nbytes]
        stream = torch.cuda.current_stream(self.device).cuda_stream
        st = _lib.apa_attn_pool_workspace_init(_ptr(ws_al), nbytes, ctypes.c_void_p(stream))
        if st != APA_OK:
            raise ApaError(st, "apa_attn_pool_workspace_init")
        self._ws[key] = (d, ws_al, ws)
        return self._ws[key]

    def __call__(self, X, want_attn=False, want_topdown=False, want_pose=False, logits=None):
        """X: float32 CUDA tensor [B*F, HW, C] (or [B*F, H, W, C]).  Returns (logits, end_points)."""
        torch = self._torch
        if not X.is_cuda or X.dtype != torch.float32:
            raise TypeError("X must be a float32 CUDA tensor (PyTorch is only the container here)")
        if X.dim() == 4:
            X = X.reshape(X.shape[0], X.shape[1] * X.shape[2], X.shape[3])
        X = X.contiguous()
        batch, hw, ch = X.shape
        if hw != self.hw or ch != self.channels:
            raise ValueError("X is [%d,%d,%d], op was built for hw=%d channels=%d" %
                             (batch, hw, ch, self.hw, self.channels))
        want = (want_attn, want_topdown, want_pose and self._has_pose_logits)
        d, ws, _ = self._workspace(batch, want)
        K, R = self.num_classes, self.rank
        M = K if (self.flags & APA_FLAG_PER_CLASS) else 1
        if logits is None:
            logits = torch.empty((batch // self.frames, K), dtype=torch.float32, device=X.device)
        tail = (R,) if R > 1 else ()
        attn = torch.empty((batch, hw, M) + tail, dtype=torch.float32, device=X.device) if want[0] else None
        topdown = torch.empty((batch, hw, K) + tail, dtype=torch.float32, device=X.device) if want[1] else None
        pose = (torch.empty((batch, hw, self.num_pose_keypoints), dtype=torch.float32, device=X.device)
                if want[2] else None)
        stream = torch.cuda.current_stream(X.device).cuda_stream
        st = _lib.apa_attn_pool_fwd(ctypes.byref(d), _ptr(X), ctypes.byref(self.weights), _ptr(logits),
                                    _ptr(attn), _ptr(topdown), _ptr(pose), _ptr(ws), ws.numel(),
                                    ctypes.byref(self.tuning) if self.tuning is not None else None,
                                    ctypes.c_void_p(stream))
        if st != APA_OK:
            raise ApaError(st, "apa_attn_pool_fwd")
        end_points = OrderedDict()
        if pose is not None:
            end_points["PoseLogits"] = pose
        if attn is not None:
            end_points["PosePrelogitsBasedAttention"] = attn
        if topdown is not None:
            end_points["TopDownAttention"] = topdown
        end_points["Logits"] = logits
        return logits, end_points

    def forward_gather(self, X, gather):
        """Batch-sharded multi-GPU call: this rank's logits are stored straight into every peer's gathered
        ring by the kernel.  `gather` is a dist.PeerGather (its host-side mirror of the device step counter is
        advanced here; returns the step of this call) or a filled-in ApaPeerGather."""
        torch = self._torch
        step = None
        if hasattr(gather, "next_step"):
            step = gather.next_step()
            gather = gather.struct
        if X.dim() == 4:
            X = X.reshape(X.shape[0], X.shape[1] * X.shape[2], X.shape[3])
        X = X.contiguous()
        d, ws, _ = self._workspace(X.shape[0], (False, False, False))
        stream = torch.cuda.current_stream(X.device).cuda_stream
        st = _lib.apa_attn_pool_fwd_gather(ctypes.byref(d), _ptr(X), ctypes.byref(self.weights),
                                           ctypes.byref(gather), _ptr(ws), ws.numel(),
                                           ctypes.byref(self.tuning) if self.tuning is not None else None,
                                           ctypes.c_void_p(stream))
        if st != APA_OK:
            raise ApaError(st, "apa_attn_pool_fwd_gather")
        return step

    def forward_host(self, X_host, logits_host=None):
        """Plugin-style call with HOST buffers (pinned for speed): H2D copy, head, D2H copy, sync."""
        torch = self._torch
        if X_host.is_cuda or X_host.dtype != torch.float32:
            raise TypeError("X_host must be a float32 CPU tensor")
        if X_host.dim() == 4:
            X_host = X_host.reshape(X_host.shape[0], X_host.shape[1] * X_host.shape[2], X_host.shape[3])
        X_host = X_host.contiguous()
        batch = X_host.shape[0]
        d, ws, _ = self._workspace(batch, (False, False, False))
        key = ("staging", batch)
        st_bufs = self._ws.get(key)
        if st_bufs is None:
            st_bufs = (torch.empty(X_host.shape, dtype=torch.float32, device=self.device),
                       torch.empty((batch // self.frames, self.num_classes), dtype=torch.float32,
                                   device=self.device))
            self._ws[key] = st_bufs
        if logits_host is None:
            logits_host = torch.empty((batch // self.frames, self.num_classes), dtype=torch.float32,
                                      pin_memory=True)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        st = _lib.apa_attn_pool_fwd_host(ctypes.byref(d), _ptr(X_host), ctypes.byref(self.weights),
                                         _ptr(logits_host), _ptr(st_bufs[0]), _ptr(st_bufs[1]), _ptr(ws),
                                         ws.numel(),
                                         ctypes.byref(self.tuning) if self.tuning is not None else None,
                                         ctypes.c_void_p(stream))
        if st != APA_OK:
            raise ApaError(st, "apa_attn_pool_fwd_host")
        return logits_host


def make_train(dropout_keep=0.2, dropout_seed=0):
    """ApaTrain: keep probability of slim.dropout(last_conv) (0.2 when cfg.NET.DROPOUT < 0, nets_factory.py:143-146)
    and the seed of this step's mask."""
    t = ApaTrain()
    t.dropout_keep, t.dropout_seed = float(dropout_keep), int(dropout_seed) & 0xFFFFFFFFFFFFFFFF
    return t


def dropout_mask(train, shape, device):
    """The mask apa_attn_pool_fwd_train / _bwd apply to X of `shape` (0 or 1/keep), materialised."""
    import torch
    n = 1
    for v in shape:
        n *= int(v)
    out = torch.empty(shape, dtype=torch.float32, device=device)
    st = _lib.apa_dropout_mask(ctypes.byref(train), n, _ptr(out),
                               ctypes.c_void_p(torch.cuda.current_stream(out.device).cuda_stream))
    if st != APA_OK:
        raise ApaError(st, "apa_dropout_mask")
    return out


class AttnPoolTrainOp(AttnPoolOp):
    """The head in training mode (train.py:399-420 builds it with is_training=True; :509-513 differentiates it):
    `forward(X, train)` -> logits and `backward(dlogits)` -> (dX, {TF variable name: gradient}).  Single-map head
    that reads X directly (SINGLE_LAYER_ATT, not PER_CLASS); everything else raises ApaError(UNSUPPORTED)."""

    def _train_buffers(self, batch):
        key = ("train", int(batch))
        hit = self._ws.get(key)
        if hit is None:
            torch = self._torch
            d = self.desc(batch)
            st = attn_pool_check(d)
            if st != APA_OK:
                raise ApaError(st, "apa_attn_pool_check")
            nsaved, nws = int(_lib.apa_attn_pool_train_saved_bytes(ctypes.byref(d))), int(
                _lib.apa_attn_pool_train_workspace(ctypes.byref(d)))
            if nsaved == 0 or nws == 0:
                raise ApaError(APA_ERR_UNSUPPORTED, "training mode")

            def aligned(n):
                raw = torch.empty(n + 256, dtype=torch.uint8, device=self.device)
                off = (-raw.data_ptr()) % 256
                return raw[off:off + n], raw
            hit = (d, aligned(nsaved), aligned(nws))
            self._ws[key] = hit
        return hit

    def forward(self, X, train=None, logits=None):
        torch = self._torch
        if X.dim() == 4:
            X = X.reshape(X.shape[0], X.shape[1] * X.shape[2], X.shape[3])
        X = X.contiguous()
        d, (saved, _), (ws, _) = self._train_buffers(X.shape[0])
        if logits is None:
            logits = torch.empty((X.shape[0] // self.frames, self.num_classes), dtype=torch.float32, device=X.device)
        st = _lib.apa_attn_pool_fwd_train(ctypes.byref(d), _ptr(X), ctypes.byref(self.weights),
                                          ctypes.byref(train) if train is not None else None, _ptr(logits),
                                          _ptr(saved), saved.numel(), _ptr(ws), ws.numel(),
                                          ctypes.c_void_p(torch.cuda.current_stream(X.device).cuda_stream))
        if st != APA_OK:
            raise ApaError(st, "apa_attn_pool_fwd_train")
        self._last = (X, train)
        return logits

    def backward(self, dlogits, want_dX=True, grads=None):
        """Gradients of sum(logits * dlogits) for the last `forward` call.  Returns (dX or None, OrderedDict of
        TF variable name -> gradient tensor shaped like the variable)."""
        torch = self._torch
        X, train = self._last
        d, (saved, _), (ws, _) = self._train_buffers(X.shape[0])
        dlogits = dlogits.contiguous()
        C, K = self.channels, self.num_classes
        if grads is None:
            grads = OrderedDict()
            for r in range(self.rank):
                grads[attn_scope(r) + "/weights"] = torch.empty((C, 1) if r == 0 else (1, 1), device=X.device)
                grads[attn_scope(r) + "/biases"] = torch.empty((1,), device=X.device)
                grads[topdown_scope(r) + "/weights"] = torch.empty((C, K), device=X.device)
                grads[topdown_scope(r) + "/biases"] = torch.empty((K,), device=X.device)
        dX = torch.empty_like(X) if want_dX is True else (want_dX if torch.is_tensor(want_dX) else None)
        g = ApaGrads()
        g.dX = _ptr(dX).value if dX is not None else None
        for r in range(self.rank):
            g.attn_w[r] = _ptr(grads[attn_scope(r) + "/weights"]).value
            g.attn_b[r] = _ptr(grads[attn_scope(r) + "/biases"]).value
            g.topdown_w[r] = _ptr(grads[topdown_scope(r) + "/weights"]).value
            g.topdown_b[r] = _ptr(grads[topdown_scope(r) + "/biases"]).value
        st = _lib.apa_attn_pool_bwd(ctypes.byref(d), _ptr(X), ctypes.byref(self.weights),
                                    ctypes.byref(train) if train is not None else None, _ptr(dlogits),
                                    _ptr(saved), saved.numel(), ctypes.byref(g), _ptr(ws), ws.numel(),
                                    ctypes.c_void_p(torch.cuda.current_stream(X.device).cuda_stream))
        if st != APA_OK:
            raise ApaError(st, "apa_attn_pool_bwd")
        return dX, grads


def softmax_xent_loss(logits, labels, weight=1.0, want_grad=True):
    """loss.py:72-79 (softmax-xentropy action loss): returns (loss 0-d tensor, dlogits or None).  `dlogits` is what
    AttnPoolTrainOp.backward takes."""
    import torch
    logits, labels = logits.contiguous(), labels.contiguous()
    if labels.dtype != torch.int32:
        raise TypeError("labels must be int32")
    rows, K = logits.shape
    row_loss = torch.empty((rows,), dtype=torch.float32, device=logits.device)
    loss = torch.empty((1,), dtype=torch.float32, device=logits.device)
    grad = torch.empty_like(logits) if want_grad else None
    st = _lib.apa_softmax_xent_loss(rows, K, _ptr(logits), _ptr(labels), float(weight), _ptr(row_loss), _ptr(loss),
                                    _ptr(grad), ctypes.c_void_p(torch.cuda.current_stream(logits.device).cuda_stream))
    if st != APA_OK:
        raise ApaError(st, "apa_softmax_xent_loss")
    return loss[0], grad


def pose_l2_loss(pose_logits, pose_labels, valid, weight=1.0, want_grad=True):
    """loss.py:29-70 (masked L2 heat-map loss, un-sampled): pose_logits / pose_labels [B, HW, Kp] (or [B, H, W, Kp]),
    valid [B, Kp] bool / uint8.  Returns (loss 0-d tensor, d loss / d pose_logits or None)."""
    import torch
    shape = pose_logits.shape
    x = pose_logits.reshape(shape[0], -1, shape[-1]).contiguous()
    y = pose_labels.reshape(shape[0], -1, shape[-1]).contiguous()
    if x.shape != y.shape:
        raise ValueError("labels must have the logits' size (the reference resizes them otherwise, loss.py:14-22)")
    v = valid.to(torch.uint8).contiguous()
    B, HW, Kp = x.shape
    row_loss = torch.empty((B,), dtype=torch.float32, device=x.device)
    loss = torch.empty((1,), dtype=torch.float32, device=x.device)
    grad = torch.empty_like(x) if want_grad else None
    st = _lib.apa_pose_l2_loss(B, HW, Kp, _ptr(x), _ptr(y), _ptr(v), float(weight), _ptr(row_loss), _ptr(loss), _ptr(grad),
                               ctypes.c_void_p(torch.cuda.current_stream(x.device).cuda_stream))
    if st != APA_OK:
        raise ApaError(st, "apa_pose_l2_loss")
    return loss[0], (grad.reshape(shape) if grad is not None else None)


APA_EVAL_SOFTMAX = 0
APA_EVAL_SIGMOID = 1
APA_AP_MAX_SAMPLES = 16384


def eval_tail(logits, labels=None, counts=None, multi_label=False, want_probs=True):
    """What eval.py:193-203 does with a batch of logits: returns (predictions int32 [rows], probs [rows, K] or None).
    With `labels` (int32 [rows]) the hits and the rows seen are added to `counts` (uint64 [2], zeroed by the caller
    before the first batch): counts[0] / counts[1] is the streaming accuracy."""
    import torch
    logits = logits.contiguous()
    rows, K = logits.shape
    pred = torch.empty((rows,), dtype=torch.int32, device=logits.device)
    probs = torch.empty_like(logits) if want_probs else None
    if (labels is None) != (counts is None):
        raise ValueError("labels and counts go together")
    if labels is not None:
        if labels.dtype != torch.int32 or counts.dtype not in (torch.int64, torch.uint64) or counts.numel() < 2:
            raise TypeError("labels must be int32 and counts a 64-bit integer tensor of 2 words")
        labels = labels.contiguous()
    st = _lib.apa_eval_tail(rows, K, _ptr(logits), APA_EVAL_SIGMOID if multi_label else APA_EVAL_SOFTMAX,
                            _ptr(probs), _ptr(pred), _ptr(labels), _ptr(counts),
                            ctypes.c_void_p(torch.cuda.current_stream(logits.device).cuda_stream))
    if st != APA_OK:
        raise ApaError(st, "apa_eval_tail")
    return pred, probs


def mean_ap(scores, labels):
    """compute_map of the reference (eval.py:305 -> eval/utils.py:4-16) on the GPU: returns (mAP, class APs float64
    [K] with -1 for classes without positives) as tensors; scores float32 [samples, K], labels int32 [samples]."""
    import torch
    scores, labels = scores.contiguous(), labels.contiguous()
    if scores.dtype != torch.float32 or labels.dtype != torch.int32:
        raise TypeError("scores must be float32 and labels int32")
    n, K = scores.shape
    aps = torch.empty((K,), dtype=torch.float64, device=scores.device)
    out = torch.empty((2,), dtype=torch.float64, device=scores.device)
    st = _lib.apa_mean_ap(n, K, _ptr(scores), _ptr(labels), _ptr(aps), _ptr(out),
                          ctypes.c_void_p(torch.cuda.current_stream(scores.device).cuda_stream))
    if st != APA_OK:
        raise ApaError(st, "apa_mean_ap")
    return out[0], aps


def peer_gather_wait(gather_struct, lag, stream):
    """Enqueue a wait on `stream` until every rank's block of call (latest - lag) has landed in this rank's ring."""
    st = _lib.apa_peer_gather_wait(ctypes.byref(gather_struct), int(lag), ctypes.c_void_p(stream))
    if st != APA_OK:
        raise ApaError(st, "apa_peer_gather_wait")


def attn_pool(X, params, num_classes, flags=APA_FLAG_SINGLE_LAYER_ATT, rank=1, frames=1, **kwargs):
    """Functional form, like the reference's `pose_to_heatmap(*args, **kwargs)` wrappers: builds the op
    for this call's shapes and runs it.  Prefer AttnPoolOp for repeated calls."""
    want = {k: kwargs.pop(k, False) for k in ("want_attn", "want_topdown", "want_pose")}
    hw = X.shape[1] if X.dim() == 3 else X.shape[1] * X.shape[2]
    op = AttnPoolOp(params, num_classes, hw, channels=X.shape[-1], frames=frames, rank=rank, flags=flags,
                    device=X.device, **kwargs)
    return op(X, **want)


def gram(X, out=None):
    """Full second-order pooling G[b] = X[b]^T X[b] (paper; not in the reference code)."""
    import torch
    if X.dim() == 4:
        X = X.reshape(X.shape[0], X.shape[1] * X.shape[2], X.shape[3])
    X = X.contiguous()
    B, HW, C = X.shape
    G = out if out is not None else torch.empty((B, C, C), dtype=torch.float32, device=X.device)
    if tuple(G.shape) != (B, C, C) or G.dtype != torch.float32 or not G.is_contiguous():
        raise ValueError("out must be a contiguous float32 [B, C, C] tensor")
    st = _lib.apa_gram_fwd(B, HW, C, _ptr(X), _ptr(G),
                           ctypes.c_void_p(torch.cuda.current_stream(X.device).cuda_stream))
    if st != APA_OK:
        raise ApaError(st, "apa_gram_fwd")
    return G
