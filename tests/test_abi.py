"""CPU: the C-ABI library loads and exports every symbol include/apa.h declares; the no-device
entry points (validation, workspace sizing, plan, strerror) behave; nothing here launches a kernel."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "apa.h")


def _declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(apa_[a-z_0-9]+)\s*\(", src)))


def test_header_symbols_are_exported_and_wrapped():
    from attentionalpoolingaction_b200.custom_ops import attn_pool_factory as ops
    declared = _declared_symbols()
    assert declared == sorted(ops.ABI_SYMBOLS)
    out = subprocess.run(["nm", "-D", "--defined-only", ops.LIB_PATH], capture_output=True, text=True, check=True)
    exported = set(re.findall(r" T (apa_[a-z_0-9]+)", out.stdout))
    assert set(declared) <= exported
    lib = ctypes.CDLL(ops.LIB_PATH)
    for s in declared:
        assert getattr(lib, s) is not None


def test_header_compiles_as_plain_c():
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-fsyntax-only", "-x", "c", HEADER], check=True)


def test_struct_layouts_match_header():
    from attentionalpoolingaction_b200.custom_ops import attn_pool_factory as ops
    prog = r'''
#include <stdio.h>
#include "apa.h"
int main(void){ printf("%zu %zu %zu %zu %zu %zu %d %d\n", sizeof(ApaDesc), sizeof(ApaWeights), sizeof(ApaTuning),
                       sizeof(ApaPeerGather), sizeof(ApaTrain), sizeof(ApaGrads), APA_ABI_VERSION, APA_MAX_RANK);
                return 0; }'''
    exe = "/tmp/apa_sizes_%d" % os.getpid()
    subprocess.run(["gcc", "-x", "c", "-I", os.path.join(ROOT, "include"), "-o", exe, "-"], input=prog,
                   text=True, check=True)
    d, w, t, g, tr, gr, ver, mr = map(int, subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split())
    os.remove(exe)
    assert (d, w, t, g) == (ctypes.sizeof(ops.ApaDesc), ctypes.sizeof(ops.ApaWeights), ctypes.sizeof(ops.ApaTuning),
                            ctypes.sizeof(ops.ApaPeerGather))
    assert (tr, gr) == (ctypes.sizeof(ops.ApaTrain), ctypes.sizeof(ops.ApaGrads))
    assert ver == ops.APA_ABI_VERSION == ops.attn_pool_lib.apa_abi_version()
    assert mr == ops.APA_MAX_RANK


def test_descriptor_validation_codes():
    from attentionalpoolingaction_b200.custom_ops import attn_pool_factory as ops
    ok = ops.make_desc(32, 196, 2048, 393)
    assert ops.attn_pool_check(ok) == ops.APA_OK
    assert ops.attn_pool_check(ops.make_desc(0, 196, 2048, 393)) == ops.APA_ERR_SHAPE
    assert ops.attn_pool_check(ops.make_desc(32, 196, 2048, 393, frames=5)) == ops.APA_ERR_SHAPE
    assert ops.attn_pool_check(ops.make_desc(30, 196, 2048, 393, frames=10)) == ops.APA_OK
    assert ops.attn_pool_check(ops.make_desc(32, 196, 2048, 393, rank=5)) == ops.APA_ERR_RANK
    assert ops.attn_pool_check(ops.make_desc(32, 196, 2048, 393, rank=0)) == ops.APA_ERR_RANK
    bad = ops.make_desc(32, 196, 2048, 393, rank=2, flags=ops.APA_FLAG_SINGLE_LAYER_ATT | ops.APA_FLAG_SOFTMAX_ATT)
    assert ops.attn_pool_check(bad) == ops.APA_ERR_SOFTMAX_RANK
    assert "nets_factory.py:278" in ops.strerror(ops.APA_ERR_SOFTMAX_RANK)
    assert ops.attn_pool_check(ops.make_desc(32, 196, 2048, 393, flags=1 << 9)) == ops.APA_ERR_UNSUPPORTED
    v = ops.make_desc(32, 196, 2048, 393)
    v.abi_version = 99
    assert ops.attn_pool_check(v) == ops.APA_ERR_UNSUPPORTED
    assert ops.attn_pool_check(ops.make_desc(4, 196, 2048, 393, flags=0, pose_channels=0)) == ops.APA_ERR_SHAPE
    for code in range(-8, 1):
        assert ops.strerror(code)


def test_workspace_sizes():
    from attentionalpoolingaction_b200.custom_ops import attn_pool_factory as ops
    d = ops.make_desc(32, 196, 2048, 393)
    base = ops.attn_pool_workspace(d)
    assert base >= 256 + 32 * 2048 * 4 and base % 256 == 0
    assert ops.attn_pool_workspace(d, want_attn=True) >= base + 32 * 196 * 4
    assert ops.attn_pool_workspace(d, want_topdown=True) < 64 * 1024   # outputs are caller-owned
    dense = ops.make_desc(2, 196, 2048, 393, flags=ops.APA_FLAG_SINGLE_LAYER_ATT | ops.APA_FLAG_PER_CLASS)
    assert ops.attn_pool_workspace(dense) >= 2 * 2 * 196 * 393 * 4
    assert ops.attn_pool_workspace(ops.make_desc(0, 1, 4, 1)) == 0
    cfg3 = ops.make_desc(128, 196, 2048, 393, rank=2, flags=0)
    assert ops.attn_pool_workspace(cfg3) >= 128 * 196 * 768 * 4


@pytest.mark.parametrize("kw,expect", [
    # (descriptor, tuning) -> (form, grid of the streaming kernel, launches per call, CONTRACT grid)
    (dict(B=32), (1, 148, 4, 16 * 13)),                       # config 2: flow form, one STREAM CTA per SM
    (dict(B=1), (1, 12, 4, 16 * 13)),                         # the eval.py batch: at least 16 rows per STREAM CTA
    (dict(B=1, HW=10), (1, 1, 4, 16 * 13)),
    (dict(B=30), (1, 148, 4, 16 * 13)),                       # any batch
    (dict(B=320, F=10, K=600), (1, 148, 4, 16 * 19)),         # HICO 10-crop
    (dict(B=32, C=1024, rank=3), (1, 148, 4, 16 * 13)),
    (dict(B=32, stream_ctas=37), (1, 37, 4, 16 * 13)),        # throughput setting: fewer STREAM CTAs per call
    (dict(B=32, form=2), (2, 64, 1, 0)),                      # cluster form on request: 8 images per cluster
    (dict(B=36, form=2), (2, 144, 1, 0)),
    (dict(B=32, form=2, cluster_shape=1), (2, 128, 1, 0)),
    (dict(B=30, form=2), (1, 148, 4, 16 * 13)),               # ... and the flow form where it does not apply
    (dict(B=32, form=2, F=2), (1, 148, 4, 16 * 13)),
    (dict(B=32, pose=True), (1, 148, 5, 16 * 13)),            # config 003: the pose GEMM in front
])
def test_dispatch_policy_without_device(kw, expect):
    """Which form of the single-map head a descriptor gets: no device needed."""
    from attentionalpoolingaction_b200.custom_ops import attn_pool_factory as ops
    kw = dict(kw)
    t = ops.make_tuning(form=kw.pop("form", 0), stream_ctas=kw.pop("stream_ctas", 0), cluster_shape=kw.pop("cluster_shape", 0))
    flags = 0 if kw.pop("pose", False) else ops.APA_FLAG_SINGLE_LAYER_ATT
    d = ops.make_desc(kw.get("B"), kw.get("HW", 196), kw.get("C", 2048), kw.get("K", 393), frames=kw.get("F", 1),
                      rank=kw.get("rank", 1), flags=flags)
    p = ops.attn_pool_plan(d, t)
    assert p["path"] == 0
    assert (p["form"], p["grid"], p["launches"], p["contract_grid"]) == expect
    assert 0 < p["smem_bytes"] <= 227 * 1024


def test_dense_path_plan_and_tuning_validation():
    from attentionalpoolingaction_b200.custom_ops import attn_pool_factory as ops
    dense = ops.attn_pool_plan(ops.make_desc(2, 196, 2048, 393, flags=ops.APA_FLAG_SINGLE_LAYER_ATT |
                                             ops.APA_FLAG_PER_CLASS))
    assert dense["path"] == 1 and dense["form"] == 0
    assert ops.attn_pool_plan(ops.make_desc(2, 196, 512, 393))["path"] == 1          # C outside {1024, 2048}
    assert ops.attn_pool_plan(ops.make_desc(2, 196, 2048, 393), want_topdown=True)["path"] == 1
    for bad in (dict(form=7), dict(stream_shape=3), dict(cluster_shape=4), dict(stream_ctas=-1)):
        with pytest.raises(ops.ApaError) as e:
            ops.attn_pool_plan(ops.make_desc(2, 196, 2048, 393), ops.make_tuning(**bad))
        assert e.value.status == ops.APA_ERR_UNSUPPORTED
    t = ops.make_tuning()
    t.reserved[2] = 1
    with pytest.raises(ops.ApaError):
        ops.attn_pool_plan(ops.make_desc(2, 196, 2048, 393), t)


def test_forward_without_gpu_fails_loudly():
    """No CPU fallback: on a box without a GPU the forward entry returns an error, never a result."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("device present")
    from attentionalpoolingaction_b200.custom_ops import attn_pool_factory as ops
    d = ops.make_desc(1, 4, 1024, 3)
    w = ops.ApaWeights()
    buf = (ctypes.c_float * 64)()
    addr = ctypes.addressof(buf)
    aligned = ctypes.c_void_p((addr + 255) // 256 * 256)
    for i in range(1):
        w.attn_w[i] = w.attn_b[i] = w.topdown_w[i] = w.topdown_b[i] = addr
    # NULL checks come first and need no device
    st = ops.attn_pool_lib.apa_attn_pool_fwd(ctypes.byref(d), None, ctypes.byref(w), None, None, None, None,
                                             None, 0, None, None)
    assert st == ops.APA_ERR_NULL
    # with plausible pointers the call must stop at the workspace / device checks, not compute anything
    st = ops.attn_pool_lib.apa_attn_pool_fwd(ctypes.byref(d), aligned, ctypes.byref(w), aligned, None, None,
                                             None, aligned, 0, None, None)
    assert st != ops.APA_OK


def test_training_entry_points_validate_without_device():
    """Sizes and argument checks of the training path (no kernel is launched for a rejected call)."""
    from attentionalpoolingaction_b200.custom_ops import attn_pool_factory as ops
    lib = ops.attn_pool_lib
    single = ops.APA_FLAG_SINGLE_LAYER_ATT | ops.APA_FLAG_RELU_ATT
    d = ops.make_desc(32, 196, 2048, 393, flags=single)
    saved, ws = lib.apa_attn_pool_train_saved_bytes(d), lib.apa_attn_pool_train_workspace(d)
    # saved: a_0 [32*196] + g [32*2048] + s [32], in 256-byte granules; workspace holds the per-chunk partials
    assert saved == 4 * (32 * 196 + 32 * 2048 + 64) and ws > 2 * 32 * 7 * 2048 * 4
    for flags in (single | ops.APA_FLAG_PER_CLASS, ops.APA_FLAG_RELU_ATT):
        bad = ops.make_desc(32, 196, 2048, 393, flags=flags)
        assert lib.apa_attn_pool_train_saved_bytes(bad) == 0 and lib.apa_attn_pool_train_workspace(bad) == 0
        assert lib.apa_attn_pool_fwd_train(bad, None, None, None, None, None, 0, None, 0, None) == ops.APA_ERR_UNSUPPORTED
    assert lib.apa_attn_pool_fwd_train(d, None, None, None, None, None, 0, None, 0, None) == ops.APA_ERR_NULL
    assert lib.apa_attn_pool_bwd(d, None, None, None, None, None, 0, None, None, 0, None) == ops.APA_ERR_NULL
    t = ops.make_train(1.5, 0)
    assert lib.apa_attn_pool_fwd_train(d, None, None, t, None, None, 0, None, 0, None) == ops.APA_ERR_UNSUPPORTED
    assert lib.apa_dropout_mask(ops.make_train(0.2, 0), 6, None, None) == ops.APA_ERR_NULL
