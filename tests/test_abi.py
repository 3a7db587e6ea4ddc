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
int main(void){ printf("%zu %zu %zu %d %d\n", sizeof(ApaDesc), sizeof(ApaWeights), sizeof(ApaTuning),
                       APA_ABI_VERSION, APA_MAX_RANK); return 0; }'''
    exe = "/tmp/apa_sizes_%d" % os.getpid()
    subprocess.run(["gcc", "-x", "c", "-I", os.path.join(ROOT, "include"), "-o", exe, "-"], input=prog,
                   text=True, check=True)
    d, w, t, ver, mr = map(int, subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split())
    os.remove(exe)
    assert (d, w, t) == (ctypes.sizeof(ops.ApaDesc), ctypes.sizeof(ops.ApaWeights), ctypes.sizeof(ops.ApaTuning))
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


@pytest.mark.parametrize("B,frames,K,expect_cluster", [(32, 1, 393, 16), (30, 1, 393, 4), (15, 1, 393, 8),
                                                      (130, 1, 393, 1), (2, 1, 393, 16)])
def test_plan_without_device(B, frames, K, expect_cluster):
    """Plan introspection falls back to the nominal 148-SM B200 when no device is visible."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("device present: plan comes from the occupancy query")
    from attentionalpoolingaction_b200.custom_ops import attn_pool_factory as ops
    plan = ops.attn_pool_plan(ops.make_desc(B, 196, 2048, K, frames=frames))
    assert plan["path"] == 0 and plan["fused"] == 1 and plan["launches"] == 1
    assert plan["cluster_size"] == expect_cluster
    assert plan["cluster_local"] == (1 if (B % 4 == 0 and frames == 1) else 0)
    assert plan["grid"] <= 148 and plan["grid"] % plan["cluster_size"] == 0
    assert plan["smem_bytes"] <= 227 * 1024
    dense = ops.attn_pool_plan(ops.make_desc(2, 196, 2048, 393, flags=ops.APA_FLAG_SINGLE_LAYER_ATT |
                                             ops.APA_FLAG_PER_CLASS))
    assert dense["path"] == 1
    # several frames per video (HICO 10-crop): the STREAM + CONTRACT pipeline, two launches
    pipe = ops.attn_pool_plan(ops.make_desc(320, 196, 2048, 600, frames=10))
    assert pipe["path"] == 0 and pipe["cluster_local"] == 2 and pipe["launches"] == 2 and pipe["fused"] == 0
    assert pipe["cluster_size"] == 1 and pipe["grid"] == 320 and pipe["contract_grid"] == 16 * 11
    assert pipe["smem_bytes"] <= 76800
    t = ops.make_tuning()
    t.form = 1      # the grid-barrier kernel stays selectable
    assert ops.attn_pool_plan(ops.make_desc(320, 196, 2048, 600, frames=10), t)["cluster_local"] == 0



@pytest.mark.parametrize("kw,expect", [
    # (descriptor, tuning code) -> (cluster_local, cluster_size, grid, launches, smem_bytes)
    (dict(B=32), (1, 16, 64, 1, 73024)),                 # config 2: cluster-local throughput shape (8 images / cluster)
    (dict(B=36), (1, 16, 144, 1, 66752)),                # B % 8 != 0: 4 images / cluster, 3 CTAs / SM
    (dict(B=32, rank=2), (1, 16, 128, 1, 122048)),       # rank chain: 2-CTA/SM shape
    (dict(B=32, K=600), (1, 16, 128, 1, 67904)),         # K = 600: the 8-image shape no longer fits 3 per SM
    (dict(B=30), (0, 4, 120, 1, 107472)),                # odd batch: grid-barrier kernel
    (dict(B=32, C=1024), (0, 4, 128, 1, 103376)),
    (dict(B=32, K=800), (0, 4, 128, 1, 107472)),         # K > 768: not cluster-local
    (dict(B=320, F=10, K=600), (2, 1, 320, 2, 74048)),   # several frames per video: STREAM + CONTRACT pipeline
    (dict(B=32, code=4), (1, 16, 128, 1, 113856)),       # forced shapes / forms
    (dict(B=32, code=5), (1, 16, 128, 1, 66752)),
    (dict(B=32, code=7), (2, 4, 128, 2, 74048)),
    (dict(B=32, code=1), (0, 4, 128, 1, 107472)),
])
def test_dispatch_policy_without_device(kw, expect):
    """Which form of the single-map head a descriptor gets, pinned on the nominal B200."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("device present: plan comes from the occupancy query")
    from attentionalpoolingaction_b200.custom_ops import attn_pool_factory as ops
    kw = dict(kw)
    t = ops.make_tuning()
    t.form = kw.pop("code", 0)
    d = ops.make_desc(kw.get("B"), 196, kw.get("C", 2048), kw.get("K", 393), frames=kw.get("F", 1), rank=kw.get("rank", 1))
    p = ops.attn_pool_plan(d, t)
    assert (p["cluster_local"], p["cluster_size"], p["grid"], p["launches"], p["smem_bytes"]) == expect
    assert p["path"] == 0


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
