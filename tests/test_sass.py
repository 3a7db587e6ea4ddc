"""The built library contains the Blackwell instructions the design claims (B200_PROFILING.md's SASS mnemonics):
tcgen05 MMAs and TMEM loads in the dense GEMM, bulk (TMA-engine) copies + mbarrier waits in the streaming kernels,
cluster barriers in the cluster form, cp.async in the training kernels.  Needs cuobjdump only (no GPU)."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "attentionalpoolingaction_b200", "custom_ops", "libapa_attn_pool.so")


@pytest.fixture(scope="module")
def sass():
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(exe):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([exe, "-sass", LIB], capture_output=True, text=True, check=True).stdout
    funcs, cur = {}, None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            funcs[cur] = []
        elif cur is not None:
            funcs[cur].append(line)
    return {k: "\n".join(v) for k, v in funcs.items()}


def _any(sass, name_part, mnemonic):
    return any(name_part in f and mnemonic in body for f, body in sass.items())


def test_library_is_sm_100a_only():
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(exe):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([exe, "-lelf", LIB], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_blackwell_instructions_are_where_the_design_says(sass):
    assert _any(sass, "gemm_tf32x3_kernel", "UTCHMMA")          # tcgen05.mma, kind::tf32
    assert _any(sass, "gemm_tf32x3_kernel", "LDTM")             # tcgen05.ld from TMEM
    assert _any(sass, "apa_flow_stream_kernel", "UBLKCP")       # cp.async.bulk global -> shared
    assert _any(sass, "apa_flow_stream_kernel", "SYNCS")        # mbarrier arrive / try_wait
    assert _any(sass, "apa_pool_local_kernel", "UCGABAR")       # cluster barrier (cluster form)
    assert _any(sass, "bwd_small_kernel", "LDGSTS")             # cp.async tile loads (training)
    assert not _any(sass, "apa_flow_stream_kernel", "HMMA")     # the HBM-bound path stays off the tensor cores
    # every kernel family is present
    for name in ("apa_flow_reduce_kernel", "apa_flow_contract_kernel", "apa_flow_finish_kernel", "train_fused_fwd_kernel",
                 "bwd_fused_kernel", "apa_class_ap_kernel", "apa_eval_tail_kernel", "xent_rows_kernel", "pose_rows_kernel"):
        assert any(name in f for f in sass), name
