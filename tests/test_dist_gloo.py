"""CPU, world_size 2, gloo: the N>1 host path — contiguous batch shard, per-rank head, ONE all-gather of
the logits — reproduces the single-process result.  The per-rank head here is the oracle's PyTorch-CPU
port (the CUDA op needs a GPU); what is under test is attentionalpoolingaction_b200.dist."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, num_videos, frames, out_dir):
    sys.path.insert(0, ROOT)
    from attentionalpoolingaction_b200 import dist as apa_dist
    from oracle import attn_pool_oracle as orc
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    K, C, HW = 9, 16, 6
    fl = dict(orc.DEFAULT_FLAGS, SINGLE_LAYER_ATT=True, RANK=2)
    params = {k: torch.from_numpy(v) for k, v in orc.make_params(K, C, flags=fl, seed=5).items()}
    X = torch.from_numpy(orc.make_features(num_videos * frames, HW, C, seed=5))
    head = lambda x: orc.attn_pool_head_torch_cpu(x, params, fl, K, frames_per_video=frames)
    sharded = apa_dist.ShardedHead(head, num_videos)
    local = apa_dist.shard_features(X, frames, rank, world)
    got = sharded(local)
    full = head(X)
    assert got.shape == full.shape
    assert torch.allclose(got, full, rtol=1e-6, atol=1e-7)
    np.save(os.path.join(out_dir, "rank%d.npy" % rank), got.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("num_videos,frames", [(8, 1), (6, 3), (5, 2)])
def test_sharded_head_world2_gloo(tmp_path, num_videos, frames):
    port = 29500 + (os.getpid() + num_videos * 7 + frames) % 2000
    mp.spawn(_worker, args=(2, port, num_videos, frames, str(tmp_path)), nprocs=2, join=True)
    a = np.load(tmp_path / "rank0.npy")
    b = np.load(tmp_path / "rank1.npy")
    assert np.array_equal(a, b)          # every rank ends with the same gathered logits
