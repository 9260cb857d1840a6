"""World-size-2 tests of the multi-GPU host path on CPU (gloo): sharding, per-bucket batching, the metric gather and
the max-over-ranks timing of `sweep.py`, with a stand-in enhancer (the CUDA path itself is covered by the gpu tests).
The data path has no collective: every utterance must be processed exactly once, by exactly one rank, and its result
must not depend on the number of ranks."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from snr_aligned_diffse_b200.shard import synthetic_lengths, tpad_of
from snr_aligned_diffse_b200.sweep import enhance_sweep, gather_metrics, pack_batch


def _waves(n=40, seed=3):
    L = synthetic_lengths(n, seed=seed) // 8          # short utterances: the test is about the plumbing
    g = torch.Generator().manual_seed(seed)
    return [torch.randn(int(l), generator=g) * 0.1 for l in L]


def _fake_enhance(y, lens):
    # deterministic, length-aware stand-in: scales every utterance by a factor that depends on its own length only
    return y * (1.0 + lens[:, None].to(y.dtype) * 1e-6)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_path, target_frames=None):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    waves = _waves()
    local = enhance_sweep(_fake_enhance, waves, rank=rank, world=world, max_batch=4, target_frames=target_frames)
    allm = gather_metrics(local, world)
    if rank == 0:
        torch.save(dict(all=allm, local_ids=local["ids"]), out_path)
    else:
        torch.save(dict(local_ids=local["ids"]), out_path + f".{rank}")
    dist.barrier()
    dist.destroy_process_group()


def test_pack_batch_shapes():
    waves = _waves(10)
    L = [w.numel() for w in waves]
    tp = tpad_of(L[0])
    idx = [i for i in range(10) if tpad_of(L[i]) == tp]
    y, lens = pack_batch(waves, idx, tp)
    assert y.shape == (len(idx), 128 * tp - 1) and lens.tolist() == [L[i] for i in idx]
    for r, i in enumerate(idx):
        assert torch.equal(y[r, :L[i]], waves[i]) and torch.count_nonzero(y[r, L[i]:]) == 0


import pytest  # noqa: E402


@pytest.mark.parametrize("target_frames", [None, 640], ids=["max-batch", "frame-budget"])
def test_sweep_world2_gloo_matches_single_rank(tmp_path, target_frames):
    waves = _waves()
    single = gather_metrics(enhance_sweep(_fake_enhance, waves, rank=0, world=1, max_batch=4), 1)
    assert single["ids"] == list(range(len(waves)))
    out = str(tmp_path / "m.pt")
    mp.spawn(_worker, args=(2, _free_port(), out, target_frames), nprocs=2, join=True)
    got = torch.load(out)
    other = torch.load(out + ".1")
    # a partition of the utterances over the two ranks
    assert sorted(got["local_ids"] + other["local_ids"]) == list(range(len(waves)))
    assert set(got["local_ids"]).isdisjoint(other["local_ids"])
    # gathered table identical to the single-rank run (per-utterance results do not depend on the sharding)
    assert got["all"]["ids"] == single["ids"] and got["all"]["samples"] == single["samples"]
    assert np.allclose(got["all"]["checksum"], single["checksum"], rtol=0, atol=1e-9)
    assert got["all"]["job_seconds"] > 0
