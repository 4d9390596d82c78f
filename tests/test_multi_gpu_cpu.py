"""World-size-2 gloo tests of the multi-GPU host logic (no GPU needed): buffers are independent, so the batch is sharded
contiguously with fb200_shard_range, every rank factorises its own shard with NO data-path collective, and the final
activations are all-gathered once (north_star).  The per-rank compute is stood in for by the CPU oracle on
tiny buffers -- what is under test is the sharding / seeding / gather plumbing bench.py uses under torchrun."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, total, out_dir):
    sys.path[:0] = [ROOT, os.path.join(ROOT, "flucoma-core_b200")]
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import flucoma_b200 as fb
    from oracle import c_oracle as co
    from tests.golden.make_golden import synth_audio

    begin, count = fb.shard_range(total, world, rank)
    n, win, hop, K, iters = 2048, 128, 32, 3, 8
    F = co.num_frames(n, win, hop)
    acts = np.zeros((count, F, K), np.float32)
    for i in range(count):
        g = begin + i                                    # global buffer index decides both the audio and the NMF seed
        r = co.bufnmf_channel(synth_audio(1000 + g, n), win, win, hop, K, iters, g)
        acts[i] = r["acts"]
    # equal-count all-gather needs padding of the short shards: pad to the largest shard
    maxc = max(fb.shard_range(total, world, r_)[1] for r_ in range(world))
    padded = torch.zeros((maxc, F, K), dtype=torch.float32)
    padded[:count] = torch.from_numpy(acts)
    gathered = torch.zeros((world * maxc, F, K), dtype=torch.float32)
    dist.all_gather_into_tensor(gathered, padded)
    # timing convention of bench.py: MAX over ranks
    t = torch.tensor([float(rank + 1)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        full = np.concatenate([gathered[r_ * maxc:r_ * maxc + fb.shard_range(total, world, r_)[1]].numpy() for r_ in range(world)])
        np.save(os.path.join(out_dir, "gathered.npy"), full)
        np.save(os.path.join(out_dir, "tmax.npy"), t.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("total", [5, 4])
def test_sharded_bufnmf_allgather_gloo(tmp_path, total):
    world = 2
    port = 29500 + (os.getpid() % 500) + total
    mp.spawn(_worker, args=(world, port, total, str(tmp_path)), nprocs=world, join=True)
    got = np.load(tmp_path / "gathered.npy")
    assert np.load(tmp_path / "tmax.npy")[0] == world
    sys.path[:0] = [ROOT]
    from oracle import c_oracle as co
    from tests.golden.make_golden import synth_audio
    assert got.shape[0] == total
    for g in range(total):  # the gathered activations equal a single-process run over the whole batch, in order
        r = co.bufnmf_channel(synth_audio(1000 + g, 2048), 128, 128, 32, 3, 8, g)
        assert np.array_equal(got[g], r["acts"])
