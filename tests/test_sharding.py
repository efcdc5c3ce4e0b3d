"""CPU tests of the multi-GPU host logic: (clip x head) partitioning and the readout gather,
exercised with world_size-2 gloo processes."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_chain_partition_covers_everything():
    from gdkvm_b200.sharding import chain_partition
    for B, H, world in [(64, 8, 8), (7, 8, 2), (512, 8, 4), (2, 8, 8), (1, 8, 4), (5, 3, 1)]:
        seen = torch.zeros(B, H, dtype=torch.int32)
        for r in range(world):
            bs, hs = chain_partition(B, H, world, r)
            seen[bs, hs] += 1
        assert torch.all(seen == 1), (B, H, world)
    with pytest.raises(ValueError):
        chain_partition(3, 8, 8, 0)


def test_packed_partition_whole_clips_balanced_tokens():
    from gdkvm_b200.sharding import packed_partition
    from oracle.gdr_ref import gdr_recurrent_varlen_ref, make_inputs
    lens = [50, 7, 120, 0, 33, 64, 90, 5]
    cu = [0]
    for x in lens:
        cu.append(cu[-1] + x)
    for world in (1, 2, 3, 4, 8):
        clips_seen, toks = [], 0
        for r in range(world):
            cs, ts = packed_partition(cu, world, r)
            clips_seen += list(range(cs.start, cs.stop))
            assert ts.start == cu[cs.start] and ts.stop == cu[cs.stop]
            toks += ts.stop - ts.start
        assert clips_seen == list(range(len(lens))) and toks == cu[-1]
    cs0, ts0 = packed_partition(cu, 2, 0)
    assert abs((ts0.stop - ts0.start) - cu[-1] / 2) <= max(lens) / 2
    # the shards reproduce the unsharded result (the checker stands in for the GPU op)
    H, K, V = 2, 16, 8
    q, k, v, g, beta, _ = make_inputs(1, cu[-1], H, K, V, seed=22)
    S0 = 0.1 * torch.randn(len(lens), H, K, V, generator=torch.Generator().manual_seed(23))
    o_ref, s_ref = gdr_recurrent_varlen_ref(q, k, v, g, beta, cu, None, S0)
    for r in range(3):
        cs, ts = packed_partition(cu, 3, r)
        cu_r = [c - cu[cs.start] for c in cu[cs.start:cs.stop + 1]]
        o_r, s_r = gdr_recurrent_varlen_ref(q[:, ts], k[:, ts], v[:, ts], g[:, ts], beta[:, ts], cu_r, None, S0[cs])
        assert torch.equal(o_r, o_ref[:, ts]) and torch.equal(s_r, s_ref[cs])
    with pytest.raises(ValueError):
        packed_partition(cu, 2, 2)


def _worker(rank, world, port, tmp):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from gdkvm_b200.sharding import chain_partition, gather_readout
    from oracle.gdr_ref import gdr_recurrent_ref, make_inputs   # the checker stands in for the GPU op here
    B, T, H, K, V = 4, 30, 2, 16, 8
    q, k, v, g, beta, S0 = make_inputs(B, T, H, K, V, seed=21)
    bs, hs = chain_partition(B, H, world, rank)
    o_loc, _ = gdr_recurrent_ref(q[bs], k[bs], v[bs], g[bs], beta[bs], None, S0[bs])
    o_all = gather_readout(o_loc)
    o_ref, _ = gdr_recurrent_ref(q, k, v, g, beta, None, S0)
    ok = torch.equal(o_all, o_ref)
    torch.save(torch.tensor(ok), os.path.join(tmp, f"ok{rank}.pt"))
    dist.destroy_process_group()


def test_two_rank_gloo_shard_and_gather(tmp_path):
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    for r in range(2):
        assert bool(torch.load(tmp_path / f"ok{r}.pt"))
