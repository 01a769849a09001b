"""world_size-2 gloo tests (CPU) of the data-parallel host logic: sharding, bucket ranges, bucketed all-reduce + 1/N scaling
== the oracle's data-parallel semantics (N replicas, per-replica BatchNorm, gradients averaged)."""
import os
import socket
from collections import OrderedDict

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import late_fusion_oracle as O


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _offsets(state, align=64):
    off, offsets = 0, OrderedDict()
    for k, v in state.items():
        if O.is_parameter(k):
            offsets[k] = off
            off = (off + v.numel() + align - 1) // align * align
    return offsets, off


def _worker(rank, world, port, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    from mml_b200 import dist as mdist

    r, lr, w = mdist.init_from_env("gloo")
    assert (r, w) == (rank, world)
    torch.manual_seed(0)
    state = O.init_avmnist_state()
    B = 4
    d = O.synthetic_batch(B, 99, (32, 32))
    lo, hi = mdist.shard_batch(B, rank, world)
    A = O.apply_missing_mask(d["audio"], d["audio_mask"])[lo:hi]
    out = O.train_step(OrderedDict((k, v.clone()) for k, v in state.items()), {}, A, d["image"][lo:hi], d["labels"][lo:hi], d["dropout_mask"][lo:hi], 0.5,
                       apply_update=False)
    offsets, total = _offsets(state)
    G = torch.zeros(total)
    for k, g in out["grads"].items():
        G[offsets[k]:offsets[k] + g.numel()] = g.reshape(-1)
    buckets = mdist.bucket_ranges(offsets, total)
    assert buckets[0][1] == total and buckets[1][0] == 0 and buckets[0][0] == buckets[1][1]  # [image+head], [audio]: a partition
    assert buckets[0][0] == offsets["image_encoder.conv1.weight"]
    for a, b in buckets:
        dist.all_reduce(G[a:b], op=dist.ReduceOp.SUM)
    G *= 1.0 / world  # the Adam kernel's grad_scale
    if rank == 0:
        torch.save({"G": G, "offsets": offsets}, tmp)
    dist.barrier()
    dist.destroy_process_group()


def test_bucketed_allreduce_equals_oracle_data_parallel(tmp_path):
    world, port, tmp = 2, _free_port(), str(tmp_path / "g.pt")
    mp.spawn(_worker, args=(world, port, tmp), nprocs=world, join=True)
    got = torch.load(tmp)
    torch.manual_seed(0)
    state = O.init_avmnist_state()
    B = 4
    d = O.synthetic_batch(B, 99, (32, 32))
    A = O.apply_missing_mask(d["audio"], d["audio_mask"])
    shards = [(A[:2], d["image"][:2], d["labels"][:2], d["dropout_mask"][:2]), (A[2:], d["image"][2:], d["labels"][2:], d["dropout_mask"][2:])]
    ref = O.data_parallel_grads(state, shards)
    for k, g in ref.items():
        o = got["offsets"][k]
        assert torch.allclose(got["G"][o:o + g.numel()], g.reshape(-1), rtol=1e-5, atol=1e-7), k


def test_shard_batch_and_ranges():
    from mml_b200 import dist as mdist

    assert mdist.shard_batch(2048, 3, 8) == (768, 1024)
    with pytest.raises(ValueError):
        mdist.shard_batch(10, 0, 4)
    assert mdist.bucket_ranges({"net.0.weight": 0}, 128) == [(0, 128)]
