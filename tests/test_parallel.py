"""CPU, world_size 2 over gloo: batch sharding + the single all-gather of outputs (host logic of the
multi-GPU path; the per-rank forward is a stub here because the product path needs a GPU)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from femasr_b200.parallel import (all_gather_outputs, shard_counts, shard_range, sharded_forward, sharded_test_tile,
                                  tile_classes)


def test_shard_ranges_partition_the_batch():
    for n in (0, 1, 2, 7, 32, 255, 256):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = shard_counts(n, world)
            assert sum(sizes) == n and max(sizes) - min(sizes) <= 1


def _stub_sr(x):                       # stands in for FeMaSRNet.test: x4 nearest "SR" with a per-image tag
    return x.repeat_interleave(4, 2).repeat_interleave(4, 3) * 2.0 + 1.0


def _worker(rank, world, port, n, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(0)
        x = torch.rand(n, 3, 8, 8, generator=g)
        out = sharded_forward(_stub_sr, x, rank, world, out_shape_fn=lambda s: (s[0], 3, s[2] * 4, s[3] * 4))
        ok = torch.equal(out, _stub_sr(x))
        # equal shards take the all_gather_into_tensor fast path
        a, b = shard_range(n, rank, world)
        out2 = all_gather_outputs(_stub_sr(x[a:b]), shard_counts(n, world)) if n % world == 0 else out
        ok = ok and torch.equal(out2, _stub_sr(x))
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n", [8, 5, 1])
def test_world2_gloo_gather(n):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, port, n, ret), nprocs=2, join=True)
    assert ret[0] and ret[1]


def _stub_tile_sr(t):                  # a per-pixel "SR" (x4 nearest) so that tiling + paste must reproduce the whole image
    return t.repeat_interleave(4, 2).repeat_interleave(4, 3) * 3.0 - 0.5


def _tile_worker(rank, world, port, shape, ts, tp, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        x = torch.rand(shape, generator=torch.Generator().manual_seed(1))
        out = sharded_test_tile(_stub_tile_sr, x, 4, ts, tp, rank, world, max_batch=4)
        ret[rank] = bool(torch.equal(out, _stub_tile_sr(x)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("shape,ts,tp", [((1, 3, 72, 56), 32, 8), ((2, 3, 40, 100), 24, 4), ((1, 3, 20, 20), 32, 8)])
def test_world2_gloo_sharded_test_tile(shape, ts, tp):
    """Tile-list sharding: every tile runs on exactly one rank, both ranks end with the whole image."""
    classes = tile_classes(shape[2], shape[3], ts, tp)
    ntiles = sum(len(v) for v in classes.values())
    assert ntiles == -(-shape[2] // ts) * -(-shape[3] // ts)
    for tiles in classes.values():            # round-robin deal: per class the ranks differ by at most one tile
        assert abs(len(tiles[0::2]) - len(tiles[1::2])) <= 1
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ret = mp.Manager().dict()
    mp.spawn(_tile_worker, args=(2, port, shape, ts, tp, ret), nprocs=2, join=True)
    assert ret[0] and ret[1]


def test_sharded_test_tile_matches_oracle_test_tile():
    """Single process, the CPU oracle as the per-tile runner: the sharded/batched schedule equals the reference's
    one-tile-at-a-time test_tile (femasr_arch.py:387-447)."""
    from femasr_b200.spec import random_state_dict
    from oracle import femasr_oracle as O
    sd = random_state_dict(4, 256, seed=41, init="perturbed")
    x = torch.rand((1, 3, 40, 24), generator=torch.Generator().manual_seed(42))
    with torch.no_grad():
        want = O.test_tile(sd, x, 4, 16, 8)
        got = sharded_test_tile(lambda t: O.test(sd, t, 4), x, 4, 16, 8, rank=0, world=1)
    assert (got - want).abs().max().item() <= 1e-5


def _peer_worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from femasr_b200.parallel import PeerGather
        shape = (3, 3, 16, 24)
        pg = PeerGather(shape, torch.float32, dev, rank, world, nbuf=2)          # includes the collective self-check
        ok = True
        for k in (0, 1, 0):
            shard = torch.full(shape, float(10 * k + rank + 1), device=dev)
            pg.push(k, shard).synchronize()
            dist.barrier()
            want = torch.cat([torch.full(shape, float(10 * k + r + 1)) for r in range(world)], 0)
            ok = ok and torch.equal(pg.full[k].cpu(), want)
            dist.barrier()
        pg.close()
        dist.barrier()
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.gpu
def test_world2_peer_gather_matches_all_gather():
    """PeerGather (copy-engine peer writes over CUDA IPC) fills every rank's buffer like all_gather_into_tensor.  Needs
    two GPUs with peer access; skipped on the single-GPU test box (bench.py --gpus 2 exercises the same path there)."""
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2 or os.environ.get("FEMASR_TEST_MULTI_GPU") != "1":
        pytest.skip("needs two GPUs and FEMASR_TEST_MULTI_GPU=1 (spawns two NCCL ranks)")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ret = mp.Manager().dict()
    mp.spawn(_peer_worker, args=(2, port, ret), nprocs=2, join=True)
    assert ret[0] and ret[1]
