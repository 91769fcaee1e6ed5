"""N > 1 host logic on CPU: world_size-2 and -3 process groups over gloo run the same exchange code the
NCCL path uses (frame sharding, owner hash, all-to-all-v of cell buckets) on host buffers."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from online_3d_reconstruction_b200 import abi, exchange


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _make_cells(rank, n, seed=0):
    rng = np.random.default_rng(seed + rank)
    c = np.zeros(n, dtype=abi.CELL)
    i = rng.integers(-300, 300, n) + (1 << 20)
    j = rng.integers(-200, 500, n) + (1 << 20)
    c["key"] = (np.uint64(1 << 20) << np.uint64(42)) | (j.astype(np.uint64) << np.uint64(21)) | i.astype(np.uint64)
    c["sx"], c["sy"], c["sz"] = rng.normal(size=(3, n)).astype(np.float32)
    c["n"] = rng.integers(1, 50, n)
    c["sr"], c["sg"], c["sb"] = rng.integers(0, 255 * 50, (3, n))
    return c


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        cells = _make_cells(rank, 4000 + 500 * rank)
        own = exchange.owner_of(cells["key"], world)
        order = np.argsort(own, kind="stable")            # what o3r_exchange_pack does on the GPU
        counts = np.bincount(own, minlength=world)
        send = torch.from_numpy(cells[order].view(np.uint8).copy())
        recv, n = exchange.exchange_cells(send, counts)
        got = recv[:n * exchange.CELL_BYTES].numpy().view(abi.CELL)
        np.save(os.path.join(out_dir, f"recv{rank}.npy"), got)
        # an empty send must work too
        recv2, n2 = exchange.exchange_cells(torch.zeros(0, dtype=torch.uint8), [0] * world)
        assert n2 == 0
        dist.barrier()
    finally:
        dist.destroy_process_group()


def _header_worker(rank, world, port, out_dir):
    """exchange_by_header: the header (counts + cell range) travels as the device would publish it; rank 1 sends nothing."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        cells = _make_cells(rank, 0 if rank == 1 else 3000 + 700 * rank, seed=5)
        own = exchange.owner_of(cells["key"], world)
        order = np.argsort(own, kind="stable")
        info = np.zeros(world + 8, dtype=np.int32)
        info[:world] = np.bincount(own, minlength=world)
        if len(cells):
            i = (cells["key"] & np.uint64(0x1fffff)).astype(np.int64) - (1 << 20)
            j = ((cells["key"] >> np.uint64(21)) & np.uint64(0x1fffff)).astype(np.int64) - (1 << 20)
            info[world:world + 6] = [i.min(), j.min(), 0, i.max(), j.max(), 0]
        else:
            info[world:world + 6] = [0x7fffffff] * 3 + [-0x80000000] * 3
        info[world + 6] = len(cells)
        send = torch.from_numpy(cells[order].view(np.uint8).copy())
        recv, n, bb, n_sent = exchange.exchange_by_header(send, torch.from_numpy(info))
        assert n_sent == len(cells)
        np.save(os.path.join(out_dir, f"hrecv{rank}.npy"), recv[:n * exchange.CELL_BYTES].numpy().view(abi.CELL))
        np.save(os.path.join(out_dir, f"hbb{rank}.npy"), np.array(bb))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_header_driven_exchange_one_round_trip(world, tmp_path):
    """The exchange the multi-GPU bench uses (exchange.exchange_cycle minus the GPU pack / merge): sizes come from the
    all-gathered headers, every rank learns the union of the cell ranges, an empty sender is handled."""
    port = _free_port()
    mp.spawn(_header_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    sent = [_make_cells(r, 0 if r == 1 else 3000 + 700 * r, seed=5) for r in range(world)]
    allc = np.concatenate(sent)
    i = (allc["key"] & np.uint64(0x1fffff)).astype(np.int64) - (1 << 20)
    j = ((allc["key"] >> np.uint64(21)) & np.uint64(0x1fffff)).astype(np.int64) - (1 << 20)
    for r in range(world):
        got = np.load(tmp_path / f"hrecv{r}.npy")
        exp = np.concatenate([s[exchange.owner_of(s["key"], world) == r] for s in sent])
        assert np.array_equal(got, exp)
        assert np.load(tmp_path / f"hbb{r}.npy").tolist() == [i.min(), j.min(), 0, i.max(), j.max(), 0]


@pytest.mark.parametrize("world", [2, 3])
def test_all_to_all_v_delivers_every_cell_to_its_owner(world, tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    sent = [_make_cells(r, 4000 + 500 * r) for r in range(world)]
    allc = np.concatenate(sent)
    for r in range(world):
        got = np.load(tmp_path / f"recv{r}.npy")
        assert np.all(exchange.owner_of(got["key"], world) == r)
        # exactly the cells this rank owns, grouped by source rank in rank order, each bucket in send order
        exp = np.concatenate([s[exchange.owner_of(s["key"], world) == r] for s in sent])
        assert np.array_equal(got, exp)
    # merging the shards equals merging everything in one place (keys / integer sums exactly)
    whole = exchange.merge_cells_host(allc)
    parts = np.concatenate([exchange.merge_cells_host(np.load(tmp_path / f"recv{r}.npy")) for r in range(world)])
    parts = parts[np.argsort(parts["key"])]
    assert np.array_equal(parts["key"], whole["key"])
    for f in ("n", "sr", "sg", "sb"):
        assert np.array_equal(parts[f], whole[f])
    for f in ("sx", "sy", "sz"):
        assert np.allclose(parts[f], whole[f], rtol=1e-5, atol=1e-6)


def test_frame_sharding_covers_the_cycle_once():
    for world in (1, 2, 4, 8):
        seen = sorted(f for r in range(world) for f in exchange.shard_frames(50, r, world))
        assert seen == list(range(50))
        assert max(len(exchange.shard_frames(50, r, world)) for r in range(world)) - \
            min(len(exchange.shard_frames(50, r, world)) for r in range(world)) <= 1


def test_owner_hash_is_uniform_and_stable():
    keys = _make_cells(0, 100000)["key"]
    for world in (2, 4, 8):
        cnt = np.bincount(exchange.owner_of(keys, world), minlength=world)
        assert cnt.min() > 0.9 * len(keys) / world
    # known answers of the splitmix64 finaliser (the device uses the same constants)
    assert int(exchange.hash64(np.uint64(1))) == 0x5692161D100B05E5
    assert int(exchange.hash64(np.uint64(0))) == 0
