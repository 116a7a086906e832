"""N>1 host logic on CPU: two processes over gloo. Each rank renders only its own bands of a cornell frame
(with the CPU oracle standing in for the GPU renderer -- the sharding code is device-agnostic torch), the
bands are gathered with the product's gather_frame, and every rank must hold the frame a single process
renders. Also checks ray-stream ranges and band plans with ragged sizes."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle_lib as ol
import scenes
from conftest import _reference_scene, load_product


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, W, H, band_rows, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    prod = load_product()
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        tris, nodes, mats = _reference_scene(scenes.CORNELL)
        plan = prod.sharding.BandPlan(W, H, world, band_rows)
        frame = np.zeros((W * H, 4), dtype=np.float32)
        for fc in (1, 2):
            for lo, hi in plan.gid_ranges(rank):
                ol.oracle_render(tris, nodes, mats, frame, W, H, fc, 3, gid0=lo, gid1=hi, threads=1)
        full = prod.sharding.gather_frame(plan, torch.from_numpy(frame), rank)
        np.save(os.path.join(out_dir, "frame_%d.npy" % rank), full.numpy())
        # ray stream: the shards tile the stream exactly
        lo, hi = prod.sharding.stream_range(1000003, world, rank)
        t = torch.tensor([hi - lo], dtype=torch.int64)
        dist.all_reduce(t)
        assert int(t.item()) == 1000003
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("W,H,band_rows", [(64, 48, 8), (50, 37, 5)])
def test_two_rank_frame_gather(tmp_path, W, H, band_rows):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, W, H, band_rows, str(tmp_path)), nprocs=2, join=True)
    tris, nodes, mats = _reference_scene(scenes.CORNELL)
    want = np.zeros((W * H, 4), dtype=np.float32)
    for fc in (1, 2):
        ol.oracle_render(tris, nodes, mats, want, W, H, fc, 3)
    for r in range(2):
        got = np.load(os.path.join(str(tmp_path), "frame_%d.npy" % r))
        assert got.shape == want.shape
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), "rank %d frame differs" % r


def test_band_plan_covers_every_pixel_once():
    prod = load_product()
    for W, H, world, rows in ((7, 5, 3, 2), (1920, 1080, 8, 8), (3840, 2160, 8, 16), (5, 1, 4, 8)):
        plan = prod.sharding.BandPlan(W, H, world, rows)
        seen = np.zeros(W * H, dtype=np.int32)
        for r in range(world):
            for lo, hi in plan.gid_ranges(r):
                seen[lo:hi] += 1
        assert (seen == 1).all()
        # the strided description handed to b2rt_execute_bands names exactly the same pixels
        seen[:] = 0
        for r in range(world):
            gid0, band, stride, n_full, tail = plan.band_launch(r)
            for k in range(n_full):
                seen[gid0 + k * stride:gid0 + k * stride + band] += 1
            if tail:
                seen[tail[0]:tail[1]] += 1
        assert (seen == 1).all()
        # pack/unpack round trip without a process group
        frames = [torch.zeros((W * H, 4)) for _ in range(world)]
        ref = torch.arange(W * H * 4, dtype=torch.float32).view(-1, 4)
        for r in range(world):
            for lo, hi in plan.gid_ranges(r):
                frames[r][lo:hi] = ref[lo:hi]
        gathered = torch.cat([plan.pack(frames[r], r) for r in range(world)])
        assert torch.equal(plan.unpack(gathered), ref)


# ---- the partition of the C ABI (b2rt_shard_bands = what b2rt_execute_shard and device-group handles render) -----------
def _abi_worker(rank, world, port, W, H, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    prod = load_product()
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        tris, nodes, mats = _reference_scene(scenes.CORNELL)
        gid0, band, stride, n_full, t0, t1 = prod.capi.shard_bands(W, H, rank, world)
        ranges = [(gid0 + k * stride, gid0 + k * stride + band) for k in range(n_full)] + ([(t0, t1)] if t1 > t0 else [])
        # root-gather like the product: every rank's finished pixels land in rank 0's image, in place (bands are contiguous)
        frame = np.zeros((W * H, 4), dtype=np.float32)
        for fc in (1, 2):
            for lo, hi in ranges:
                ol.oracle_render(tris, nodes, mats, frame, W, H, fc, 3, gid0=lo, gid1=hi, threads=1)
        mine = torch.from_numpy(frame)
        if rank == 0:
            for src in range(1, world):
                g0, b, s, n, a0, a1 = prod.capi.shard_bands(W, H, src, world)
                for lo, hi in [(g0 + k * s, g0 + k * s + b) for k in range(n)] + ([(a0, a1)] if a1 > a0 else []):
                    dist.recv(mine[lo:hi], src=src)
            np.save(os.path.join(out_dir, "abi_frame.npy"), mine.numpy())
        else:
            for lo, hi in ranges:
                dist.send(mine[lo:hi].contiguous(), dst=0)
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("W,H", [(64, 48), (50, 37)])
def test_two_rank_root_gather_with_the_abi_partition(tmp_path, W, H):
    """world_size 2 over gloo: the bands b2rt_shard_bands deals out, rendered per rank (the CPU oracle standing in for the
    kernels) and received band by band into place on rank 0 -- the send/recv gather of b2rt_execute_shard -- give the frame
    one process renders."""
    port = _free_port()
    mp.spawn(_abi_worker, args=(2, port, W, H, str(tmp_path)), nprocs=2, join=True)
    tris, nodes, mats = _reference_scene(scenes.CORNELL)
    want = np.zeros((W * H, 4), dtype=np.float32)
    for fc in (1, 2):
        ol.oracle_render(tris, nodes, mats, want, W, H, fc, 3)
    got = np.load(os.path.join(str(tmp_path), "abi_frame.npy"))
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


def test_abi_partition_covers_every_pixel_once():
    prod = load_product()
    for W, H, world in ((7, 5, 3), (1920, 1080, 8), (3840, 2160, 8), (5, 1, 4), (203, 131, 3), (64, 8, 2), (33, 17, 5)):
        seen = np.zeros(W * H, dtype=np.int32)
        for r in range(world):
            g0, b, s, n, t0, t1 = prod.capi.shard_bands(W, H, r, world)
            assert b == 8 * W and s == world * b
            for k in range(n):
                seen[g0 + k * s:g0 + k * s + b] += 1
            if t1 > t0:
                seen[t0:t1] += 1
        assert (seen == 1).all(), (W, H, world)
