"""N>1 host path on CPU: world_size-2 gloo run of the stream sharding and of the throughput aggregation that
bench.py uses (units summed over ranks, time = max over ranks). Each rank replays its own shard of streams
through the CPU oracle, so the test also shows that sharded results equal the unsharded run stream by stream."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
import workload  # noqa: E402


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_total, n_scans, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tfe_slam_ucl_b200 import sharding          # host logic only: libhsgpu.so is never mapped here
    from oracle import pyoracle
    first, count = sharding.shard_range(n_total, world, rank)
    sc = workload.scene_preset(workload.SCENE_RPLIDAR_ROOM)
    ranges, _ = workload.scangen_batch(sc, 4000 + first, count, 0, n_scans, threads=1, truth=False)
    pts, cnt = workload.scan_to_points_host(ranges, sc, 20.0)
    procs = [pyoracle.CpuProcessor("port", 0.05, 256, 256, (0.5, 0.5), 2) for _ in range(count)]
    secs, poses, _ = pyoracle.run_streams(procs, pts, cnt, threads=1)
    local_seconds = 1.0 + rank              # deterministic stand-in for the device timer
    units, seconds = sharding.aggregate_throughput(count * n_scans, local_seconds)
    final = sharding.gather_rows(torch.from_numpy(poses[-1].copy()))
    if rank == 0:
        np.save(os.path.join(out_dir, "gathered.npy"), final.numpy())
        np.save(os.path.join(out_dir, "agg.npy"), np.array([units, seconds]))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_range_partitions_exactly():
    sys.path.insert(0, ROOT)
    from tfe_slam_ucl_b200 import sharding
    for n_total in (0, 1, 7, 8, 1024, 8192, 8193):
        for world in (1, 2, 3, 4, 8):
            seen = []
            for r in range(world):
                first, count = sharding.shard_range(n_total, world, r)
                seen.extend(range(first, first + count))
                for g in range(first, first + count):
                    assert sharding.owner_of(g, n_total, world) == r
            assert seen == list(range(n_total))
            counts = [sharding.shard_range(n_total, world, r)[1] for r in range(world)]
            assert max(counts) - min(counts) <= 1
    with pytest.raises(ValueError):
        sharding.shard_range(8, 2, 2)


def test_two_rank_gloo_sharded_run_matches_unsharded(tmp_path):
    world, n_total, n_scans = 2, 5, 6
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n_total, n_scans, str(tmp_path)), nprocs=world, join=True)
    units, seconds = np.load(tmp_path / "agg.npy")
    assert units == n_total * n_scans and seconds == 2.0       # sum of units, MAX of times (ranks took 1 s and 2 s)
    gathered = np.load(tmp_path / "gathered.npy")
    # unsharded reference run in this process
    sys.path.insert(0, ROOT)
    from oracle import pyoracle
    sc = workload.scene_preset(workload.SCENE_RPLIDAR_ROOM)
    ranges, _ = workload.scangen_batch(sc, 4000, n_total, 0, n_scans, threads=1, truth=False)
    pts, cnt = workload.scan_to_points_host(ranges, sc, 20.0)
    procs = [pyoracle.CpuProcessor("port", 0.05, 256, 256, (0.5, 0.5), 2) for _ in range(n_total)]
    _, poses, _ = pyoracle.run_streams(procs, pts, cnt, threads=2)
    assert gathered.shape == (n_total, 3)
    assert np.array_equal(gathered.view(np.int32), poses[-1].view(np.int32))


class _FakeEngine:
    """Stands in for hs.Engine in the hypothesis-sharding host logic (no GPU here): maps are small arrays, the 'kernel' writes
    a deterministic function of (pose, map checksum) through the raw pointers it is given, like the C ABI does."""

    def __init__(self, seed):
        rng = np.random.default_rng(seed)
        self.maps = [np.zeros(16 * 16 >> (2 * l), dtype=np.dtype([("logOddsVal", np.float32), ("updateIndex", np.int32)])) for l in range(3)]
        for m in self.maps:
            m["logOddsVal"] = rng.normal(size=m.shape).astype(np.float32)
            m["updateIndex"] = rng.integers(0, 100, m.shape)

    def level_dims(self, l):
        return 16 >> l, 16 >> l, 0.05 * (1 << l)

    def download_level(self, stream, l):
        return self.maps[l].copy()

    def upload_level(self, stream, l, cells):
        self.maps[l] = np.array(cells, copy=True)

    def hessian_derivs_batch_device(self, stream, level, poses_ptr, n, pts_ptr, n_pts, H_ptr, d_ptr):
        import ctypes
        poses = np.ctypeslib.as_array((ctypes.c_float * (3 * n)).from_address(poses_ptr)).reshape(n, 3)
        H = np.ctypeslib.as_array((ctypes.c_float * (9 * n)).from_address(H_ptr)).reshape(n, 9)
        d = np.ctypeslib.as_array((ctypes.c_float * (3 * n)).from_address(d_ptr)).reshape(n, 3)
        key = np.float32(self.maps[level]["logOddsVal"].sum() + self.maps[level]["updateIndex"].sum())
        H[:] = (poses.sum(axis=1, keepdims=True) * np.arange(1, 10, dtype=np.float32)[None, :] + key).astype(np.float32)
        d[:] = (poses * np.float32(2.0) + key).astype(np.float32)


def _hyp_worker(rank, world, port, n_hyp, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tfe_slam_ucl_b200 import hypothesis_sharding as hsd
    eng = _FakeEngine(seed=100 + rank)              # every rank starts with DIFFERENT maps; rank 1 owns the truth
    nbytes = hsd.replicate_stream_maps(eng, torch, dist, 1, rank, device="cpu")
    poses = torch.from_numpy(np.random.default_rng(5).normal(size=(n_hyp, 3)).astype(np.float32))
    pts = torch.zeros((4, 2))
    H = torch.zeros((n_hyp, 9)); d = torch.zeros((n_hyp, 3))
    hsd.evaluate_sharded(eng, torch, dist, rank, world, 0, poses, n_hyp, pts, 4, H, d)
    np.save(os.path.join(out_dir, "hyp_%d.npy" % rank), np.concatenate([H.numpy(), d.numpy()], axis=1))
    if rank == 0:
        np.save(os.path.join(out_dir, "hyp_bytes.npy"), np.array([nbytes]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_hypothesis_sharding_replicates_and_gathers(tmp_path):
    """SURVEY 8(e) row 2 on CPU: the owner's maps reach the other rank, every rank evaluates its own block of the hypotheses
    and ends up with all results, equal to one rank evaluating everything against the owner's maps."""
    world, n_hyp = 2, 64
    port = _free_port()
    mp.spawn(_hyp_worker, args=(world, port, n_hyp, str(tmp_path)), nprocs=world, join=True)
    r0, r1 = np.load(tmp_path / "hyp_0.npy"), np.load(tmp_path / "hyp_1.npy")
    assert np.array_equal(r0.view(np.int32), r1.view(np.int32))
    owner = _FakeEngine(seed=101)
    poses = np.random.default_rng(5).normal(size=(n_hyp, 3)).astype(np.float32)
    H = np.zeros((n_hyp, 9), np.float32); d = np.zeros((n_hyp, 3), np.float32)
    owner.hessian_derivs_batch_device(0, 0, poses.ctypes.data, n_hyp, 0, 4, H.ctypes.data, d.ctypes.data)
    assert np.array_equal(r0.view(np.int32), np.concatenate([H, d], axis=1).view(np.int32))
    assert int(np.load(tmp_path / "hyp_bytes.npy")[0]) == (256 + 64 + 16) * 8
