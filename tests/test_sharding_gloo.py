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
