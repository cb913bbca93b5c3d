"""BASELINE configs[3] across GPUs: ONE scan, thousands of pose hypotheses.

The hypotheses are independent but all read one stream's maps, so -- unlike the stream sharding of the hot path -- this
mode has a real exchange step: the owner's level grids (10.5 MiB for 3 levels of 1024^2) are replicated to every rank once
(an NCCL broadcast here, one process per GPU; cudaMemcpyPeerAsync in the single-process C++ wrapper,
include/hsgpu/ShardedEngine.h), every rank evaluates its contiguous block of the hypotheses with the device-pointer entry
point, and the n_hyp x 12 floats come back with an all-gather. Python/torch.distributed is plumbing for bench.py and the
tests; the arithmetic is hs_hessian_derivs_batch_device / hs_match_hypotheses.
"""
import time

import numpy as np

from . import sharding


def replicate_stream_maps(eng, torch, dist, src_rank, rank, levels=3, stream=0, device="cuda"):
    """Copies every level of `stream` from rank `src_rank`'s engine into the same stream of every other rank's engine.
    Returns the bytes broadcast."""
    total = 0
    for l in range(levels):
        sx, sy, _ = eng.level_dims(l)
        if rank == src_rank:
            cells = eng.download_level(stream, l)
            t = torch.from_numpy(cells.view(np.uint8).copy()).to(device)
        else:
            t = torch.empty(sx * sy * 8, dtype=torch.uint8, device=device)
        if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
            dist.broadcast(t, src=src_rank)
        if rank != src_rank:
            eng.upload_level(stream, l, t.cpu().numpy().view(eng_cell_dtype()))
        total += sx * sy * 8
    return total


def eng_cell_dtype():
    return np.dtype([("logOddsVal", np.float32), ("updateIndex", np.int32)])


def evaluate_sharded(eng, torch, dist, rank, world, level, d_poses_all, n_hyp, d_pts, n_pts, d_H_all, d_d_all):
    """This rank's block of the hypotheses through hs_hessian_derivs_batch_device, written in place into the [n_hyp] result
    tensors, then an all-gather of the blocks (n_hyp must be divisible by world)."""
    first, count = sharding.shard_range(n_hyp, world, rank)
    eng.hessian_derivs_batch_device(0, level, d_poses_all.data_ptr() + first * 12, count, d_pts.data_ptr(), n_pts,
                                    d_H_all.data_ptr() + first * 36, d_d_all.data_ptr() + first * 12)
    if world > 1:
        dist.all_gather_into_tensor(d_H_all, d_H_all[first:first + count].clone())
        dist.all_gather_into_tensor(d_d_all, d_d_all[first:first + count].clone())


def bench_record(hs, torch, dist, local_rank, rank, world, n_hyp=4096, reps=20):
    """Sub-record for bench.py at N > 1: replicate one stream's maps, shard n_hyp hypotheses, gather. All ranks call this."""
    import workload
    dev = torch.device("cuda", local_rank)
    sc = workload.scene_preset(workload.SCENE_RPLIDAR_ROOM)
    ranges, _ = workload.scangen_batch(sc, 0xC0FFEE, 1, 0, 201, truth=False)
    pts, cnt = workload.scan_to_points_host(ranges, sc, 20.0)
    # one explicit CUDA stream for the engine's kernels, the collectives (torch issues them behind the CURRENT stream) and the
    # timing events -- the legacy default stream's handle is 0, which hs_set_cuda_stream reads as "use the engine's own stream"
    side = torch.cuda.Stream(device=dev)
    with hs.Engine(n_streams=1, max_points=360, device=local_rank) as eng, torch.cuda.stream(side):
        eng.set_update_factor_free(0.4); eng.set_update_factor_occupied(0.9)
        eng.set_map_update_min_dist_diff(0.0); eng.set_map_update_min_angle_diff(0.0)
        eng.set_cuda_stream(side.cuda_stream)
        pose = np.zeros((1, 3), np.float32)
        if rank == 0:
            for k in range(200):
                pose, _ = eng.update_batch(pts[k], cnt[k], want_cov=False)
        t_pose = torch.from_numpy(pose.copy()).to(dev)
        dist.broadcast(t_pose, src=0)
        pose = t_pose.cpu().numpy()
        torch.cuda.synchronize(); dist.barrier()
        t0 = time.perf_counter()
        rep_bytes = replicate_stream_maps(eng, torch, dist, 0, rank, device=dev)
        torch.cuda.synchronize(); dist.barrier()
        rep_s = time.perf_counter() - t0
        rng = np.random.default_rng(1)
        base = np.array([20.0 * pose[0, 0] + 512.0, 20.0 * pose[0, 1] + 512.0, pose[0, 2]], np.float32)
        noise = rng.uniform(-1, 1, (n_hyp, 3)).astype(np.float32)
        poses = (base[None, :] + noise * np.array([20.0, 20.0, 0.5], np.float32)).astype(np.float32)
        p = np.ascontiguousarray(pts[200, 0, :cnt[200, 0]])
        d_pose = torch.from_numpy(poses).to(dev); d_pts = torch.from_numpy(p).to(dev)
        d_H = torch.zeros((n_hyp, 9), device=dev); d_d = torch.zeros((n_hyp, 3), device=dev)
        for _ in range(3):
            evaluate_sharded(eng, torch, dist, rank, world, 0, d_pose, n_hyp, d_pts, p.shape[0], d_H, d_d)
        torch.cuda.synchronize(); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(side)
        for _ in range(reps):
            evaluate_sharded(eng, torch, dist, rank, world, 0, d_pose, n_hyp, d_pts, p.shape[0], d_H, d_d)
        e1.record(side)
        torch.cuda.synchronize()
        _, t_sharded = sharding.aggregate_throughput(0, e0.elapsed_time(e1) * 1e-3 / reps, device=dev)
        H_sh, d_sh = d_H.cpu().numpy(), d_d.cpu().numpy()
        # the same on rank 0 alone
        d_H1 = torch.zeros((n_hyp, 9), device=dev); d_d1 = torch.zeros((n_hyp, 3), device=dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(side)
        for _ in range(reps):
            eng.hessian_derivs_batch_device(0, 0, d_pose.data_ptr(), n_hyp, d_pts.data_ptr(), p.shape[0], d_H1.data_ptr(), d_d1.data_ptr())
        e1.record(side)
        torch.cuda.synchronize()
        t_single = e0.elapsed_time(e1) * 1e-3 / reps
        same = bool(np.array_equal(H_sh.view(np.int32), d_H1.cpu().numpy().view(np.int32)) and
                    np.array_equal(d_sh.view(np.int32), d_d1.cpu().numpy().view(np.int32)))
        # full multi-start match, sharded: every rank matches its block (host call), results all-gathered
        hints = (pose[0][None, :] + noise * np.array([1.0, 1.0, 0.5], np.float32)).astype(np.float32)
        first, count = sharding.shard_range(n_hyp, world, rank)
        g_pose = torch.zeros((n_hyp, 3), device=dev); g_cov = torch.zeros((n_hyp, 9), device=dev)

        def full_sharded():
            got, cov = eng.match_hypotheses(0, p, hints[first:first + count])
            g_pose[first:first + count] = torch.from_numpy(got).to(dev)
            g_cov[first:first + count] = torch.from_numpy(cov.reshape(count, 9)).to(dev)
            dist.all_gather_into_tensor(g_pose, g_pose[first:first + count].clone())
            dist.all_gather_into_tensor(g_cov, g_cov[first:first + count].clone())
            torch.cuda.synchronize()

        for _ in range(2):
            full_sharded()
        dist.barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            full_sharded()
        _, t_full_sh = sharding.aggregate_throughput(0, (time.perf_counter() - t0) / reps, device=dev)
        t0 = time.perf_counter()
        for _ in range(reps):
            ref_pose, ref_cov = eng.match_hypotheses(0, p, hints)
        t_full_1 = (time.perf_counter() - t0) / reps
        same_full = bool(np.array_equal(g_pose.cpu().numpy().view(np.int32), ref_pose.view(np.int32)))
        ok = torch.tensor([1.0 if (same and same_full) else 0.0], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    return {"workload": "BASELINE configs[3] sharded: one stream's 3-level maps replicated to %d GPUs, %d hypotheses in contiguous blocks, results all-gathered (NCCL)" % (world, n_hyp),
            "n_gpus": world, "hypotheses": n_hyp,
            "replicate": {"bytes": rep_bytes, "seconds": rep_s, "how": "download on the owner, NCCL broadcast, upload on the others (one-time per map state)"},
            "single_evaluation": {"sharded_ms": t_sharded * 1e3, "one_gpu_ms": t_single * 1e3, "hypotheses_per_s": n_hyp / t_sharded,
                                  "note": "kernel + 2 all-gathers per call, CUDA events, max over ranks; at this size one GPU needs microseconds, so the collective's latency decides"},
            "full_match": {"sharded_ms": t_full_sh * 1e3, "one_gpu_ms": t_full_1 * 1e3, "hypotheses_per_s": n_hyp / t_full_sh,
                           "note": "hs_match_hypotheses on each rank's block (host call) + all-gather, host clock, max over ranks"},
            "bit_identical_to_one_gpu_on_every_rank": bool(ok.item() == 1.0)}
