"""The C++ drop-in facade (include/hsgpu/HectorSlamProcessor.h): compiles against the mini-Eigen shim here (a user
builds it against real Eigen), links libhsgpu.so, and -- on a GPU -- reproduces the CPU oracle bit for bit when
driven like HectorMappingRos::scanCallback drives the reference."""
import os
import subprocess

import numpy as np
import pytest

from hs_testutil import bits, make_streams

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def build_facade_test(tmp_path, hs):
    exe = str(tmp_path / "facade_test")
    libdir = os.path.dirname(hs.LIB_PATH)
    cmd = ["g++", "-std=c++14", "-O2", "-ffp-contract=off", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "oracle", "shim"),
           "-o", exe, os.path.join(ROOT, "tests", "cpp", "facade_test.cpp"), "-L", libdir, "-lhsgpu", "-Wl,-rpath," + libdir, "-pthread"]
    subprocess.run(cmd, check=True)
    return exe


def test_facade_compiles_and_fails_loudly_without_gpu(tmp_path, hs):
    exe = build_facade_test(tmp_path, hs)
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by test_facade_matches_oracle")
    inp = tmp_path / "in.bin"
    inp.write_bytes(np.array([0, 4], np.int32).tobytes())
    res = subprocess.run([exe, str(inp), str(tmp_path / "out.bin")], capture_output=True, text=True)
    assert res.returncode != 0 and "no usable CUDA device" in res.stderr     # constructor throws: no CPU fallback


@pytest.mark.gpu
def test_facade_matches_oracle(tmp_path, hs, oracle):
    exe = build_facade_test(tmp_path, hs)
    n_scans, mp = 40, 360
    sc, ranges, pts, cnt, truth = make_streams(hs, 1, n_scans, seed0=808)
    with open(tmp_path / "in.bin", "wb") as f:
        f.write(np.array([n_scans, mp], np.int32).tobytes())
        for k in range(n_scans):
            f.write(np.array([cnt[k, 0]], np.int32).tobytes())
            f.write(pts[k, 0].astype(np.float32).tobytes())
    res = subprocess.run([exe, str(tmp_path / "in.bin"), str(tmp_path / "out.bin")], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    raw = open(tmp_path / "out.bin", "rb").read()
    rows = np.frombuffer(raw[: n_scans * 48], np.float32).reshape(n_scans, 12)
    p = oracle.CpuProcessor("port")
    p.set_update_factor_free(0.4); p.set_update_factor_occupied(0.9)
    p.set_map_update_min_dist_diff(0.4); p.set_map_update_min_angle_diff(0.9)
    for k in range(n_scans):
        if k == n_scans // 2:
            p.reset()
        pose = p.update(pts[k, 0], cnt[k, 0], hint=p.pose())
        assert np.array_equal(bits(rows[k, :3]), bits(pose)), k
        if k not in (0, n_scans // 2):
            assert np.array_equal(bits(rows[k, 3:]), bits(p.cov().reshape(-1))), k
    off = n_scans * 48
    for l in range(3):
        sx, sy, ui = np.frombuffer(raw[off: off + 12], np.int32)
        off += 12
        cells = np.frombuffer(raw[off: off + sx * sy * 8], oracle.CELL_DTYPE)
        off += sx * sy * 8
        ref = p.download_level(l)
        assert (sx, sy) == p.level_dims(l)[:2] and ui == p.update_index(l)
        assert np.array_equal(cells["updateIndex"], ref["updateIndex"]) and np.array_equal(bits(cells["logOddsVal"]), bits(ref["logOddsVal"]))


def build_batch_test(tmp_path, hs):
    exe = str(tmp_path / "batch_test")
    libdir = os.path.dirname(hs.LIB_PATH)
    cmd = ["g++", "-std=c++14", "-O2", "-I", os.path.join(ROOT, "include"), "-o", exe, os.path.join(ROOT, "tests", "cpp", "batch_test.cpp"),
           "-L", libdir, "-lhsgpu", "-Wl,-rpath," + libdir, "-pthread"]
    subprocess.run(cmd, check=True)
    return exe


def test_batch_engine_compiles_and_fails_loudly_without_gpu(tmp_path, hs):
    """include/hsgpu/BatchEngine.h (C++ RAII wrapper over the C ABI) needs nothing but the C++ standard library."""
    exe = build_batch_test(tmp_path, hs)
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by test_batch_engine_matches_oracle")
    inp = tmp_path / "in.bin"
    inp.write_bytes(np.array([2, 0, 4], np.int32).tobytes())
    res = subprocess.run([exe, str(inp), str(tmp_path / "out.bin")], capture_output=True, text=True)
    assert res.returncode == 1 and "no usable CUDA device" in res.stderr


@pytest.mark.gpu
def test_batch_engine_matches_oracle(tmp_path, hs, oracle):
    exe = build_batch_test(tmp_path, hs)
    n, n_scans, mp = 5, 9, 360
    sc, ranges, pts, cnt, truth = make_streams(hs, n, n_scans, seed0=4040)
    with open(tmp_path / "in.bin", "wb") as f:
        f.write(np.array([n, n_scans, mp], np.int32).tobytes())
        for k in range(n_scans):
            f.write(cnt[k].astype(np.int32).tobytes())
            f.write(pts[k].astype(np.float32).tobytes())
    res = subprocess.run([exe, str(tmp_path / "in.bin"), str(tmp_path / "out.bin")], capture_output=True, text=True)
    assert res.returncode == 0, (res.returncode, res.stderr)
    raw = open(tmp_path / "out.bin", "rb").read()
    poses = np.frombuffer(raw[: n_scans * n * 12], np.float32).reshape(n_scans, n, 3)
    procs = [oracle.CpuProcessor("port") for _ in range(n)]
    for p in procs:
        p.set_update_factor_free(0.4); p.set_update_factor_occupied(0.9)
        p.set_map_update_min_dist_diff(0.0); p.set_map_update_min_angle_diff(0.0)
    for k in range(n_scans):
        for s, p in enumerate(procs):
            assert np.array_equal(bits(poses[k, s]), bits(p.update(pts[k, s], cnt[k, s], hint=p.pose()))), (k, s)
    off = n_scans * n * 12
    for s in (0, n - 1):
        ui = np.frombuffer(raw[off: off + 4], np.int32)[0]
        off += 4
        cells = np.frombuffer(raw[off: off + 1024 * 1024 * 8], oracle.CELL_DTYPE)
        off += 1024 * 1024 * 8
        ref = procs[s].download_level(0)
        assert ui == procs[s].update_index(0)
        assert np.array_equal(cells["updateIndex"], ref["updateIndex"]) and np.array_equal(bits(cells["logOddsVal"]), bits(ref["logOddsVal"]))


def build_maprep_test(tmp_path, hs):
    exe = str(tmp_path / "maprep_test")
    libdir = os.path.dirname(hs.LIB_PATH)
    cmd = ["g++", "-std=c++14", "-O2", "-ffp-contract=off", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "oracle", "shim"),
           "-o", exe, os.path.join(ROOT, "tests", "cpp", "maprep_test.cpp"), "-L", libdir, "-lhsgpu", "-Wl,-rpath," + libdir, "-pthread"]
    subprocess.run(cmd, check=True)
    return exe


def test_maprep_gpu_compiles_and_fails_loudly_without_gpu(tmp_path, hs):
    """include/hsgpu/MapRepGpu.h implements the reference's MapRepresentationInterface (the seam below the processor)."""
    exe = build_maprep_test(tmp_path, hs)
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by test_maprep_gpu_matches_oracle")
    inp = tmp_path / "in.bin"
    inp.write_bytes(np.array([0, 4], np.int32).tobytes())
    res = subprocess.run([exe, str(inp), str(tmp_path / "out.bin")], capture_output=True, text=True)
    assert res.returncode != 0 and "no usable CUDA device" in res.stderr


@pytest.mark.gpu
def test_maprep_gpu_matches_oracle(tmp_path, hs, oracle):
    """matchData + host-side gate + updateByScan through MapRepresentationInterface == the reference's update(), bit for bit
    (node thresholds 0.4 m / 0.9 rad, so some scans are matched but not integrated)."""
    exe = build_maprep_test(tmp_path, hs)
    n_scans, mp = 120, 360
    sc, ranges, pts, cnt, truth = make_streams(hs, 1, n_scans, seed0=909)
    with open(tmp_path / "in.bin", "wb") as f:
        f.write(np.array([n_scans, mp], np.int32).tobytes())
        for k in range(n_scans):
            f.write(np.array([cnt[k, 0]], np.int32).tobytes())
            f.write(pts[k, 0].astype(np.float32).tobytes())
    res = subprocess.run([exe, str(tmp_path / "in.bin"), str(tmp_path / "out.bin")], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    raw = open(tmp_path / "out.bin", "rb").read()
    rows = np.frombuffer(raw[: n_scans * 48], np.float32).reshape(n_scans, 12)
    p = oracle.CpuProcessor("port")
    p.set_update_factor_free(0.4); p.set_update_factor_occupied(0.9)
    p.set_map_update_min_dist_diff(0.4); p.set_map_update_min_angle_diff(0.9)
    for k in range(n_scans):
        pose = p.update(pts[k, 0], cnt[k, 0], hint=p.pose())
        assert np.array_equal(bits(rows[k, :3]), bits(pose)), k
        if k > 0:
            assert np.array_equal(bits(rows[k, 3:]), bits(p.cov().reshape(-1))), k
    assert 1 <= p.update_index(0) < n_scans - 1         # the gate both passed (more than the first scan) and held back scans
    off = n_scans * 48
    for l in range(3):
        sx, sy, ui = np.frombuffer(raw[off: off + 12], np.int32)
        off += 12
        cells = np.frombuffer(raw[off: off + sx * sy * 8], oracle.CELL_DTYPE)
        off += sx * sy * 8
        ref = p.download_level(l)
        assert (sx, sy) == p.level_dims(l)[:2] and ui == p.update_index(l)
        assert np.array_equal(cells["updateIndex"], ref["updateIndex"]) and np.array_equal(bits(cells["logOddsVal"]), bits(ref["logOddsVal"]))


@pytest.mark.gpu
def test_maprep_gpu_map_without_matching_integrates_stale_containers(tmp_path, hs, oracle):
    """The reference's HectorSlamProcessor::update(..., map_without_matching = true) on top of MapRepGpu: updateByScan of a scan
    that was NOT matched integrates, on the coarse levels, the containers the last matchData left behind
    (MapRepMultiMap.h:134-147). Checked against the reference itself (oracle/_ref) when present, else its C restatement."""
    exe = build_maprep_test(tmp_path, hs)
    n_scans, mp, period = 60, 360, 3
    sc, ranges, pts, cnt, truth = make_streams(hs, 1, n_scans, seed0=1717)
    with open(tmp_path / "in.bin", "wb") as f:
        f.write(np.array([n_scans, mp], np.int32).tobytes())
        for k in range(n_scans):
            f.write(np.array([cnt[k, 0]], np.int32).tobytes())
            f.write(pts[k, 0].astype(np.float32).tobytes())
    res = subprocess.run([exe, str(tmp_path / "in.bin"), str(tmp_path / "out.bin"), str(period)], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    raw = open(tmp_path / "out.bin", "rb").read()
    rows = np.frombuffer(raw[: n_scans * 48], np.float32).reshape(n_scans, 12)
    p = oracle.CpuProcessor("ref" if oracle.have("ref") else "port")
    p.set_update_factor_free(0.4); p.set_update_factor_occupied(0.9)
    p.set_map_update_min_dist_diff(0.4); p.set_map_update_min_angle_diff(0.9)
    for k in range(n_scans):
        pose = p.update(pts[k, 0], cnt[k, 0], hint=p.pose(), map_without_matching=(k % period == period - 1))
        assert np.array_equal(bits(rows[k, :3]), bits(pose)), k
    off = n_scans * 48
    for l in range(3):
        sx, sy, ui = np.frombuffer(raw[off: off + 12], np.int32)
        off += 12
        cells = np.frombuffer(raw[off: off + sx * sy * 8], oracle.CELL_DTYPE)
        off += sx * sy * 8
        ref = p.download_level(l)
        assert ui == p.update_index(l)
        assert np.array_equal(cells["updateIndex"], ref["updateIndex"]) and np.array_equal(bits(cells["logOddsVal"]), bits(ref["logOddsVal"])), l


def build_sharded_test(tmp_path, hs):
    exe = str(tmp_path / "sharded_test")
    libdir = os.path.dirname(hs.LIB_PATH)
    cmd = ["g++", "-std=c++14", "-O2", "-I", os.path.join(ROOT, "include"), "-o", exe, os.path.join(ROOT, "tests", "cpp", "sharded_test.cpp"),
           "-L", libdir, "-lhsgpu", "-Wl,-rpath," + libdir, "-pthread"]
    subprocess.run(cmd, check=True)
    return exe


def test_sharded_engine_compiles_and_fails_loudly_without_gpu(tmp_path, hs):
    """include/hsgpu/ShardedEngine.h (C++ multi-GPU host layer: one engine + one host thread per shard) needs nothing but the
    C ABI and the C++ standard library."""
    exe = build_sharded_test(tmp_path, hs)
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by test_sharded_engine_equals_one_engine")
    inp = tmp_path / "in.bin"
    inp.write_bytes(np.array([2, 0, 4], np.int32).tobytes())
    res = subprocess.run([exe, str(inp), str(tmp_path / "out.bin")], capture_output=True, text=True)
    assert res.returncode == 1 and "no usable CUDA device" in res.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("n_shards", [0, 3])
def test_sharded_engine_equals_one_engine(tmp_path, hs, oracle, n_shards):
    """Streams sharded over engines (one per visible GPU, or 3 shards round-robin over the visible GPUs) give the poses, maps and
    hypothesis results of ONE engine holding all streams, bit for bit; the poses also equal the CPU oracle's."""
    exe = build_sharded_test(tmp_path, hs)
    n, n_scans, mp = 7, 8, 360
    sc, ranges, pts, cnt, truth = make_streams(hs, n, n_scans, seed0=6060)
    with open(tmp_path / "in.bin", "wb") as f:
        f.write(np.array([n, n_scans, mp], np.int32).tobytes())
        for k in range(n_scans):
            f.write(cnt[k].astype(np.int32).tobytes())
            f.write(pts[k].astype(np.float32).tobytes())
    res = subprocess.run([exe, str(tmp_path / "in.bin"), str(tmp_path / "out.bin"), str(n_shards)], capture_output=True, text=True)
    assert res.returncode == 0, (res.returncode, res.stderr)
    assert "sharded ok" in res.stderr
    poses = np.frombuffer(open(tmp_path / "out.bin", "rb").read(), np.float32).reshape(n_scans, n, 3)
    procs = [oracle.CpuProcessor("port") for _ in range(n)]
    for p in procs:
        p.set_update_factor_free(0.4); p.set_update_factor_occupied(0.9)
        p.set_map_update_min_dist_diff(0.0); p.set_map_update_min_angle_diff(0.0)
    for k in range(n_scans):
        for s, p in enumerate(procs):
            assert np.array_equal(bits(poses[k, s]), bits(p.update(pts[k, s], cnt[k, s], hint=p.pose()))), (k, s)
