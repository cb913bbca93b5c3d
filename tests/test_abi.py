"""The drop-in boundary: libhsgpu.so loads on a machine without a GPU, exports every symbol that
include/hsgpu.h declares, keeps the LogOddsCell layout, and fails loudly (no CPU fallback) without a device."""
import ctypes
import os
import re
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    text = open(os.path.join(ROOT, "include", "hsgpu.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = re.findall(r"^\s*(?:const\s+)?(?:int|float|uint64_t|char\*|const char\*|const float\*)\s*\*?\s*(hs_[a-z0-9_]+)\s*\(", text, flags=re.M)
    return sorted(set(names))


def test_header_declares_expected_surface():
    names = declared_functions()
    for must in ["hs_create", "hs_destroy", "hs_reset", "hs_update_batch", "hs_update_batch_device", "hs_update_batch_ranges",
                 "hs_hessian_derivs_batch", "hs_raycast_batch", "hs_match_batch", "hs_download_level", "hs_upload_level",
                 "hs_get_poses", "hs_get_covariances", "hs_set_update_factor_free", "hs_set_update_factor_occupied",
                 "hs_set_map_update_min_dist_diff", "hs_set_map_update_min_angle_diff", "hs_get_scale_to_map", "hs_get_map_levels",
                 "hs_classify_level", "hs_get_update_index", "hs_status_string", "hs_integrate_batch", "hs_scan_to_points_host"]:
        assert must in names, must
    assert len(names) >= 35


def test_library_exports_every_declared_symbol(hs):
    L = ctypes.CDLL(hs.LIB_PATH)
    missing = [n for n in declared_functions() if not hasattr(L, n)]
    assert not missing, missing


def test_cell_layout_matches_logoddscell(hs):
    # LogOddsCell = {float logOddsVal; int updateIndex} (GridMapLogOdds.h:99-100): 8 bytes, val first
    assert hs.CELL_DTYPE.itemsize == 8
    assert hs.CELL_DTYPE.fields["logOddsVal"][1] == 0 and hs.CELL_DTYPE.fields["updateIndex"][1] == 4


def test_no_device_fails_loudly_and_arguments_are_checked(hs):
    import torch
    if torch.cuda.is_available():
        pytest.skip("this check is for GPU-less machines")
    with pytest.raises(hs.HsError) as e:
        hs.Engine(n_streams=1)
    assert e.value.status == hs.HS_ERR_NO_DEVICE
    assert "no CPU fallback" in str(e.value)


def test_invalid_config_rejected(hs):
    L = hs.lib()
    h = ctypes.c_void_p()
    for bad in [dict(levels=0), dict(levels=9), dict(n_streams=0), dict(max_points=0), dict(map_size_x=1), dict(map_resolution=0.0)]:
        kw = dict(map_resolution=0.05, map_size_x=64, map_size_y=64, start_x=0.5, start_y=0.5, levels=2, n_streams=1, max_points=16, device=0)
        kw.update(bad)
        cfg = hs.HsConfig(kw["map_resolution"], kw["map_size_x"], kw["map_size_y"], kw["start_x"], kw["start_y"], kw["levels"],
                          kw["n_streams"], kw["max_points"], kw["device"])
        assert L.hs_create(ctypes.byref(cfg), ctypes.byref(h)) == hs.HS_ERR_INVALID_ARG
    assert L.hs_create(None, ctypes.byref(h)) == hs.HS_ERR_INVALID_ARG
    assert hs.status_string(hs.HS_ERR_NO_DEVICE).startswith("no usable CUDA device")
    assert hs.status_string(0) == "ok"


def test_product_does_not_link_the_oracle(hs):
    """The shipped library must not depend on, or contain symbols of, the CPU oracle."""
    import subprocess
    out = subprocess.run(["nm", "-D", hs.LIB_PATH], capture_output=True, text=True).stdout
    assert "hso_" not in out and "hsref_" not in out
    ldd = subprocess.run(["ldd", hs.LIB_PATH], capture_output=True, text=True).stdout
    assert "hsoracle" not in ldd and "hsref" not in ldd
    for root, _, files in os.walk(os.path.join(ROOT, "tfe-slam-ucl_b200")):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".c", ".h")):
                assert "oracle" not in open(os.path.join(root, fn)).read().replace("no oracle", ""), fn


def test_product_library_holds_no_workload_generator(hs):
    """The synthetic scan generator lives in workload/libhsscangen.so (include/hsscangen.h), not in the product; the CPU-only
    arms (bench.py --impl reference, golden generation, the gloo test) load it without mapping libhsgpu.so."""
    import subprocess
    import workload
    out = subprocess.run(["nm", "-D", hs.LIB_PATH], capture_output=True, text=True).stdout
    assert "hs_scangen_batch" not in out and "hs_scene_preset" not in out
    wl = subprocess.run(["nm", "-D", workload.lib() and workload.LIB_PATH], capture_output=True, text=True).stdout
    assert "hs_scangen_batch" in wl and "hs_scan_to_points_host_batch" in wl
    ldd = subprocess.run(["ldd", workload.LIB_PATH], capture_output=True, text=True).stdout
    assert "cuda" not in ldd.lower() and "hsgpu" not in ldd
    code = ("import sys; sys.path.insert(0, %r); import workload; sc = workload.scene_preset(0); "
            "r, _ = workload.scangen_batch(sc, 1, 1, 0, 1, truth=False); p, c = workload.scan_to_points_host(r, sc, 20.0); "
            "maps = open('/proc/self/maps').read(); assert 'libhsgpu' not in maps and 'libcudart' not in maps; print(int(c.sum()))" % ROOT)
    res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert res.returncode == 0 and int(res.stdout.strip()) > 100, res.stderr


def test_profile_phase_anchors_resolve():
    """tests/tools/ncu_phase_tables.py finds the phases of k_match / k_raycast by anchor strings in hs_kernels.cuh: the
    anchors must still be there (and in order) after an edit of the kernels, or the committed phase tables cannot be regenerated."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("ncu_phase_tables", os.path.join(ROOT, "tests", "tools", "ncu_phase_tables.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)      # resolves every anchor at import (SystemExit if one is missing)
    for name, table in (("k_raycast", mod.ray), ("k_match", mod.match)):
        for phase, lo, hi in table:
            assert 0 < lo <= hi, (name, phase, lo, hi)
