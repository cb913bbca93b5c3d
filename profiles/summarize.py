"""Turns the ncu artefacts brought back in gpurun_out/ into the committed summaries under profiles/.

    python profiles/summarize.py <tag> <launches.csv> <prof.ncu-rep> [<prof2.ncu-rep> ...]

Writes profiles/<tag>_launches.csv (copy), profiles/<tag>_launches.md (per-kernel shares) and
profiles/<tag>_<kernel>.md (key raw metrics, top stall reasons, hottest SASS lines) for every kernel found.
"""
import collections
import csv
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__warps_active.avg.per_cycle_active", "smsp__warps_eligible.avg.per_cycle_active", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_warps", "launch__waves_per_multiprocessor", "launch__grid_size", "launch__block_size",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "lts__t_sectors_srcunit_tex_op_read.sum",
        "lts__t_sectors_srcunit_tex_op_write.sum", "lts__t_sectors_srcunit_tex_op_atom.sum", "lts__t_sectors_srcunit_tex_op_red.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]


def ncu_csv(rep, page, extra=()):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"] + list(extra), capture_output=True, text=True).stdout
    return list(csv.reader(out.splitlines()))


def launches(tag, path):
    shutil.copy(path, os.path.join(HERE, tag + "_launches.csv"))
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    idx = {h: i for i, h in enumerate(rows[0])}
    # bench.py runs four legs on one engine, each after a reset (k_reset_state + k_clear_cells): inputs resident in HBM (the
    # leg `value` and `roofline` are measured on), hs_update_batch from pinned host buffers (the matcher then fetches its scans
    # over PCIe inside the kernel), hs_update_batch_ranges (adds k_scan_to_points) and the pipelined hs_submit_batch. Shares are
    # reported per leg.
    legs, cur = [], None
    for r in rows[1:]:
        name = r[idx["Kernel Name"]].split("(")[0]
        us = float(r[idx["Metric Value"]]) / 1e3
        if name.startswith("k_reset_state"):
            cur = collections.OrderedDict()
            legs.append(cur)
        if cur is None:
            cur = collections.OrderedDict()
            legs.append(cur)
        cur.setdefault(name, []).append(us)
    leg_names = ["inputs resident in HBM (bench `value`, `roofline`)", "hs_update_batch, pinned host buffers, deferred integration (bench `e2e`: warm-up + R repeats of the K-step loop)",
                 "hs_update_batch_ranges (bench `e2e.ranges_api`)", "hs_submit_batch / hs_wait, DMA on a copy stream (bench `e2e.pipelined_api`)"]
    with open(os.path.join(HERE, tag + "_launches.md"), "w") as f:
        f.write("# %s: every launch with its device time (ncu --metrics gpu__time_duration.sum --clock-control none)\n\n" % tag)
        f.write("Times are cold-cache and serialised by the profiler: compare SHARES with bench.py's CUDA-event shares, not absolutes.\n")
        for li, agg in enumerate(legs):
            hot = {k: v for k, v in agg.items() if k.replace("void ", "").startswith(("k_match", "k_raycast", "k_scan_to_points", "k_cloud_to_points"))}
            tot = sum(sum(v) for v in hot.values()) or 1.0
            f.write("\n## leg %d: %s\n\n" % (li + 1, leg_names[li] if li < len(leg_names) else "further launches"))
            f.write("| kernel | launches | avg us | min us | max us | share of hot-path time |\n|---|---|---|---|---|---|\n")
            for k, v in agg.items():
                f.write("| %s | %d | %.1f | %.1f | %.1f | %s |\n" % (k, len(v), sum(v) / len(v), min(v), max(v), "%.3f" % (sum(v) / tot) if k in hot else "-"))


def kernels(tag, rep):
    raw = ncu_csv(rep, "raw")
    hdr, units = raw[0], raw[1]
    idx = {h: i for i, h in enumerate(hdr)}
    seen = set()
    for r in raw[2:]:
        name = r[idx["Kernel Name"]].split("(")[0].replace("void ", "").replace("<", "_").replace(">", "").replace(", ", "_").replace(",", "_")
        if name in seen:
            continue
        seen.add(name)
        with open(os.path.join(HERE, "%s_%s.md" % (tag, name)), "w") as f:
            f.write("# %s -- %s (ncu --set full --clock-control none --import-source on, one launch)\n\n" % (tag, r[idx["Kernel Name"]]))
            f.write("| metric | value | unit |\n|---|---|---|\n")
            for k in KEYS:
                if k in idx and r[idx[k]] not in ("", "n/a"):
                    f.write("| %s | %s | %s |\n" % (k, r[idx[k]], units[idx[k]]))
            stalls = sorted([(float(r[idx[h]]), h) for h in hdr if "issue_stalled" in h and h.endswith("per_issue_active.ratio") and r[idx[h]] not in ("", "n/a")], reverse=True)[:8]
            f.write("\nTop warp-stall reasons (warps stalled per issue-active cycle):\n\n")
            for v, h in stalls:
                f.write("* %.2f  %s\n" % (v, h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")))
            src = ncu_csv(rep, "source", ["--kernel-name", "regex:" + r[idx["Kernel Name"]].split("(")[0].replace("void ", "").split("<")[0], "--launch-count", "1"])
            if len(src) > 2:
                sh = src[1]
                si = {h: i for i, h in enumerate(sh)}
                data = [x for x in src[2:] if len(x) > si["Instructions Executed"] and x[si["# Samples"]].isdigit()]
                tot = sum(int(x[si["# Samples"]]) for x in data) or 1
                f.write("\nHottest SASS instructions (stall samples, %d total):\n\n```\n" % tot)
                for x in sorted(data, key=lambda x: -int(x[si["# Samples"]]))[:20]:
                    f.write("%5.1f%%  exec %10s  %s\n" % (100.0 * int(x[si["# Samples"]]) / tot, x[si["Instructions Executed"]], x[si["Source"]].strip()[:100]))
                f.write("```\n")


def traffic(tag, reps, streams_per_launch=1024):
    """profiles/traffic.json: dram__bytes_read.sum + dram__bytes_write.sum per launch of the hot kernels (bench.py's
    roofline.traffic reads it when the workload matches)."""
    import json
    out = {"source": "profiles/%s_*.md: ncu --set full --clock-control none, bench.py --steps 4 --warmup 3 (%d streams, one launch each)" % (tag, streams_per_launch),
           "streams_per_launch": streams_per_launch, "kernels": {}}
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    for rep in reps:
        raw = ncu_csv(rep, "raw")
        hdr, units = raw[0], raw[1]
        idx = {h: i for i, h in enumerate(hdr)}
        for r in raw[2:]:
            name = r[idx["Kernel Name"]].split("(")[0].replace("void ", "").split("<")[0]
            key = "k_match" if name.startswith("k_match") else name
            if key in out["kernels"]:
                continue
            rd = float(r[idx["dram__bytes_read.sum"]]) * scale[units[idx["dram__bytes_read.sum"]]]
            wr = float(r[idx["dram__bytes_write.sum"]]) * scale[units[idx["dram__bytes_write.sum"]]]
            out["kernels"][key] = {"dram_bytes_per_launch": rd + wr, "dram_read": rd, "dram_write": wr,
                                   "gpu_time_us_under_ncu": float(r[idx["gpu__time_duration.sum"]])}
    with open(os.path.join(HERE, "traffic.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    tag = sys.argv[1]
    launches(tag, sys.argv[2])
    for rep in sys.argv[3:]:
        kernels(tag, rep)
    traffic(tag, sys.argv[3:])
