"""Per-source-line instruction and stall-sample counts of one kernel from an `ncu --set full --import-source on` report.

ncu's CSV source page is SASS-only; the line of every SASS instruction comes from `nvdisasm --print-line-info` on the cubin
of the SAME build (extracted with `cuobjdump -xelf all libhsgpu.so`), joined by instruction offset.

usage: ncu_lines.py <report.ncu-rep> <kernel-name regex> <mangled-name substring> [--so path] [--src path] [--top N] [--ranges a-b:name,...]
"""
import argparse, collections, csv, os, re, subprocess, sys, tempfile

ap = argparse.ArgumentParser()
ap.add_argument("rep"); ap.add_argument("kernel"); ap.add_argument("mangled")
ap.add_argument("--so", default=os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "tfe-slam-ucl_b200", "libhsgpu.so"))
ap.add_argument("--src", default=None)
ap.add_argument("--top", type=int, default=40)
ap.add_argument("--ranges", default="")
ap.add_argument("--launch", type=int, default=0, help="which matching launch of the report (0 = first)")
a = ap.parse_args()

tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(a.so)], cwd=tmp, capture_output=True)
cubin = max((os.path.join(tmp, f) for f in os.listdir(tmp)), key=os.path.getsize)
dis = subprocess.run(["nvdisasm", "--print-line-info", cubin], capture_output=True, text=True).stdout.splitlines()
off2line, cur, infn = {}, None, False
for ln in dis:
    if ln.startswith(".text."):
        infn = a.mangled in ln
        continue
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", ln)
    if m and cur:
        off2line[int(m.group(1), 16)] = cur
out = subprocess.run(["ncu", "-i", a.rep, "--page", "source", "--csv", "--kernel-name", "regex:" + a.kernel], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
# several launches are concatenated: each starts with a "Kernel Name" row
blocks, curb = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        curb = []
        blocks.append((r[1], curb))
    elif curb is not None:
        curb.append(r)
name, blk = blocks[a.launch]
hdr = blk[0]
ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in blk[1:] if len(r) > ix["Instructions Executed"] and r[0].startswith("0x")]
base = int(data[0][0], 16)
per = collections.defaultdict(lambda: [0, 0, 0])
tot_i = tot_s = 0
for r in data:
    off = int(r[0], 16) - base
    key = off2line.get(off, ("?", 0))
    ie, sa, ti = int(r[ix["Instructions Executed"]] or 0), int(r[ix["# Samples"]] or 0), int(r[ix["Thread Instructions Executed"]] or 0)
    per[key][0] += ie; per[key][1] += sa; per[key][2] += ti
    tot_i += ie; tot_s += sa
src = {}
if a.src:
    for i, l in enumerate(open(a.src).read().splitlines(), 1):
        src[i] = l
print("# %s\n# %d SASS instructions, %d warp-instructions executed, %d stall samples" % (name, len(data), tot_i, tot_s))
if a.ranges:
    print("\n| phase | lines | warp-instructions | share | stall samples | share | lanes/instr |\n|---|---|---|---|---|---|---|")
    for spec in a.ranges.split(","):
        rng, nm = spec.split(":")
        files = None
        if "@" in rng:
            rng, files = rng.split("@")
        lo, hi = [int(x) for x in rng.split("-")]
        sel = [(k, v) for k, v in per.items() if lo <= k[1] <= hi and (files is None or k[0] == files) and (files is not None or k[0].endswith("hs_kernels.cuh"))]
        i = sum(v[0] for _, v in sel); s = sum(v[1] for _, v in sel); t = sum(v[2] for _, v in sel)
        print("| %s | %s | %d | %.1f%% | %d | %.1f%% | %.1f |" % (nm, rng + ("@" + files if files else ""), i, 100.0 * i / tot_i, s, 100.0 * s / max(tot_s, 1), t / max(i, 1)))
    other = [(k, v) for k, v in per.items() if not k[0].endswith("hs_kernels.cuh")]
    i = sum(v[0] for _, v in other); s = sum(v[1] for _, v in other)
    print("| (other files: libm etc.) | - | %d | %.1f%% | %d | %.1f%% | |" % (i, 100.0 * i / tot_i, s, 100.0 * s / max(tot_s, 1)))
print("\n| file:line | warp-instructions | share | stall samples | share | lanes/instr | source |\n|---|---|---|---|---|---|---|")
for k, v in sorted(per.items(), key=lambda kv: -kv[1][0])[: a.top]:
    print("| %s:%d | %d | %.1f%% | %d | %.1f%% | %.1f | `%s` |" % (k[0], k[1], v[0], 100.0 * v[0] / tot_i, v[1], 100.0 * v[1] / max(tot_s, 1), v[2] / max(v[0], 1),
                                                            (src.get(k[1], "") if k[0].endswith("hs_kernels.cuh") else "").strip()[:90]))
