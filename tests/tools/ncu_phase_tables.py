"""Regenerates profiles/r02_c2_k_match_lines.md, r02_c2_k_raycast_lines.md (and r02_c3_k_raycast_lines.md when the report is
there) from the ncu reports in gpurun_out/: the phase ranges are found by anchor strings in hs_kernels.cuh, so the tables
follow the source as it moves. usage: python tests/tools/ncu_phase_tables.py"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
SRC = os.path.join(ROOT, "tfe-slam-ucl_b200", "csrc", "hs_kernels.cuh")
src = open(SRC).read().splitlines()


def line(pat, start=0):
    for i, l in enumerate(src[start:], start + 1):
        if pat in l:
            return i
    raise SystemExit("anchor not found: " + pat)


def run(rep, kern, mangled, out, ranges):
    if not os.path.exists(rep):
        print("skip", out, "(no", rep + ")")
        return
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests/tools/ncu_lines.py"), rep, kern, mangled, "--src", SRC, "--top", "12", "--ranges",
                        ",".join("%d-%d:%s" % (lo, hi, nm) for nm, lo, hi in ranges)], capture_output=True, text=True)
    if r.returncode != 0:
        raise SystemExit(r.stderr[-2000:])
    open(os.path.join(ROOT, "profiles", out), "w").write(r.stdout)
    print(out, "ok")


k = line("__global__ void __launch_bounds__(THREADS, MIN_CTAS) k_raycast(") - 1
ray = [("prologue + launch order + marker base", k, line("// clear the hash table, reset the box and the allocators", k) - 1),
       ("clear hash + map", line("// clear the hash table, reset the box and the allocators", k), line("// phases A1 + A2: one thread per beam", k) - 1),
       ("A1+A2 beams / box / hash", line("// phases A1 + A2: one thread per beam", k), line("// phase A3: walk list", k) - 1),
       ("A3 owners", line("// phase A3: walk list", k), line("auto run_bands", k) - 1),
       ("band set-up", line("auto run_bands", k), line("// phase B1: free cells with the reference's bresenham2D", k) - 1),
       ("B1 walk", line("// phase B1: free cells with the reference's bresenham2D", k), line("// phase B2: updateSetFree exactly once", k) - 1),
       ("B2 sweep", line("// phase B2: updateSetFree exactly once", k), line("hs_stamp(P, 4);", k) - 1),
       ("C end points", line("hs_stamp(P, 4);", k), line("// maintenance kernels", k) - 1),
       ("helpers (beam / hash / udiv / shared-memory ops / fold)", line("#define HS_HASH_EMPTY"), k - 1)]
pp = line("__device__ __forceinline__ void hs_points_phase_wc(")
dense = line("if (!PB && total) {   // warp-uniform", pp)
last = line("hs_store_products(prod + i, pitch, gx, gy, rot, fv);", dense)
s3 = max(i for i in range(dense, last) if src[i - 1].strip() == "#pragma unroll" and "for (int j = 0; j < PPT; ++j) {" in src[i] and "const int i = i0 + j * istride;" in src[i + 1])
ev = line("__device__ __forceinline__ HsAcc hs_eval_cta(")
chain = line("if ((threadIdx.x >> 5) == serial_warp) {", ev)
gn = line("__device__ __forceinline__ bool hs_gauss_newton_step(")
tail = line("__device__ __forceinline__ void hs_match_tail(")
km = line("k_match(HsParams P, int n_slots, const int32_t* __restrict__ stream_ids,") - 1
kmend = line("// k_match2: the same computation as k_match", km) - 2
sp = line("__device__ __forceinline__ void hs_store_products(")
hl = line("__device__ __forceinline__ int hs_slot_stream(")
hl_end = line("// k_scan_to_points: HectorMappingRos::rosLaserScanToDataContainer") - 2
match = [("points step 1 (transform / key / ballot)", pp - 1, dense - 1), ("points dense miss pass (gather expf division)", dense, s3 - 1),
         ("points step 3 (bilinear / products)", s3, ev - 5), ("evaluation glue + barrier", ev, chain - 1), ("strict chain", chain, gn - 4),
         ("Gauss-Newton step", gn, tail - 4), ("tail", tail, km - 14), ("kernel body (prologue / level loop / barriers)", km, kmend),
         ("store products", sp, sp + 11), ("small helpers", hl, hl_end)]


def main():
    G = os.path.join(ROOT, "gpurun_out")
    run(os.path.join(G, "r02_c2.ncu-rep"), "k_raycast", "k_raycastILi384ELi32768ELi4E", "r02_c2_k_raycast_lines.md", ray)
    run(os.path.join(G, "r02_c2.ncu-rep"), "k_match", "k_matchILb1ELi3ELi1ELi388E", "r02_c2_k_match_lines.md", match)
    run(os.path.join(G, "r02_c3.ncu-rep"), "k_raycast", "k_raycastILi1024ELi147456ELi1E", "r02_c3_k_raycast_lines.md", ray)


if __name__ == "__main__":
    main()
