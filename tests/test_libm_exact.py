"""hs_libm_exact.h (the sinf/cosf/expf the kernels use) against the host libm the CPU reference uses: a strided
sweep over all float bit patterns of the supported domain must match bit for bit. (The exhaustive sweep,
stride 1, takes ~25 s on 8 threads: `tests/tools/libm_exact_check 1 8`; 0 mismatches on glibc 2.39.)"""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_libm_exact_matches_host_libm(tmp_path):
    exe = str(tmp_path / "libm_exact_check")
    src = os.path.join(ROOT, "tests", "tools", "libm_exact_check.c")
    inc = os.path.join(ROOT, "tfe-slam-ucl_b200", "csrc")
    subprocess.run(["gcc", "-O2", "-ffp-contract=off", "-mfma", "-pthread", "-I", inc, "-o", exe, src, "-lm"], check=True)
    res = subprocess.run([exe, "257", str(min(8, os.cpu_count() or 1))], capture_output=True, text=True)
    print(res.stdout, res.stderr)
    assert res.returncode == 0, res.stdout + res.stderr
    for fn in ("sinf", "cosf", "expf"):
        assert "%s checked" % fn in res.stdout and "mismatches 0" in res.stdout
