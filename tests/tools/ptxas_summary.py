"""Registers / spills / stack per kernel from tfe-slam-ucl_b200/csrc/ptxas.log (written by the csrc Makefile)."""
import os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
t = open(os.path.join(ROOT, "tfe-slam-ucl_b200", "csrc", "ptxas.log")).read()
ents = re.findall(r"Compiling entry function '(\S+)' for 'sm_100a'\nptxas info\s+: Function properties for \S+\n\s+(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\nptxas info\s+: Used (\d+) registers", t)
flt = sys.argv[1] if len(sys.argv) > 1 else ""
for name, stack, ss, sl, regs in ents:
    dn = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
    if flt in dn:
        print(regs.rjust(4), "regs  spill", ss, sl, " stack", stack, " ", dn.split("(")[0][:100])
