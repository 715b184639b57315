#!/usr/bin/env python
"""Registers / spills of the codec kernels from `nvcc -Xptxas -v` (cross-compiled, no GPU needed).
    python tools/regs.py [extra nvcc flags...]"""
import re, subprocess, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xptxas=-v", "-c", "-o", "/dev/null",
       *sys.argv[1:], str(ROOT / "idencomp_b200/csrc/idn_gpu.cu")]
out = subprocess.run(cmd, capture_output=True, text=True).stderr
if "error" in out:
    print(out[-3000:]); sys.exit(1)
ent = None
for ln in out.splitlines():
    m = re.search(r"Compiling entry function '(\S+)'", ln)
    if m:
        ent = m.group(1); continue
    if ent and "bytes stack frame" in ln:
        spill = re.search(r"(\d+) bytes spill stores, (\d+) bytes spill loads", ln); stack = re.search(r"(\d+) bytes stack frame", ln)
    m = re.search(r"Used (\d+) registers", ln)
    if m and ent:
        name = subprocess.run(["c++filt", ent], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"idn::", "", name); name = re.sub(r"\(.*", "", name); name = name.replace("void ", "")
        if any(k in name for k in ("encode", "decode", "score", "assemble")):
            print(f"{name[:95]:95s} regs {m.group(1):>3s} stack {stack.group(1)} spill {spill.group(1)}/{spill.group(2)}")
        ent = None
