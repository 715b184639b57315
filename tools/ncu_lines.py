#!/usr/bin/env python
"""Per-source-line instruction and stall totals of one kernel from an ncu report.

    python tools/ncu_lines.py <report.ncu-rep> <kernel regex> [lib.so] [top N]

ncu's CSV source page is SASS-only; this joins it (by instruction offset) with `nvdisasm -g` line info of the cubin
inside the shared library, which must be the build that was profiled.
"""
import collections
import csv
import io
import re
import subprocess
import sys
import tempfile
from pathlib import Path

rep, kre = sys.argv[1], sys.argv[2]
lib = sys.argv[3] if len(sys.argv) > 3 else str(Path(__file__).resolve().parent.parent / "idencomp_b200" / "libidn_gpu.so")
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40

out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kre}"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
if not starts:
    sys.exit("kernel not found in report")
seg = rows[starts[0]:starts[1] if len(starts) > 1 else len(rows)]
kname = seg[0][1]
hdr = seg[1]
ia, isrc, ist, iaddr = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Address")
ithr = hdr.index("Thread Instructions Executed")
sass = [(int(r[iaddr], 16) if r[iaddr].startswith("0x") else int(r[iaddr]), r[isrc], int(r[ia] or 0), int(r[ist] or 0), int(r[ithr] or 0)) for r in seg[2:] if len(r) > ia and r[iaddr]]
base = sass[0][0]

with tempfile.TemporaryDirectory() as td:
    subprocess.run(["cuobjdump", "-xelf", "all", str(Path(lib).resolve())], cwd=td, check=True, capture_output=True)
    cub = next(Path(td).glob("*.cubin"))
    dis = subprocess.run(["nvdisasm", "-g", "-c", str(cub)], capture_output=True, text=True).stdout

fn = re.sub(r"<.*", "", re.sub(r"\(.*", "", kname)).split("::")[-1]
tmpl = "ILb1E" if "<(bool)1>" in kname else ("ILb0E" if "<(bool)0>" in kname else "")
lines = {}
cur, infn = None, False
for ln in dis.splitlines():
    if ln.startswith("//---") and ".text." in ln:
        infn = fn in ln and tmpl in ln
        continue
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (Path(m.group(1)).name, int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m:
        lines[int(m.group(1), 16)] = cur

agg = collections.defaultdict(lambda: [0, 0, 0, 0])
tot_i = tot_s = 0
for addr, src, n, st, thr in sass:
    key = lines.get(addr - base, ("?", 0))
    a = agg[key]
    a[0] += n
    a[1] += st
    a[2] += 1
    a[3] += thr
    tot_i += n
    tot_s += st
print(f"{kname[:80]}: {len(sass)} SASS instrs, {tot_i} warp instrs executed, {tot_s} stall samples")
srcs = {}
for (f, l), (n, st, k, thr) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    if f not in srcs:
        cand = [c for c in Path(__file__).resolve().parent.parent.rglob(f) if c.is_file()] if f not in ('?', '') else []
        srcs[f] = cand[0].read_text().splitlines() if cand else []
    text = srcs[f][l - 1].strip()[:90] if 0 < l <= len(srcs[f]) else ""
    print(f"{n / max(tot_i, 1):6.3f} instr {st / max(tot_s, 1):6.3f} stall {k:4d} sass  {f}:{l:<5d} {text}")
