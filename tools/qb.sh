#!/bin/bash
# quick device-resident bench: prints compress/decompress GB/s and the top kernels
IDN_BENCH_NOVERIFY=$NOVERIFY python bench.py --no-cpu-baseline --no-other-mode --no-fastq --no-e2e --steps 3 "$@" > /tmp/qb.json 2> /tmp/qb.err || { tail -5 /tmp/qb.err; exit 1; }
python - <<PY
import json
d=json.load(open("/tmp/qb.json"))
k=d["roofline"]["kernels_ms_per_step"]
print("value %.1f comp %.1f decomp %.1f ok %s | "%(d["value"], d["compress_GBps"], d["decompress_GBps"], d["verified_round_trip"]) + " ".join("%s %.2f"%(n,v) for n,v in list(k.items())[:6]))
PY
