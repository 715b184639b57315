#!/bin/bash
# occupancy experiment: __launch_bounds__ min blocks per SM of the encode / decode kernels
for cfg in "1 1" "10 1" "12 1" "1 12" "1 14"; do
  set -- $cfg
  IDN_NVCC_EXTRA="-DIDN_ENC_MINB=$1 -DIDN_DEC_MINB=$2" python -c "from idencomp_b200 import build; build.build_gpu(force=True)" 2>/dev/null
  python bench.py --no-cpu-baseline --no-other-mode --no-fastq --no-e2e --steps 3 2>/dev/null > /tmp/m.json
  python - "$1" "$2" <<'PY'
import json, sys
d = json.load(open("/tmp/m.json")); k = d["roofline"]["kernels_ms_per_step"]
print("enc minB", sys.argv[1], "dec minB", sys.argv[2], "encode", round(k["encode"], 1), "decode", round(k["decode"], 1))
PY
done
