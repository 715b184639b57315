#!/bin/bash
# native-mode lane quantum sweep: container size and kernel times
for q in 512 1024 2048 4096 16384; do
  python bench.py --mode native --lane-symbols $q --no-cpu-baseline --no-other-mode --no-fastq --no-e2e --steps 3 2>/dev/null > /tmp/lane_$q.json
  python - "$q" <<'PY'
import json, sys
q = sys.argv[1]
d = json.load(open(f"/tmp/lane_{q}.json")); k = d["roofline"]["kernels_ms_per_step"]
print(q, round(d["compress_GBps"], 1), round(d["decompress_GBps"], 1), round(d["container_bytes_per_read"], 2), round(k["encode_lane"], 1), round(k["decode_lane"], 1))
PY
done
