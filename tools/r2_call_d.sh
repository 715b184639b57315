#!/bin/bash
# round 2, GPU call D: whole parity suite (no -x), e2e arrangements
mkdir -p gpurun_out
python -c "from idencomp_b200 import build; build.build_gpu(force=True)"
( time python -m pytest tests -q -m gpu -p no:cacheprovider 2>&1 | tail -40 ) > gpurun_out/d_pytest.log 2>&1
B="python bench.py --no-extra-workloads --no-cpu-baseline --no-other-mode --no-fastq --steps 3"
for cfg in "" "--e2e-threads 3 --e2e-chunk-blocks 32" "--e2e-threads 2" "--e2e-pipe-blocks 16" "--e2e-pipe-blocks 64" "--e2e-threads 3"; do
  echo "== e2e [$cfg]"; $B $cfg --e2e-profile 2> /tmp/e.err | python -c "import json,sys; d=json.load(sys.stdin); e=d['e2e']; print('value %.1f e2e %.1f c %.1f d %.1f'%(d['value'], e['value'], e['compress_GBps'], e['decompress_GBps']))"; grep "e2e phases" /tmp/e.err | cut -c1-600
done > gpurun_out/d_e2e.log 2>&1
echo done
