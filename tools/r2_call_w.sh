#!/bin/bash
# round 2, GPU call W: the compressor's sink on a writer thread
mkdir -p gpurun_out
( time python -m pytest tests -q -m gpu -p no:cacheprovider 2>&1 | tail -8 ) > gpurun_out/w_pytest.log 2>&1
for k in 1 2; do
  IDN_HOST_TRACE=1 python bench.py --no-extra-workloads --no-cpu-baseline --no-other-mode --steps 3 2> gpurun_out/w_trace_$k.err | python -c "import json,sys; d=json.load(sys.stdin); f=d['e2e_file']; print('run $k', {a:{b:round(c,2) for b,c in f[a].items() if 'GBps' in b} for a in ('no_identifiers','with_identifiers')}, 'e2e', round(d['e2e']['value'],1))"
done > gpurun_out/w_runs.log 2>&1
timeout 100 python tools/fuzz_gpu.py --seconds 60 --seed 7000000 > gpurun_out/fuzz4.log 2>&1
echo "exit $?" >> gpurun_out/fuzz4.log
echo done
