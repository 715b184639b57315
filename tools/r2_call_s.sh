#!/bin/bash
# round 2, GPU call S: run-to-run spread of the file-to-file leg (context cache, resident models, pool classes)
mkdir -p gpurun_out
( time python -m pytest tests/test_gpu_host_api.py tests/test_gpu_text_path.py tests/test_gpu_pipeline.py -q -m gpu -p no:cacheprovider 2>&1 | tail -6 ) > gpurun_out/s_pytest.log 2>&1
for k in 1 2 3; do
  IDN_HOST_TRACE=1 python bench.py --no-extra-workloads --no-cpu-baseline --no-other-mode --steps 3 2> gpurun_out/s_trace_$k.err | python -c "import json,sys; d=json.load(sys.stdin); f=d['e2e_file']; print('run $k', {a:{b:round(c,2) for b,c in f[a].items() if 'GBps' in b} for a in ('no_identifiers','with_identifiers')}, 'e2e', round(d['e2e']['value'],1))"
done > gpurun_out/s_runs.log 2>&1
echo done
