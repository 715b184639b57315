#!/bin/bash
# round 2, GPU call I: page-locked pooled host buffers + cached contexts in the host mirror; where does the file-to-file time go
mkdir -p gpurun_out
( time python -m pytest tests/test_gpu_text_path.py tests/test_gpu_pipeline.py tests/test_gpu_host_api.py -q -m gpu -p no:cacheprovider -x 2>&1 | tail -30 ) > gpurun_out/i_pytest.log 2>&1
for cfg in "256 32 24000000" "64 16 24000000" "128 32 8000000"; do set -- $cfg
  IDN_HOST_TRACE=1 python bench.py --no-extra-workloads --no-cpu-baseline --no-other-mode --steps 3 --text-chunk-mb $1 --file-batch-blocks $2 --e2e-file-reads $3 2> gpurun_out/i_trace_$1_$2_$3.err | python -c "import json,sys; d=json.load(sys.stdin); print('chunk $1 MB batch $2 reads $3', json.dumps(d.get('e2e_file')))"
done > gpurun_out/i_chunks.log 2>&1
echo done
