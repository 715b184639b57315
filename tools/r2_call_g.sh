#!/bin/bash
# round 2, GPU call G: text path tests, file-to-file e2e leg
mkdir -p gpurun_out
( time python -m pytest tests -q -m gpu -p no:cacheprovider -x 2>&1 | tail -40 ) > gpurun_out/g_pytest.log 2>&1
( time python bench.py --no-extra-workloads --no-cpu-baseline --no-other-mode --steps 3 > gpurun_out/g_bench.json 2> gpurun_out/g_bench.err ) > gpurun_out/g_bench.time 2>&1
tail -5 gpurun_out/g_bench.err >> gpurun_out/g_bench.time
for mb in 64 512; do python bench.py --no-extra-workloads --no-cpu-baseline --no-other-mode --steps 3 --text-chunk-mb $mb 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print('chunk $mb MB', json.dumps(d.get('e2e_file')))"; done > gpurun_out/g_chunks.log 2>&1
echo done
