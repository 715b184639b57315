#!/bin/bash
# round 2, GPU call V: the clock sampler (samples stamped inside the timed region), short and driver-style runs
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 3 --no-extra-workloads --no-cpu-baseline --no-e2e --no-fastq --no-other-mode 2>gpurun_out/v_short.err | python -c "import json,sys; d=json.load(sys.stdin); print('steps 2:', d['clocks'], d['value'])" > gpurun_out/v_clocks.log 2>&1
( time python bench.py --steps 20 --warmup 5 > gpurun_out/v_bench.json 2> gpurun_out/v_bench.err ) > gpurun_out/v_bench.time 2>&1
python -c "import json; d=json.load(open('gpurun_out/v_bench.json')); print('steps 20:', d['clocks'], d['value'], d['e2e']['value']); print({k:v['clocks'] for k,v in d['workloads'].items()})" >> gpurun_out/v_clocks.log 2>&1
echo done
