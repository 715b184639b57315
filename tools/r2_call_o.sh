#!/bin/bash
# round 2, GPU call O: software-pipelined scorer; second randomised parity run
mkdir -p gpurun_out
( time python -m pytest tests -q -m gpu -p no:cacheprovider -x 2>&1 | tail -8 ) > gpurun_out/o_pytest.log 2>&1
Q="--steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-other-mode --no-fastq --no-extra-workloads"
P="import json,sys; d=json.load(sys.stdin); print(round(d['value'],1), round(d['compress_GBps'],1), round(d['decompress_GBps'],1), d['gpu_launches'], {k:round(v,2) for k,v in list(d['roofline']['kernels_ms_per_step'].items())[:8]})"
( echo "== select4"; python bench.py --workload hiseq100_select4 $Q | python -c "$P"
  echo "== select4 native"; python bench.py --workload hiseq100_select4 --mode native $Q | python -c "$P"
) > gpurun_out/o_bench.log 2>&1
timeout 200 python tools/fuzz_gpu.py --seconds 150 --seed 500000 > gpurun_out/fuzz2.log 2>&1
echo "exit $?" >> gpurun_out/fuzz2.log
echo done
