#!/bin/bash
# round 2, GPU call M: lanes bucketed by model pair (native format with per-lane selection)
mkdir -p gpurun_out
( time python -m pytest tests -q -m gpu -p no:cacheprovider -x 2>&1 | tail -8 ) > gpurun_out/m_pytest.log 2>&1
Q="--steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-other-mode --no-fastq --no-extra-workloads"
P="import json,sys; d=json.load(sys.stdin); print(round(d['value'],1), round(d['compress_GBps'],1), round(d['decompress_GBps'],1), d['gpu_launches'], {k:round(v,2) for k,v in list(d['roofline']['kernels_ms_per_step'].items())[:8]})"
( echo "== select4 native, buckets"; python bench.py --workload hiseq100_select4 --mode native $Q | python -c "$P"
  echo "== select4 native, no buckets"; IDN_NO_BUCKETS=1 python bench.py --workload hiseq100_select4 --mode native $Q | python -c "$P"
  echo "== nova native 12M (uniform, unchanged?)"; python bench.py --workload novaseq150_native --reads 12000000 $Q | python -c "$P"
  echo "== main"; python bench.py $Q | python -c "$P"
) > gpurun_out/m_bench.log 2>&1
echo done
