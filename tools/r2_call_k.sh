#!/bin/bash
# round 2, GPU call K: reads bucketed by model pair (per-read selection)
mkdir -p gpurun_out
( time python -m pytest tests -q -m gpu -p no:cacheprovider -x 2>&1 | tail -8 ) > gpurun_out/k_pytest.log 2>&1
Q="--steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-other-mode --no-fastq --no-extra-workloads"
( echo "== buckets"; python bench.py --workload hiseq100_select4 $Q | python -c "import json,sys; d=json.load(sys.stdin); print(d['value'], d['compress_GBps'], d['decompress_GBps'], d['gpu_launches'], d['roofline']['kernels_ms_per_step'])"
  echo "== no buckets"; IDN_NO_BUCKETS=1 python bench.py --workload hiseq100_select4 $Q | python -c "import json,sys; d=json.load(sys.stdin); print(d['value'], d['compress_GBps'], d['decompress_GBps'], d['gpu_launches'], d['roofline']['kernels_ms_per_step'])"
  echo "== main"; python bench.py $Q | python -c "import json,sys; d=json.load(sys.stdin); print(d['value'], d['compress_GBps'], d['decompress_GBps'], d['gpu_launches'], d['roofline']['kernels_ms_per_step'])"
) > gpurun_out/k_select.log 2>&1
python bench.py --workload hiseq100_select4 --reads 4000000 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-other-mode --no-fastq --no-extra-workloads > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'^score_multi_kernel' -s 10 -c 1 -o gpurun_out/prof_r2_score2 python bench.py --workload hiseq100_select4 --reads 4000000 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-other-mode --no-fastq --no-extra-workloads > gpurun_out/k_ncu_score.log 2>&1
echo done
