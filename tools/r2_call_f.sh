#!/bin/bash
# round 2, GPU call F: lane kernels after the register diet, pipeline with the aux stream, full default bench
mkdir -p gpurun_out
( time python -m pytest tests -q -m gpu -p no:cacheprovider -x 2>&1 | tail -12 ) > gpurun_out/f_pytest.log 2>&1
( tools/qb.sh --reads 12000000 --workload novaseq150_native; tools/qb.sh --workload pacbio_native
  echo "== LANE_MINB=7"; IDN_NVCC_EXTRA="-DIDN_LANE_MINB=7" python -c "from idencomp_b200 import build; build.build_gpu(force=True)"
  tools/qb.sh --reads 12000000 --workload novaseq150_native
  echo "== LANE_MINB=9"; IDN_NVCC_EXTRA="-DIDN_LANE_MINB=9" python -c "from idencomp_b200 import build; build.build_gpu(force=True)"
  tools/qb.sh --reads 12000000 --workload novaseq150_native
  python -c "from idencomp_b200 import build; build.build_gpu(force=True)" ) > gpurun_out/f_qb.log 2>&1
B="python bench.py --no-extra-workloads --no-cpu-baseline --no-other-mode --no-fastq --steps 3"
for cfg in "" "--e2e-threads 3 --e2e-chunk-blocks 32" "--e2e-threads 2" "--e2e-pipe-blocks 16" "--e2e-pipe-blocks 64"; do
  echo "== e2e [$cfg]"; $B $cfg 2> /tmp/e.err | python -c "import json,sys; d=json.load(sys.stdin); e=d['e2e']; print('value %.1f e2e %.1f c %.1f d %.1f'%(d['value'], e['value'], e['compress_GBps'], e['decompress_GBps']))"; tail -2 /tmp/e.err
done > gpurun_out/f_e2e.log 2>&1
( time python bench.py --steps 5 --warmup 3 > gpurun_out/f_bench.json 2> gpurun_out/f_bench.err ) > gpurun_out/f_bench.time 2>&1
echo done
