#!/bin/bash
# round 2, GPU call Y: sub-chunk size of the pipelined host-pointer calls for long reads in the compat format
mkdir -p gpurun_out
for pb in 59 110 160 320; do
  python bench.py --workload pacbio --no-extra-workloads --no-cpu-baseline --no-other-mode --no-fastq --steps 3 --e2e-pipe-blocks $pb 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); e=d['e2e']; print('compat pipe $pb', round(e['value'],1), round(e['compress_GBps'],1), round(e['decompress_GBps'],1), e.get('blocks_per_call'))"
done > gpurun_out/y_pipe.log 2>&1
for pb in 59 160; do
  python bench.py --workload pacbio_native --no-extra-workloads --no-cpu-baseline --no-other-mode --no-fastq --steps 3 --e2e-pipe-blocks $pb 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); e=d['e2e']; print('native pipe $pb', round(e['value'],1), round(e['compress_GBps'],1), round(e['decompress_GBps'],1), e.get('blocks_per_call'))"
done >> gpurun_out/y_pipe.log 2>&1
echo done
