#!/bin/bash
# round 2, GPU call L: 16-byte loads in crc_read, warp-cooperative assemble
mkdir -p gpurun_out
( time python -m pytest tests -q -m gpu -p no:cacheprovider -x 2>&1 | tail -8 ) > gpurun_out/l_pytest.log 2>&1
Q="--steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-other-mode --no-fastq --no-extra-workloads"
P="import json,sys; d=json.load(sys.stdin); print(round(d['value'],1), round(d['compress_GBps'],1), round(d['decompress_GBps'],1), {k:round(v,2) for k,v in list(d['roofline']['kernels_ms_per_step'].items())[:8]})"
( echo "== main"; python bench.py $Q | python -c "$P"
  echo "== main, thread-per-read assemble"; IDN_ASSEMBLE_THREAD=1 python bench.py $Q | python -c "$P"
  echo "== nova native 12M"; python bench.py --workload novaseq150_native --reads 12000000 $Q | python -c "$P"
  echo "== pacbio"; python bench.py --workload pacbio $Q | python -c "$P"
  echo "== pacbio native"; python bench.py --workload pacbio_native $Q | python -c "$P"
  echo "== select4"; python bench.py --workload hiseq100_select4 $Q | python -c "$P"
) > gpurun_out/l_bench.log 2>&1
echo done
