#!/bin/bash
# round 2, GPU call P: per-read CRC partials computed by the encoder (backward CRC), A/B against the crc_read pass
mkdir -p gpurun_out
( time python -m pytest tests -q -m gpu -p no:cacheprovider -x 2>&1 | tail -8 ) > gpurun_out/p_pytest.log 2>&1
Q="--steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-other-mode --no-fastq --no-extra-workloads"
P="import json,sys; d=json.load(sys.stdin); print(round(d['value'],1), round(d['compress_GBps'],1), round(d['decompress_GBps'],1), d['gpu_launches'], {k:round(v,2) for k,v in list(d['roofline']['kernels_ms_per_step'].items())[:8]})"
( echo "== main, encoder CRC"; python bench.py $Q | python -c "$P"
  echo "== main, crc_read pass"; IDN_NO_ENC_CRC=1 python bench.py $Q | python -c "$P"
  echo "== pacbio, encoder CRC"; python bench.py --workload pacbio $Q | python -c "$P"
  echo "== select4"; python bench.py --workload hiseq100_select4 $Q | python -c "$P"
) > gpurun_out/p_bench.log 2>&1
timeout 150 python tools/fuzz_gpu.py --seconds 100 --seed 900000 > gpurun_out/fuzz3.log 2>&1
echo "exit $?" >> gpurun_out/fuzz3.log
echo done
