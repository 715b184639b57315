#!/bin/bash
# round 2, GPU call Z: automatic sub-chunk size for long reads (library default and bench); pipeline tests
mkdir -p gpurun_out
( time python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_parity.py -q -m gpu -p no:cacheprovider 2>&1 | tail -6 ) > gpurun_out/z_pytest.log 2>&1
( python bench.py --workload pacbio --no-extra-workloads --no-cpu-baseline --no-other-mode --no-fastq --steps 3 2>gpurun_out/z_err1.log | python -c "import json,sys; d=json.load(sys.stdin); e=d['e2e']; print('pacbio compat', round(e['value'],1), round(e['compress_GBps'],1), round(e['decompress_GBps'],1), e.get('pipeline_blocks'))"
  python bench.py --no-extra-workloads --no-cpu-baseline --no-other-mode --no-fastq --steps 3 2>gpurun_out/z_err2.log | python -c "import json,sys; d=json.load(sys.stdin); e=d['e2e']; print('hiseq', round(d['value'],1), round(e['value'],1), e.get('pipeline_blocks'))"
) > gpurun_out/z_e2e.log 2>&1
echo done
