#!/bin/bash
# round 2, GPU call A: parity suite on the new kernels, build-variant sweep, ncu capture of the default build
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/a_gpu.txt
( time python -m pytest tests -x -q -m gpu -p no:cacheprovider 2>&1 | tail -25 ) > gpurun_out/a_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/a_pytest.log
export QB_ARGS="--reads 12000000"
( tools/var_sweep.sh "" "-DIDN_NO_SYM16 -DIDN_NO_AENC -DIDN_QROW352" "-DIDN_ENC_MINB=8 -DIDN_DEC_MINB=8" "-DIDN_NO_AENC" "-DIDN_QROW352" "-DIDN_NO_SYM16" "-DIDN_ENC_MINB=8 -DIDN_DEC_MINB=8 -DIDN_NO_AENC" ) > gpurun_out/a_sweep.log 2>&1
# native + novaseq with the default build
python -c "from idencomp_b200 import build; build.build_gpu(force=True)"
( QB_ARGS="" tools/qb.sh --reads 8000000 --workload novaseq150; tools/qb.sh --reads 8000000 --workload novaseq150 --mode compat; tools/qb.sh --reads 50000 --workload pacbio ) > gpurun_out/a_other.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'^(decode|encode)_kernel' -s 6 -c 2 -o gpurun_out/prof_r2a python bench.py --reads 4000000 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-other-mode --no-fastq > gpurun_out/a_ncu.log 2>&1
echo done
