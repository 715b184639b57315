#!/bin/bash
# round 2, GPU call B: position locality of the q-score tables
mkdir -p gpurun_out
export QB_ARGS="--reads 12000000"
( tools/var_sweep.sh "" "-DIDN_QWIN_ALL" "-DIDN_QWIN_NONE" "-DIDN_ENC_MINB=9 -DIDN_DEC_MINB=9 -DIDN_NO_SYM16" "-DIDN_ENC_MINB=7 -DIDN_DEC_MINB=7"
  echo "== no row sort"; python -c "from idencomp_b200 import build; build.build_gpu(force=True)"; IDN_NO_ROW_SORT=1 tools/qb.sh --reads 12000000 ) > gpurun_out/b_sweep.log 2>&1
python -c "from idencomp_b200 import build; build.build_gpu(force=True)"
( tools/qb.sh --reads 8000000 --workload novaseq150; tools/qb.sh --reads 8000000 --workload novaseq150 --mode compat; tools/qb.sh --reads 50000 --workload pacbio; tools/qb.sh --reads 12000000 --mode native; tools/qb.sh ) > gpurun_out/b_other.log 2>&1
( time python -m pytest tests -x -q -m gpu -p no:cacheprovider 2>&1 | tail -15 ) > gpurun_out/b_pytest.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'^(decode|encode)_kernel' -s 6 -c 2 -o gpurun_out/prof_r2b python bench.py --reads 4000000 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-other-mode --no-fastq > gpurun_out/b_ncu.log 2>&1
echo done
