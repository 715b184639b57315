#!/bin/bash
# like var_sweep.sh, plus DRAM bytes of one launch of kernel $KERNEL (regex) from ncu
for cfg in "$@"; do
  IDN_NVCC_EXTRA="$cfg" python -c "from idencomp_b200 import build; build.build_gpu(force=True)" 2>/dev/null
  echo "== [$cfg]"; tools/qb.sh $QB_ARGS
  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:"$KERNEL" -s 2 -c 1 --csv python bench.py --reads 4000000 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-other-mode --no-fastq 2>/dev/null | grep -E "dram__|gpu__time" | awk -F, '{print $(NF-2), $(NF-1), $NF}'
done
