#!/bin/bash
# build-variant sweep on the GPU box: each argument is one set of extra nvcc flags (quote it)
for cfg in "$@"; do
  IDN_NVCC_EXTRA="$cfg" python -c "from idencomp_b200 import build; build.build_gpu(force=True)" 2>/dev/null
  echo "== [$cfg]"; tools/qb.sh $QB_ARGS
done
