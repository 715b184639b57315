#!/bin/bash
# round 2, GPU call C: full parity suite (pipelined host calls, host mirror workers), cache-hint variants, the new bench
mkdir -p gpurun_out
( time python -m pytest tests -x -q -m gpu -p no:cacheprovider 2>&1 | tail -25 ) > gpurun_out/c_pytest.log 2>&1
export QB_ARGS="--reads 12000000"
( tools/var_sweep.sh "-DIDN_ACID_NOALLOC" "-DIDN_ACID_NOALLOC -DIDN_QWIN_ALL" ) > gpurun_out/c_sweep.log 2>&1
python -c "from idencomp_b200 import build; build.build_gpu(force=True)"
( time python bench.py --steps 5 --warmup 3 > gpurun_out/c_bench.json 2> gpurun_out/c_bench.err ) > gpurun_out/c_bench.time 2>&1
echo "bench rc=$?" >> gpurun_out/c_bench.time
( time python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/c_ref.json 2> gpurun_out/c_ref.err ) >> gpurun_out/c_bench.time 2>&1
echo done
