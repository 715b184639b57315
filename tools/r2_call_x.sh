#!/bin/bash
# round 2, GPU call X: driver-style lines of the final build (both arms) and smoke
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/x_smoke.log 2>&1
( time python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/x_bench_ref.json 2> gpurun_out/x_bench_ref.err ) > gpurun_out/x_bench_ref.time 2>&1
( time python bench.py --steps 20 --warmup 5 > gpurun_out/x_bench.json 2> gpurun_out/x_bench.err ) > gpurun_out/x_bench.time 2>&1
echo done
