#!/bin/bash
# round 2, final GPU call: the whole GPU suite, smoke, and the driver-style line of the final build
mkdir -p gpurun_out
( time python -m pytest tests -q -m gpu -p no:cacheprovider 2>&1 | tail -6 ) > gpurun_out/x_pytest.log 2>&1
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/x_smoke.log 2>&1
( time python bench.py --steps 20 --warmup 5 > gpurun_out/x_bench.json 2> gpurun_out/x_bench.err ) > gpurun_out/x_bench.time 2>&1
echo done
