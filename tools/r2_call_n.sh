#!/bin/bash
# round 2: randomised parity run (tools/fuzz_gpu.py) on the final build
mkdir -p gpurun_out
timeout 260 python tools/fuzz_gpu.py --seconds 200 --seed 31000000 > gpurun_out/fuzz5.log 2>&1
echo "exit $?" >> gpurun_out/fuzz5.log
echo done
