#!/bin/bash
# round 2, GPU call N: randomised parity run (tools/fuzz_gpu.py)
mkdir -p gpurun_out
timeout 420 python tools/fuzz_gpu.py --seconds 330 --seed 1000 > gpurun_out/fuzz.log 2>&1
echo "exit $?" >> gpurun_out/fuzz.log
echo done
