#!/bin/bash
# round 2, GPU call T: model-pair sweep on the final build
mkdir -p gpurun_out
bash tools/sweep_pairs.sh > gpurun_out/r2_sweep_pairs.md 2> gpurun_out/t_sweep.err
echo done
