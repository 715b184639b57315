#!/bin/bash
# round 2, GPU call H (8 GPUs): multi-device tests, PCIe matrix, bench under torchrun at N = 2 and 8
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/h_topo.txt 2>&1
( time timeout 240 python -m pytest tests/test_gpu_pipeline.py -q -m gpu -p no:cacheprovider 2>&1 | tail -15 ) > gpurun_out/h_pytest.log 2>&1
timeout 180 python tools/pcie_matrix.py > gpurun_out/h_pcie.json 2> gpurun_out/h_pcie.err
for n in 8 2; do
  ( time timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) bench.py --gpus $n --steps 3 --warmup 3 > gpurun_out/h_bench_$n.json 2> gpurun_out/h_bench_$n.err ) > gpurun_out/h_bench_$n.time 2>&1
  tail -3 gpurun_out/h_bench_$n.err >> gpurun_out/h_bench_$n.time
done
echo done
