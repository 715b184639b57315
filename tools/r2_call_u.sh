#!/bin/bash
# round 2, GPU call U (8 GPUs): multi-device tests and the default bench under torchrun at N = 8 on the final build
mkdir -p gpurun_out
( time timeout 240 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_host_api.py -q -m gpu -p no:cacheprovider 2>&1 | tail -8 ) > gpurun_out/u_pytest.log 2>&1
( time timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29508 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/u_bench_8.json 2> gpurun_out/u_bench_8.err ) > gpurun_out/u_bench_8.time 2>&1
tail -3 gpurun_out/u_bench_8.err >> gpurun_out/u_bench_8.time
echo done
