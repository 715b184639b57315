#!/bin/bash
# round 2, GPU call Q: the final build as the driver will run it: GPU suite, smoke, default bench, reference arm, launch list
mkdir -p gpurun_out
( time python -m pytest tests -q -m gpu -p no:cacheprovider 2>&1 | tail -8 ) > gpurun_out/q_pytest.log 2>&1
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/q_smoke.log 2>&1
( time python bench.py > gpurun_out/q_bench.json 2> gpurun_out/q_bench.err ) > gpurun_out/q_bench.time 2>&1
( time python bench.py --impl reference > gpurun_out/q_bench_ref.json 2> gpurun_out/q_bench_ref.err ) > gpurun_out/q_bench_ref.time 2>&1
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra-workloads > gpurun_out/q_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r2_launches_final.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra-workloads > gpurun_out/q_ncu_list.log 2>&1
Q="--steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-other-mode --no-fastq --no-extra-workloads"
python bench.py --reads 4000000 $Q > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'^(decode|encode)_kernel' -s 6 -c 2 -o gpurun_out/prof_r2_final python bench.py --reads 4000000 $Q > gpurun_out/q_ncu_codec.log 2>&1
echo done
