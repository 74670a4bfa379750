#!/bin/bash
# Everything one gpurun call should bring back: tests, the bench line, ncu evidence.
# usage: gpu_round.sh [tests] [bench] [bench4] [launches] [full:<kernel regex>,...]
mkdir -p gpurun_out
for what in "$@"; do
case $what in
tests)
  python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -25 > gpurun_out/pytest_gpu.txt; cat gpurun_out/pytest_gpu.txt;;
bench)
  python bench.py --steps 20 --warmup 3 --batch 1 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json;;
bench8)
  python bench.py --steps 10 --warmup 3 --batch 8 --no-cpu > gpurun_out/bench_b8.json 2> gpurun_out/bench_b8.err; tail -3 gpurun_out/bench_b8.err; cat gpurun_out/bench_b8.json;;
bench4)
  python bench.py --steps 10 --warmup 3 --batch 4 --no-cpu > gpurun_out/bench_b4.json 2> gpurun_out/bench_b4.err; tail -3 gpurun_out/bench_b4.err; cat gpurun_out/bench_b4.json;;
launches)
  timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches.csv python scripts/profile_step.py > gpurun_out/prof_step.log 2>&1; tail -2 gpurun_out/prof_step.log;;
launches4)
  timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches_b4.csv python scripts/profile_step.py 4 > gpurun_out/prof_step4.log 2>&1; tail -2 gpurun_out/prof_step4.log;;
traffic)
  timeout 1200 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
    --log-file gpurun_out/traffic.csv python scripts/profile_step.py 1 > gpurun_out/prof_traffic.log 2>&1; tail -2 gpurun_out/prof_traffic.log;;
gemm4)
  timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on \
    -k regex:gemm_tc05 -s 1 -c 4 -f -o gpurun_out/prof_gemm_b4 python scripts/profile_step.py 4 > gpurun_out/prof_gemm_b4.log 2>&1;;
full:*)
  for k in $(echo ${what#full:} | tr ',' ' '); do
    timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on \
      -k regex:$k -s 20 -c 2 -f -o gpurun_out/prof_$k python scripts/profile_step.py > gpurun_out/prof_$k.log 2>&1
  done;;
esac
done
