#!/bin/bash
# Everything one gpurun call should bring back: tests, the bench line, ncu evidence.
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -15 > gpurun_out/pytest_gpu.txt
cat gpurun_out/pytest_gpu.txt
python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches.csv python scripts/profile_step.py > gpurun_out/prof_step.log 2>&1
tail -2 gpurun_out/prof_step.log
if [ "$1" == "full" ]; then
for k in gemm_tc05 attention_kernel blend_kernel preprocess_kernel; do
  timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on \
    -k regex:$k -s 20 -c 2 -f -o gpurun_out/prof_$k python scripts/profile_step.py > gpurun_out/prof_$k.log 2>&1
done
fi
