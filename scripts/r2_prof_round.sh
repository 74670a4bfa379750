mkdir -p gpurun_out
python scripts/prof_wgrad.py 16 > gpurun_out/r2_wgrad_cases.txt 2>&1; cat gpurun_out/r2_wgrad_cases.txt
VS_PROFILE_STEP=1 timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -f -o gpurun_out/r2_full_wgrad python scripts/prof_wgrad.py 4 > gpurun_out/r2_prof_wgrad.log 2>&1; tail -2 gpurun_out/r2_prof_wgrad.log
VS_PROFILE_STEP=1 timeout 1200 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_train_b8.csv python scripts/profile_train.py 8 > gpurun_out/r2_prof_train.log 2>&1; tail -2 gpurun_out/r2_prof_train.log
python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err; tail -3 gpurun_out/r2_bench_default.err; head -c 400 gpurun_out/r2_bench_default.json
