# final evidence round (one GPU): tests, smoke, bench line, launch lists
mkdir -p gpurun_out
python -m pytest tests -q -m gpu -rP > gpurun_out/r2_gpu_tests_full.txt 2>&1; tail -1 gpurun_out/r2_gpu_tests_full.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err; tail -1 gpurun_out/r2_bench_default.err
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err; head -c 600 gpurun_out/r2_bench_reference.json
python scripts/launch_times.py gpurun_out/lt_final.json 5
VS_RASTER_STREAMS=1 VS_PROFILE_STEP=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r2_launches_bench_b8.csv python bench.py --no-cpu --steps 2 --warmup 3 --train-batch 0 --stress-steps 0 > gpurun_out/r2_prof_bench.log 2>&1
VS_RASTER_STREAMS=1 VS_PROFILE_STEP=1 timeout 1200 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_train_b8.csv python scripts/profile_train.py 8 > gpurun_out/r2_prof_train.log 2>&1; tail -1 gpurun_out/r2_prof_train.log
python scripts/profile_train.py 8 > gpurun_out/r2_profile_train.log 2>&1
python scripts/prof_lpips.py > gpurun_out/r2_prof_lpips.txt 2>&1; grep images gpurun_out/r2_prof_lpips.txt
