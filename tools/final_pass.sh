#!/bin/bash
# Round-end evidence on one B200 (run through gpurun from the repo root): GPU test suite, bench lines of both arms, ncu launch
# list + one --set full capture of the same bench command. Outputs land in gpurun_out/ (copied into profiles/ afterwards).
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -3 > gpurun_out/final_tests.log
python bench.py --gpus 1 --steps 50 --warmup 5 > gpurun_out/final_bench_n1.json 2> gpurun_out/final_bench_n1.err
python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/final_bench_ref.json 2> gpurun_out/final_bench_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/final_launches.csv \
    python bench.py --gpus 1 --steps 2 --warmup 3 --no-cpu > gpurun_out/final_ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"^k_" --launch-skip 60 -c 12 -o gpurun_out/final_prof -f \
    python bench.py --gpus 1 --steps 2 --warmup 3 --no-cpu > gpurun_out/final_ncu_full.log 2>&1
cat gpurun_out/final_tests.log
tail -c 600 gpurun_out/final_bench_n1.json
tail -c 400 gpurun_out/final_bench_ref.json
tail -2 gpurun_out/final_ncu_full.log
