#!/bin/bash
# One GPU call for the round-end evidence: tests, bench (both arms), launch list, full ncu capture.
python -m pytest tests -x -q -m gpu > gpurun_out/pytest_full_final.log 2>&1; tail -2 gpurun_out/pytest_full_final.log
python bench.py > gpurun_out/bench_final_r1.log 2>&1; tail -1 gpurun_out/bench_final_r1.log | cut -c1-200
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:engine -c 12 --csv --log-file gpurun_out/launches_r1_final.csv python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --no-breakdown > gpurun_out/ncu_l_final.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:engine_f32 -c 1 -s 3 -o gpurun_out/prof_r1_engine_v4d -f python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-breakdown > gpurun_out/ncu_v4d.log 2>&1; tail -1 gpurun_out/ncu_v4d.log
