python -m pytest tests -x -q -m gpu > gpurun_out/pytest_full_v4.log 2>&1; tail -3 gpurun_out/pytest_full_v4.log
python bench.py > gpurun_out/bench_full_v4.log 2>&1; tail -1 gpurun_out/bench_full_v4.log | cut -c1-300
python bench.py --no-e2e --no-cpu --no-breakdown --times 16 > gpurun_out/bench_t16_v4.log 2>&1; tail -1 gpurun_out/bench_t16_v4.log | cut -c1-200
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:engine -c 12 --csv --log-file gpurun_out/launches_r1_v4.csv python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --no-breakdown > gpurun_out/ncu_l_v4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:engine_f32 -c 1 -s 3 -o gpurun_out/prof_r1_engine_v4c -f python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-breakdown > gpurun_out/ncu_v4c.log 2>&1; tail -1 gpurun_out/ncu_v4c.log
