ncu --metrics gpu__time_duration.sum --clock-control none -k regex:engine -c 12 --csv --log-file gpurun_out/launches_r1_v4.csv python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --no-breakdown > gpurun_out/ncu_l_v4.log 2>&1
for args in "--config paper64 --times 100" "--config hera350 --precision complex128" "--config ska512 --precision complex128" "--config ska512"; do
  python bench.py $args --no-e2e --no-cpu > gpurun_out/tmp.log 2>&1; tail -1 gpurun_out/tmp.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('$args', 'ms', round(d['ms_per_step'],3), 'frac', round(d['roofline']['frac'],4), 'value %.3e'%d['value'], {k:round(v['frac'],3) for k,v in d['roofline'].get('legs',{}).items()})" >> gpurun_out/bench_configs_v4.log
done
cat gpurun_out/bench_configs_v4.log
