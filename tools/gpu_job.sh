cd $GRAFT_REPO_ROOT
python bench.py --workload C5 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_c5_1gpu.json 2> gpurun_out/bench.err || tail -5 gpurun_out/bench.err; python -c "
import json; d=json.load(open('gpurun_out/r02_bench_c5_1gpu.json')); print(d['ms_per_step'], d['ms_each_step'], d['e2e']['ms_per_step'], d['stage_ms'], d['host_timers_ms'], d['result_checksum'])"
