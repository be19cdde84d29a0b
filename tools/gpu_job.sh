cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_c4_1gpu.json 2> gpurun_out/bench.err; python -c "
import json; d=json.load(open('gpurun_out/r02_bench_c4_1gpu.json')); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['stage_ms'], d['result_checksum'], d['roofline']['frac'], d['cpu_baseline']['value'])"
python bench.py --workload C3 --steps 20 --warmup 5 > gpurun_out/r02_bench_c3_1gpu.json 2>> gpurun_out/bench.err; python -c "
import json; d=json.load(open('gpurun_out/r02_bench_c3_1gpu.json')); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['stage_ms'], d['result_checksum'])"
python bench.py --workload C2 --steps 20 --warmup 5 > gpurun_out/r02_bench_c2_1gpu.json 2>> gpurun_out/bench.err; python -c "
import json; d=json.load(open('gpurun_out/r02_bench_c2_1gpu.json')); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['stage_ms'], d['result_checksum'])"
