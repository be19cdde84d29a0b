cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench_c4_1gpu.json 2> gpurun_out/bench.err; python -c "
import json; d=json.load(open('gpurun_out/r02_bench_c4_1gpu.json')); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['stage_ms'], d['result_checksum'], d['roofline']['frac'], d['roofline']['traffic'])"
