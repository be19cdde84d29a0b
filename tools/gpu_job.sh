cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/b.json 2> gpurun_out/bench.err || tail -5 gpurun_out/bench.err; python -c "
import json; d=json.load(open('gpurun_out/b.json')); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['stage_ms'], d['result_checksum'])"
python bench.py --workload C3 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/b3.json 2>> gpurun_out/bench.err; python -c "
import json; d=json.load(open('gpurun_out/b3.json')); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['stage_ms'], d['result_checksum'])"
