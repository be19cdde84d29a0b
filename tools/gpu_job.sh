# What a round's end-of-work check runs on a 1-GPU box (gpurun -- 'bash tools/gpu_job.sh'): the GPU parity suite, the smoke
# test and one short bench line.  (During development this file is overwritten with the job of the moment.)
cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python __graft_entry__.py smoke 2>&1 | tail -1
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/b.json 2> gpurun_out/bench.err || tail -5 gpurun_out/bench.err
python -c "
import json; d=json.load(open('gpurun_out/b.json')); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['stage_ms'], d['result_checksum'])"
