cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/j26_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/j26_pytest.log
tail -3 gpurun_out/j26_pytest.log
timeout 200 python tools/prof_tree.py halo 10000000 2>&1 | tail -1 | cut -c1-80
timeout 200 python tools/prof_tree.py gmm 1000000 2>&1 | tail -1 | cut -c1-80
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/j26_bench1.json 2> gpurun_out/j26_bench1.err
timeout 600 python bench.py --workload C2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/j26_c2.json 2> gpurun_out/j26_c2.err
for f in j26_bench1 j26_c2; do python - <<PY
import json
d=json.load(open('gpurun_out/$f.json'))
print('$f', 'ms/step', round(d['ms_per_step'],2), 'e2e', d['e2e'].get('ms_per_step'), d.get('result_checksum'), d.get('stage_ms'), d['host_timers_ms'])
PY
done
