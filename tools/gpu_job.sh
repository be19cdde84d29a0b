cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/j18_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/j18_pytest.log
tail -4 gpurun_out/j18_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/j18_bench1.json 2> gpurun_out/j18_bench1.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/j18_bench2.json 2> gpurun_out/j18_bench2.err
for n in 1 2; do python - <<PY
import json
d=json.load(open('gpurun_out/j18_bench$n.json'))
print($n, 'ms/step', round(d['ms_per_step'],2), d['ms_each_step'], 'e2e', round(d['e2e']['ms_per_step'],2), 'checksum', d['result_checksum'], 'knn', d['knn_ms_over_ranks'])
print(d['stage_ms'], d['host_timers_ms'])
PY
done
