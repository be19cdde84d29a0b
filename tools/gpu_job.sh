cd $GRAFT_REPO_ROOT
run() { n=$1; wl=$2; k=$3; w=$4; out=$5; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $n --workload $wl --steps $k --warmup $w > gpurun_out/$out.json 2> gpurun_out/$out.err; python - <<PY
import json
try:
    d=json.load(open('gpurun_out/$out.json'))
    print('$out', 'ms/step', round(d['ms_per_step'],2), [round(x,1) for x in d['ms_each_step']][:6], 'e2e', round(d['e2e']['ms_per_step'],2), d['result_checksum'], d['knn_ms_over_ranks'])
    print('   ', d['stage_ms'], d['host_timers_ms'])
except Exception as e:
    print('$out failed', e); print(open('gpurun_out/$out.err').read()[-1500:])
PY
}
run 8 C4 20 5 r02_bench_c4_8gpu
run 4 C4 20 5 r02_bench_c4_4gpu
run 2 C4 20 5 r02_bench_c4_2gpu
