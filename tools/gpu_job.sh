cd $GRAFT_REPO_ROOT
R=/tmp/ncu_r02; mkdir -p $R
ncu --profile-from-start off --clock-control none --import-source on --set full -k 'regex:agg_kernel|DeviceThreeWayPartitionKernel' -c 300 -f -o $R/aggregate python tools/prof_stage.py aggregate > gpurun_out/r02_ncu_aggregate.log 2>&1
python tools/ncu_summary.py $R/aggregate.ncu-rep --by-name > gpurun_out/r02_aggregate_ncu_summary.txt; head -30 gpurun_out/r02_aggregate_ncu_summary.txt | cut -c1-120
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_c4_1gpu.json 2> gpurun_out/bench.err; python -c "
import json; d=json.load(open('gpurun_out/r02_bench_c4_1gpu.json')); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['stage_ms'], d['result_checksum'], d['roofline']['frac'], d['cpu_baseline']['value'])"
python bench.py --workload C3 --steps 20 --warmup 5 > gpurun_out/r02_bench_c3_1gpu.json 2>> gpurun_out/bench.err; python -c "
import json; d=json.load(open('gpurun_out/r02_bench_c3_1gpu.json')); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['stage_ms'], d['result_checksum'])"
python bench.py --workload C2 --steps 20 --warmup 5 > gpurun_out/r02_bench_c2_1gpu.json 2>> gpurun_out/bench.err; python -c "
import json; d=json.load(open('gpurun_out/r02_bench_c2_1gpu.json')); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['stage_ms'], d['result_checksum'])"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_c4_reference_arm.json 2>> gpurun_out/bench.err; cut -c1-200 gpurun_out/r02_bench_c4_reference_arm.json
