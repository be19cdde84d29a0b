cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/j10_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/j10_pytest.log
tail -3 gpurun_out/j10_pytest.log
timeout 600 python tools/knn_tournament.py 10000000 base > gpurun_out/j10_tour.log 2>&1
cat gpurun_out/j10_tour.log
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/j10_bench.json 2> gpurun_out/j10_bench.err; tail -2 gpurun_out/j10_bench.err
python -c "
import json; d=json.load(open('gpurun_out/j10_bench.json')); print({k:d[k] for k in ('value','ms_per_step','result_checksum','stage_ms','hbm_passes')}); print(d['e2e']); print(d['roofline']['frac'], d['cpu_baseline'])"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:knn_fast -s 3 -c 1 -o gpurun_out/r02_knn_fast_c4_v3 -f python tools/prof_knn.py 10000000 > gpurun_out/j10_ncu.log 2>&1
tail -2 gpurun_out/j10_ncu.log
