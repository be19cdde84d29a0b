set -x
cd $GRAFT_REPO_ROOT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/j1_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/j1_pytest.log
tail -5 gpurun_out/j1_pytest.log
timeout 300 python tools/knn_tournament.py 10000000 base > gpurun_out/j1_tour_fast.log 2>&1
AL_KNN_GENERIC=1 timeout 300 python tools/knn_tournament.py 10000000 base > gpurun_out/j1_tour_generic.log 2>&1
cat gpurun_out/j1_tour_fast.log gpurun_out/j1_tour_generic.log
AL_LIB=$PWD/astrolink_b200/libastrolink_b200_stats.so timeout 300 python tools/prof_knn.py 10000000 > gpurun_out/j1_stats.log 2>&1
cat gpurun_out/j1_stats.log
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/j1_bench.json 2> gpurun_out/j1_bench.err
cat gpurun_out/j1_bench.json
timeout 900 ncu --set full --clock-control none --import-source on -k regex:knn_fast -s 3 -c 1 -o gpurun_out/r02_knn_fast_c4 -f python tools/prof_knn.py 10000000 > gpurun_out/j1_ncu.log 2>&1
tail -3 gpurun_out/j1_ncu.log
