cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/j24_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/j24_pytest.log
tail -4 gpurun_out/j24_pytest.log
timeout 600 python tools/knn_tournament.py 10000000 base heap2 > gpurun_out/j24_tour.log 2>&1
cat gpurun_out/j24_tour.log
