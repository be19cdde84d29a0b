cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/j12_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/j12_pytest.log
tail -30 gpurun_out/j12_pytest.log
timeout 600 python tools/knn_tournament.py 10000000 base > gpurun_out/j12_tour.log 2>&1
cat gpurun_out/j12_tour.log
