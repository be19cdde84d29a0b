cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/j17_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/j17_pytest.log
tail -4 gpurun_out/j17_pytest.log
AL_PROFILE=1 timeout 300 python tools/prof_agg.py 10000000 2>&1 | tail -14
