set -x
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/j3_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/j3_pytest.log
tail -4 gpurun_out/j3_pytest.log
timeout 600 python tools/knn_tournament.py 10000000 base o0 s0 os0 > gpurun_out/j3_tour.log 2>&1
cat gpurun_out/j3_tour.log
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/j3_launches_c4.csv python tools/prof_pipeline.py C4 > gpurun_out/j3_ncu_pipeline.log 2>&1
tail -3 gpurun_out/j3_ncu_pipeline.log
AL_PROFILE=1 timeout 300 python tools/prof_agg.py 10000000 2>&1 | tail -14 > gpurun_out/j3_agg.log
cat gpurun_out/j3_agg.log
