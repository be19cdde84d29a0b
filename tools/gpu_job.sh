cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/j14_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/j14_pytest.log
tail -5 gpurun_out/j14_pytest.log
timeout 600 python tools/knn_tournament.py 10000000 base sm10 sm14 > gpurun_out/j14_tour.log 2>&1
cat gpurun_out/j14_tour.log
for lib in "" _bot5 _bot6; do AL_LIB=$PWD/astrolink_b200/libastrolink_b200$lib.so timeout 200 python tools/prof_tree.py halo 10000000 2>&1 | tail -1 | cut -c1-80 | sed "s/^/lib$lib: /"; done > gpurun_out/j14_tree.log
AL_TREE_BOT_CBITS=8 timeout 200 python tools/prof_tree.py halo 10000000 2>&1 | tail -1 | cut -c1-80 | sed "s/^/botcbits8: /" >> gpurun_out/j14_tree.log
cat gpurun_out/j14_tree.log
timeout 300 python tools/prof_agg.py 10000000 2>&1 | tail -3
