cd $GRAFT_REPO_ROOT
python tools/knn_tournament.py 10000000 base D1 D2 D4 2>&1 | grep -v Warning
