cd $GRAFT_REPO_ROOT
python tools/knn_tournament.py 10000000 base F1 F2 F3 2>&1 | grep -v Warning
