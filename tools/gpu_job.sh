cd $GRAFT_REPO_ROOT
for nf in 1 2 4; do echo -n "nf $nf: "; AL_TREE_NF=$nf python tools/prof_tree.py halo 10000000 2>&1 | tail -1 | cut -c30-100; done
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
