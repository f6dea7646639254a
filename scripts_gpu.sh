cd $GRAFT_REPO_ROOT
RPT2=1 python tools/gather_timing.py 3 2>&1 | tail -9
