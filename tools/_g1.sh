cd $GRAFT_REPO_ROOT
export PYTHONPATH=$GRAFT_REPO_ROOT
timeout 600 python tools/perf_e2e.py first_half=2 first_half=1 > gpurun_out/e2e8.txt 2>&1; cat gpurun_out/e2e8.txt
