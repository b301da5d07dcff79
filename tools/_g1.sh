cd $GRAFT_REPO_ROOT
export PYTHONPATH=$GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/gpu_tests.txt 2>&1; tail -5 gpurun_out/gpu_tests.txt
timeout 600 python tools/perf_e2e.py > gpurun_out/e2e5.txt 2>&1; cat gpurun_out/e2e5.txt
