cd $GRAFT_REPO_ROOT
export PYTHONPATH=$GRAFT_REPO_ROOT
( timeout 300 python tools/ab_np.py ; timeout 300 python tools/ab_np.py grid=0 ) > gpurun_out/ab9.txt 2>&1
cat gpurun_out/ab9.txt
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/gpu_tests.txt 2>&1; tail -3 gpurun_out/gpu_tests.txt
