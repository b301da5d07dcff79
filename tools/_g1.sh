cd $GRAFT_REPO_ROOT
export PYTHONPATH=$GRAFT_REPO_ROOT
( timeout 300 python tools/ab_np.py; timeout 300 python tools/ab_np.py qnode=0 ) > gpurun_out/ab18.txt 2>&1; grep "^libmptcc\|Error" gpurun_out/ab18.txt
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -x -q -m gpu > gpurun_out/np3_parity.txt 2>&1; tail -2 gpurun_out/np3_parity.txt
