cd $GRAFT_REPO_ROOT
export PYTHONPATH=$GRAFT_REPO_ROOT
( timeout 300 python tools/ab_np.py; timeout 300 python tools/ab_np.py always_replay=1 ) > gpurun_out/ab14.txt 2>&1; cat gpurun_out/ab14.txt
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -x -q -m gpu > gpurun_out/np3_parity.txt 2>&1; tail -2 gpurun_out/np3_parity.txt
