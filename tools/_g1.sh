cd $GRAFT_REPO_ROOT
export PYTHONPATH=$GRAFT_REPO_ROOT
(
timeout 300 python tools/ab_np.py
) > gpurun_out/ab7.txt 2>&1
cat gpurun_out/ab7.txt
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -x -q -m gpu > gpurun_out/np3_parity.txt 2>&1; tail -3 gpurun_out/np3_parity.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:narrowphase3 -s 3 -c 1 -f -o gpurun_out/np3_c python tools/ab_np.py > gpurun_out/ncu_np3.log 2>&1; tail -2 gpurun_out/ncu_np3.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:broadphase -s 3 -c 1 -f -o gpurun_out/bp_c python tools/ab_np.py > gpurun_out/ncu_bp.log 2>&1; tail -2 gpurun_out/ncu_bp.log
