cd $GRAFT_REPO_ROOT
export PYTHONPATH=$GRAFT_REPO_ROOT
V=motionplanningtoolkit_b200/variants
(
timeout 300 python tools/ab_np.py np3=1
timeout 300 python tools/ab_np.py np3=1 agent_top=0
timeout 300 python tools/ab_np.py np3=1 agent_top=511
for v in w12 w18 w20 w24; do MPTCC_LIB=$V/libmptcc_$v.so timeout 300 python tools/ab_np.py np3=1 agent_top=0; done
MPTCC_LIB=$V/libmptcc_w20.so timeout 300 python tools/ab_np.py np3=1
) > gpurun_out/ab3.txt 2>&1
cat gpurun_out/ab3.txt
MPTCC_OPTIONS=np3=1 timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/np3_parity.txt 2>&1; tail -3 gpurun_out/np3_parity.txt
