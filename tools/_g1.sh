cd $GRAFT_REPO_ROOT
export PYTHONPATH=$GRAFT_REPO_ROOT
V=motionplanningtoolkit_b200/variants
( for v in c2 c4; do MPTCC_LIB=$V/libmptcc_$v.so timeout 300 python tools/ab_np.py;  MPTCC_LIB=$V/libmptcc_$v.so timeout 300 python tools/ab_np.py 262144; done ) > gpurun_out/ab11.txt 2>&1; cat gpurun_out/ab11.txt
( timeout 600 python tools/perf_e2e.py ramp=0; for v in c2 c4; do MPTCC_LIB=$V/libmptcc_$v.so timeout 600 python tools/perf_e2e.py ramp=0; done ) > gpurun_out/e2e3.txt 2>&1; cat gpurun_out/e2e3.txt
