cd $GRAFT_REPO_ROOT
export PYTHONPATH=$GRAFT_REPO_ROOT
V=motionplanningtoolkit_b200/variants
( timeout 300 python tools/ab_np.py; timeout 300 python tools/ab_np.py 262144; for v in a8 a64 a24p2k off; do MPTCC_LIB=$V/libmptcc_$v.so timeout 300 python tools/ab_np.py;  MPTCC_LIB=$V/libmptcc_$v.so timeout 300 python tools/ab_np.py 262144; done ) > gpurun_out/ab13.txt 2>&1; cat gpurun_out/ab13.txt
