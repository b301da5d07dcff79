cd $GRAFT_REPO_ROOT
export PYTHONPATH=$GRAFT_REPO_ROOT
V=motionplanningtoolkit_b200/variants
( timeout 300 python tools/ab_np.py; for v in s0 s4 s5 s8; do MPTCC_LIB=$V/libmptcc_$v.so timeout 300 python tools/ab_np.py; done; timeout 300 python tools/ab_np.py ) > gpurun_out/ab17.txt 2>&1; grep "^libmptcc\|Error" gpurun_out/ab17.txt
