cd $GRAFT_REPO_ROOT
export PYTHONPATH=$GRAFT_REPO_ROOT
V=motionplanningtoolkit_b200/variants
( timeout 300 python tools/ab_np.py; for v in w17 w18 w20; do MPTCC_LIB=$V/libmptcc_$v.so timeout 300 python tools/ab_np.py; done ) > gpurun_out/ab16.txt 2>&1; grep "^libmptcc\|Error" gpurun_out/ab16.txt
