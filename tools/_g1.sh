cd $GRAFT_REPO_ROOT
export PYTHONPATH=$GRAFT_REPO_ROOT
timeout 900 ncu --set full --clock-control none --import-source on -k regex:narrowphase3 -s 3 -c 1 -f -o gpurun_out/prof_r2f_narrowphase python bench.py --steps 2 --warmup 3 --no-other-configs > gpurun_out/ncu_np.log 2>&1; tail -1 gpurun_out/ncu_np.log
