cd $GRAFT_REPO_ROOT
export PYTHONPATH=$GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/gpu_tests.txt 2>&1; tail -2 gpurun_out/gpu_tests.txt
timeout 900 python bench.py --steps 100 --warmup 5 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -c 200 gpurun_out/bench_final.json; tail -2 gpurun_out/bench_final.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2f.csv python bench.py --steps 5 --warmup 3 --no-other-configs > gpurun_out/ncu_launches.log 2>&1; tail -1 gpurun_out/ncu_launches.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:narrowphase3 -s 3 -c 1 -f -o gpurun_out/prof_r2f_narrowphase python bench.py --steps 2 --warmup 3 --no-other-configs > gpurun_out/ncu_np.log 2>&1; tail -1 gpurun_out/ncu_np.log
