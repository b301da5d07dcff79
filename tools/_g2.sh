cd $GRAFT_REPO_ROOT
export PYTHONPATH=$GRAFT_REPO_ROOT
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 50 --warmup 5 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; tail -c 300 gpurun_out/bench_n2.json
timeout 600 python tools/perf_edges.py > gpurun_out/edges_r2f.txt 2>&1; tail -5 gpurun_out/edges_r2f.txt
timeout 600 python tools/perf_latency.py > gpurun_out/latency_r2f.txt 2>&1; tail -12 gpurun_out/latency_r2f.txt
timeout 600 python bench.py --config C5-terrain > gpurun_out/terrain_r2f.txt 2>&1; tail -c 700 gpurun_out/terrain_r2f.txt
