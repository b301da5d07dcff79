#!/bin/bash
# GPU-side capture behind profiles/<tag>_*: usage (from the repository root, on a B200):
#   gpurun --timeout 1500 -- 'bash tools/capture_profiles.sh r2g'    then here:   python tools/make_profiles.py r2g
# Every ncu pass runs only after the same command has exited 0 without ncu; numbers printed under ncu are not bench values.
set -u
TAG=${1:-r2g}
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > $O/gpu_tests.txt
python bench.py > $O/bench_$TAG.json 2> $O/bench_$TAG.err || exit 1
python bench.py --impl reference > $O/bench_ref_$TAG.json 2>> $O/bench_$TAG.err
python bench.py --steps 5 --warmup 3 --no-other-configs > /dev/null 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/launches_$TAG.csv \
    python bench.py --steps 5 --warmup 3 --no-other-configs > $O/ncu_launches.log 2>&1
python bench.py --steps 2 --warmup 3 --no-other-configs > /dev/null 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:narrowphase3_kernel -s 3 -c 1 -f -o $O/prof_${TAG}_narrowphase \
    python bench.py --steps 2 --warmup 3 --no-other-configs > $O/ncu_np.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:broadphase_grid_kernel -s 3 -c 1 -f -o $O/prof_${TAG}_broadphase \
    python bench.py --steps 2 --warmup 3 --no-other-configs > $O/ncu_bp.log 2>&1
tail -c 600 $O/bench_$TAG.json
