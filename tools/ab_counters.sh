#!/bin/bash
# like ab_variants but also prints the counters
cd "$(dirname "$0")/../motionplanningtoolkit_b200/csrc" || exit 1
for v in "$@"; do
  rm -f collide_kernel.o split_kernels.o mptcc_api.o
  make EXTRA_NVFLAGS="$v" CXXFLAGS="-std=c++17 -O2 -fPIC -ffp-contract=off -fno-fast-math -I/usr/local/cuda/include $v" > /tmp/ab_build.log 2>&1 || { tail -5 /tmp/ab_build.log; continue; }
  echo "== $v"
  (cd ../..; python bench.py --steps 20 --warmup 3 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('   Mposes/s', round(d['value']/1e6,1), 'ms', round(d['ms_per_step'],3), d['roofline']['pipeline_ms'], 'mismatch', d['parity']['mismatches_vs_oracle'])"; python tools/gpu_counters.py 131072 2>&1 | grep "gpu counters")
done
rm -f collide_kernel.o split_kernels.o mptcc_api.o; make > /dev/null 2>&1
