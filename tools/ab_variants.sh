#!/bin/bash
# A/B build variants of the kernel on the GPU box: tools/ab_variants.sh "<flags A>" "<flags B>" ...
# OPTS="poses_per_warp=64" sets runtime options for every variant.
# Each variant rebuilds libmptcc.so with the given -D flags and runs a short bench.
cd "$(dirname "$0")/../motionplanningtoolkit_b200/csrc" || exit 1
for v in "$@"; do
  rm -f collide_kernel.o split_kernels.o mptcc_api.o
  make EXTRA_NVFLAGS="$v" CXXFLAGS="-std=c++17 -O2 -fPIC -ffp-contract=off -fno-fast-math -I/usr/local/cuda/include $v" > /tmp/ab_build.log 2>&1 || { tail -5 /tmp/ab_build.log; continue; }
  echo "== $v  $(grep -E 'Used [0-9]+ registers' /tmp/ab_build.log | head -1 | grep -oE 'Used [0-9]+ registers')"
  (cd ../..; MPTCC_OPTS="$OPTS" python bench.py --steps 10 --warmup 3 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('   Mposes/s', round(d['value']/1e6,1), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']/1e6,1), 'mismatch', d['parity']['mismatches_vs_oracle'])")
done
rm -f collide_kernel.o split_kernels.o mptcc_api.o; make > /dev/null 2>&1
