#!/bin/bash
# Development aid: device-resident C3 timings (tools/quick_bench.py) of several builds of the library, one after the other.
# usage: tools/variant_bench.sh name=path.so ...   (results in gpurun_out/variants.log)
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
: > gpurun_out/variants.log
for v in "$@"; do
  name="${v%%=*}"; path="${v#*=}"
  for cf in 3 1; do
    echo "== $name cf=$cf" >> gpurun_out/variants.log
    PNFFT_B200_LIB="$PWD/$path" timeout 300 python tools/quick_bench.py 256 16777216 $cf 2>&1 | tail -3 >> gpurun_out/variants.log
  done
done
grep -E "==|b_kernel" gpurun_out/variants.log | sed -e "s/'binning.*//" 
