#!/bin/bash
# Development aid (GPU box): rebuild the library with tuning macros and run the decode sweep for each variant.
# usage: tools/ab_build_sweep.sh "name1|flags1" "name2|flags2" ...
for v in "$@"; do
  name="${v%%|*}"; flags="${v#*|}"
  RD_EXTRA_NVCC_FLAGS="$flags" python -m radialog_b200.build --force > /dev/null 2>gpurun_out/ab_build_$name.err || { echo "$name: build failed"; tail -3 gpurun_out/ab_build_$name.err; continue; }
  echo "== $name ($flags)"
  RD_EXTRA_NVCC_FLAGS="$flags" timeout 300 python tools/decode_sweep.py 32 96 2>&1 | grep "ms/step" | head -2
done
python -m radialog_b200.build --force > /dev/null 2>&1
