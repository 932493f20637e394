#!/bin/bash
# Builds A/B variants of libkh_b200.so into gpurun_variants/libkh_<name>.so (they travel to the GPU box; tools/ab.sh runs them).
#   tools/build_variants.sh name1:"-DFLAG=1 -DOTHER=2" name2:"..."
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
mkdir -p "$ROOT/gpurun_variants"
for v in "$@"; do
  name=${v%%:*}; flags=${v#*:}
  (
    d=$(mktemp -d /tmp/khvar_${name}_XXXX)
    mkdir -p $d/keyhunt_b200 $d/include
    cp -r "$ROOT/keyhunt_b200/csrc" $d/keyhunt_b200/csrc
    cp "$ROOT/include/keyhunt_b200.h" $d/include/
    find $d/keyhunt_b200/csrc -name '*.o' -delete
    make -s -j4 -C $d/keyhunt_b200/csrc ../libkh_b200.so EXTRA="$flags" >/dev/null
    cp $d/keyhunt_b200/libkh_b200.so "$ROOT/gpurun_variants/libkh_${name}.so"
    python3 "$ROOT/tools/ptxas_summary.py" $d/keyhunt_b200/csrc | grep -E "kernel<0, false|kernel<3, false, false|kernel<1, false, false|kernel<4, false|giant|baby" | sed "s/^/[$name] /"
  ) &
done
wait
ls -la "$ROOT/gpurun_variants"
