#!/bin/bash
# A/B of compile-time variants of the normal-search kernel (tools/build_variant.py) on the bench workload, one process each
mkdir -p gpurun_out
out=gpurun_out/ab_normals_variants.log
: > $out
for v in "" "$@"; do
  if [ -z "$v" ]; then lib=""; else lib="tools/_bin/libfm3d_$v.so"; fi
  echo "=== variant '${v:-base}'" >> $out
  FM3D_LIB=$lib AB_REPS=5 timeout 300 python tools/gpu_ab_normals.py 3995 '{}' >> $out 2>&1
done
# base once more at the end (drift check)
echo "=== variant 'base' (again)" >> $out
AB_REPS=5 timeout 300 python tools/gpu_ab_normals.py 3995 '{}' >> $out 2>&1
