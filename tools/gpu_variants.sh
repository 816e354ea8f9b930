#!/bin/bash
# gpurun helper: time the batch-4096 training step with each experimental library variant (tools/build_variant.sh).
# usage: tools/gpu_variants.sh "<fuse bits>" tag1 tag2 ...   ("base" = the regular build)
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
fuse=$1; shift
cp streamz_b200/lib/libstreamz_b200.so /tmp/base.so
for tag in "$@"; do
  if [ "$tag" = base ]; then cp /tmp/base.so streamz_b200/lib/libstreamz_b200.so; else cp streamz_b200/lib_exp/libstreamz_b200_$tag.so streamz_b200/lib/libstreamz_b200.so; fi
  for f in $fuse; do
    SZB_STEP_FUSE=$f timeout 120 python tools/gpu_mlp_step.py 3xtf32 60 2>&1 | tail -1 | sed "s/^/$tag fuse=$f /" | tee -a gpurun_out/variants.txt
  done
done
cp /tmp/base.so streamz_b200/lib/libstreamz_b200.so
