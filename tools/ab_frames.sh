#!/bin/bash
# A/B of alternative library builds on back-to-back frame time: tools/ab_frames.sh <frames> lib1.so[:--no-pdl] lib2.so ...
frames=$1; shift
for spec in "$@"; do
  so=${spec%%:*}; extra=""; [[ "$spec" == *:* ]] && extra=${spec#*:}
  echo "== $spec"
  FOVGS_LIB_PATH=$PWD/fov-3dgs_b200/lib/$so timeout ${AB_TIMEOUT:-150} python tools/frame_time.py --frames $frames $extra 2>&1 | tail -1 | cut -c1-600
done
