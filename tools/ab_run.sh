#!/bin/bash
# A/B measurements of alternative builds: tools/ab_run.sh <variant> <frames> lib1.so lib2.so ...   (libs under fov-3dgs_b200/lib)
variant=$1; frames=$2; shift 2
for so in "$@"; do
  echo "== $so"
  FOVGS_LIB_PATH=$PWD/fov-3dgs_b200/lib/$so python tools/stage_times.py --variant $variant --frames $frames 2>&1 | tail -1 | grep -o "stages {[^}]*}"
done
