#!/bin/bash
# A/B measurements of alternative builds: tools/ab_run.sh <variant> <frames> lib1.so lib2.so ...   (libs under fov-3dgs_b200/lib,
# built with `make -C fov-3dgs_b200 VARIANT=_x EXTRA=-D...`).  Every run is bounded (a hung experimental kernel must not eat the box).
variant=$1; frames=$2; shift 2
for so in "$@"; do
  echo "== $so"
  FOVGS_LIB_PATH=$PWD/fov-3dgs_b200/lib/$so timeout ${AB_TIMEOUT:-150} python tools/stage_times.py --variant $variant --frames $frames --size ${AB_SIZE:-big} 2>&1 | tail -1 | grep -o "stages {[^}]*}" || echo "   FAILED or timed out"
done
