for so in libfovgs.so libfovgs_g256_d2048.so libfovgs_g1024_d2048.so libfovgs_g512_d1024.so libfovgs_g256_d768.so; do
  echo "== $so"
  FOVGS_LIB_PATH=$PWD/fov-3dgs_b200/lib/$so python tools/stage_times.py --variant fov --frames 18 2>&1 | tail -1 | grep -o "stages {[^}]*}"
  FOVGS_LIB_PATH=$PWD/fov-3dgs_b200/lib/$so python tools/stage_times.py --variant obb --frames 9 2>&1 | tail -1 | grep -o "stages {[^}]*}"
done
