#!/bin/bash
# One GPU session's worth of evidence for profiles/: tests, both bench arms, launch list of the bench command, DRAM traffic of the
# bench frames, one `ncu --set full` capture of the frame kernels.  Usage (on the GPU box): bash tools/round_run.sh <tag>
tag=${1:-r2}; out=gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > $out/${tag}_gpu_tests.log
timeout 600 python bench.py > $out/${tag}_bench_ours.json 2> $out/${tag}_bench_ours.err
timeout 600 python bench.py --impl reference > $out/${tag}_bench_reference.json 2> $out/${tag}_bench_reference.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_" -c 600 --csv --log-file $out/${tag}_launches_bench_fov_6M_ours.csv \
    python bench.py --steps 2 --warmup 3 --no-extra > $out/${tag}_bench_under_ncu.log 2>&1
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:"k_" -c 400 --csv \
    --log-file $out/${tag}_traffic.csv python tools/stage_times.py --variant fov --first 5 --frames 18 > $out/${tag}_traffic.log 2>&1
python tools/traffic.py $out/${tag}_traffic.csv 5 18 > $out/${tag}_traffic_summary.log 2>&1
cp profiles/roofline_traffic.json $out/${tag}_roofline_traffic.json 2>/dev/null
# third frame of tools/one_frame.py (gaze centre): k_pre, k_tile_scan, k_color_tma, k_scatter, k_lazy_blend x 2
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_pre|k_tile_scan|k_color_tma|k_scatter|k_lazy_blend" --launch-skip 12 -c 6 \
    -o $out/${tag}_ncu_full -f python tools/one_frame.py --size big --variant fov --frames 3 > $out/${tag}_ncu_full.log 2>&1
tail -3 $out/${tag}_gpu_tests.log; cat $out/${tag}_bench_ours.json | cut -c1-400; cat $out/${tag}_bench_reference.json | cut -c1-300
