#!/bin/bash
# Runs on the GPU box (gpurun): plain bench, then the ncu launch list of the same command, then
# one `--set full` capture per kernel family.  Outputs go to gpurun_out/; tools/write_profiles.py
# turns them into the summaries under profiles/.
TAG=${1:-r01}
set -x
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-other-configs --no-orbit --no-tiled > gpurun_out/${TAG}_bench_plain.log 2> gpurun_out/${TAG}_bench_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-other-configs --no-orbit --no-tiled > gpurun_out/${TAG}_ncu_launches.log 2>&1
python tools/prof_fhd.py fhd 0 3 > gpurun_out/${TAG}_prof_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:'raymarch_persistent|band_list|retrace' -s 6 -c 3 \
    -o gpurun_out/${TAG}_raymarch -f python tools/prof_fhd.py fhd 0 3 > gpurun_out/${TAG}_ncu_raymarch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'bloom|composite|flare_add' -s 6 -c 3 \
    -o gpurun_out/${TAG}_post -f python tools/prof_fhd.py fhd 0 3 > gpurun_out/${TAG}_ncu_post.log 2>&1
python tools/prof_fhd.py 4k 0 3 aa > gpurun_out/${TAG}_prof_4k_plain.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'raymarch_persistent' -s 1 -c 1 \
    -o gpurun_out/${TAG}_raymarch_4k_aa -f python tools/prof_fhd.py 4k 0 3 aa > gpurun_out/${TAG}_ncu_raymarch_4k.log 2>&1
python tools/prof_png.py > gpurun_out/${TAG}_png_plain.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'png_' -s 3 -c 3 \
    -o gpurun_out/${TAG}_png -f python tools/prof_png.py > gpurun_out/${TAG}_ncu_png.log 2>&1
[ -n "$SKIP_TEXTURE" ] && exit 0
python tools/video_breakdown.py > gpurun_out/${TAG}_video_plain.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'background_kernel|entity_accumulate|entity_table|compose_kernel|stats_rows|select_hist' -s 8 -c 6 \
    -o gpurun_out/${TAG}_texture -f python tools/video_breakdown.py > gpurun_out/${TAG}_ncu_texture.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_texture.log
