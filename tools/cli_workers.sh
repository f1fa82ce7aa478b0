#!/bin/bash
# render.py --video with files: frames/s against the number of file-writer threads (BHR_PNG_WORKERS)
OUT=/tmp/bhr_cli_video; mkdir -p $OUT
for WK in 2 4 8 16; do
  rm -rf $OUT/* $OUT/.frames_*
  BHR_PNG_WORKERS=$WK python render.py --video --orbit --n_frames 1500 --fps 36 -r fhd -o $OUT/orbit.mp4 > $OUT/log.txt 2>&1
  echo "workers $WK: $(grep 'frames/s' $OUT/log.txt | tail -1)"
done
