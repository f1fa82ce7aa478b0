#!/bin/bash
# The command BASELINE.json names for configs[4], with frame files: render.py --video --orbit -r fhd.
# Device-side PNG streams (default) against the host-side encoder (BHR_PNG_DEVICE=0).  Without imageio the frame files
# are muxed into a QuickTime 'png ' movie (mov.py); the movie is read back with OpenCV and compared with the files.
N=${1:-1800}
OUT=/tmp/bhr_cli_video
rm -rf $OUT; mkdir -p $OUT
df -h /tmp | tail -1
for MODE in 1 0; do
  rm -rf $OUT/*  $OUT/.frames_*
  T0=$(date +%s.%N)
  BHR_PNG_DEVICE=$MODE python render.py --video --orbit --n_frames $N --fps 36 -r fhd -o $OUT/orbit.mp4 > $OUT/log_$MODE.txt 2> $OUT/time_$MODE.txt
  echo "=== BHR_PNG_DEVICE=$MODE rc=$? wall $(python -c "print(round($(date +%s.%N) - $T0, 2))") s for $N frames (process start to exit) ==="
  grep -E "frames/s|Session rendered|imageio|Warning|Muxed|Video saved" $OUT/log_$MODE.txt | tail -5
  tail -3 $OUT/time_$MODE.txt
  D=$(ls -d $OUT/.frames_* | head -1)
  echo "files: $(ls $D/*.png | wc -l), bytes: $(du -sb $D | cut -f1)"
  python - "$D" $OUT/orbit.mp4 <<'PY'
import os, sys, numpy as np
from PIL import Image
d = sys.argv[1]
if os.path.exists(sys.argv[2]):
    import cv2
    cap = cv2.VideoCapture(sys.argv[2])
    n, same = int(cap.get(cv2.CAP_PROP_FRAME_COUNT)), 0
    for f in range(min(n, 120)):
        ok, bgr = cap.read()
        same += bool(ok) and np.array_equal(bgr[..., ::-1], np.array(Image.open(f"{d}/frame_{f:04d}.png")))
    print(f"movie: {os.path.getsize(sys.argv[2])} bytes, {n} frames at {cap.get(cv2.CAP_PROP_FPS)} fps, "
          f"first {min(n, 120)} decoded frames identical to the PNG files: {same}")
for f in (0, 59, 60, 599):
    a = np.array(Image.open(f"{d}/frame_{f:04d}.png"))
    print(f"frame {f}: {a.shape} mean {a.mean():.3f} sha {hash(a.tobytes()) & 0xffffffff:08x}")
PY
done
