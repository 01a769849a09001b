#!/bin/bash
# warm-cache per-kernel durations of the image-only and audio-only steps (ncu --cache-control none): kernel time vs launch gaps
mkdir -p gpurun_out
for enc in audio image; do
  MML_SKIP_ENCODER=$enc ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 3000 --csv --log-file gpurun_out/s2_warm_skip_${enc}.csv python tools/step_once.py > gpurun_out/s2_warm_skip_${enc}.log 2>&1
  echo "ncu rc=$? ($enc skipped)"; tail -1 gpurun_out/s2_warm_skip_${enc}.log
done
