#!/bin/bash
# staging parity tests, then compute-sanitizer (memcheck, racecheck) over the kernel-level tests -> gpurun_out/s2_sanitizer_*.log
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_staging_gpu.py -x -q > gpurun_out/s2_staging_tests.log 2>&1; echo "staging rc=$?"; tail -3 gpurun_out/s2_staging_tests.log
which compute-sanitizer || export PATH=$PATH:/usr/local/cuda/bin
CS="compute-sanitizer --error-exitcode 9 --launch-timeout 120"
t0=$(date +%s)
timeout 900 $CS --tool memcheck python -m pytest tests/test_fused_gpu.py tests/test_staging_gpu.py -x -q -p no:cacheprovider > gpurun_out/s2_sanitizer_memcheck_fused.log 2>&1; echo "memcheck fused rc=$? ($(( $(date +%s) - t0 )) s)"; tail -4 gpurun_out/s2_sanitizer_memcheck_fused.log
t0=$(date +%s)
timeout 900 $CS --tool memcheck python -m pytest tests/test_conv_gpu.py -x -q -p no:cacheprovider -k "not 256x28 and not 256x14x14x128x128 and not 256x7" > gpurun_out/s2_sanitizer_memcheck_conv.log 2>&1; echo "memcheck conv rc=$? ($(( $(date +%s) - t0 )) s)"; tail -4 gpurun_out/s2_sanitizer_memcheck_conv.log
t0=$(date +%s)
timeout 600 $CS --tool memcheck python -m pytest tests/test_utt_gpu.py -x -q -p no:cacheprovider -k "kernels" > gpurun_out/s2_sanitizer_memcheck_utt.log 2>&1; echo "memcheck utt rc=$? ($(( $(date +%s) - t0 )) s)"; tail -4 gpurun_out/s2_sanitizer_memcheck_utt.log
t0=$(date +%s)
timeout 900 $CS --tool racecheck python -m pytest tests/test_fused_gpu.py tests/test_staging_gpu.py -x -q -p no:cacheprovider > gpurun_out/s2_sanitizer_racecheck_fused.log 2>&1; echo "racecheck fused rc=$? ($(( $(date +%s) - t0 )) s)"; tail -4 gpurun_out/s2_sanitizer_racecheck_fused.log
t0=$(date +%s)
timeout 600 $CS --tool racecheck python -m pytest tests/test_utt_gpu.py -x -q -p no:cacheprovider -k "kernels" > gpurun_out/s2_sanitizer_racecheck_utt.log 2>&1; echo "racecheck utt rc=$? ($(( $(date +%s) - t0 )) s)"; tail -4 gpurun_out/s2_sanitizer_racecheck_utt.log
TAG="only_audio" MML_SKIP_ENCODER=image python tools/step_time.py 2>&1 | tail -5
