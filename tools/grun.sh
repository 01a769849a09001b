#!/bin/bash
# build first (refuse to spend GPU time on a stale library), then run the given command on the GPU box
set -e
cd "$(dirname "$0")/.."
make -C task-specific-pretraining-multimodal_b200/csrc -j8 > /tmp/mml_build.log 2>&1 || { grep -E "error" -A3 /tmp/mml_build.log | head -30; echo "BUILD FAILED"; exit 1; }
exec /usr/local/graft/bin/gpurun --timeout "${GRUN_TIMEOUT:-1800}" -- "$@"
