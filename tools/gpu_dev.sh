#!/bin/bash
# dev tool: gpurun without the reference test-suite data (80 MB) in the snapshot; the tracked .gpurunignore is restored
# afterwards whatever happens.  usage: tools/gpu_dev.sh [gpurun options] -- 'command'
cd "$(dirname "$0")/.."
cp .gpurunignore /tmp/.gpurunignore.keep
trap 'cp /tmp/.gpurunignore.keep .gpurunignore' EXIT
printf 'oracle/_ref/test_run\noracle/_ref/test_ray_cuda\n' >> .gpurunignore
/usr/local/graft/bin/gpurun "$@"
