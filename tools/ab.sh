#!/bin/bash
# dev tool: A/B of kernel builds on the hall-250k workload (oracle-built BVH); usage: tools/ab.sh lib1.so lib2.so ...
mkdir -p gpurun_out
for lib in "$@"; do
  for fm in ${FIN_MINS:-8}; do
    RC_TRACE_FIN_MIN=$fm RC_DEV_CUDA_LIB=$lib timeout 600 python tools/profile_hall.py --spp ${SPP:-8} ${EXTRA} 2>&1 | tail -1 | tee -a gpurun_out/ab.log
  done
done
