#!/bin/bash
# development build with per-phase clock64() accounting (tools/phase_probe.py loads it)
mkdir -p tools/_build
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --shared -Xcompiler -fPIC -Xcompiler -fvisibility=default \
  -DSPW_PHASE_TIMING -o tools/_build/libspwgnn_phase.so spwgnn_b200/csrc/spwgnn.cu 2>&1 | grep -E "error" 
ls -la tools/_build/libspwgnn_phase.so
