#!/bin/bash
# TEST HARNESS ONLY: compile the kernel sources for the host against the CUDA shim.
set -e
here="$(cd "$(dirname "$0")" && pwd)"
root="$(cd "$here/../.." && pwd)"
g++ -O2 -std=c++20 -DGNSSACQ_EMU_BUILD -ffp-contract=off -fPIC -shared -pthread \
    -I "$here" -x c++ "$root/gnss-dsp-tools_b200/csrc/gnssacq.cu" -x c++ "$here/cuda_emu.cpp" \
    -o "$here/libgnssacq_emu.so"
