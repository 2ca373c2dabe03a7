#!/bin/bash
# TEST HARNESS ONLY: compile the kernel sources for the host against the CUDA shim.
set -e
here="$(cd "$(dirname "$0")" && pwd)"
root="$(cd "$here/../.." && pwd)"
obj="$here/obj"
mkdir -p "$obj"
flags="-O2 -std=c++20 -DGNSSACQ_EMU_BUILD -ffp-contract=off -fPIC -pthread -I $here"
g++ $flags -c -x c++ "$root/gnss-dsp-tools_b200/csrc/gnssacq.cu" -o "$obj/gnssacq.o" &
g++ $flags -c -x c++ "$here/cuda_emu.cpp" -o "$obj/cuda_emu.o" &
for k in 0 1 2 3 4 5 6 7 8 9 10; do
  g++ $flags -DGNSSACQ_REG_PART=$k -c -x c++ "$root/gnss-dsp-tools_b200/csrc/registry.cu" -o "$obj/registry_$k.o" &
done
wait
g++ -shared -pthread "$obj"/*.o -o "$here/libgnssacq_emu.so" -ldl
