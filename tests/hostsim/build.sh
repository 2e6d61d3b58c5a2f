#!/bin/bash
# TEST-ONLY: host emulation of the device math (tile FFT engine, small linear algebra).  Never linked into the product.
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
if [ ! -f libhostsim.so ] || [ hostsim.cu -nt libhostsim.so ] || [ -n "$(find ../../triqs_b200/csrc -name '*.cuh' -newer libhostsim.so)" ] || [ ../../triqs_b200/csrc/tq_fft_plan.cu -nt libhostsim.so ]; then
  $NVCC -O2 -std=c++17 -arch=sm_100a -shared -Xcompiler -fPIC -o libhostsim.so hostsim.cu ../../triqs_b200/csrc/tq_fft_plan.cu ../../triqs_b200/csrc/tq_ctx.cu -ldl
fi
echo "built tests/hostsim/libhostsim.so"
