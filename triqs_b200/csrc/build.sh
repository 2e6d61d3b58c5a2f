#!/bin/bash
# Builds libtqgpu.so (sm_100a only) in-tree.  Usage: build.sh [extra nvcc flags]
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=default --fmad=true"
OUT=../libtqgpu.so
SRCS="tq_ctx.cu tq_fft_plan.cu fourier_matsubara.cu fourier_pipe.cu fourier_pipe_planned0.cu fourier_pipe_planned1.cu fourier_pipe_planned2.cu fourier_pipe_planned3.cu fourier_pipe_heavy0.cu fourier_pipe_heavy1.cu fourier_pipe_heavy2.cu fourier_pipe_heavy3.cu fourier_col.cu tail_fit.cu lattice.cu lattice_pipe.cu dyson.cu interp.cu imfreq_ops.cu tight_binding.cu fourier_real.cu legendre.cu"
mkdir -p build
pids=()
for s in $SRCS; do
  o=build/${s%.cu}.o
  if [ ! -f "$o" ] || [ "$s" -nt "$o" ] || [ -n "$(find . -maxdepth 1 -name '*.cuh' -newer "$o" 2>/dev/null)" ] || [ ../../include/tqgpu.h -nt "$o" ]; then
    $NVCC $FLAGS "$@" -c "$s" -o "$o" &
    pids+=($!)
  fi
done
for p in "${pids[@]}"; do wait "$p"; done
$NVCC -shared -o $OUT build/*.o -lcudart -ldl
echo "built $(readlink -f $OUT)"
