#!/bin/bash
# Builds a library variant for A/B runs: tools/build_variant.sh <name> [extra nvcc flags] -> tools/ab/libtqgpu_<name>.so
# (run with TQ_LIB_PATH=tools/ab/libtqgpu_<name>.so; only the pipelined sources are rebuilt unless ALL=1)
set -e
name=$1; shift
cd "$(dirname "$0")/../triqs_b200/csrc"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=default --fmad=true"
B=../../tools/ab/build_$name
mkdir -p $B
PIPE="fourier_pipe.cu fourier_pipe_planned0.cu fourier_pipe_planned1.cu fourier_pipe_planned2.cu fourier_pipe_planned3.cu fourier_col.cu lattice_pipe.cu ${EXTRA_SRCS}"
if [ -n "$ALL" ]; then PIPE=$(ls *.cu); fi
pids=()
for s in $PIPE; do
  $NVCC $FLAGS "$@" -c "$s" -o $B/${s%.cu}.o &
  pids+=($!)
done
for p in "${pids[@]}"; do wait "$p"; done
objs=""
for o in build/*.o; do
  b=$(basename $o)
  if [ -f $B/$b ]; then objs="$objs $B/$b"; else objs="$objs $o"; fi
done
$NVCC -shared -o ../../tools/ab/libtqgpu_$name.so $objs -lcudart -ldl
echo "built tools/ab/libtqgpu_$name.so"
