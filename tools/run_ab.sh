# A/B of library variants on the C2 kernels: tools/run_ab.sh <out> <lib variants...>
out=$1; shift
mkdir -p gpurun_out
echo "variant default" >> $out
python tools/c2_kernels.py >> $out 2>&1
python tools/c2_kernels.py >> $out 2>&1
for v in "$@"; do
  echo "variant $v" >> $out
  TQ_LIB_PATH=tools/ab/libtqgpu_$v.so python tools/c2_kernels.py >> $out 2>&1
  TQ_LIB_PATH=tools/ab/libtqgpu_$v.so python tools/c2_kernels.py >> $out 2>&1
done
