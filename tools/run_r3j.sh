mkdir -p gpurun_out/r3j
out=gpurun_out/r3j/ab.log
: > $out
TQ_LIB_PATH=tools/ab/libtqgpu_dual7.so python -m pytest tests/test_gpu_parity.py -m gpu -q -k "c2_blockgf or long_mesh_pipelined" 2>&1 | tail -3 >> $out
for v in default nw7 dual7; do
  echo "variant $v" >> $out
  if [ $v = default ]; then unset TQ_LIB_PATH; else export TQ_LIB_PATH=tools/ab/libtqgpu_$v.so; fi
  python tools/c2_kernels.py 2>&1 | tail -1 >> $out
  python tools/c2_kernels.py 2>&1 | tail -1 >> $out
done
