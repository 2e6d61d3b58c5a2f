mkdir -p gpurun_out/r3c
out=gpurun_out/r3c/knock.log
: > $out
for k in 0 1 16 128 256 401 2 4 6 40 46 447; do
  echo "knock $k" >> $out
  TQ_KNOCK=$k TQ_LIB_PATH=tools/ab/libtqgpu_ko.so python tools/c2_kernels.py >> $out 2>&1
done
