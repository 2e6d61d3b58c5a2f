mkdir -p gpurun_out/r3h
out=gpurun_out/r3h/knock_alias.log
: > $out
for k in 0 1 144 401 4 40 46; do
  echo "alias knock $k" >> $out
  TQ_SCRATCH_ALIAS=1 TQ_KNOCK=$k TQ_LIB_PATH=tools/ab/libtqgpu_ko.so python tools/c2_kernels.py >> $out 2>&1
done
echo "noalias knock 144" >> $out
TQ_KNOCK=144 TQ_LIB_PATH=tools/ab/libtqgpu_ko.so python tools/c2_kernels.py >> $out 2>&1
