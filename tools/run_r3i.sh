mkdir -p gpurun_out/r3i
out=gpurun_out/r3i/c1.log
: > $out
for k in 0 8 1 2 4 3 7 15; do
  echo "col skip $k" >> $out
  TQ_COL_SKIP=$k TQ_LIB_PATH=tools/ab/libtqgpu_ko.so python tools/c1_probe.py 2>&1 | tail -1 >> $out
done
