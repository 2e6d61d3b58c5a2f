mkdir -p gpurun_out/r4a
export TQ_LIB_PATH=tools/ab/libtqgpu_dev.so
for c in 65536,8192,25 10000,1025,25 200000,20000,25 6000,1000,25; do
  python tools/plan_scan.py $c >> gpurun_out/r4a/scan.log 2>&1
done
