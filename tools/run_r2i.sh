python tools/c4_probe.py > gpurun_out/c4_ab.log 2>&1
TQ_LIB_PATH=tools/ab/libtqgpu_dy3.so python tools/c4_probe.py >> gpurun_out/c4_ab.log 2>&1
TQ_LIB_PATH=tools/ab/libtqgpu_dy4.so python tools/c4_probe.py >> gpurun_out/c4_ab.log 2>&1
