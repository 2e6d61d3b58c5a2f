mkdir -p gpurun_out/r3a
python -m pytest tests -m gpu -x -q > gpurun_out/r3a/gputests.log 2>&1
python tools/c2_kernels.py > gpurun_out/r3a/c2.log 2>&1
TQ_LIB_PATH=tools/ab/libtqgpu_dev.so python tools/phase_probe.py 16 > gpurun_out/r3a/phase.log 2>&1
