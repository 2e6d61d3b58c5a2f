mkdir -p gpurun_out/r3k
python -m pytest tests/test_gpu_gf_mirror.py -m gpu -q > gpurun_out/r3k/t.log 2>&1
