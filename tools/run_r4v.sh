mkdir -p gpurun_out/r4v
python -m pytest tests/test_gpu_cpp_api.py -m gpu -q > gpurun_out/r4v/t.log 2>&1
