mkdir -p gpurun_out/r4n
python -m pytest tests/test_gpu_parity.py -m gpu -q -k "large_pageable or pageable_buffers" > gpurun_out/r4n/t.log 2>&1
