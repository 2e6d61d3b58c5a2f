mkdir -p gpurun_out/r2b
(time timeout 900 python -m pytest tests/test_gpu_planned.py -x -q) > gpurun_out/r2b/planned.log 2>&1
echo "rc=$?" >> gpurun_out/r2b/planned.log
(time timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_gpu_planned.py) > gpurun_out/r2b/pytest.log 2>&1
echo "rc=$?" >> gpurun_out/r2b/pytest.log
echo done
