mkdir -p gpurun_out/r4c
python -m pytest tests/test_gpu_multi.py -m gpu -q > gpurun_out/r4c/multi.log 2>&1
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r4c/bench_n2.json 2> gpurun_out/r4c/bench_n2.err
