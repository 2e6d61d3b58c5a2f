mkdir -p gpurun_out/final
( time python -m pytest tests -m gpu -q ) > gpurun_out/final/gputests.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final/smoke.log 2>&1
( time python bench.py ) > gpurun_out/final/bench_n1.json 2> gpurun_out/final/bench_n1.err
( time python bench.py --impl reference ) > gpurun_out/final/bench_ref.json 2> gpurun_out/final/bench_ref.err
python bench.py --steps 2 --warmup 1 --no-extras > gpurun_out/final/bench_plain.json 2>gpurun_out/final/bench_plain.err && ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 200 --csv --log-file gpurun_out/final/launches.csv python bench.py --steps 2 --warmup 1 --no-extras > gpurun_out/final/bench_ncu.log 2>&1
