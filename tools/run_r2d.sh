bash tools/run_ab.sh gpurun_out/ab1.log B C D
for lag in 1 0; do echo "lag0=$lag" >> gpurun_out/lag.log; TQ_LAG=$lag python tools/bench_lengths.py matsubara >> gpurun_out/lag.log 2>&1; done
