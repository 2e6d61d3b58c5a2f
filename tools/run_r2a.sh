mkdir -p gpurun_out/r2a
nvidia-smi -L > gpurun_out/r2a/gpus.txt
(time python -m pytest tests -m gpu -q) > gpurun_out/r2a/pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2a/pytest.log
python tools/lsq_tail_error.py > gpurun_out/r2a/lsq.json 2> gpurun_out/r2a/lsq.err
./tools/fp64_peak > gpurun_out/r2a/fp64_peak.json 2>&1
(time python bench.py --steps 10 --warmup 3) > gpurun_out/r2a/bench.json 2> gpurun_out/r2a/bench.err
python tools/ncu_configs.py > gpurun_out/r2a/ncu_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'dyson|lattice_pipe|matsubara_col|interp|fit_tail|density|matsubara_pipe|invert' -c 40 -o gpurun_out/r2a/prof_configs python tools/ncu_configs.py > gpurun_out/r2a/ncu.log 2>&1
ncu -i gpurun_out/r2a/prof_configs.ncu-rep --page raw --csv > gpurun_out/r2a/prof_configs_raw.csv 2>/dev/null
ls -la gpurun_out/r2a > gpurun_out/r2a/ls.txt
sz=$(stat -c %s gpurun_out/r2a/prof_configs.ncu-rep 2>/dev/null || echo 0)
if [ "$sz" -gt 45000000 ]; then rm -f gpurun_out/r2a/prof_configs.ncu-rep; fi
echo done
