mkdir -p gpurun_out/r2c
python tools/ncu_planned.py > gpurun_out/r2c/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'matsubara_pipe_kernel' -s 4 -c 4 -o gpurun_out/r2c/planned python tools/ncu_planned.py > gpurun_out/r2c/ncu.log 2>&1
ncu -i gpurun_out/r2c/planned.ncu-rep --page raw --csv > gpurun_out/r2c/planned_raw.csv 2>/dev/null
ncu -i gpurun_out/r2c/planned.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/r2c/planned_source.csv 2>/dev/null
gzip -9 gpurun_out/r2c/planned_source.csv
ls -la gpurun_out/r2c > gpurun_out/r2c/ls.txt
sz=$(stat -c %s gpurun_out/r2c/planned.ncu-rep 2>/dev/null || echo 0)
if [ "$sz" -gt 40000000 ]; then rm -f gpurun_out/r2c/planned.ncu-rep; fi
(timeout 600 python -m pytest tests/test_gpu_parity.py -q -k "invert or dyson" ) > gpurun_out/r2c/inv.log 2>&1
echo done
