mkdir -p gpurun_out/r3u
out=gpurun_out/r3u/blue_vs_planned.log
C="9999,1000,25 8633,1400,25 6150,1025,25 7979,1300,25 10000,1025,25 5353,890,25 12345,2000,25"
echo planned > $out
python tools/len_probe.py $C 2>&1 | cut -c1-140 >> $out
echo bluestein >> $out
FORCE_BLUESTEIN=1 python tools/len_probe.py $C 2>&1 | cut -c1-140 >> $out
