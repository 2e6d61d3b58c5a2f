mkdir -p gpurun_out/r3p
out=gpurun_out/r3p/rader.log
: > $out
C="6150,1025,25 61500,10250,25 6600,1100,25 7800,1300,25 10200,1700,25"
for v in default norader41 norader; do
  echo "variant $v" >> $out
  if [ $v = default ]; then unset TQ_LIB_PATH; else export TQ_LIB_PATH=tools/ab/libtqgpu_$v.so; fi
  python tools/len_probe.py $C 2>&1 | tail -5 >> $out
done
