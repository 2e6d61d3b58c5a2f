mkdir -p gpurun_out/r4w
: > gpurun_out/r4w/sweep.log
for tg in "4 2" "6 2" "8 2" "4 4" "3 4" "2 8" "5 2" "4 2"; do
  set -- $tg
  python bench.py --no-extras --e2e-threads $1 --e2e-group $2 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('threads $1 group $2', round(d['e2e']['value']/1e6, 1), 'M pts/s', 'value', round(d['value']/1e9, 3))" >> gpurun_out/r4w/sweep.log
done
