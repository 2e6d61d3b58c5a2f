"""Developer tool: SASS-level profile from an `ncu --page source --csv --print-source cuda,sass` export.
Collects every SASS row of one kernel (deduplicated by address), orders by address and prints samples / executed counts
in address buckets so phases of a long unrolled kernel can be told apart.  usage: ncu_sass_profile.py file.csv kernel-substring [bucket]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
sel = sys.argv[2]
bucket = int(sys.argv[3]) if len(sys.argv) > 3 else 64
ins = {}
func = None
hdr = None
for i, r in enumerate(rows):
    if r and r[0] == "Function Name":
        func = r[1]
        hdr = rows[i + 1]
        continue
    if func is None or sel not in func or not r or r[0] != "" or len(r) < 8:
        continue
    if not r[2].startswith("0x"):
        continue
    addr = int(r[2], 16)
    try:
        ins[addr] = (r[3].strip(), int(r[6]), int(r[7]), r[4])
    except ValueError:
        pass
addrs = sorted(ins)
base = addrs[0]
tot_s = sum(v[1] for v in ins.values())
tot_i = sum(v[2] for v in ins.values())
print(f"{len(addrs)} SASS instructions, {tot_s} samples, {tot_i} warp-instr executed")
for b0 in range(0, len(addrs), bucket):
    chunk = addrs[b0:b0 + bucket]
    s = sum(ins[a][1] for a in chunk)
    n = sum(ins[a][2] for a in chunk)
    ops = {}
    for a in chunk:
        op = ins[a][0].split()[0]
        if op.startswith("@"):
            op = ins[a][0].split()[1]
        op = op.split(".")[0]
        ops[op] = ops.get(op, 0) + 1
    top = sorted(ops.items(), key=lambda kv: -kv[1])[:5]
    hot = max(chunk, key=lambda a: ins[a][1])
    print(f"[{(chunk[0]-base)//16:5d}] samples {100.0*s/max(tot_s,1):5.1f}%  exec {100.0*n/max(tot_i,1):5.1f}%  {' '.join(f'{k}:{v}' for k,v in top):50s} hot: {ins[hot][0][:60]} ({ins[hot][1]})")
