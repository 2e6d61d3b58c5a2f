"""Developer tool: phases of a long unrolled kernel from an `ncu --page source --csv --print-source sass` export.
Instructions in address order, bucketed; per bucket: share of stall samples, share of executed warp instructions, opcode mix, the
dominant stall reasons.  usage: ncu_sass_phases.py file.csv kernel-substring [bucket]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
sel = sys.argv[2]
bucket = int(sys.argv[3]) if len(sys.argv) > 3 else 64
i = 0
while i < len(rows):
    r = rows[i]
    if r and r[0] == "Kernel Name":
        name = r[1]; hdr = rows[i + 1]; i += 2
        body = []
        while i < len(rows) and not (rows[i] and rows[i][0] == "Kernel Name"):
            if rows[i]: body.append(rows[i])
            i += 1
        if sel not in name: continue
        H = {n: j for j, n in enumerate(hdr)}
        stall_cols = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
        def g(r, n):
            try: return float(r[H[n]])
            except Exception: return 0.0
        tot_s = sum(g(r, "# Samples") for r in body); tot_i = sum(g(r, "Instructions Executed") for r in body)
        print(f"=== {name[:110]}\n  {len(body)} SASS instr, samples {tot_s:.0f}, warp-instr {tot_i:.0f}")
        allst = collections.Counter()
        for b0 in range(0, len(body), bucket):
            ch = body[b0:b0 + bucket]
            s = sum(g(r, "# Samples") for r in ch); n = sum(g(r, "Instructions Executed") for r in ch)
            ops = collections.Counter()
            st = collections.Counter()
            for r in ch:
                t = r[H["Source"]].split()
                if not t: continue
                op = t[1] if t[0].startswith("@") and len(t) > 1 else t[0]
                ops[op.split(".")[0]] += 1
                for c in stall_cols: st[c[6:]] += g(r, c)
            allst.update(st)
            hot = max(ch, key=lambda r: g(r, "# Samples"))
            print(f"[{b0:5d}] smp {100*s/max(tot_s,1):5.1f}% exec {100*n/max(tot_i,1):5.1f}%  {' '.join(f'{k}:{v}' for k,v in ops.most_common(4)):36s} | {' '.join(f'{k}={100*v/max(tot_s,1):.1f}' for k,v in st.most_common(4)):52s} | hot {hot[H['Source']].strip()[:44]} ({g(hot,'# Samples'):.0f})")
        print("  all stalls: " + "  ".join(f"{k}={100*v/max(tot_s,1):.1f}%" for k, v in allst.most_common(12)))
    else:
        i += 1
