"""Developer tool: LSU picture of a kernel from an `ncu --page source --csv --print-source sass` export: per memory opcode
(LDS / STS / LDG / STG / ATOMS ...) executed warp instructions, shared wavefronts (actual / ideal), global tag requests and
sectors, and stall samples.  usage: ncu_lsu_breakdown.py file.csv [kernel-substring]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
sel = sys.argv[2] if len(sys.argv) > 2 else ""
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
        def g(r, n):
            try: return float(r[H[n]])
            except Exception: return 0.0
        agg = collections.defaultdict(lambda: [0.0] * 8)
        tot_inst = tot_samp = 0
        for r in body:
            src = r[H["Source"]].strip()
            toks = src.split()
            if not toks: continue
            op = toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]
            inst = g(r, "Instructions Executed"); samp = g(r, "# Samples")
            tot_inst += inst; tot_samp += samp
            base = op.split(".")[0]
            if base in ("LDS", "STS", "LDG", "STG", "ATOMS", "LDSM", "LD", "ST", "LDL", "STL", "RED", "ATOMG", "SHFL", "SYNCS", "UTMALDG", "UBLKCP", "LDC", "LDCU"):
                key = op if base in ("LDS", "STS", "LDG", "STG") else base
                a = agg[key]
                a[0] += inst; a[1] += g(r, "L1 Wavefronts Shared"); a[2] += g(r, "L1 Wavefronts Shared Ideal")
                a[3] += g(r, "L1 Tag Requests Global"); a[4] += g(r, "L2 Theoretical Sectors Global"); a[5] += g(r, "L2 Theoretical Sectors Global Ideal")
                a[6] += samp; a[7] += 1
        print(f"=== {name[:100]}\n  warp-instr {tot_inst:.0f}  samples {tot_samp:.0f}")
        print(f"  {'op':22s} {'sites':>5s} {'warp-instr':>12s} {'smem wavef':>12s} {'ideal':>12s} {'glob tagreq':>12s} {'sectors':>12s} {'ideal':>12s} {'samples%':>8s}")
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0]):
            print(f"  {k:22s} {a[7]:5.0f} {a[0]:12.0f} {a[1]:12.0f} {a[2]:12.0f} {a[3]:12.0f} {a[4]:12.0f} {a[5]:12.0f} {100*a[6]/max(tot_samp,1):8.1f}")
    else:
        i += 1
