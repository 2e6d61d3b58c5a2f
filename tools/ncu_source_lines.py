"""Aggregates an `ncu --page source --csv --print-source cuda,sass` export per CUDA source line:
samples (stall sampling) and executed warp instructions.  usage: ncu_source_lines.py file.csv [kernel-substring] [top]"""
import csv
import sys

fn = sys.argv[1]
sel = sys.argv[2] if len(sys.argv) > 2 else ""
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
rows = list(csv.reader(open(fn)))
i = 0
sections = []
cur = None
while i < len(rows):
    r = rows[i]
    if r and r[0] == "File Path":
        cur = {"file": r[1], "func": rows[i + 1][1], "hdr": rows[i + 2], "lines": []}
        sections.append(cur)
        i += 3
        continue
    if cur is not None and r:
        cur["lines"].append(r)
    i += 1
seen = set()
for s in sections:
    if sel not in s["func"]:
        continue
    key = (s["file"], s["func"])
    if key in seen:
        continue
    seen.add(key)
    h = s["hdr"]
    si, ii = h.index("# Samples"), h.index("Instructions Executed")
    stall_cols = [j for j, n in enumerate(h) if n.startswith("stall_") and "Not Issued" not in n]
    agg = []
    tot_s = tot_i = 0
    for r in s["lines"]:
        if r[0] == "":
            continue  # SASS row (already summed in the CUDA line row)
        try:
            smp, ins = int(r[si]), int(r[ii])
        except ValueError:
            continue
        tot_s += smp
        tot_i += ins
        st = sorted(((int(r[j]), h[j]) for j in stall_cols if r[j].isdigit() and int(r[j]) > 0), reverse=True)[:3]
        agg.append((smp, ins, r[0], r[1].strip()[:100], st))
    print(f"=== {s['func'][:90]}  [{s['file'].split('/')[-1]}]  samples={tot_s} warp-instr={tot_i}")
    for smp, ins, ln, src, st in sorted(agg, reverse=True)[:top]:
        print(f"{100.0*smp/max(tot_s,1):5.1f}%s {100.0*ins/max(tot_i,1):5.1f}%i  L{ln:>4s} {src:100s} {' '.join(f'{n[6:]}={v}' for v,n in st)}")
