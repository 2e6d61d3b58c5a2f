"""ncu `--page raw --csv` export -> compact per-kernel JSON for profiles/ (one entry per distinct kernel: the LAST profiled launch).
usage: ncu_profile_json.py raw.csv out.json [note]"""
import csv
import json
import sys

KEEP = {
    "gpu__time_duration.sum": "duration", "dram__bytes_read.sum": "dram_read", "dram__bytes_write.sum": "dram_write",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct_of_peak", "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct_of_peak",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct", "launch__registers_per_thread": "registers",
    "launch__grid_size": "grid", "launch__block_size": "block", "smsp__inst_executed.sum": "warp_instructions",
    "sm__inst_executed_pipe_fp64.sum": "fp64_warp_instructions", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active": "fp64_pipe_active_pct",
    "sm__inst_executed_pipe_lsu.sum": "lsu_warp_instructions", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum": "smem_bank_conflicts",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum": "smem_wavefronts", "l1tex__data_pipe_lsu_wavefronts.sum": "lsu_wavefronts",
    "l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_active": "lsu_writeback_active_pct",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct", "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active": "fp64_pipe_inst_pct", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active": "lsu_pipe_inst_pct",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active": "alu_pipe_inst_pct", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active": "fma_pipe_inst_pct",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed": "lsu_data_pipe_pct",
    "sass__thread_inst_executed_per_opcode_category": "thread_inst_per_opcode_category",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.avg.pct_of_peak_sustained_elapsed": "smem_wavefront_pct_of_peak",
}
rows = list(csv.reader(open(sys.argv[1])))
h, units = rows[0], rows[1]
out = {}
for r in rows[2:]:
    name = r[h.index("Kernel Name")]
    d = {}
    for k, short in KEEP.items():
        if k in h:
            i = h.index(k)
            try:
                d[short] = float(r[i])
            except ValueError:
                d[short] = r[i]
            if short in ("duration", "dram_read", "dram_write"):
                d[short + "_unit"] = units[i]
    stalls = {}
    for i, n in enumerate(h):
        if n.startswith("smsp__average_warps_issue_stalled_") and n.endswith("_per_issue_active.ratio"):
            try:
                v = float(r[i])
            except ValueError:
                continue
            if v >= 0.1:
                stalls[n[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]] = round(v, 2)
    d["stalled_warps_per_issue_active"] = stalls
    if d.get("smem_wavefronts"):
        d["smem_bank_conflict_frac"] = d["smem_bank_conflicts"] / d["smem_wavefronts"]
    if d.get("warp_instructions") and d.get("fp64_warp_instructions") is not None:
        d["fp64_instruction_frac"] = d["fp64_warp_instructions"] / d["warp_instructions"]
    out[name] = d
json.dump({"note": sys.argv[3] if len(sys.argv) > 3 else "", "kernels": out}, open(sys.argv[2], "w"), indent=1)
print(len(out), "kernels")
