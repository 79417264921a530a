"""Compact markdown summary of one `ncu --set full` report. Usage: ncu_summary.py report.ncu-rep [title]"""
import csv, subprocess, sys
rep = sys.argv[1]
title = sys.argv[2] if len(sys.argv) > 2 else rep
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
d = dict(zip(hdr, vals)); u = dict(zip(hdr, units))
def g(k, fmt="{}"):
    v = d.get(k, "")
    try:
        return fmt.format(float(v))
    except Exception:
        return v or "n/a"
def scaled(k):   # value with its unit
    return f"{g(k, '{:.4g}')} {u.get(k, '')}".strip()
print(f"### {title}\n")
print(f"`{d.get('Kernel Name', '')[:110]}`  grid {d.get('launch__grid_size')} x block {d.get('launch__block_size')}, "
      f"{g('launch__registers_per_thread', '{:.0f}')} regs/thread, {scaled('launch__shared_mem_per_block_dynamic')} dyn smem, "
      f"{scaled('gpu__time_duration.sum')} (ncu: serialised, clocks not locked)\n")
print("| metric | value |\n|---|---|")
for name, k in (("DRAM read", "dram__bytes_read.sum"), ("DRAM write", "dram__bytes_write.sum"),
                ("DRAM throughput % of peak", "dram__throughput.avg.pct_of_peak_sustained_elapsed"),
                ("L2 -> L1/TMA read bytes", "l1tex__m_xbar2l1tex_read_bytes.sum"),
                ("SM throughput % of peak", "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
                ("tensor pipe (hmma subpipe) cycles active %", "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active"),
                ("tensor-core busy incl. operand fetch (pipe_tc) %", "sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_active"),
                ("XU (SFU) pipe % of peak", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
                ("FMA pipe %", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
                ("ALU pipe %", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active"),
                ("issue slots busy %", "sm__issue_active.avg.pct_of_peak_sustained_elapsed"),
                ("achieved occupancy (warps/SM)", "sm__warps_active.avg.per_cycle_active"),
                ("shared-memory bank conflicts (st)", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum"),
                ("local-memory (spill) load requests", "l1tex__t_requests_pipe_lsu_mem_local_op_ld.sum")):
    if k in d and d[k] != "":
        print(f"| {name} | {scaled(k)} |")
st = sorted(((float(v), k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""))
             for k, v in d.items() if "issue_stalled" in k and k.endswith("per_issue_active.ratio") and v), reverse=True)[:5]
print("\nTop warp stall reasons (warps stalled per issue-active cycle): " + ", ".join(f"{n} {v:.2f}" for v, n in st) + "\n")
