"""Pipe utilisation, stall reasons and a per-opcode / per-source-line split of ONE kernel of an .ncu-rep (developer tool)."""
import csv, io, re, subprocess, sys, collections
rep = sys.argv[1]; elems = float(sys.argv[2]) if len(sys.argv) > 2 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw))); hdr, vals = rows[0], rows[2]
d = dict(zip(hdr, vals))
for k in ["gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
          "launch__registers_per_thread", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
          "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
          "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
          "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum",
          "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum",
          "dram__bytes_read.sum", "dram__bytes_write.sum"]:
    print(f"{k:90s} {d.get(k)}")
for k in sorted(d):
    if "issue_stalled" in k and k.endswith("per_issue_active.ratio"):
        print(f"{k.split('issue_stalled_')[1].split('_per_issue')[0]:28s} {float(d[k]):.3f}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src))); hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
ops = collections.Counter(); st = collections.Counter(); tot = tots = 0
for r in rows[2:]:
    if len(r) < len(hdr): continue
    s = r[ix["Source"]].strip()
    m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_]+(\.[A-Z0-9_]+)*)", s)
    op = m.group(2) if m else s
    base = op.split(".")[0]
    key = op if base in ("MUFU", "F2F", "I2F", "F2I", "FRND", "IMAD", "LDS", "STS", "LDL", "STL") else base
    n = int(r[ix["Instructions Executed"]]); s_ = int(r[ix["Warp Stall Sampling (All Samples)"]])
    ops[key] += n; st[key] += s_; tot += n; tots += s_
div = (elems / 32) if elems else 1.0
print(f"total warp instructions {tot}  per 32 elements {tot / div:.1f}")
for k, v in ops.most_common(60):
    print(f"{k:28s} {v / div:8.2f}   stall {100 * st[k] / max(tots, 1):5.1f}%")
