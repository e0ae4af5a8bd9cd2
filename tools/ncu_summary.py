"""Summarise an .ncu-rep (read here with `ncu -i`) into a small text table for profiles/."""
import csv, subprocess, sys, io

KEYS = [("gpu__time_duration.sum", "dur"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%"),
        ("sm__inst_executed_pipe_uniform.sum", None),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
        ("lts__t_bytes.sum", "l2_bytes"), ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%"),
        ("smsp__inst_executed.sum", "warp_inst"), ("smsp__thread_inst_executed_per_inst_executed.ratio", "thr/inst"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"), ("launch__registers_per_thread", "regs"),
        ("launch__grid_size", "grid"), ("gpc__cycles_elapsed.avg.per_second", "clk")]

def main(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    cols = [(k, n) for k, n in KEYS if n and k in idx]
    print("# " + rep)
    print("id  kernel                               " + "  ".join(f"{n:>10s}" for _, n in cols))
    print("                                         " + "  ".join(f"{units[idx[k]][:10]:>10s}" for k, _ in cols))
    for r in rows[2:]:
        if len(r) < len(hdr):
            continue
        name = r[idx["Kernel Name"]].split("(")[0].replace("void ", "")[-36:]
        vals = []
        for k, _ in cols:
            v = r[idx[k]]
            try:
                f = float(v.replace(",", ""))
                vals.append(f"{f:10.3f}" if abs(f) < 1e5 else f"{f:10.3e}")
            except ValueError:
                vals.append(f"{v[:10]:>10s}")
        print(f"{r[idx['ID']]:>3s} {name:36s} " + "  ".join(vals))

if __name__ == "__main__":
    main(sys.argv[1])
