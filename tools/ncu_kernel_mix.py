"""Per-opcode executed-instruction mix of the kernels of an .ncu-rep whose name matches a substring (developer tool).
usage: ncu_kernel_mix.py rep.ncu-rep "substring" units [listing.txt]"""
import csv, io, re, subprocess, sys, collections
rep, sub, units = sys.argv[1], sys.argv[2], float(sys.argv[3])
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
secs, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "rows": []}; secs.append(cur)
    elif cur is not None and cur["hdr"] is None:
        cur["hdr"] = r
    elif cur is not None:
        cur["rows"].append(r)
for s in secs:
    if sub not in s["name"]:
        continue
    ix = {h: i for i, h in enumerate(s["hdr"])}
    ops = collections.Counter(); st = collections.Counter(); tot = tots = 0; L = []
    for r in s["rows"]:
        if len(r) < len(s["hdr"]): continue
        t = r[ix["Source"]].strip()
        m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_]+(\.[A-Z0-9_]+)*)", t)
        op = (m.group(2) if m else t).split(".")[0]
        n = int(r[ix["Instructions Executed"]]); w = int(r[ix["Warp Stall Sampling (All Samples)"]])
        ops[op] += n; st[op] += w; tot += n; tots += w; L.append((n, w, t))
    print(s["name"][:120]); print(f"total warp instructions {tot}, per unit {tot / units:.1f}")
    for k, v in ops.most_common(36): print(f"{k:14s} {v / units:8.1f}  stall {100 * st[k] / max(tots, 1):5.1f}%")
    if len(sys.argv) > 4:
        open(sys.argv[4], "w").write("\n".join(f"{i:5d} {a / units:8.2f} {b:6d}  {c}" for i, (a, b, c) in enumerate(L)))
    break
