"""ncu_src_stalls.py <report.ncu-rep> — headline pipe metrics and warp-stall samples per SASS opcode (source page)."""
import csv, subprocess, sys, io
from collections import Counter
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, vals = rows[0], rows[2]
d = dict(zip(hdr, vals))
print(d["Kernel Name"][:90])
for k in ("gpu__time_duration.sum", "sm__cycles_elapsed.max", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
          "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active" , "sm__inst_executed_pipe_tensor.sum",
          "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
          "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
          "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread"):
    if k in d: print(f"  {k} = {d[k]}")
for k in hdr:
    if "tensor" in k and "pct" in k and d[k] not in ("", "0"): print(f"  {k} = {d[k]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = Counter(); byop = {}
for r in rows[2:]:
    s = r[ix["Source"]].strip()
    op = s.split()[0] if not s.startswith("@") else s.split()[1]
    c = byop.setdefault(op, Counter())
    c["samples"] += int(r[ix["# Samples"]] or 0); c["exec"] += int(r[ix["Instructions Executed"]] or 0)
    for st in stalls:
        v = int(r[ix[st]] or 0); c[st] += v; tot[st] += v
print("  stall totals:", [(k, v) for k, v in tot.most_common() if v])
for op, c in sorted(byop.items(), key=lambda x: -x[1]["samples"])[:12]:
    print(f"  {op:28s} samples {c['samples']:6d} exec {c['exec']:9d}", [(k[6:], v) for k, v in c.most_common(7) if k.startswith("stall_") and v])
