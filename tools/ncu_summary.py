"""Extracts the judged metrics from a .ncu-rep into a small CSV + top-stall listing (profiles/)."""
import csv, io, subprocess, sys

rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, body = rows[0], rows[1], rows[2:]
want = ["Kernel Name", "gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum",
        "sm__cycles_elapsed.avg", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed"]
idx = [hdr.index(w) for w in want if w in hdr]
with open(out + ".csv", "w", newline="") as f:
    w = csv.writer(f)
    w.writerow([hdr[i] for i in idx]); w.writerow([units[i] for i in idx])
    for r in body:
        w.writerow([r[i] for i in idx])
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
srows = list(csv.reader(io.StringIO(src)))
hi = [i for i, r in enumerate(srows) if r and r[0] == "Address"]
if hi:
    h = srows[hi[0]]
    end = hi[1] - 1 if len(hi) > 1 else len(srows)
    b = [r for r in srows[hi[0] + 1:end] if len(r) == len(h)]
    ci = {n: i for i, n in enumerate(h)}
    fl = lambda r, n: float(r[ci[n]] or 0) if r[ci[n]].replace(".", "", 1).isdigit() else 0.0
    stalls = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
    with open(out + "_stalls.txt", "w") as f:
        tot = sum(fl(r, "# Samples") for r in b)
        f.write(f"kernel: {srows[0][1] if srows and len(srows[0]) > 1 else ''}\ntotal samples {tot:.0f}\n")
        agg = sorted(((sum(fl(r, s) for r in b), s) for s in stalls), reverse=True)[:8]
        for v, s in agg:
            f.write(f"{s:28s} {v:10.0f} {100 * v / max(tot, 1):5.1f}%\n")
        f.write("\ntop instructions by samples: addr samples executed sass [top stall]\n")
        for r in sorted(b, key=lambda r: -fl(r, "# Samples"))[:25]:
            st = max(stalls, key=lambda s: fl(r, s))
            f.write(f"{r[ci['Address']][-5:]} {fl(r, '# Samples'):8.0f} {fl(r, 'Instructions Executed'):12.0f} {r[ci['Source']][:70]:70s} [{st}]\n")
print("wrote", out)
