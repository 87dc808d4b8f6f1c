"""Top SASS lines of an ncu report by stall samples / executed instructions (dev tool)."""
import csv, subprocess, sys, io
rep = sys.argv[1]; n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, vals = rows[0], rows[2]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__grid_size", "launch__registers_per_thread",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.avg.per_cycle_active", "smsp__issue_active.avg.pct",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__occupancy_limit", "launch__waves_per_multiprocessor", "smsp__inst_executed.sum"]
for i, h in enumerate(hdr):
    if any(h == w or (w.endswith("limit") and h.startswith(w)) for w in want):
        print(f"{h} = {vals[i]} {rows[1][i]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]; data = rows[2:]
iS, iI, iSrc = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Source")
stall = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[iS]) for r in data)
print("total samples", tot, "warp instr", sum(int(r[iI]) for r in data))
agg = {}
for r in data:
    for i in stall:
        if r[i] not in ("", "0"): agg[hdr[i]] = agg.get(hdr[i], 0) + int(r[i])
print(sorted(agg.items(), key=lambda x: -x[1])[:8])
for idx, r in sorted(enumerate(data), key=lambda x: -int(x[1][iS]))[:n]:
    st = sorted([(int(r[i]), hdr[i][6:]) for i in stall if r[i] not in ("", "0")], reverse=True)[:2]
    print(f"{idx:5d} {r[iS]:>6s} {r[iI]:>9s}  {r[iSrc][:64]:64s} {st}")
