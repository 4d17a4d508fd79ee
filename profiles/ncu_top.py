"""Summarise an ncu report's source page: top SASS instructions by stall samples (with the dominant stall reasons)
and per-64-instruction-window sample totals.  usage: python profiles/ncu_top.py report.ncu-rep [n_top [kernel_index]]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
n_top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0            # a report may hold several kernels: pick one
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"] or [0]
lo = starts[which]
hi = starts[which + 1] if which + 1 < len(starts) else len(rows)
rows = rows[lo:hi]
hdr, data = rows[1], [r for r in rows[2:] if len(r) >= len(rows[1])]
ia, isrc, isamp, iex = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[isamp] or 0) for r in data)
print(rows[0][1][:100])
print("total samples", tot, " instructions executed", sum(int(r[iex] or 0) for r in data), " SASS lines", len(data))
agg = {hdr[i]: sum(int(r[i] or 0) for r in data) for i in stall_cols}
print("stall totals:", sorted(agg.items(), key=lambda kv: -kv[1])[:8])
for r in sorted(data, key=lambda r: -int(r[isamp] or 0))[:n_top]:
    st = sorted([(hdr[i][6:], int(r[i] or 0)) for i in stall_cols], key=lambda kv: -kv[1])[:2]
    print(r[ia][-5:], r[isamp].rjust(7), r[iex].rjust(10), r[isrc][:64].ljust(64), st)
print("--- windows of 64 SASS instructions: first address, samples, instructions executed")
for i in range(0, len(data), 64):
    w = data[i:i + 64]
    s = sum(int(r[isamp] or 0) for r in w)
    if s > tot * 0.01:
        print(w[0][ia][-5:], str(s).rjust(7), str(sum(int(r[iex] or 0) for r in w)).rjust(11), "|", w[0][isrc][:50])
