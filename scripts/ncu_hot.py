"""Top stall-sample SASS lines of an ncu report:  python scripts/ncu_hot.py rep.ncu-rep [N] [kernel-regex]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; N = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = None; data = []
for r in rows:
    if r and r[0] == "Address":
        hdr = r; continue
    if hdr and len(r) == len(hdr):
        data.append(r)
iS = hdr.index("# Samples"); iSrc = hdr.index("Source"); iEx = hdr.index("Instructions Executed")
tot = sum(int(r[iS]) for r in data)
print("total samples", tot, "instructions", len(data))
# stall reason columns
reason_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") ]
order = sorted(range(len(data)), key=lambda i: -int(data[i][iS]))[:N]
for i in sorted(order):
    r = data[i]
    rs = {hdr[c]: int(r[c]) for c in reason_cols if r[c] not in ("", "0")}
    top = sorted(rs.items(), key=lambda kv: -kv[1])[:3]
    print(f"{i:5d} {int(r[iS]):6d} {100*int(r[iS])/tot:5.1f}%  ex={r[iEx]:>8}  {r[iSrc].strip()[:90]:90s} {top}")
