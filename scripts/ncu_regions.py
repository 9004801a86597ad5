"""Aggregates an `ncu --page source --csv` export by SASS address and prints the hottest instructions.
Usage: python scripts/ncu_regions.py src.csv [n_atoms] [lo:hi:name ...]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ia, isrc, isamp, iex = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
data = []
for r in rows[2:]:
    try:
        data.append((int(r[ia], 16), r[isrc], int(r[isamp]), int(r[iex])))
    except Exception:
        pass
base = data[0][0]
na = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
ts, te = sum(d[2] for d in data), sum(d[3] for d in data)
print("samples", ts, "instr/atom", te / na)
for spec in sys.argv[3:]:
    lo, hi, name = spec.split(":")
    lo, hi = int(lo, 16), int(hi, 16)
    s = sum(d[2] for d in data if lo <= d[0] - base < hi)
    e = sum(d[3] for d in data if lo <= d[0] - base < hi)
    print(f"{name:20s} samples {100*s/ts:5.1f}%  instr {100*e/te:5.1f}%  instr/atom {e/na:8.1f}")
for d in sorted(data, key=lambda d: -d[2])[:30]:
    print(hex(d[0] - base), d[2], round(100 * d[2] / ts, 2), d[3] / na, d[1][:80])
