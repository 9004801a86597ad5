"""Splits an `ncu --page source --csv` export into runs of SASS with the same execution count per
atom and prints each run's share of stall samples and instructions.  Usage: ncu_segments.py src.csv n_atoms"""
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
base = data[0][0]; ts = sum(d[2] for d in data); na = float(sys.argv[2])
print("samples", ts, "instr/atom", sum(d[3] for d in data) / na)
seg = []; cur = None
for a, src, s, e in data:
    key = round(e / na, 1)
    if cur is None or abs(key - cur[0]) > 0.3 * max(key, cur[0], 1):
        if cur: seg.append(cur)
        cur = [key, a - base, a - base, 0, 0, 0]
    cur[2] = a - base; cur[3] += s; cur[4] += e; cur[5] += 1
seg.append(cur)
for k, lo, hi, s, e, nn in seg:
    if s / ts > 0.004 or e / na > 50:
        print(f"{lo:#7x}-{hi:#7x} exec/atom {k:6.1f} ninstr {nn:5d} samples {100*s/ts:5.1f}% instr/atom {e/na:8.1f}")
