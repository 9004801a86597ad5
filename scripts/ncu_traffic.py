"""DRAM traffic of the descriptor stage from an ncu capture of one 1M-atom pass.

    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
        -k regex:symfun -o gpurun_out/traffic_1m python scripts/profile_1m.py 50
    python scripts/ncu_traffic.py gpurun_out/traffic_1m.ncu-rep 50 [out.json]

Sums dram__bytes_read.sum + dram__bytes_write.sum over the symfun_prep / symfun_sweep* launches of the
capture (= one step of bench.py's workload) and writes profiles/traffic_1M.json, which bench.py reports
as roofline.traffic (a number cannot be taken under a profiler inside the bench run itself)."""
import csv
import io
import json
import os
import subprocess
import sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main():
    rep, nx = sys.argv[1], int(sys.argv[2])
    out = sys.argv[3] if len(sys.argv) > 3 else os.path.join(os.path.dirname(__file__), "..", "profiles", "traffic_1M.json")
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    h, units = rows[0], rows[1]
    ik, ir, iw, it = (h.index(c) for c in ("Kernel Name", "dram__bytes_read.sum", "dram__bytes_write.sum",
                                            "gpu__time_duration.sum"))
    tot, per = 0.0, []
    for r in rows[2:]:
        if "symfun_prep" not in r[ik] and "symfun_sweep" not in r[ik]:
            continue
        rd = float(r[ir].replace(",", "")) * UNIT[units[ir]]
        wr = float(r[iw].replace(",", "")) * UNIT[units[iw]]
        tot += rd + wr
        per.append({"kernel": r[ik].split("(")[0].replace("void <unnamed>::", ""), "read": rd, "write": wr,
                    "time": r[it] + " " + units[it]})
    n = 8 * nx**3
    d = {"nx": nx, "atoms": n, "dram_bytes_per_step": tot, "dram_bytes_per_atom": tot / n, "launches": per,
         "source": "ncu capture " + os.path.basename(rep) + " of scripts/profile_1m.py (dram__bytes_read.sum + "
                   "dram__bytes_write.sum over the symfun_prep / symfun_sweep launches of one step)"}
    json.dump(d, open(out, "w"), indent=1)
    print(json.dumps({k: d[k] for k in ("atoms", "dram_bytes_per_step", "dram_bytes_per_atom")}))


if __name__ == "__main__":
    main()
