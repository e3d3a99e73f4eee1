"""Build profiles/r1_ncu_traffic.json (DRAM bytes per launch of K1 and K2, feeds bench.py's roofline.traffic) from an
`ncu --set full` capture of bench.py:  python tools/ncu_traffic.py capture.ncu-rep profiles/r1_ncu_traffic.json"""
import csv
import json
import subprocess
import sys


def main():
    rep, out_path = sys.argv[1], sys.argv[2]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]

    def val(r, key):
        i = hdr.index(key)
        x = float(r[i].replace(",", ""))
        u = units[i].lower()
        scale = {"gbyte": 1e9, "mbyte": 1e6, "kbyte": 1e3, "byte": 1.0, "ms": 1e3, "us": 1.0, "ns": 1e-3}.get(u, 1.0)
        return x * scale

    kernels = {}
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        kind = "vote" if "vote" in name else "numeric" if "numeric" in name else None
        if kind is None or kind in kernels:
            continue
        rd, wr = val(r, "dram__bytes_read.sum"), val(r, "dram__bytes_write.sum")
        kernels[kind] = {"kernel": name[:80], "dram_bytes_read": rd, "dram_bytes_write": wr,
                         "duration_us_under_ncu": val(r, "gpu__time_duration.sum"),
                         "inst_executed": val(r, "inst_executed"), "registers": val(r, "launch__registers_per_thread"),
                         "ipc": val(r, "sm__inst_executed.avg.per_cycle_active"), "traffic_bytes": rd + wr}
    doc = {"source": "ncu --set full --clock-control none, bench.py --steps 2 --warmup 3 --no-cpu --no-e2e "
                     "(1M records x 32 fields, n=16), one launch each", "kernels": kernels}
    with open(out_path, "w") as f:
        json.dump(doc, f, indent=1)
    print(json.dumps(doc, indent=1))


if __name__ == "__main__":
    main()
