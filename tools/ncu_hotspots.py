"""Bucket a kernel's executed instructions / stall samples along the SASS (ncu --page source --csv) to find its hot regions."""
import csv
import subprocess
import sys


def main():
    rep = sys.argv[1]
    bucket = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    start = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[start]
    ci, si, sm = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("# Samples")
    agg = []
    for r in rows[start + 1:]:
        try:
            agg.append((int(r[ci]), int(r[sm]), r[si].strip()))
        except (ValueError, IndexError):
            continue
    tot, tots = sum(a for a, _, _ in agg), sum(s for _, s, _ in agg)
    print(f"{len(agg)} SASS instructions, {tot} warp-instructions executed, {tots} samples")
    for i in range(0, len(agg), bucket):
        chunk = agg[i:i + bucket]
        e, s = sum(a for a, _, _ in chunk), sum(b for _, b, _ in chunk)
        ops = " ".join(c[2].split()[0] for c in chunk[:6])
        print(f"{i:5d} exec {100 * e / tot:5.1f}%  samples {100 * s / max(tots, 1):5.1f}%  {ops}")


if __name__ == "__main__":
    main()
