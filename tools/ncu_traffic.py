"""Writes profiles/ncu_traffic.json: DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of the kernel the
benchmark's `roofline` names, taken from an `ncu --set full` report and keyed by a hash of the kernel sources, so that
bench.py reports `traffic` only for the build it was measured on (never a constant).
usage: python tools/ncu_traffic.py report.ncu-rep KEY KERNEL_REGEX [capture description]"""
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import kernel_source_hash  # noqa: E402


def main():
    rep, key, pat = sys.argv[1], sys.argv[2], re.compile(sys.argv[3])
    note = sys.argv[4] if len(sys.argv) > 4 else os.path.basename(rep)
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    vals = []
    for r in data:
        if not pat.search(r[ix["Kernel Name"]]):
            continue
        tot = 0.0
        for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            tot += float(r[ix[k]].replace(",", "")) * scale.get(units[ix[k]], 1)
        vals.append(tot)
    if not vals:
        raise SystemExit(f"no launch matches {pat.pattern}")
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    db = json.load(open(path)) if os.path.exists(path) else {}
    db[key] = {"dram_bytes": sum(vals) / len(vals), "launches": len(vals), "source_hash": kernel_source_hash(), "capture": note}
    json.dump(db, open(path, "w"), indent=1)
    print(json.dumps(db[key]))


if __name__ == "__main__":
    main()
