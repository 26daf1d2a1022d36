"""Markdown summary of an `ncu --set full` report: one table per captured launch with the metrics the roofline argument uses.
usage: python tools/ncu_summary.py report.ncu-rep "title / command line" > profiles/rNN_xxx_ncu_summary.md
(reads the report with `ncu -i report --page raw --csv`)"""
import csv
import io
import subprocess
import sys

METRICS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor pipe active % (elapsed)"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("l1tex__m_xbar2l1tex_read_bytes.sum", "L2->SM bytes"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem / block"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("sm__cycles_elapsed.max", "SM cycles"),
]


def main():
    rep, title = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    print(f"# {title}\n")
    print(f"Source: `{rep.split('/')[-1]}` (read with `ncu -i … --page raw --csv`). Durations under ncu are cold-cache, serialised and at "
          "profiler clocks: use them for shares and counters, not as benchmark numbers.\n")
    for n, r in enumerate(data):
        name = r[ix["Kernel Name"]]
        print(f"## launch {n}: `{name[:150]}`\n")
        print("| metric | value |\n|---|---|")
        rd = wr = None
        for key, label in METRICS:
            if key not in ix:
                continue
            v, u = r[ix[key]], units[ix[key]]
            print(f"| {label} (`{key}`) | {v} {u} |")
            if key == "dram__bytes_read.sum":
                rd = (float(v.replace(",", "")), u)
            if key == "dram__bytes_write.sum":
                wr = (float(v.replace(",", "")), u)
        if rd and wr:
            scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            tot = rd[0] * scale.get(rd[1], 1) + wr[0] * scale.get(wr[1], 1)
            print(f"| **traffic (read + write)** | {tot / 1e6:.1f} MB |")
        print()


if __name__ == "__main__":
    main()
