"""Compact per-launch table from an `ncu --set full` report: python scripts/ncu_summary.py rep.ncu-rep [> profiles/x.txt]
(reads the report with `ncu -i ... --page raw --csv`; no GPU needed)."""
import csv
import io
import json
import subprocess
import sys

WANT = [("gpu__time_duration.sum", "dur_us"), ("dram__bytes_read.sum", "dram_rd_MB"), ("dram__bytes_write.sum", "dram_wr_MB"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pipe_%"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_%"), ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_%"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_%"), ("sm__cycles_active.avg", "sm_active_cyc"),
        ("sm__cycles_elapsed.max", "elapsed_cyc"), ("launch__grid_size", "grid"), ("launch__registers_per_thread", "regs"),
        ("launch__shared_mem_per_block_dynamic", "dyn_smem"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy_%")]
UNIT = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    out = []
    print(f"# {rep}: {len(rows) - 2} launches (ncu --set full --clock-control none; cold-cache, serialised: compare shares, not absolutes)")
    print("kernel".ljust(44) + " ".join(n.rjust(13) for _, n in WANT))
    for r in rows[2:]:
        name = r[col["Kernel Name"]].split("(")[0].replace("void ", "").replace("srw::", "")[:43]
        vals = {}
        for m, n in WANT:
            if m not in col:
                vals[n] = None
                continue
            v = r[col[m]].replace(",", "")
            try:
                f = float(v) * UNIT.get(units[col[m]], 1.0)
            except ValueError:
                f = None
            vals[n] = f
        out.append(dict(kernel=name, **vals))
        print(name.ljust(44) + " ".join(("-" if vals[n] is None else f"{vals[n]:.2f}").rjust(13) for _, n in WANT))
    if out:
        tr = [o["dram_rd_MB"] + o["dram_wr_MB"] for o in out if o["dram_rd_MB"] is not None]
        print(f"# mean DRAM traffic per launch: {sum(tr) / len(tr):.2f} MB (read+write); mean duration {sum(o['dur_us'] for o in out) / len(out):.2f} us")
        print("# json: " + json.dumps(dict(mean_traffic_bytes=sum(tr) / len(tr) * 1e6, launches=len(out))))


if __name__ == "__main__":
    main()
