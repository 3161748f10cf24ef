"""Every captured launch of an .ncu-rep (--set full) as one CSV row: duration, DRAM bytes, achieved DRAM GB/s and the
main pipe utilisations.  Used for the HBM-bound kernels (norms, DDIM update) where several launches are captured at once.

    python tools/ncu_table.py gpurun_out/prof_hbm.ncu-rep profiles/r2_hbm_kernels.csv
"""
import csv
import subprocess
import sys

COLS = [
    ("gpu__time_duration.sum", "dur_us", 1e-3),   # ns -> us
    ("dram__bytes_read.sum", "dram_rd_bytes", 1.0),
    ("dram__bytes_write.sum", "dram_wr_bytes", 1.0),
    ("launch__grid_size", "grid", 1.0), ("launch__block_size", "block", 1.0), ("launch__registers_per_thread", "regs", 1.0),
    ("lts__t_sector_hit_rate.pct", "l2_hit_pct", 1.0),
    ("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pct", 1.0),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "xu_pct", 1.0),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_pct", 1.0),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct", 1.0),
]
UNIT_SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}


def num(s):
    try:
        return float(s.replace(",", ""))
    except ValueError:
        return float("nan")


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["kernel"] + [c[1] for c in COLS] + ["dram_gbs"])
        for vals in rows[2:]:
            if len(vals) != len(hdr):
                continue
            name = vals[hdr.index("Kernel Name")].split("(")[0]
            rec = []
            for key, _, scale in COLS:
                if key in hdr:
                    i = hdr.index(key)
                    rec.append(num(vals[i]) * UNIT_SCALE.get(units[i], 1.0) * scale)
                else:
                    rec.append(float("nan"))
            dur_s = rec[0] * 1e-6
            gbs = (rec[1] + rec[2]) / dur_s / 1e9 if dur_s > 0 else float("nan")
            w.writerow([name] + [f"{v:.4g}" for v in rec] + [f"{gbs:.1f}"])
    print(open(out).read())


if __name__ == "__main__":
    main()
