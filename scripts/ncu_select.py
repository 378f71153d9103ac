"""`ncu -i X.ncu-rep --page raw --csv` -> the handful of columns profiles/*_ncu_full*.csv keep, one row per launch.
usage: python scripts/ncu_select.py raw.csv out.csv"""
import csv
import re
import sys

COLS = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram_read"), ("dram__bytes_write.sum", "dram_write"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pipe_pct"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"),
        ("sm__issue_active.avg.pct_of_peak_sustained_elapsed", "issue_active_pct"), ("smsp__inst_executed.sum", "warp_inst"),
        ("launch__registers_per_thread", "regs"), ("lts__t_sector_hit_rate.pct", "l2_hit_pct"),
        ("launch__occupancy_limit_registers", "occ_lim_regs")]


def short(name):
    """st3d::k_foo<1, 0>(args...) -> k_foo<1, 0>"""
    name = re.sub(r"^void\s+", "", name)
    name = name.split("(")[0]
    return re.sub(r"^(st3d::|tc::)+", "", name)


def main(src, dst):
    with open(src, newline="") as fh:
        rows = list(csv.reader(l for l in fh if not l.startswith("==")))
    hdr, units, body = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    with open(dst, "w", newline="") as out:
        w = csv.writer(out)
        w.writerow(["kernel", "grid"] + [f"{s} [{units[idx[c]]}]" for c, s in COLS])
        for r in body:
            w.writerow([short(r[idx["Kernel Name"]]), r[idx["Grid Size"]]] + [r[idx[c]] for c, _ in COLS])
    print(f"{dst}: {len(body)} launches")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
