#!/usr/bin/env python
"""Compact per-launch summary of an `ncu --page raw --csv` export (the judged copy goes under profiles/).
   tools/ncu_summary.py gpurun_out/conv_raw.csv profiles/r01f_ncu_conv_gemm.csv"""
import csv
import sys

KEEP = [
    ("Kernel Name", "kernel"),
    ("launch__grid_size", "grid"),
    ("launch__registers_per_thread", "regs"),
    ("gpu__time_duration.sum", "duration_us"),
    ("sm__cycles_elapsed.avg.per_second", "sm_ghz"),
    ("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pipe_active_pct"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_throughput_pct"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_throughput_pct"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_throughput_pct"),
    ("dram__bytes_read.sum", "dram_read"),
    ("dram__bytes_write.sum", "dram_write"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"),
]


def main(src, dst):
    rows = list(csv.reader(open(src)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    cols = [(ix[k], n, units[ix[k]]) for k, n in KEEP if k in ix]
    with open(dst, "w", newline="") as fh:
        w = csv.writer(fh)
        w.writerow(["launch"] + [f"{n} [{u}]" if u else n for _, n, u in cols])
        for k, r in enumerate(data):
            if len(r) < len(hdr):
                continue
            w.writerow([k] + [r[i][:60] for i, _, _ in cols])
    print(f"{dst}: {len(data)} launches")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
