#!/usr/bin/env python
"""Cut the LAST guided step out of an `ncu --metrics gpu__time_duration.sum --csv` log of bench.py --ncu and print the
per-kernel-family shares (the check that the conv kernel's share of the step agrees with bench.py's live events).
   tools/ncu_launch_list.py launches_all.csv out_last_step.csv"""
import csv
import sys

src, dst = sys.argv[1], sys.argv[2]
rows = [r for r in csv.reader(l for l in open(src) if not l.startswith("==")) if r]
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
recs = []
for r in rows[1:]:
    if len(r) < len(hdr):
        continue
    try:
        v = float(r[ix["Metric Value"]].replace(",", ""))
    except ValueError:
        continue
    unit = r[ix["Metric Unit"]]
    us = v / 1e3 if unit in ("nsecond", "ns") else v * 1e3 if unit in ("msecond", "ms") else v
    recs.append((r[ix["Kernel Name"]], us))
# a step = [prep_x ... ddpm/ddim step kernel]; take the last complete one
starts = [i for i, (k, _) in enumerate(recs) if "prep_x" in k]
ends = [i for i, (k, _) in enumerate(recs) if "ddpm_step_kernel" in k or "ddim_step_kernel" in k]
end = ends[-1]
start = max(s for s in starts if s < end)
step = recs[start:end + 1]
fam = {}
for k, us in step:
    name = ("conv_gemm_kernel" if "conv_gemm" in k else "gn_apply" if "gn_apply" in k else "gn_finalize/stats" if "gn_" in k
            else "attention" if "attn" in k else "sampler update" if "_step_kernel" in k else "prologue/misc")
    f = fam.setdefault(name, [0.0, 0])
    f[0] += us
    f[1] += 1
tot = sum(us for _, us in step)
with open(dst, "w", newline="") as fh:
    w = csv.writer(fh)
    w.writerow(["launch", "kernel", "gpu__time_duration_us"])
    for i, (k, us) in enumerate(step):
        w.writerow([i, k[:90], f"{us:.2f}"])
    w.writerow([])
    w.writerow(["family", "launches", "us", "share"])
    for name, (us, n) in sorted(fam.items(), key=lambda kv: -kv[1][0]):
        w.writerow([name, n, f"{us:.1f}", f"{us / tot:.4f}"])
    w.writerow(["total", len(step), f"{tot:.1f}", "1.0"])
print(f"{dst}: last step = launches {start}..{end} of {len(recs)} ({len(step)} launches, {tot / 1e3:.3f} ms serialised)")
for name, (us, n) in sorted(fam.items(), key=lambda kv: -kv[1][0]):
    print(f"  {name:20s} {n:4d} launches {us / 1e3:8.3f} ms  {100 * us / tot:5.1f} %")
