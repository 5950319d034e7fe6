"""profiles/r02_kernels.md from the per-kernel ncu capture of scripts/ncu_kernels.py.

usage: python scripts/summarize_kernels.py gpurun_out/r02_kernels_raw.csv gpurun_out/r02_ncu_kernels.log profiles/r02_kernels.md

Input = `ncu -i r02_kernels.ncu-rep --page raw --csv` (one row per profiled launch, metrics as columns) and the driver
script's stdout (section order).  For every launch: duration, DRAM bytes, achieved DRAM GB/s and its fraction of the
measured HBM peak (MEASURED_PEAKS.json hbm_gbs), and — where SURVEY.md 8(d) defines algorithmic bytes for the kernel —
the algorithmic GB/s and the traffic / algorithmic ratio.
"""
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
raw_csv, log, out_md = sys.argv[1], sys.argv[2], sys.argv[3]
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
HBM = peaks["hbm_gbs"]

P_PPO = 266755
P_RB = 2996382
P_AX = 3292837
# (kernel regex, section regex) -> (algorithmic bytes per launch, what the figure is)
ALG = [
    (r"cartpole_step", r"n4096", 4096 * 128, "128 B/env-step in our layout (f64 physics state r+w 64, obs + next_obs 32, action 8, reward/done/counters 24); SURVEY 8d: 89 B with f32 state"),
    (r"cartpole_step", r"n1048576", (1 << 20) * 128, "128 B/env-step (as above)"),
    (r"frames_step", r"", 1024 * 77616, "77 616 B/env-step: read 3 kept frames 21 168, write obs stack 28 224 + next_obs stack 28 224 (reference layout; SURVEY 8d: 28 229 B + the stack shift)"),
    (r"gae_kernel", r"4096x128", 4096 * 128 * 32, "32 B/transition (r, d, V read 12 + V' via shift; adv, ret written 8; standardise re-reads + writes adv 8; SURVEY 8d: 32)"),
    (r"gae_kernel", r"8192x2048", 8192 * 2048 * 32, "32 B/transition"),
    (r"adam_kernel", r"ppo_minibatch", P_PPO * 28, "28 B/param (p, g, m, v read; p, m, v written), P = 266 755"),
    (r"adam_kernel", r"rainbow_learn", P_RB * 28, "28 B/param, P = 2 996 382"),
    (r"rmsprop_centered_kernel", r"apex_learn", P_AX * 32, "32 B/param (p, g, sq, ga read; p, sq, ga written), P = 3 292 837"),
    (r"grad_sumsq_kernel", r"ppo_minibatch", P_PPO * 4, "4 B/param"),
    (r"grad_sumsq_kernel", r"rainbow_learn", P_RB * 4, "4 B/param"),
    (r"grad_sumsq_kernel", r"apex_learn", P_AX * 4, "4 B/param"),
    (r"copy_kernel", r"target_copy", P_RB * 8, "8 B/param (target := online)"),
    (r"per_update_levels", r"B32", 32 * 16 * 21, "16 d B/update, d = 21 levels at 1 M slots (SURVEY 8d: 336 B)"),
    (r"per_update_levels", r"B512", 512 * 16 * 21, "16 d B/update"),
    (r"per_update_leaves", r"B32", 32 * 16, "16 B/update (leaf r+w)"),
    (r"per_update_leaves", r"B512", 512 * 16, "16 B/update"),
    (r"per_sample_kernel", r"B32", 32 * 344, "2 x 8 d + 8 B/draw (SURVEY 8d: 344 B)"),
    (r"per_sample_kernel", r"B512", 512 * 344, "344 B/draw"),
    (r"c51_loss_kernel", r"rainbow_learn", 32 * (3 * 4 * 51 * 4 + 4 * 51 * 4 + 24 + 16), "3 A K 4 B logits in + A K 4 B gradient out per sample (A = 4, K = 51; SURVEY 8d: ~2.5 KB)"),
    (r"td_loss_kernel", r"apex_learn", 512 * (3 * 4 * 4 + 4 * 4 + 24 + 16), "(3 A + A) 4 B + n-step reward/done + weight per sample"),
    (r"im2col_u8", r"rainbow_learn", None, None),
]


def unit_scale(u):
    u = u.strip().lower()
    return {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "usecond": 1e-6, "nsecond": 1e-9,
            "ms": 1e-3, "msecond": 1e-3, "second": 1.0, "s": 1.0, "%": 1, "": 1}.get(u, 1)


rows = list(csv.reader(open(raw_csv)))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
names, units = rows[hdr], rows[hdr + 1]
col = {n: i for i, n in enumerate(names)}
data = [r for r in rows[hdr + 2:] if r and r[0].isdigit()]
sections = [m.group(1) for m in re.finditer(r"^section (\S+) done", open(log).read(), flags=re.M)]


def val(r, name):
    i = col[name]
    return float(r[i].replace(",", "")) * unit_scale(units[i])


# launches appear in section order; the NVTX column (if present) names the section, else cut by the known launch order
nvtx_col = next((i for i, n in enumerate(names) if "NVTX" in n or "nvtx" in n), None)
out = []
for r in data:
    kname = re.sub(r"^void ", "", r[col["Kernel Name"]])
    kname = re.sub(r"\(.*$", "", kname).replace("(anonymous namespace)::", "").replace("<unnamed>::", "")
    sec = ""
    m = re.search(r"jbsec_(\w+)", " ".join(r))
    sec = m.group(1) if m else ""
    dur = val(r, "gpu__time_duration.sum")
    dram = val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum")
    out.append({"section": sec, "kernel": kname[:70], "grid": r[col["Grid Size"]], "block": r[col["Block Size"]],
                "regs": int(val(r, "launch__registers_per_thread")), "us": dur * 1e6, "dram": dram,
                "l2": val(r, "lts__t_bytes.sum") if "lts__t_bytes.sum" in col else 0.0,
                "sm_pct": val(r, "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
                "dram_pct": val(r, "dram__throughput.avg.pct_of_peak_sustained_elapsed")})

with open(out_md, "w") as f:
    f.write("# Round 2 — one ncu capture per kernel of the hot path, at its BASELINE size\n\n"
            "Command (one B200, under gpurun): `ncu --clock-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,"
            "dram__bytes_write.sum,lts__t_bytes.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,dram__throughput.avg.pct_of_peak_sustained_elapsed,"
            "launch__registers_per_thread,launch__grid_size,launch__block_size --profile-from-start off -o gpurun_out/r02_kernels "
            "python scripts/ncu_kernels.py` then `ncu -i ... --page raw --csv`; table built by `scripts/summarize_kernels.py`.\n\n"
            f"ncu times are cold-cache and serialised (caches flushed before every replay pass).  HBM peak = {HBM:.1f} GB/s "
            "(MEASURED_PEAKS.json, driver-measured copy bandwidth).  `DRAM GB/s` = (dram read + write bytes) / duration; "
            "`alg GB/s` = SURVEY.md 8(d) algorithmic bytes / duration; `traffic/alg` > 1 means re-reads or write-allocate "
            "traffic, < 1 means part of the working set stayed in L2 (126 MB) across the flush.  Kernels whose work is a few "
            "hundred KB are launch-latency bound: their GB/s says so, it is not a defect of the access pattern.\n\n"
            "| section (workload) | kernel | grid x block | regs | us | DRAM MB | DRAM GB/s | % of HBM peak | alg MB | alg GB/s | % of peak (alg) | traffic/alg | SM % |\n"
            "|---|---|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|\n")
    for o in out:
        alg = note = None
        for kre, sre, b, n in ALG:
            if re.search(kre, o["kernel"]) and re.search(sre, o["section"]):
                alg, note = b, n
                break
        gbs = o["dram"] / (o["us"] * 1e-6) / 1e9 if o["us"] > 0 else 0
        if alg:
            ags = alg / (o["us"] * 1e-6) / 1e9
            acell = f"{alg / 1e6:.3f} | {ags:.1f} | {100 * ags / HBM:.1f}% | {o['dram'] / alg:.2f}"
        else:
            acell = "– | – | – | –"
        f.write(f"| {o['section']} | `{o['kernel']}` | {o['grid']} x {o['block']} | {o['regs']} | {o['us']:.1f} | {o['dram'] / 1e6:.3f} | "
                f"{gbs:.1f} | {100 * gbs / HBM:.1f}% | {acell} | {o['sm_pct']:.1f} |\n")
    f.write("\n## Algorithmic bytes used above\n\n")
    seen = set()
    for kre, sre, b, n in ALG:
        if n and (kre, n) not in seen:
            seen.add((kre, n))
            f.write(f"* `{kre}` ({sre or 'all sizes'}): {n}\n")
print(f"{len(out)} launches -> {out_md}")
