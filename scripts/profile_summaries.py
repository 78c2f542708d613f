"""Turns the ncu outputs of a round (scripts/r2_profile.sh) into the committed summaries under profiles/.
usage: profile_summaries.py <round> <launch_list.csv> <spread.ncu-rep> <interp.ncu-rep> [libibk.so]"""
import collections, csv, io, json, os, re, shutil, subprocess, sys

rnd, lst, rep_s, rep_i = int(sys.argv[1]), sys.argv[2], sys.argv[3], sys.argv[4]
lib = sys.argv[5] if len(sys.argv) > 5 else None
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = f"r{rnd:02d}"
out = lambda name: os.path.join(root, "profiles", f"{tag}_{name}")
shutil.copy(lst, out("launch_list_bench_steps2.csv"))

# ---- the launches of ONE timed step, aggregated per kernel
rows = list(csv.reader(open(lst)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
h = rows[hi]
idx = {k: h.index(k) for k in ("ID", "Kernel Name", "Metric Name", "Metric Unit", "Metric Value")}
scale = {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
per, order = {}, []
for r in rows[hi + 1:]:
    if len(r) < len(h):
        continue
    k = int(r[idx["ID"]])
    if k not in per:
        per[k] = {"name": r[idx["Kernel Name"]]}
        order.append(k)
    per[k][r[idx["Metric Name"]]] = float(r[idx["Metric Value"]].replace(",", "")) * scale.get(r[idx["Metric Unit"]], 1.0)
names = [per[k]["name"].split("(")[0].replace("void ", "") for k in order]
# a step starts with the ghost zero of ibk_spread_begin (region_items_kernel<2>) and ends after the interpolation kernel
starts = [i for i, n in enumerate(names) if n.startswith("region_items_kernel<2>")]
ends = [i for i, n in enumerate(names) if n.startswith("interp_")]
s0 = starts[-3]
s1 = next(e for e in ends if e > s0) + 1
agg = collections.OrderedDict()
for i in range(s0, s1):
    p, n = per[order[i]], names[i]
    a = agg.setdefault(n, {"launches": 0, "time_us": 0.0, "dram_read_bytes": 0.0, "dram_write_bytes": 0.0})
    a["launches"] += 1
    a["time_us"] += p["gpu__time_duration.sum"] * 1e6
    a["dram_read_bytes"] += p["dram__bytes_read.sum"]
    a["dram_write_bytes"] += p["dram__bytes_write.sum"]
tot = sum(a["time_us"] for a in agg.values())
for n, a in agg.items():
    a["share_of_step"] = round(a["time_us"] / tot, 4)
    a["time_us"] = round(a["time_us"], 1)
    print(f"{n:40s} x{a['launches']:3d} {a['time_us']:9.1f} us {a['share_of_step'] * 100:5.1f}%  rd {a['dram_read_bytes'] / 1e6:8.1f} MB wr {a['dram_write_bytes'] / 1e6:8.1f} MB")
json.dump({"round": rnd,
           "source": "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 (cold-cache, serialised "
                     "launches; shares are comparable, absolute times are not bench values) of: python bench.py --steps 2 --warmup 1 --no-cpu-baseline "
                     f"--no-sample-parity --e2e-steps 1 (raw list: {tag}_launch_list_bench_steps2.csv)",
           "what": "one timed step at N = 1 = spreadForce (ghost zero, face park, ONE persistent march launch, fix-up, face sync, halo accumulate, face "
                   "restore) + interpolateVelocity (halo fill, interp)",
           "per_kernel": agg, "step_total_us": round(tot, 1)}, open(out("step_kernels_dram.json"), "w"), indent=1)

# ---- selected counters and stall shares of the two full-set captures
want = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "memory_l1_wavefronts_shared_ideal", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "lts__t_sector_hit_rate.pct",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"]


def summary(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(io.StringIO(raw)))
    hh, u, v = r[0], r[1], r[2]
    d = {"kernel": v[hh.index("Kernel Name")]}
    for k in want:
        if k in hh:
            d[k] = {"value": v[hh.index(k)], "unit": u[hh.index(k)]}
    st = {}
    for i, x in enumerate(hh):
        if x.startswith("smsp__pcsamp_warps_issue_stalled_") and not x.endswith("_not_issued"):
            try:
                st[x.replace("smsp__pcsamp_warps_issue_stalled_", "")] = float(v[i].replace(",", ""))
            except ValueError:
                pass
    t = sum(st.values())
    d["stall_share_of_pc_samples"] = {k: round(x / t, 4) for k, x in sorted(st.items(), key=lambda kv: -kv[1]) if x / t > 0.01}
    return d


full = {"how": "ncu --set full --clock-control none --import-source on -k regex:<kernel> -s 1 -c 1 python bench.py --steps 1 --warmup 1 "
               "--no-cpu-baseline --no-sample-parity --e2e-steps 0 (C5 shard, one B200)"}
for key, rep, nm in (("spread", rep_s, "spread_march_kernel"), ("interp", rep_i, "interp_rot_kernel")):
    full[key] = summary(rep)
    det = subprocess.run(["ncu", "-i", rep, "--page", "details"], capture_output=True, text=True).stdout.split("\n")
    open(out(f"{nm}_details.txt"), "w").write("\n".join(det[:230]) + "\n")
json.dump(full, open(out("ncu_full_summary.json"), "w"), indent=1)
for key in ("spread", "interp"):
    d = full[key]
    print(key, d["kernel"][:60], {k: d[k]["value"] for k in ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
                                                              "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed")})
    print("   stalls", d["stall_share_of_pc_samples"])

# ---- SASS evidence of TMA in the shipped binary: UTMA* mnemonics per kernel
if lib:
    sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    counts, cur = collections.OrderedDict(), None
    for line in sass.split("\n"):
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
            continue
        m = re.search(r"\b(UTMALDG|UTMASTG|UTMAREDG|UTMAPF|UBLKCP|UBLKRED|SYNCS|REDG\.E\.ADD\.F64|RED\.E\.ADD\.F64|ATOMG\S*F64)\S*", line)
        if m and cur:
            counts.setdefault(cur, collections.Counter())[m.group(1)] += 1
    with open(out("sass_tma_counts.txt"), "w") as f:
        f.write(f"cuobjdump -sass {os.path.basename(lib)} | mnemonic counts per kernel (TMA tensor loads / stores / reducing stores, bulk copies,\n"
                "mbarrier SYNCS; a line with REDG/RED/ATOMG ...F64 would be a global floating-point atomic: there is none)\n")
        for k, c in counts.items():
            f.write(f"{k:70s} " + "  ".join(f"{m}={n}" for m, n in sorted(c.items())) + "\n")
    print(open(out("sass_tma_counts.txt")).read())
