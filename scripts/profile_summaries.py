"""Turns the ncu outputs of a round into the committed summaries under profiles/.
usage: profile_summaries.py <launch_list.csv> <spread.ncu-rep> <round> [<interp.ncu-rep>]"""
import csv, json, shutil, subprocess, sys, io
lst, rep, rnd = sys.argv[1], sys.argv[2], int(sys.argv[3])
root = __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__)))
shutil.copy(lst, f"{root}/profiles/r{rnd:02d}_launch_list_bench_steps2.csv")
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
starts = [i for i, n in enumerate(names) if n.startswith("halo_zero_ghost") and (i == 0 or not names[i - 1].startswith("halo_zero_ghost"))]
s0, s1 = starts[-3], starts[-2]  # one timed step (the last segment holds the e2e step's uploads and re-bin as well)
agg = {}
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
                     f"--e2e-steps 1 (raw list: r{rnd:02d}_launch_list_bench_steps2.csv)",
           "what": "one timed step = spreadForce (ghost zero, face park, 8 tile-colour launches, fix-up, halo accumulate, face sync/restore) + "
                   "interpolateVelocity (halo fill, interp)",
           "per_kernel": agg, "step_total_us": round(tot, 1)}, open(f"{root}/profiles/r{rnd:02d}_step_kernels_dram.json", "w"), indent=1)
# full-set summary of the spread kernel
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(io.StringIO(out)))
hh, u, v = r[0], r[1], r[2]
want = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "lts__t_sector_hit_rate.pct", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"]
summ = {}
for k in want:
    for i, x in enumerate(hh):
        if x == k:
            summ[k] = {"value": v[i], "unit": u[i]}
st = {}
for i, x in enumerate(hh):
    if x.startswith("smsp__pcsamp_warps_issue_stalled_") and not x.endswith("_not_issued"):
        try:
            st[x.replace("smsp__pcsamp_warps_issue_stalled_", "")] = float(v[i].replace(",", ""))
        except ValueError:
            pass
t = sum(st.values())
summ["stall_share_of_pc_samples"] = {k: round(x / t, 4) for k, x in sorted(st.items(), key=lambda kv: -kv[1]) if x / t > 0.01}
path = f"{root}/profiles/r{rnd:02d}_ncu_full_summary.json"
try:
    prof = json.load(open(path))
except Exception:
    prof = {}
key = ("spread_tile_kernel<3,IB_4> (current), ONE of the 8 tile-colour launches of a spread; ncu --set full --clock-control none --import-source on "
       "-k regex:spread_tile -s 10 -c 1 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 1")
new = {key: summ}
for k, vv in prof.items():
    if k.startswith("spread_tile_kernel<3,IB_4> (current"):
        new["(older) " + k.replace("(current", "(earlier")] = vv
    else:
        new[k] = vv
if len(sys.argv) > 4:  # the interpolation kernel's full-set capture
    out = subprocess.run(["ncu", "-i", sys.argv[4], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(io.StringIO(out)))
    hh, u, v = r[0], r[1], r[2]
    isum = {}
    for k in want + ["l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "memory_l1_wavefronts_shared_ideal"]:
        for i, x in enumerate(hh):
            if x == k:
                isum[k] = {"value": v[i], "unit": u[i]}
    st = {}
    for i, x in enumerate(hh):
        if x.startswith("smsp__pcsamp_warps_issue_stalled_") and not x.endswith("_not_issued"):
            try:
                st[x.replace("smsp__pcsamp_warps_issue_stalled_", "")] = float(v[i].replace(",", ""))
            except ValueError:
                pass
    t = sum(st.values())
    isum["stall_share_of_pc_samples"] = {k: round(x / t, 4) for k, x in sorted(st.items(), key=lambda kv: -kv[1]) if x / t > 0.01}
    ikey = ("interp_rot_kernel<IB_4,320> (current); ncu --set full --clock-control none --import-source on -k regex:interp_rot -s 3 -c 1 "
            "python bench.py --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 1")
    new = {k: vv for k, vv in new.items() if not k.startswith("interp_rot_kernel")}
    new = {ikey: isum, **new}
json.dump(new, open(path, "w"), indent=1)
print(json.dumps(summ["stall_share_of_pc_samples"]))
print({k: summ[k]["value"] for k in ("gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum")})
