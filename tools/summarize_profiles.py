#!/usr/bin/env python
"""Turn the raw ncu outputs that tools/make_profiles.sh left in gpurun_out/ into the committed text
summaries under profiles/ (run in the build container: ncu -i works without a GPU).
    python tools/summarize_profiles.py [tag=r1] [round=r01]"""
import collections, csv, io, os, shutil, subprocess, sys
tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
rnd = sys.argv[2] if len(sys.argv) > 2 else "r01"
os.makedirs("profiles", exist_ok=True)

# ---- launch list
rows = list(csv.reader(open(f"gpurun_out/launches_{tag}.csv")))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
hdr, data = rows[hi], rows[hi + 1:]
ik, iv, im = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
ig = hdr.index("Grid Size") if "Grid Size" in hdr else None
agg = collections.OrderedDict()
for r in data:
    if len(r) <= iv or r[im] != "gpu__time_duration.sum":
        continue
    key = (r[ik].split("(")[0], r[ig] if ig is not None else "")
    a = agg.setdefault(key, [0, 0.0]); a[0] += 1; a[1] += float(r[iv].replace(",", ""))
tot = sum(v[1] for v in agg.values())
lines = ["# ncu launch list of `python bench.py --steps 20 --warmup 10 --no-cpu-baseline` (first 400 launches)",
         "# command: ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv (tools/make_profiles.sh)",
         "# per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes",
         "kernel | grid | launches | total_us | mean_us | share"]
for (name, grid), (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    lines.append(f"{name} | {grid} | {n} | {t/1000:.1f} | {t/1000/n:.2f} | {t/tot:.3f}")
open(f"profiles/{rnd}_launches_bench.txt", "w").write("\n".join(lines) + "\n")

# ---- full capture
raw = subprocess.run(["ncu", "-i", f"gpurun_out/prof_fwd_l4_{tag}.ncu-rep", "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum",
        "sm__cycles_elapsed.max", "smsp__cycles_active.avg", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
out = ["# ncu --set full --clock-control none --import-source on -k regex:warp_corr_fwd -s 15 -c 2  python tools/profile_level.py 4 0",
       "# kernel: cerb::warp_corr_fwd_kernel<float,8,32,1,4,3> (finest PWC level: B=1, C=32, 128x256, warped, rotating buffers -> inputs from HBM)",
       "# two launches captured; values per launch", ""]
for w in want:
    if w in hdr:
        i = hdr.index(w)
        out.append(f"{w} [{units[i]}]: {', '.join(r[i] for r in data)}")
out += ["", "warp stall reasons (issue-stalled warps per issue-active cycle, launch 0):"]
for i, h in enumerate(hdr):
    if "smsp__average_warps_issue_stalled" in h and h.endswith("_per_issue_active.ratio") and "not_issued" not in h:
        v = float(data[0][i])
        if v >= 0.05:
            out.append(f"  {h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')}: {v:.2f}")
i_r, i_w = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
out += ["", f"dram traffic per launch: read {data[0][i_r]} {units[i_r]} + write {data[0][i_w]} {units[i_w]}",
        "algorithmic bytes per launch (SURVEY 8d): 128*256*4*(2*32+81+2) = 19.268 MB (8.65 MB read + 10.62 MB write)",
        "reads match the algorithmic input bytes (no over-fetch from HBM: raw-box overlap between tiles is served by L2);",
        "the 10.6 MB of output is still dirty in the 126 MB L2 when the kernel ends, so ncu sees ~0 DRAM write bytes in-kernel."]
open(f"profiles/{rnd}_ncu_fwd_finest_level.txt", "w").write("\n".join(out) + "\n")
for src, dst in ((f"gpurun_out/bench_{tag}.json", f"profiles/{rnd}_bench_n1.json"), (f"gpurun_out/bench_ref_{tag}.json", f"profiles/{rnd}_bench_reference_arm.json")):
    if os.path.exists(src):
        shutil.copy(src, dst)
print(open(f"profiles/{rnd}_launches_bench.txt").read())
print("\n".join(out[4:12]))
