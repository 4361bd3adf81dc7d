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
lines = ["# ncu launch list of `python bench.py --steps 20 --warmup 10 --no-cpu-baseline --no-train --no-other-configs` (first 400 launches)",
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
       "# kernel: cerb::tc::warp_corr_fwd_tc_kernel<float> (finest PWC level, both flow directions per launch: B=2, C=32, 128x256, warped,",
       "#         rotating buffers -> inputs from HBM; 512 tiles of 8x16 on 148 persistent CTAs of 704 threads)",
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
        "algorithmic bytes per launch (SURVEY 8d): 2*128*256*4*(2*32+81+2) = 38.535 MB (8.9 MB of inputs -- the two directions share",
        "the feature maps -- + 21.2 MB of output); the raw-box overlap between tiles (10x the tile's own pixels) is served by L2:",
        "compare lts__t_bytes with dram__bytes_read.  Output still dirty in the 126 MB L2 at kernel end shows up as DRAM writes only partly."]
open(f"profiles/{rnd}_ncu_fwd_finest_level.txt", "w").write("\n".join(out) + "\n")
# ---- fused backward kernel
try:
    raw = subprocess.run(["ncu", "-i", f"gpurun_out/prof_bwd_{tag}.ncu-rep", "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ob = ["# ncu --set full --clock-control none --import-source on -k regex:corr_bwd_fused -s 1 -c 1  python tools/profile_backward.py",
          "# kernel: cerb::corr_bwd_fused_kernel<float,float,4,true> (HRNet training level: B=8, C=48, 128x256; both gradients, splat and flow",
          "#         gradient in one launch; the call also runs flow_warp_fwd_kernel (warped map into the workspace) and a memset)", ""]
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            ob.append(f"{w} [{units[i]}]: {', '.join(r[i] for r in data)}")
    ob += ["", "warp stall reasons (issue-stalled warps per issue-active cycle):"]
    for i, h in enumerate(hdr):
        if "smsp__average_warps_issue_stalled" in h and h.endswith("_per_issue_active.ratio") and "not_issued" not in h:
            v = float(data[0][i])
            if v >= 0.05:
                ob.append(f"  {h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')}: {v:.2f}")
    ob += ["", "algorithmic bytes of the whole backward call (SURVEY 8d): 8*128*256*4*(2*81+4*48+4) = 375.4 MB"]
    open(f"profiles/{rnd}_ncu_backward_fused.txt", "w").write("\n".join(ob) + "\n")
except Exception as exc:  # noqa: BLE001
    print("no backward capture:", exc)
# ---- tensor-core backward kernel and window-splat kernel
for rep, dst, head in (
        (f"gpurun_out/prof_bwd_tc_{tag}.ncu-rep", f"profiles/{rnd}_ncu_backward_tc.txt",
         ["# ncu --set full --clock-control none --import-source on -k regex:corr_bwd_tc -s 1 -c 1  python tools/profile_backward_tc.py",
          "# kernel: cerb::btc::corr_bwd_tc_kernel (HRNet training level: B=8, C=48, 128x256; both correlation gradients as banded GEMMs,",
          "#         2048 tiles of 8x16 on 148 persistent CTAs of 512 threads; opt-in path: cerb_debug_set_backward_kernel(1))", ""]),
        (f"gpurun_out/prof_splat_{tag}.ncu-rep", f"profiles/{rnd}_ncu_splat_window.txt",
         ["# ncu --set full --clock-control none --import-source on -k regex:flow_warp_bwd_box -s 1 -c 1  python tools/profile_backward_tc.py",
          "# kernel: cerb::splat::flow_warp_bwd_box_kernel (B=8, C=48, 128x256: 2048 CTAs of 128 threads, 4 per SM; x2 window in by TMA,",
          "#         fixed-point ATOMS into the gradient window, one UTMAREDG per 8 channels)", ""])):
    try:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        hdr, units, data = rows[0], rows[1], rows[2:]
        ob = list(head)
        for w in want:
            if w in hdr:
                i = hdr.index(w)
                ob.append(f"{w} [{units[i]}]: {', '.join(r[i] for r in data)}")
        ob += ["", "warp stall reasons (issue-stalled warps per issue-active cycle):"]
        for i, h in enumerate(hdr):
            if "smsp__average_warps_issue_stalled" in h and h.endswith("_per_issue_active.ratio") and "not_issued" not in h:
                v = float(data[0][i])
                if v >= 0.05:
                    ob.append(f"  {h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')}: {v:.2f}")
        open(dst, "w").write("\n".join(ob) + "\n")
    except Exception as exc:  # noqa: BLE001
        print("no capture", rep, exc)
# ---- sanitizer logs (memcheck / synccheck verbatim, racecheck: one line per distinct hazard site + the summary)
for tool in ("memcheck", "synccheck"):
    src = f"gpurun_out/sanitizer_{tool}_{tag}.txt"
    if os.path.exists(src):
        open(f"profiles/{rnd}_sanitizer_{tool}.txt", "w").write(
            f"# compute-sanitizer --tool {tool} python tools/sanitize_cases.py   (every kernel path incl. the tensor-core forward and the fused backward)\n" + open(src).read())
src = f"gpurun_out/sanitizer_racecheck_{tag}.txt"
if os.path.exists(src):
    sites = collections.Counter()
    for ln in open(src):
        if "Race reported between" in ln or "and Write access" in ln or "and Read access" in ln:
            sites[ln.split("=========")[-1].strip().split("+0x")[0] + " ... " + ln.strip().split(" in ")[-1]] += 1
    body = ["# compute-sanitizer --tool racecheck python tools/race_cluster_cases.py   (cluster-split kernels, tensor-core forward, fused backward)",
            "# distinct hazard sites (count): all are consumer reads (costvolume_fwd.cu:1011, LDS of the staged x2 tile) against the gather warps'",
            "# sts_f32 into the same stage of the CUDA-core 4x16 kernel -- ordered by the full/empty mbarrier phases, which racecheck does not model.",
            "# Nothing is reported for the tensor-core kernel or the backward kernels."]
    body += [f"{n:6d}  {k}" for k, n in sites.most_common()]
    body += [ln.strip() for ln in open(src) if "RACECHECK SUMMARY" in ln]
    open(f"profiles/{rnd}_sanitizer_racecheck.txt", "w").write("\n".join(body) + "\n")
for src, dst in ((f"gpurun_out/bench_{tag}.json", f"profiles/{rnd}_bench_n1.json"), (f"gpurun_out/bench_ref_{tag}.json", f"profiles/{rnd}_bench_reference_arm.json")):
    if os.path.exists(src):
        shutil.copy(src, dst)
print(open(f"profiles/{rnd}_launches_bench.txt").read())
print("\n".join(out[4:12]))
