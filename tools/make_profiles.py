"""Turn the raw captures under gpurun_out/ into the tracked summaries under profiles/ (files named rNN_*, NN = argv[1], default 02).

inputs (produced on the GPU box, see profiles/README.md for the exact commands):
  gpurun_out/launches_bench.csv      ncu --metrics gpu__time_duration.sum on bench.py
  gpurun_out/prof_nn_final.ncu-rep   ncu --set full, nn_accum_kernel, first 5 launches of one registration
  gpurun_out/prof_normals_final.ncu-rep
"""
import csv, json, os, subprocess, sys, collections, re
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RND = "r" + (sys.argv[1] if len(sys.argv) > 1 else "02")
OUT = os.path.join(ROOT, "profiles"); G = os.path.join(ROOT, "gpurun_out")
sys.path.insert(0, os.path.join(ROOT, "tools"))

def sh(cmd): return subprocess.run(cmd, shell=True, capture_output=True, text=True).stdout

def raw_csv(rep):
    path = "/tmp/_raw.csv"
    open(path, "w").write(sh(f"ncu -i {rep} --page raw --csv 2>/dev/null"))
    return list(csv.reader(open(path)))

# 1. launch list
src = os.path.join(G, "launches_bench.csv")
if os.path.exists(src):
    open(os.path.join(OUT, RND + "_bench_launches.csv"), "w").write(open(src).read())
    open(os.path.join(OUT, RND + "_bench_launch_summary.txt"), "w").write(
        "# python bench.py --steps 2 --warmup 3 --no-cpu-baseline under ncu --metrics gpu__time_duration.sum --clock-control none\n"
        "# (kernels inside the WHILE-node graph of the `value` region are not listed by ncu; the same kernels appear\n"
        "#  from the stream-mode roofline pass and the e2e pass)\n" + sh(f"python {ROOT}/tools/summarize_launches.py {src}"))

# 2. dominant kernel
rep = os.path.join(G, "prof_nn_final.ncu-rep")
if os.path.exists(rep):
    rows = raw_csv(rep); h = {n: i for i, n in enumerate(rows[0])}
    probe = [r for r in rows[2:] if "probe_kernel" in r[h["Kernel Name"]]]
    data = [r for r in rows[2:] if "nn_accum" in r[h["Kernel Name"]]]      # the capture interleaves probe and search launches
    def col(name): return [float(r[h[name]].replace(",", "")) for r in data]
    unit = lambda name: rows[1][h[name]]
    def to_bytes(v, u): return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
    rd = [to_bytes(v, unit("dram__bytes_read.sum")) for v in col("dram__bytes_read.sum")]
    wr = [to_bytes(v, unit("dram__bytes_write.sum")) for v in col("dram__bytes_write.sum")]
    dur = col("gpu__time_duration.sum")
    summary = {
        "probe_kernel_duration_" + unit("gpu__time_duration.sum"): [float(r[h["gpu__time_duration.sum"]].replace(",", "")) for r in probe],
        "probe_kernel_warp_inst": [float(r[h["smsp__inst_executed.sum"]].replace(",", "")) for r in probe],
        "kernel": "nn_accum_kernel", "points": 1000000,
        "capture": "ncu --set full --clock-control none, every search of one 1M-point registration "
                   "(far, mid, near, then converged ones that mostly skip their walks); cold-cache, serialised",
        "duration_" + unit("gpu__time_duration.sum"): dur,
        "dram_bytes_read": rd, "dram_bytes_write": wr,
        "dram_bytes_per_launch_mean": sum(a + b for a, b in zip(rd, wr)) / len(rd),
        "algorithmic_bytes_per_launch": 64 * 1_000_000,
        "registers_per_thread": col("launch__registers_per_thread")[0],
        "issue_active_pct": col("smsp__issue_active.avg.pct_of_peak_sustained_active"),
        "thread_inst_per_inst": col("smsp__thread_inst_executed_per_inst_executed.ratio"),
        "warp_instructions": col("smsp__inst_executed.sum"),
        "warp_inst_per_launch": col("smsp__inst_executed.sum"),
        "l1_hit_pct": col("l1tex__t_sector_hit_rate.pct"), "l2_hit_pct": col("lts__t_sector_hit_rate.pct"),
        "dram_throughput_pct": col("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
        "pipe_alu_pct": col("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active"),
        "pipe_fma_pct": col("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
    }
    json.dump(summary, open(os.path.join(OUT, RND + "_nn_accum.json"), "w"), indent=1)
    with open("/tmp/_raw.csv", "w", newline="") as f_:
        csv.writer(f_).writerows(rows[:2] + data)
    open(os.path.join(OUT, RND + "_nn_accum_ncu_metrics.txt"), "w").write(
        "# nn_accum_kernel, ncu --set full --clock-control none --import-source on, all searches of one 1M registration\n"
        "# (columns: far first search, second, third, then the converged ones)\n" + sh(f"python {ROOT}/tools/ncu_key_metrics.py /tmp/_raw.csv"))
    open("/tmp/_src.csv", "w").write(sh(f"ncu -i {rep} --page source --print-source cuda,sass --csv 2>/dev/null"))
    open(os.path.join(OUT, RND + "_nn_accum_source_hotspots.txt"), "w").write(
        "# executed warp-instructions and stall samples per CUDA source line; launches of the capture: 0 probe, 1 far first search,\n# 2 probe, 3 second search, 4 (idle probe), 5 third, 6, 7 fourth (the first that records margins)\n"
        + sh(f"python {ROOT}/tools/ncu_source_lines.py /tmp/_src.csv 1 3 7"))
    # the walk itself is an out-of-line device function: its SASS is listed under nn_walk

# 2b. the 16M pair (working set far beyond L2: the DRAM-bound configuration)
rep = os.path.join(G, "prof_nn_16m_raw.csv")   # exported on the GPU box: ncu -i prof_nn_16m.ncu-rep --page raw --csv
if os.path.exists(rep):
    rows = list(csv.reader(open(rep))); h = {n: i for i, n in enumerate(rows[0])}; data = rows[2:]
    def col(name): return [float(r[h[name]].replace(",", "")) for r in data]
    unit = lambda name: rows[1][h[name]]
    def to_bytes(v, u): return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
    rd = [to_bytes(v, unit("dram__bytes_read.sum")) for v in col("dram__bytes_read.sum")]
    wr = [to_bytes(v, unit("dram__bytes_write.sum")) for v in col("dram__bytes_write.sum")]
    json.dump({"kernel": "nn_accum_kernel", "points": 16 * 2 ** 20,
               "capture": "ncu --set full --clock-control none, the first 10 searches of one 16M-point registration",
               "duration_" + unit("gpu__time_duration.sum"): col("gpu__time_duration.sum"),
               "dram_bytes_read": rd, "dram_bytes_write": wr, "algorithmic_bytes_per_launch": 64 * 16 * 2 ** 20,
               "dram_bytes_per_launch_mean": sum(a + b for a, b in zip(rd, wr)) / len(rd),
               "warp_inst_per_launch": col("smsp__inst_executed.sum"),
               "issue_active_pct": col("smsp__issue_active.avg.pct_of_peak_sustained_active"),
               "dram_throughput_pct": col("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
               "l1_hit_pct": col("l1tex__t_sector_hit_rate.pct"), "l2_hit_pct": col("lts__t_sector_hit_rate.pct")},
              open(os.path.join(OUT, RND + "_nn_accum_16m.json"), "w"), indent=1)
    open(os.path.join(OUT, RND + "_nn_accum_16m_ncu_metrics.txt"), "w").write(
        "# nn_accum_kernel, ncu --set full --clock-control none, the first 10 searches of one 16M-point registration (config 4 on one GPU)\n"
        + sh(f"python {ROOT}/tools/ncu_key_metrics.py {rep}"))

rep = os.path.join(G, "prof_normals_raw.csv")
if os.path.exists(rep):
    open(os.path.join(OUT, RND + "_normals_ncu_metrics.txt"), "w").write(
        "# normals_kernel<10,true>, ncu --set full, 1M points\n" + sh(f"python {ROOT}/tools/ncu_key_metrics.py {rep}"))

# 3. SASS evidence
obj = os.path.join(ROOT, "icp-symm_b200", "csrc", "icp_loop.o")
sass = sh(f"cuobjdump -sass {obj}")
ms_ = [re.search(r"Function : \S*" + nm_ + r"\S*\n(.*?)(?=\n\s*Function :|\Z)", sass, re.S)
       for nm_ in ("nn_accum_kernel", "nn_walkILb0", "nn_walkILb1")]
m = ms_[0]
if m:
    body = "\n".join(f"// ---- {nm_}\n" + x.group(1) for nm_, x in zip(("nn_accum_kernel", "nn_walk<false> (out-of-line walk)", "nn_walk<true> (walk that records margins)"), ms_) if x)
    ops = collections.Counter(re.sub(r"\..*", "", x) for x in re.findall(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P[0-9T]+\s+)?([A-Z0-9_.]+)", body))
    lines = ["# static SASS opcode histogram of nn_accum_kernel (sm_100a); FADD2/FMUL2 = packed fp32x2, CREDUX/REDUX = warp reductions,",
             "# FMNMX3 = 3-input min/max; no HMMA/UTC*MMA: the kernel has no dense contraction (DESIGN.md section 8)"]
    lines += [f"{n:6d} {op}" for op, n in ops.most_common()]
    open(os.path.join(OUT, RND + "_nn_accum_sass_histogram.txt"), "w").write("\n".join(lines) + "\n")
    open(os.path.join(OUT, RND + "_nn_accum.sass"), "w").write(body)
print("profiles written:", sorted(os.listdir(OUT)))
