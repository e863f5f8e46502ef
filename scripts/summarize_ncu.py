"""Turns gpurun_out/*.ncu-rep + launches_<tag>.csv into small tracked summaries under profiles/ (run here, no GPU)."""
import collections, csv, io, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")
METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
           "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.avg", "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
           "smsp__inst_executed.sum", "sm__inst_executed_pipe_xu.sum", "lts__t_sector_hit_rate.pct",
           "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.per_cycle_active",
           "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
           "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
           "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum", "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum",
           "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum", "memory_l2_theoretical_sectors_local", "launch__stack_size",
           "smsp__average_warp_latency_per_inst_issued.ratio", "smsp__pcsamp_sample_count"]
SCALE = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}


def short(name):
    return re.sub(r"\(.*", "", name).split("::")[-1]


def launch_list(tag, cmd):
    p = os.path.join(OUT, f"launches_{tag}.csv")
    if not os.path.exists(p):
        return
    rows = list(csv.reader(open(p)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr, data = rows[hi], rows[hi + 1:]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.defaultdict(lambda: [0, 0.0])
    per = []
    for r in data:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", "")) * SCALE.get(r[ui], 1.0)
        agg[short(r[ki])][0] += 1
        agg[short(r[ki])][1] += v
        per.append((r[0], short(r[ki]), v))
    tot = sum(v[1] for v in agg.values())
    with open(os.path.join(PROF, f"{tag}_launch_list_summary.md"), "w") as f:
        f.write(f"# ncu launch list ({tag})\n\n`ncu --metrics gpu__time_duration.sum --clock-control none` over 2 steps of `{cmd}`\n"
                f"(cold-cache, serialised: compare SHARES with bench.py's live `roofline.share_of_step`, not absolutes)\n\n"
                f"| kernel | launches | total us | share | avg us |\n|---|---:|---:|---:|---:|\n")
        for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write(f"| {k} | {v[0]} | {v[1]:.1f} | {v[1] / tot:.3f} | {v[1] / v[0]:.1f} |\n")
        f.write(f"\ntotal {tot:.0f} us over {sum(v[0] for v in agg.values())} launches\n")
    with open(os.path.join(PROF, f"{tag}_launches.csv"), "w") as f:
        w = csv.writer(f)
        w.writerow(["id", "kernel", "duration_us"])
        for r in per:
            w.writerow([r[0], r[1], f"{r[2]:.3f}"])


def full(tag, name):
    p = os.path.join(OUT, f"{name}_{tag}.ncu-rep")
    if not os.path.exists(p):
        return
    raw = subprocess.run(["ncu", "-i", p, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    n = len(rows) - 2
    with open(os.path.join(PROF, f"{tag}_{name}_full_summary.md"), "w") as f:
        f.write(f"# ncu --set full --clock-control none --import-source on ({tag}, {name}); one column per captured launch\n\n| metric | unit | " +
                " | ".join(f"launch {i}" for i in range(n)) + " |\n|---|---|" + "---:|" * n + "\n")
        f.write("| kernel | | " + " | ".join(short(r[hdr.index("Kernel Name")]) for r in rows[2:]) + " |\n")
        for m in METRICS:
            if m in hdr:
                i = hdr.index(m)
                f.write(f"| {m} | {units[i]} | " + " | ".join(r[i] for r in rows[2:]) + " |\n")
        # warp-stall sampling: the five largest reasons of each launch, as a share of its samples
        pref = "smsp__pcsamp_warps_issue_stalled_"
        st = [(h[len(pref):], i) for i, h in enumerate(hdr) if h.startswith(pref) and not h.endswith("_not_issued")]
        cells = []
        for r in rows[2:]:
            v = sorted(((float(r[i] or 0), n) for n, i in st), reverse=True)
            tot = sum(x for x, _ in v) or 1.0
            cells.append(", ".join(f"{n} {100 * x / tot:.0f}%" for x, n in v[:5]))
        f.write("| top warp-stall reasons (pc sampling) | share of samples | " + " | ".join(cells) + " |\n")


if __name__ == "__main__":
    tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
    os.makedirs(PROF, exist_ok=True)
    launch_list(tag, os.environ.get("NCU_CMD", "python bench.py --quick --windows 8 --steps 2 --warmup 3 --no-cpu-baseline"))
    for n in sys.argv[2:] or ["gemm", "attn", "melln"]:
        full(tag, n)
    print(sorted(os.listdir(PROF)))
