"""Turn gpurun_out/ captures into the committed summaries under profiles/.
  python tools/summarize_profiles.py launches <launches.csv> <out.txt> [first_step=4 n_steps=2]
  python tools/summarize_profiles.py ncu <report.ncu-rep> <out.txt> "<title>"
"""
import csv, io, subprocess, sys, collections

KEEP = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_red.sum",
        "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_red.sum",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "lts__t_requests_srcunit_tex_op_red.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
        "sm__inst_executed.sum.per_cycle_active", "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio")


def ncu(rep, out, title):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    lines = [f"# {title}", f"# kernel: {vals[hdr.index('Kernel Name')][:140]}"]
    for h, u, v in sorted(zip(hdr, units, vals)):
        stall = "issue_stalled" in h and "per_issue_active" in h and "not_issued" not in h
        try:
            if h in KEEP or (stall and float(v) > 0.25):
                lines.append(f"{h:95s} {v} {u}")
        except ValueError:
            pass
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


def launches(path, out, first_step=4, n_steps=2):
    rows = [r for r in csv.reader(open(path)) if len(r) > 14 and r[0].isdigit()]
    names = [r[4] for r in rows]; ns = [float(r[14]) for r in rows]
    starts = [i for i, n in enumerate(names) if "intersect_kernel" in n]
    # a training step = intersect ... last adam_kernel before the next intersect
    a, b = starts[first_step - 1], starts[first_step - 1 + n_steps]
    agg = collections.OrderedDict()
    for n, t in zip(names[a:b], ns[a:b]):
        k = n.split("(")[0][:74]
        c = agg.setdefault(k, [0, 0.0]); c[0] += 1; c[1] += t
    tot = sum(v[1] for v in agg.values())
    lines = [f"# ncu --metrics gpu__time_duration.sum --clock-control none, command: python bench.py --steps 2 --warmup 3 --pretrain 0 --no-render",
             f"# the {n_steps} TIMED steps (launches {a}..{b - 1} of {len(rows)}); per-launch times are cold-cache and serialised: compare SHARES",
             f"# total {tot / n_steps / 1e6:.3f} ms/step under ncu"]
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append(f"{k:76s}{c:3d} launches {t / n_steps:12.1f} ns/step {100 * t / tot:5.1f}%")
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines[:32]))


if __name__ == "__main__":
    if sys.argv[1] == "ncu":
        ncu(sys.argv[2], sys.argv[3], sys.argv[4])
    else:
        launches(sys.argv[2], sys.argv[3], *[int(x) for x in sys.argv[4:]])
