"""Turn ncu outputs under gpurun_out/ into the small text summaries committed under profiles/.
  python tools/summarize_ncu.py launches <csv> <out.txt> [title]
  python tools/summarize_ncu.py full <ncu-rep> <out.txt> [title]"""
import collections, csv, io, subprocess, sys

mode, src, dst = sys.argv[1:4]
title = sys.argv[4] if len(sys.argv) > 4 else ""
out = []
if mode == "launches":
    lines = [l for l in open(src) if l.startswith('"')]
    rows = list(csv.DictReader(lines))
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows:
        try:
            v = float(r["Metric Value"].replace(",", ""))
        except ValueError:
            continue
        name = r["Kernel Name"].split("(")[0].replace("void ", "").replace("unnamed>::", "").replace("iir::<", "")
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    out.append(f"{title}\nncu launch list (gpu__time_duration.sum, --clock-control none; cold-cache, serialised: compare SHARES): "
               f"{sum(v[0] for v in agg.values())} launches, {tot / 1e6:.2f} ms")
    for k, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"{ns / 1e6:9.3f} ms {100 * ns / tot:5.1f}%  n={n:6d}  avg {ns / n / 1e3:8.2f} us  {k[:100]}")
else:
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    h, u = rows[0], rows[1]
    want = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
            "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
            "sm__cycles_elapsed.avg", "sm__cycles_active.avg", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor",
            "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
            "l1tex__t_requests_pipe_lsu_mem_local_op_ld.sum"]
    out.append(f"{title}\nncu --set full --clock-control none: {len(rows) - 2} launch(es) from {src}")
    for w in want:
        if w in h:
            i = h.index(w)
            out.append(f"{w} [{u[i]}]: {[r[i][:70] for r in rows[2:]]}")
open(dst, "w").write("\n".join(out) + "\n")
print("\n".join(out[:40]))
