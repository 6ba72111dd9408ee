"""Summarise gpurun_out/launches.csv (ncu --metrics gpu__time_duration.sum) and a full .ncu-rep into
profiles/<tag>_launches.txt and profiles/<tag>_kernels.txt.   usage: summarize_ncu.py <tag> [rep ...]"""
import collections, csv, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
reps = sys.argv[2:]
out = os.path.join(ROOT, "profiles")
os.makedirs(out, exist_ok=True)
lcsv = os.path.join(ROOT, "gpurun_out", "launches.csv")
if os.path.exists(lcsv):
    lines = [l for l in open(lcsv) if l.startswith('"')]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        k = row["Kernel Name"][:90]
        v = float(row["Metric Value"].replace(",", ""))
        a = agg.setdefault(k, [0, 0.0, 0.0]); a[0] += 1; a[1] += v; a[2] = max(a[2], v)
    tot = sum(v[1] for v in agg.values())
    with open(os.path.join(out, f"{tag}_launches.txt"), "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)\n")
        f.write(f"# command: python bench.py --steps 2 --warmup 3 --skip-e2e   total {tot/1e3:.1f} us over {sum(v[0] for v in agg.values())} launches\n")
        f.write(f"{'total_us':>10} {'n':>5} {'avg_us':>9} {'max_us':>9} {'share':>6}  kernel\n")
        for k, (n, v, mx) in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write(f"{v/1e3:10.1f} {n:5d} {v/n/1e3:9.1f} {mx/1e3:9.1f} {100*v/tot:5.1f}%  {k}\n")
first = True
for rep in reps:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(raw.splitlines()))
    hdr, units = r[0], r[1]
    want = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
            "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum", "l1tex__t_bytes.sum",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__shared_mem_config_size", "sm__cycles_elapsed.max",
            "smsp__average_warp_latency_issue_stalled_long_scoreboard.pct", "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
            "smsp__warp_issue_stalled_barrier_per_warp_active.pct", "smsp__warp_issue_stalled_membar_per_warp_active.pct"]
    with open(os.path.join(out, f"{tag}_kernels.txt"), "w" if first else "a") as f:
        f.write(f"# ncu --set full --clock-control none --import-source on ; report {os.path.basename(rep)}\n")
        for row in r[2:]:
            f.write("---\n")
            for w in want:
                ii = [i for i, h in enumerate(hdr) if h == w]
                if ii:
                    f.write(f"{w:72s} {row[ii[0]][:70]} [{units[ii[0]]}]\n")
    first = False
print("written", os.listdir(out))
