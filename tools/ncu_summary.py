"""Short summary of an .ncu-rep: python tools/ncu_summary.py <rep> [--stalls N]  (reads with `ncu -i`, no GPU needed)."""
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "sm__cycles_elapsed.max", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
        "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__occupancy_limit_shared_mem", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum", "lts__t_bytes.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio"]


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[0]
    units = rows[1]
    for r in rows[2:]:
        yield dict(zip(hdr, r)), dict(zip(hdr, units))


def main():
    rep = sys.argv[1]
    for r, u in raw(rep):
        print("==", r.get("Kernel Name", "?")[:100], "grid", r.get("Grid Size"), "block", r.get("Block Size"))
        for k in KEYS:
            if k in r:
                print(f"  {k:80s} {r[k]:>16s} {u.get(k, '')}")
        stalls = sorted(((float(v.replace(',', '')), k) for k, v in r.items()
                         if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("per_issue_active.ratio") and v not in ("", "n/a")),
                        reverse=True)
        for v, k in stalls[:8]:
            print(f"  stall {k[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]:40s} {v:.2f}")
    if "--stalls" in sys.argv:
        n = int(sys.argv[sys.argv.index("--stalls") + 1])
        out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(out.splitlines()))
        # find the header row
        hi = next(i for i, r in enumerate(rows) if "Source" in r and any("Sampl" in c for c in r))
        hdr = rows[hi]
        si = hdr.index("Source")
        ci = next(i for i, c in enumerate(hdr) if c.startswith("# Samples") or c == "Warp Stall Sampling (All Samples)")
        def num(v):
            try:
                return float(v or 0)
            except ValueError:      # a second kernel's header row inside a multi-kernel report
                return None
        body = [r for r in rows[hi + 1:] if len(r) > ci and num(r[ci]) is not None]
        tot = sum(float(r[ci] or 0) for r in body)
        top = sorted(body, key=lambda r: -float(r[ci] or 0))[:n]
        print(f"-- top {n} SASS lines by stall samples (total {tot:.0f}); column '{hdr[ci]}'")
        for r in top:
            print(f"  {float(r[ci] or 0) / max(tot, 1) * 100:5.1f}%  {r[si][:150]}")


if __name__ == "__main__":
    main()
