"""Condenses an ncu report (.ncu-rep, `--set full`) into the small CSV kept under profiles/:
selected raw metrics of the first kernel in the report plus the executed-instruction mix and the stall-reason
shares from the source page.  Usage: python tools/ncu_summary.py report.ncu-rep out.csv [units]
(`units` = warp-level work items per launch, e.g. rays/32*surfaces, for the per-unit instruction counts)."""
import collections
import csv
import re
import subprocess
import sys
import pathlib

KEEP = ("gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum",
        "lts__t_sectors_srcunit_tex_op_red.sum", "lts__t_sectors_op_red.sum", "lts__t_sectors_op_atom.sum")


def main():
    rep, out = sys.argv[1], sys.argv[2]
    units = float(sys.argv[3]) if len(sys.argv) > 3 else None
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(raw.splitlines()))
    names, unit_row, vals = rr[0], rr[1], rr[2]
    rows = [("metric", "unit", "value"), ("Kernel Name", "", vals[names.index("Kernel Name")])]
    for n, u, v in zip(names, unit_row, vals):
        if n in KEEP:
            rows.append((n, u, v))
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                         capture_output=True, text=True).stdout
    sr = list(csv.reader(src.splitlines()))
    if len(sr) > 2:
        h = sr[1]
        ia, ie, isamp = h.index("Source"), h.index("Instructions Executed"), h.index("# Samples")
        byop, tot, nsamp = collections.Counter(), 0, 0
        stalls = collections.Counter()
        scol = [(i, n) for i, n in enumerate(h) if n.startswith("stall_") and "Not" not in n]
        for r in sr[2:]:
            if len(r) <= ie:
                continue
            n = int(r[ie] or 0)
            m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)", r[ia])
            byop[m.group(2) if m else "?"] += n
            tot += n
            nsamp += int(r[isamp] or 0)
            for i, nm in scol:
                stalls[nm] += int(r[i] or 0)
        rows.append(("sass_warp_instructions_executed", "inst", str(tot)))
        if units:
            rows.append(("sass_warp_instructions_per_unit", "inst", f"{tot/units:.1f}"))
        for op, n in byop.most_common(16):
            rows.append((f"sass_op.{op}", "share" if not units else "inst/unit", f"{(n/units if units else n/tot):.3f}"))
        for nm, v in stalls.most_common():
            if v:
                rows.append((f"pc_sampling.{nm}", "% of samples", f"{100*v/max(1, nsamp):.1f}"))
    # digest of the engine sources this summary was made with: bench.py reports whether a capture is still current
    sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
    from optrace_b200 import build
    rows.append(("engine_source_digest", "", build.source_digest()))
    with open(out, "w", newline="") as f:
        csv.writer(f).writerows(rows)
    print(f"wrote {out}: {len(rows)} rows")


if __name__ == "__main__":
    main()
