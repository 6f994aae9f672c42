"""Summarise one kernel of an .ncu-rep (made with `ncu --set full`) as `metric [unit] = value` lines.

usage: python scripts/ncu_summary.py report.ncu-rep [units_per_launch]
``units_per_launch`` (e.g. particles x steps) adds the FP64 thread-instructions per unit.  Runs where ncu is installed
(the build container is enough: `ncu -i` needs no GPU)."""
import csv, io, subprocess, sys

KEEP = ("gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit", "launch__shared_mem_per_block", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__average_warps_issue_stalled", "smsp__average_warp_latency_per_inst_issued", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum",
        "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum", "sm__cycles_elapsed.max", "sm__cycles_elapsed.avg")


def main():
    rep = sys.argv[1]
    units = float(sys.argv[2]) if len(sys.argv) > 2 else None
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, unit = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print(f"# kernel: {d.get('Kernel Name', '?')}  grid {d.get('Grid Size', '?')} block {d.get('Block Size', '?')}")
        vals = {}
        for h, u_, v in zip(hdr, unit, r):
            if any(h.startswith(k) for k in KEEP) and v not in ("", "no data"):
                if "per_cycle" in h and "sass_thread_inst" not in h:
                    continue
                print(f"{h} [{u_}] = {v}")
                try:
                    vals[h] = float(v.replace(",", ""))
                except ValueError:
                    pass
        cyc = vals.get("sm__cycles_elapsed.max") or vals.get("sm__cycles_elapsed.avg")
        if units:
            tot = {}
            for op_ in ("dfma", "dmul", "dadd"):
                k = f"smsp__sass_thread_inst_executed_op_{op_}_pred_on.sum"
                if k in vals:
                    tot[op_] = vals[k]
                elif cyc and f"{k}.per_cycle_elapsed" in vals:
                    tot[op_] = vals[f"{k}.per_cycle_elapsed"] * cyc
            if tot:
                per = {k: v / units for k, v in tot.items()}
                flop = 2 * per.get("dfma", 0) + per.get("dmul", 0) + per.get("dadd", 0)
                print("# FP64 thread-instructions per unit (%g units): " % units + ", ".join(f"{k} {v:.2f}" for k, v in per.items())
                      + f"  => {sum(per.values()):.1f} instructions, {flop:.1f} flop")


if __name__ == "__main__":
    main()
