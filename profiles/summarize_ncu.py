"""Turn an .ncu-rep (brought back in gpurun_out/) into the text summary committed under profiles/.

    python profiles/summarize_ncu.py gpurun_out/<name>.ncu-rep > profiles/<name>.txt
"""
import collections
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "sm__cycles_active.avg", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "smsp__cycles_elapsed.avg.per_second", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct",
        "smsp__warp_issue_stalled_wait_per_warp_active.pct", "smsp__warp_issue_stalled_not_selected_per_warp_active.pct",
        "smsp__warp_issue_stalled_barrier_per_warp_active.pct", "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct"]


def ncu(rep, page, extra=()):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv", *extra], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main(rep):
    rows = ncu(rep, "raw")
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print("=" * 100)
        print("kernel:", d.get("Kernel Name"), "| grid", d.get("Grid Size"), "block", d.get("Block Size"))
        for k in KEYS:
            if k in d and d[k] != "":
                print(f"  {k:78s} {d[k]:>18s} {units[hdr.index(k)]}")
    # executed-instruction mix of each profiled kernel (SASS page)
    src = ncu(rep, "source", ("--print-source", "sass"))
    hdr, ops, name = None, None, None

    def flush():
        if ops:
            tot = sum(ops.values())
            print("-" * 100)
            print("executed warp-instructions by opcode:", name, "| total", tot)
            for op, n in ops.most_common(16):
                print(f"  {op:10s} {n:12d} {100.0 * n / tot:6.2f}%")
    for r in src:
        if r and r[0] == "Kernel Name":
            flush()
            name, ops, hdr = r[1], collections.Counter(), None
            continue
        if r and r[0] == "Address":
            hdr = r
            continue
        if hdr is None or len(r) < 6:
            continue
        d = dict(zip(hdr, r))
        tok = d["Source"].split()
        if not tok:
            continue
        op = tok[1] if tok[0].startswith("@") and len(tok) > 1 else tok[0]
        ops[op.split(".")[0]] += int(d.get("Instructions Executed") or 0)
    flush()


if __name__ == "__main__":
    main(sys.argv[1])
