"""Digest of an ncu report: headline metrics per launch + stall samples aggregated by SASS opcode.

    python tools/ncu_digest.py gpurun_out/prof.ncu-rep [launch_index]
"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
which = int(sys.argv[2]) if len(sys.argv) > 2 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
want = ["Kernel Name", "gpu__time_duration.sum", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "smsp__inst_executed.sum", "lts__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__cycles_active.avg"]
want += [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio") and "not_issued" not in h]
for w in want:
    if w in hdr:
        i = hdr.index(w)
        vals = [r[i] for r in data]
        try:
            if all(abs(float(v)) < 0.05 for v in vals) and w.startswith("smsp__average"):
                continue
        except ValueError:
            pass
        print(f"{w.replace('smsp__average_warps_issue_stalled_','stall:').replace('_per_issue_active.ratio','')} [{units[i]}]", vals)
n = len(data)
for li in range(n):
    if which is not None and li != which:
        continue
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", str(li), "--launch-count", "1"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    h = rows[1]
    ix = {k: i for i, k in enumerate(h)}
    body = rows[2:]
    # rows come twice in some ncu versions: dedupe on (address)
    seen, uniq = set(), []
    for r in body:
        if r[ix["Address"]] in seen:
            continue
        seen.add(r[ix["Address"]]); uniq.append(r)
    def f(r, k):
        try:
            return float(r[ix[k]])
        except (ValueError, KeyError, IndexError):
            return 0.0
    keys = ["# Samples", "stall_long_sb", "stall_wait", "stall_short_sb", "stall_selected", "stall_mio", "stall_math", "stall_barrier", "stall_membar", "stall_lg", "stall_not_selected", "stall_no_inst", "stall_dispatch", "stall_branch_resolving"]
    tot = {k: sum(f(r, k) for r in uniq) for k in keys}
    print(f"\n== launch {li}: {rows[0][1][:80]}  samples {tot['# Samples']:.0f}")
    print("   totals:", {k.replace('stall_', ''): int(v) for k, v in tot.items() if v > 0})
    byop = collections.defaultdict(collections.Counter)
    for r in uniq:
        toks = [t for t in r[ix["Source"]].split() if not t.startswith("@")]
        op = toks[0].split(".")[0] if toks else "?"
        for k in keys:
            byop[op][k] += f(r, k)
        byop[op]["n"] += 1
        byop[op]["exec"] += f(r, "Instructions Executed")
    print("   %-10s %5s %9s %8s %7s %7s %7s %7s %7s %7s" % ("op", "n", "exec/1e6", "samples", "longsb", "wait", "shortsb", "sel", "mio", "membar"))
    for op, c in sorted(byop.items(), key=lambda kv: -kv[1]["# Samples"])[:22]:
        print("   %-10s %5d %9.1f %8d %7d %7d %7d %7d %7d %7d" % (op, c["n"], c["exec"] / 1e6, c["# Samples"], c["stall_long_sb"], c["stall_wait"],
                                                               c["stall_short_sb"], c["stall_selected"], c["stall_mio"], c["stall_membar"]))
    print("   top instructions:")
    for r in sorted(uniq, key=lambda r: -f(r, "# Samples"))[:12]:
        print("     %6d  %s" % (f(r, "# Samples"), r[ix["Source"]][:100]))
