"""Turns ncu reports / launch lists brought back in gpurun_out/ into the small text summaries committed here.

  python profiles/summarize.py rep  gpurun_out/prof_x.ncu-rep  > profiles/r1_x.txt     (ncu --set full capture)
  python profiles/summarize.py list gpurun_out/launches.csv    > profiles/r1_launches.txt (gpu__time_duration pass)
"""
import csv
import io
import subprocess
import sys
from collections import OrderedDict

KEYS = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__occupancy_limit_shared_mem",
        "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__lsu_writeback_active_mem_lgds.sum.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active"]


def rep(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        print("# ncu --set full --clock-control none; source:", path)
        for k in KEYS:
            if k in d:
                print(f"{k} = {d[k]} {u.get(k, '')}".rstrip())
        print()


def launches(path):
    text = open(path).read()
    start = text.index('"ID"')
    rows = list(csv.DictReader(io.StringIO(text[start:])))
    agg = OrderedDict()
    for r in rows:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = r["Kernel Name"].split("(")[0]
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        ns = v * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += ns
    tot = sum(a[1] for a in agg.values())
    print("# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES); source:", path)
    print(f"{'kernel':60s} {'launches':>8s} {'total_us':>12s} {'share':>7s}")
    for name, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{name[:60]:60s} {n:8d} {ns / 1e3:12.1f} {100 * ns / tot:6.2f}%")


def _rows(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        yield dict(zip(hdr, r)), dict(zip(hdr, units))


def _num(x):
    return float(str(x).replace(",", ""))


def _bytes(v, unit):
    return _num(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def facts(burst_rep, colsum_rep, streams_profiled, n_iq):
    """roofline_facts.json for bench.py: numbers that only an ncu capture can give (fp64 pipe utilisation and DRAM bytes per launch of the
    burst kernels at the profiled size, DRAM bytes of the column-sum kernel against its algorithmic bytes)."""
    import json
    stage_of = {"fine_core8_kernel": "fine_peak", "sch_corr_kernel": "sch"}
    pipe, dram, l1pipe, issue, seen_tone = {}, {}, {}, {}, 0
    for d, u in _rows(burst_rep):
        name = d["Kernel Name"].split("(")[0].replace("void ", "").split("<")[0]
        if name == "tone8_kernel":
            seen_tone += 1
            stage = "fine_tone" if seen_tone == 1 else "post"
        elif name in stage_of:
            stage = stage_of[name]
        else:
            continue
        pipe[stage] = _num(d["sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"])
        l1pipe[stage] = _num(d["l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"])     # shared-memory + global wavefronts on the L1 data pipe
        issue[stage] = _num(d["smsp__issue_active.avg.pct_of_peak_sustained_active"])
        dram[stage] = int(_bytes(d["dram__bytes_read.sum"], u["dram__bytes_read.sum"]) + _bytes(d["dram__bytes_write.sum"], u["dram__bytes_write.sum"]))
    col = {}
    for d, u in _rows(colsum_rep):
        if d["Kernel Name"].startswith("colsum_u8_kernel"):
            rd = _bytes(d["dram__bytes_read.sum"], u["dram__bytes_read.sum"]); wr = _bytes(d["dram__bytes_write.sum"], u["dram__bytes_write.sum"])
            grid = d.get("launch__grid_size", "")
            col = {"dram_bytes": int(rd + wr)}
            break
    out = {"source": f"ncu --set full --clock-control none, {os.path.basename(burst_rep)} ({streams_profiled} streams x {n_iq} IQ in one synchronous call, one launch per kernel)",
           "streams_profiled": streams_profiled, "fp64_pipe_active_pct": pipe, "l1_data_pipe_pct": l1pipe, "issue_active_pct": issue, "dram_bytes_per_launch": dram,
           "colsum_source": f"ncu --set full --clock-control none, {os.path.basename(colsum_rep)} (32-stream launch of the bench command: 1,386,666,688 algorithmic bytes)",
           "colsum_dram_ratio": (col.get("dram_bytes", 0) / 1386666688.0) if col else None}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    import os
    if sys.argv[1] == "facts":
        facts(sys.argv[2], sys.argv[3], int(sys.argv[4]), int(sys.argv[5]))
    else:
        {"rep": rep, "list": launches}[sys.argv[1]](sys.argv[2])
