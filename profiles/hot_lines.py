"""Per-CUDA-line stall samples and executed warp instructions of one kernel from an ncu report (needs -lineinfo and
--import-source on).  Sorted by stall samples, or by instructions with a 4th argument "inst" - the per-burst kernels are
issue-slot bound, so the instruction view is the one that found the expensive block reductions and the rolled FIR.
Instructions of inlined device functions are listed under the callee line AND the call-site line, so the totals run up to 2x
smsp__inst_executed; compare shares.
   python profiles/hot_lines.py gpurun_out/prof.ncu-rep kernel_name [top_n] [inst]"""
import csv, io, subprocess, sys
from collections import defaultdict
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
by_inst = len(sys.argv) > 4 and sys.argv[4] == "inst"
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", kern], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
launch = 0
i = 0
while i < len(rows):
    r = rows[i]
    if r and r[0] == "Line No":
        launch += 1
        hdr = r
        ln, src, samp = hdr.index("Line No"), 1, hdr.index("Warp Stall Sampling (All Samples)")
        ins = hdr.index("Instructions Executed")
        agg = defaultdict(lambda: [0.0, 0.0, ""])
        cur = None
        i += 1
        while i < len(rows) and rows[i] and rows[i][0] not in ("Line No", "File Path", "Function Name"):
            x = rows[i]
            if x[ln].strip():
                cur = int(x[ln]); agg[cur][2] = x[src].strip()[:130]
            try:
                if not x[ln].strip():      # SASS rows only: the CUDA source row repeats the sum of its SASS rows
                    agg[cur][0] += float(x[samp] or 0); agg[cur][1] += float(x[ins] or 0)
            except (ValueError, TypeError):
                pass
            i += 1
        tot = sum(v[0] for v in agg.values()) or 1
        tot_i = sum(v[1] for v in agg.values()) or 1
        print(f"--- {kern} launch {launch}: {tot:.0f} stall samples, {tot_i:.0f} warp instructions")
        for line, (s, n, text) in sorted(agg.items(), key=lambda kv: -kv[1][1 if by_inst else 0])[:top]:
            print(f"{100 * s / tot:5.1f}% stalls  {100 * n / tot_i:5.1f}% inst ({n:10.0f})  L{line}: {text}")
        continue
    i += 1
