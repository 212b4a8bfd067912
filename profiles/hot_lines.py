"""Per-CUDA-line stall samples of one kernel from an ncu report (needs -lineinfo and --import-source on).
   python profiles/hot_lines.py gpurun_out/prof.ncu-rep kernel_name [top_n]"""
import csv, io, subprocess, sys
from collections import defaultdict
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
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
                agg[cur][0] += float(x[samp] or 0); agg[cur][1] += float(x[ins] or 0)
            except (ValueError, TypeError):
                pass
            i += 1
        tot = sum(v[0] for v in agg.values()) or 1
        print(f"--- {kern} launch {launch}: {tot:.0f} stall samples")
        for line, (s, n, text) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
            print(f"{100 * s / tot:5.1f}%  inst {n:10.0f}  L{line}: {text}")
        continue
    i += 1
