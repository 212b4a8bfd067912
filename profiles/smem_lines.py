"""Per-CUDA-line shared-memory wavefronts (and L1 tag requests of global accesses) of one kernel from an ncu report taken with
-lineinfo and --import-source on.  The per-burst kernels keep the L1/shared data pipe busier than the FP64 pipe
(l1tex__data_pipe_lsu_wavefronts 59-68 % vs 34-53 %), so this is the view that says where the next cycles are.
   python profiles/smem_lines.py gpurun_out/prof.ncu-rep kernel_name [top_n] [launch_no]"""
import csv, io, subprocess, sys
from collections import defaultdict
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
only = int(sys.argv[4]) if len(sys.argv) > 4 else 0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", kern], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
launch, i = 0, 0
while i < len(rows):
    r = rows[i]
    if r and r[0] == "Line No" and len(r) > 5:
        launch += 1
        hdr = r
        ln = hdr.index("Line No")
        cw, ci, cg, ce = hdr.index("L1 Wavefronts Shared"), hdr.index("L1 Wavefronts Shared Ideal"), hdr.index("L1 Tag Requests Global"), hdr.index("Instructions Executed")
        agg = defaultdict(lambda: [0.0, 0.0, 0.0, 0.0, ""])
        cur = None
        i += 1
        while i < len(rows) and rows[i] and rows[i][0] not in ("Line No", "File Path", "Function Name", "File Name"):
            x = rows[i]
            if x[ln].strip():
                cur = int(x[ln]); agg[cur][4] = x[1].strip()[:120]
            elif cur is not None:
                for k, c in enumerate((cw, ci, cg, ce)):
                    try: agg[cur][k] += float(x[c] or 0)
                    except (ValueError, TypeError): pass
            i += 1
        if only and launch != only: continue
        tw = sum(v[0] for v in agg.values()) or 1
        tg = sum(v[2] for v in agg.values()) or 1
        print(f"--- {kern} launch {launch}: {tw:.0f} shared wavefronts ({sum(v[1] for v in agg.values()):.0f} ideal), {tg:.0f} global tag requests")
        for line, (w, idl, g, n, text) in sorted(agg.items(), key=lambda kv: -(kv[1][0] + kv[1][2]))[:top]:
            print(f"{100 * w / tw:5.1f}% smem wf ({w:9.0f}, ideal {idl:9.0f})  {100 * g / tg:5.1f}% glob ({g:8.0f})  L{line}: {text}")
        continue
    i += 1
