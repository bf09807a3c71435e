"""Fused MMFS sampler on the cfg-3 layer shape: generic kernel vs the specialised kernel (fp32 / 16-bit tap weights) over
rows-per-warp settings; CUDA events, L2 flushed, median of `reps`.  Prints one JSON object (-> profiles/)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
reps = sys.argv[1] if len(sys.argv) > 1 else "20"
rows = []
for masked in ("masked", "all"):
    for mode, rpws in (("generic", (0,)), ("generic_w16", (0,)), ("exact", (0, 1, 2, 4, 8)), ("v2", (0, 1, 2, 4, 8))):
        for rpw in rpws:
            out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "sampler_one.py"), "4", reps, masked, mode, str(rpw)],
                                 capture_output=True, text=True).stdout.strip().splitlines()
            line = out[-1] if out else ""
            us = float(line.split(":")[1].split("us")[0]) if " us" in line else None
            rows.append(dict(masked=masked, mode=mode, rows_per_warp=rpw, us=us, line=line))
            print(line, flush=True)
print(json.dumps(rows))
