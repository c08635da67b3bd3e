#!/usr/bin/env python
"""Run ON the GPU box (through gpurun): the two ncu passes the profiling recipe asks for, on one
CG iteration of the 256^3 solve (every kernel of the step appears in it).

  pass 1  ncu --metrics gpu__time_duration.sum   every launch of two solves -> launches.csv
  pass 2  ncu --set full --import-source on      the launches of ONE iteration (the second
          iteration of the first solve; found in pass 1's list) -> iter.ncu-rep

    python tools/ncu_capture.py <out-prefix> [n] [relax] [krylov]

Summaries for profiles/ are written here (CPU box) by tools/summarize_ncu.py from the files this
leaves under gpurun_out/.  Numbers printed under ncu are never bench values."""
import csv
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = sys.argv[1]
n = sys.argv[2] if len(sys.argv) > 2 else "256"
relax = sys.argv[3] if len(sys.argv) > 3 else "damped_jacobi"
krylov = sys.argv[4] if len(sys.argv) > 4 else "cg"
target = [sys.executable, os.path.join(ROOT, "tools", "profile_target.py"), n, "1", relax, krylov]
marker = "CgUpdateF" if krylov == "cg" else "BicgUpdateRF"     # last kernel of an iteration

lst = out + "_launches.csv"
subprocess.run(["ncu", "--metrics", "gpu__time_duration.sum", "--clock-control", "none", "-c", "4000",
                "--csv", "--log-file", lst] + target, check=False, stdout=subprocess.DEVNULL)
rows = list(csv.reader(open(lst)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr = rows[hi]
data = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
kn = hdr.index("Kernel Name")
ends = [i for i, r in enumerate(data) if marker in r[kn]]
if len(ends) < 3:
    print("could not find two iterations in the launch list (%d markers)" % len(ends))
    sys.exit(1)
skip, count = ends[0] + 1, ends[1] - ends[0]
print("launches:", len(data), "iteration = launches", skip, "..", skip + count - 1)
# the report itself stays on the box (hundreds of MB with --set full + sources; gpurun_out/ is
# capped at 64 MiB): only the raw metric table travels
rep = os.path.join("/tmp", os.path.basename(out) + "_iter")
subprocess.run(["ncu", "--set", "full", "--clock-control", "none", "--import-source", "on",
                "-s", str(skip), "-c", str(count), "-f", "-o", rep] + target, check=False,
               stdout=subprocess.DEVNULL)
raw = subprocess.run(["ncu", "-i", rep + ".ncu-rep", "--page", "raw", "--csv"], stdout=subprocess.PIPE,
                     stderr=subprocess.DEVNULL, text=True).stdout
open(out + "_iter_raw.csv", "w").write(raw)
print("raw csv bytes:", len(raw))
