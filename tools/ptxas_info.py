"""registers / spills per kernel from `nvcc -Xptxas -v` (no GPU needed):  python tools/ptxas_info.py spmv.cu [regex]"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = sys.argv[1]
pat = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "--expt-relaxed-constexpr",
       "--extended-lambda", "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(ROOT, "sparsebench_b200", "csrc"),
       "-Xptxas", "-v", "-c", os.path.join(ROOT, "sparsebench_b200", "csrc", src), "-o", "/dev/null"]
out = subprocess.run(cmd, capture_output=True, text=True).stderr
name = None
for line in out.splitlines():
    m = re.search(r"Compiling entry function '(\S+)'", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(.*", "", name)
        info = []
        continue
    if name and ("spill" in line or "registers" in line):
        info.append(line.replace("ptxas info    :", "").strip())
        if "registers" in line:
            if not pat or pat.search(name):
                print(name, "|", " | ".join(info))
            name = None
