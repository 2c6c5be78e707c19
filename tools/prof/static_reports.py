"""Regenerate the static reports under profiles/ from the product library's sources / binary (no GPU needed):
   python tools/prof/static_reports.py
   profiles/r2_ptxas_v.txt       nvcc -Xptxas -v of every translation unit, demangled
   profiles/r2_sass_histogram.txt  SASS opcode histogram per kernel (cuobjdump -sass)"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
LIB = os.path.join(ROOT, "rna_algos_b200", "librna_algos_b200.so")
FLAGS = ["-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-fmad=false", "-Xptxas", "-v", "-c", "-o", "/dev/null"]


def demangle(text):
    return subprocess.run(["c++filt"], input=text, capture_output=True, text=True).stdout


out = ["# nvcc -Xptxas -v of rna_algos_b200/librna_algos_b200.so (sm_100a), demangled: registers, spills, shared memory per kernel and per",
       "# ABI-called chain function (round 2, final state; tools/prof/static_reports.py)"]
for tu in ("rna_abi.cu", "fold_fastnum.cu"):
    r = subprocess.run(["nvcc"] + FLAGS + [os.path.join(ROOT, "rna_algos_b200", "csrc", tu)], capture_output=True, text=True)
    out.append(f"\n## {tu}")
    for line in demangle(r.stderr).splitlines():
        line = line.replace("ptxas info    : ", "")
        if line.startswith("Compiling entry function") or line.startswith("0 bytes gmem"):
            continue
        m = re.match(r"Function properties for (.*)", line)
        out.append(("\n" + m.group(1)) if m else "    " + line.strip())
open(os.path.join(ROOT, "profiles", "r2_ptxas_v.txt"), "w").write("\n".join(out) + "\n")

sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
hist, order, name = {}, [], None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = m.group(1)
        hist[name] = collections.Counter()
        order.append(name)
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_]+)", line)
    if m and name:
        hist[name][m.group(1)] += 1
KEYS = ["FADD", "FMUL", "FFMA", "FMNMX", "FSETP", "FSEL", "MUFU", "SHFL", "LDS", "STS", "LDG", "STG", "LDGSTS", "LDL", "STL", "BAR", "IMAD", "LOP3", "ATOMG",
        "DADD", "DMUL", "DFMA", "UTMALDG", "UBLKCP"]
lines = ["# SASS opcode histogram per kernel of rna_algos_b200/librna_algos_b200.so (cuobjdump -sass, sm_100a), round 2, final state.",
         "# Reference-exact kernels (namespace rna): 0 FFMA (no contraction: every product and sum is a separate IEEE operation).",
         "# FAST kernels (fast_fold_kernel, namespace rna_fastnum): MUFU (ex2.approx / lg2.approx), SHFL (warp-shuffle reductions), FFMA allowed.",
         "# LDGSTS = cp.async (the operand rings of the HBM-resident mode).  LDL/STL: callee-saved register saves at entry / exit of the ABI-called",
         "# chain functions and the few bytes of spill of profiles/r2_ptxas_v.txt; none inside a fold loop.", ""]
names = demangle("\n".join(order)).splitlines()
for mangled, pretty in zip(order, names):
    h = hist[mangled]
    lines.append(pretty)
    lines.append(f"    {sum(h.values())} instructions: " + " ".join(f"{k}={h[k]}" for k in KEYS if h[k]))
open(os.path.join(ROOT, "profiles", "r2_sass_histogram.txt"), "w").write("\n".join(lines) + "\n")
print("written")
