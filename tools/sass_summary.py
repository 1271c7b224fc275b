"""Per-kernel SASS opcode summary of the shipped library (what proves the FP64 tensor path, TMA and the
absence of spills): `python tools/sass_summary.py > profiles/r2_sass_opcodes.md` (cuobjdump, no GPU needed)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "anml_b200", "lib", "libanml_b200.so")
KEYS = ["DMMA", "UTCIMMA", "LDTM", "UTCBAR", "DFMA", "DMUL", "DADD", "MUFU", "UTMALDG", "UBLKCP", "SYNCS", "USETMAXREG", "LDS", "STS", "LDG", "STG", "LDL", "STL",
        "ST.E.STRONG.SYS", "LD.E.STRONG.SYS", "BAR", "SHFL"]

out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
kernels = collections.OrderedDict()
name = None
for ln in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", ln)
    if m:
        name = m.group(1)
        kernels[name] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
    if m and name:
        op = m.group(2)
        kernels[name]["_total"] += 1
        for k in KEYS:
            if op == k or op.startswith(k + "."):
                kernels[name][k] += 1


def demangle(n):
    return subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip().replace("anml::", "")


WANT = ["fused_hess_i8_kernel<0>", "fused_hess_i8_kernel<1>", "fused_hess_kernel<4, 2, false, 0, true, 32>", "fused_hess_kernel<4, 2, true, 3, true, 32>", "fused_hess_kernel<4, 2, true, 1, true, 32>",
        "fused_hess_kernel<1, 4, false, 0, true, 64>", "fused_eval_kernel<4, false, 0>", "wide_hessian_kernel", "spline_banded_bulk_kernel<3, true, false>", "spline_banded_eval_kernel<3, true, false>",
        "spline_design_kernel<3>", "finish_fused_kernel", "banded_finish_kernel", "comm_allreduce_kernel"]
print("# SASS opcode counts per kernel (static instructions; `cuobjdump -sass` of the shipped library)\n")
print("DMMA = `mma.sync.m8n8k4.f64` (the FP64 tensor-core instruction of sm_100a); UTCIMMA = `tcgen05.mma.kind::i8` (INT8 tensor cores, accumulators in TMEM), LDTM = `tcgen05.ld`, UTCBAR = `tcgen05.commit`; UTMALDG = `cp.async.bulk.tensor` (TMA); "
      "UBLKCP = `cp.async.bulk`; SYNCS = mbarrier operations; USETMAXREG = `setmaxnreg`; LDL / STL = local-memory (spill) traffic.\n")
hdr = ["kernel", "instr"] + KEYS
print("| " + " | ".join(hdr) + " |")
print("|" + "---|" * len(hdr))
rows = {demangle(n): c for n, c in kernels.items()}
for w in WANT:
    for dn, c in rows.items():
        if w in dn:
            short = re.sub(r"\(.*", "", dn).replace("void ", "")
            print("| `" + short + "` | " + str(c["_total"]) + " | " + " | ".join(str(c[k]) if c[k] else "" for k in KEYS) + " |")
            break
print(f"\n{len(kernels)} kernels in the library; totals over all of them: " + ", ".join(f"{k} {sum(c[k] for c in kernels.values())}" for k in KEYS if sum(c[k] for c in kernels.values())))
