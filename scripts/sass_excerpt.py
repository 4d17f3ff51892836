"""profiles/r02_sass_excerpt_shipped_so.txt: per-kernel counts of the SASS instructions that identify the
machine features in use (DMMA, UBLKCP = TMA bulk copy, SYNCS = mbarrier, LDGSTS = cp.async, BAR.ARV) and a
few excerpts of k_resident, from `cuobjdump -sass gprf_b200/libgprf_b200.so`."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
txt = subprocess.run(["cuobjdump", "-sass", os.path.join(ROOT, "gprf_b200", "libgprf_b200.so")], capture_output=True, text=True).stdout
funcs = re.split(r'\n\s*Function : ', txt)
out = ["SASS of the shipped gprf_b200/libgprf_b200.so (cuobjdump -sass), sm_100a cubins; counts per kernel of the",
       "instructions that identify the machine features the design relies on:",
       "  DMMA.8x8x4  fp64 tensor-core MMA        UBLKCP  bulk (TMA) global->shared copy        SYNCS  mbarrier ops",
       "  LDGSTS      cp.async 16-byte copies     BAR.ARV named-barrier arrive (producer/consumer)   LDS/STS shared memory", ""]
keys = ("DMMA", "UBLKCP", "SYNCS", "LDGSTS", "BAR.ARV", "LDS", "STS")
for f in funcs[1:]:
    name = f.split('\n', 1)[0].strip()
    c = collections.Counter()
    for m in re.finditer(r'^\s+/\*[0-9a-f]{4,5}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', f, re.M):
        op = m.group(1)
        for k in keys:
            if op.startswith(k):
                c[k] += 1
        c["total"] += 1
    if c["DMMA"] == 0 and c["UBLKCP"] == 0:
        continue
    short = name if len(name) <= 70 else name[:67] + "..."
    out.append("%-70s total %6d  DMMA %5d  UBLKCP %3d  SYNCS %3d  LDGSTS %4d  BAR.ARV %2d  LDS %5d  STS %5d"
               % (short, c["total"], c["DMMA"], c["UBLKCP"], c["SYNCS"], c["LDGSTS"], c["BAR.ARV"], c["LDS"], c["STS"]))
out.append("")
for f in funcs[1:]:
    name = f.split('\n', 1)[0].strip()
    if 'k_residentILi0ELi0' not in name:
        continue
    lines = [l for l in f.split('\n') if re.match(r'^\s+/\*[0-9a-f]{4,5}\*/', l)]

    def clean(l):
        return re.sub(r'\s*/\*[0-9a-fx]+\*/\s*$', '', l).rstrip()
    i = [k for k, l in enumerate(lines) if 'UBLKCP' in l][0]
    out.append("## k_resident<euclidean, se>: a TMA bulk copy into shared memory with its mbarrier (tma_issue, resident.cuh)")
    out += [clean(l) for l in lines[max(0, i - 8):i + 3]] + [""]
    j = [k for k, l in enumerate(lines) if 'SYNCS.PHASECHK' in l][0]
    out.append("## ... the matching mbarrier wait (tma_wait)")
    out += [clean(l) for l in lines[max(0, j - 2):j + 4]] + [""]
    best = max(range(0, len(lines) - 40, 5), key=lambda s: sum('DMMA' in l for l in lines[s:s + 40]))
    out.append("## ... a block-product loop body (mk_loop: shared-memory fragments by window address, DMMA pairs)")
    out += [clean(l) for l in lines[best:best + 40]] + [""]
    a = [k for k, l in enumerate(lines) if 'BAR.ARV' in l][-1]
    out.append("## ... the producer side of the Cholesky look-ahead (named barrier 1: BAR.ARV by warp 0, BAR.SYNC by the workers)")
    out += [clean(l) for l in lines[max(0, a - 3):a + 3]] + [""]
    d = [k for k, l in enumerate(lines) if 'CCTL.E.RML2' in l]
    if d:
        out.append("## ... dead scratch dropped from the L2 (discard.global.L2 = CCTL.E.RML2, %d of them)" % len(d))
        out += [clean(l) for l in lines[max(0, d[0] - 2):d[0] + 3]]
    break
open(os.path.join(ROOT, "profiles", "r02_sass_excerpt_shipped_so.txt"), "w").write("\n".join(out) + "\n")
print("\n".join(out[-22:]))
