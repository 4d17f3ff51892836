"""Per-phase summary of an ncu SASS source page of k_resident: instructions executed, DMMAs, stall
samples by reason, shared-memory wavefronts - attributed to the ph_* function (resident.cuh) each
SASS instruction was inlined from.

    python scripts/ncu_phases.py gpurun_out/<tag>.source.csv gprf_b200/csrc/build/gprf_resident_00.o
"""
import csv
import os
import re
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import ncu_lines  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "..", "gprf_b200", "csrc", "resident.cuh")


def phase_ranges():
    lines = open(SRC).read().splitlines()
    marks = []
    for i, ln in enumerate(lines, 1):
        m = re.match(r"(?:static )?__device__ (?:__noinline__|__forceinline__) \w+ (ph_\w+|run_unit|chol_diag_block|dbg_dump)\(", ln)
        if m:
            marks.append((i, m.group(1)))
        m = re.match(r"__global__ void .*(k_resident)\(", ln)
        if m:
            marks.append((i, "k_resident(main)"))
    return marks


def main():
    src_csv, obj = sys.argv[1:3]
    marks = phase_ranges()
    # frames above the first phase function are helpers (fragment loads, mk_loop, cov_frag ...): an
    # instruction is attributed to the innermost frame that lies in a phase function
    ncu_lines.MAIN_FROM = min(l for l, nm in marks if nm.startswith("ph_"))
    table = ncu_lines.line_table(obj, "k_resident")

    def phase(line):
        name = "helpers"
        for l0, nm in marks:
            if line >= l0:
                name = nm
        return name
    rows = list(csv.reader(open(src_csv)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
    hdr = rows[hi]
    col = {h: i for i, h in enumerate(hdr)}
    reasons = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    body = [r for r in rows[hi + 1:] if len(r) > col["stall_wait"]]
    base = int(body[0][0], 16)
    agg = {}
    for r in body:
        off = int(r[0], 16) - base
        f, l = table.get(off, ("?", 0))
        ph = phase(l) if f == "resident.cuh" else ("smem_chol.cuh" if f == "smem_chol.cuh" else "other:" + f)
        a = agg.setdefault(ph, {"ins": 0, "smp": 0, "dmma": 0, "wave": 0, "wave_ideal": 0, "r": {k: 0 for k in reasons}})
        ins = int(r[col["Instructions Executed"]] or 0)
        a["ins"] += ins
        a["smp"] += int(r[col["# Samples"]] or 0)
        if "DMMA" in r[col["Source"]]:
            a["dmma"] += ins
        a["wave"] += int(r[col["L1 Wavefronts Shared"]] or 0)
        a["wave_ideal"] += int(r[col["L1 Wavefronts Shared Ideal"]] or 0)
        for k in reasons:
            a["r"][k] += int(r[col[k]] or 0)
    ti = sum(a["ins"] for a in agg.values()) or 1
    ts = sum(a["smp"] for a in agg.values()) or 1
    print("# %d warp instructions, %d samples" % (ti, ts))
    print("%-18s %6s %6s %8s %6s %9s  top stall reasons (%% of the phase's samples)" % ("phase", "ins%", "smp%", "ins/DMMA", "DMMA%", "smem wf/ideal"))
    for ph, a in sorted(agg.items(), key=lambda kv: -kv[1]["smp"]):
        top = sorted(a["r"].items(), key=lambda kv: -kv[1])[:5]
        tot = sum(a["r"].values()) or 1
        print("%-18s %6.1f %6.1f %8.1f %6.1f %9.2f  %s" % (
            ph, 100.0 * a["ins"] / ti, 100.0 * a["smp"] / ts, a["ins"] / max(1, a["dmma"]),
            100.0 * a["dmma"] / max(1, sum(x["dmma"] for x in agg.values())),
            a["wave"] / max(1, a["wave_ideal"]),
            "  ".join("%s %.0f" % (k.replace("stall_", ""), 100.0 * v / tot) for k, v in top)))


if __name__ == "__main__":
    main()
