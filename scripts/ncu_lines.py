"""Attribute an ncu SASS source page (--page source --csv) to CUDA source lines.

    python scripts/ncu_lines.py gpurun_out/<tag>.source.csv gprf_b200/csrc/build/<object>.o <kernel-substring> [top]

The line table comes from `nvdisasm -g` of the object the library was linked from (same build),
joined with ncu's per-instruction rows by code offset.  Prints the source lines with the most
warp-stall samples and the most executed instructions.
"""
import csv
import os
import re
import subprocess
import sys
import tempfile


def line_table(obj, kern):
    tmp = tempfile.mkdtemp()
    subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, stdout=subprocess.DEVNULL)
    cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    txt = subprocess.run(["nvdisasm", "-gi", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
    # an instruction is preceded by its inline chain, innermost first; attribute it to the first
    # frame that is not one of the small helpers (fragment loads, mma2, staging loops)
    table, chain, inside = {}, [], False
    for ln in txt.splitlines():
        if ln.startswith("\t.section\t.text."):
            inside = kern in ln
            continue
        if not inside:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            chain.append((os.path.basename(m.group(1)), int(m.group(2))))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(\S.*);", ln)
        if m:
            if chain:
                pick = chain[0]
                for f, l in chain:
                    if f == MAIN and l >= MAIN_FROM:
                        pick = (f, l)
                        break
                table[int(m.group(1), 16)] = pick
                last = pick
                chain = []
            elif table:
                table[int(m.group(1), 16)] = last
    return table


MAIN = "resident.cuh"
MAIN_FROM = 330          # first line of run_unit: frames above it are helpers


def main():
    src_csv, obj, kern = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 25
    table = line_table(obj, kern)
    rows = list(csv.reader(open(src_csv)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
    hdr = rows[hi]
    cs, ce = hdr.index("# Samples"), hdr.index("Instructions Executed")
    body = [r for r in rows[hi + 1:] if len(r) > ce]
    base = int(body[0][0], 16)
    agg = {}
    for r in body:
        off = int(r[0], 16) - base
        key = table.get(off, ("?", 0))
        a = agg.setdefault(key, [0, 0])
        a[0] += int(r[cs] or 0)
        a[1] += int(r[ce] or 0)
    ts = sum(a[0] for a in agg.values()) or 1
    ti = sum(a[1] for a in agg.values()) or 1
    srcs = {}

    def text(key):
        f, l = key
        if f not in srcs:
            p = os.path.join(os.path.dirname(os.path.abspath(obj)), "..", f)
            srcs[f] = open(p).read().splitlines() if os.path.exists(p) else []
        return srcs[f][l - 1].strip()[:90] if 0 < l <= len(srcs[f]) else ""
    print("# %d stall samples, %d warp instructions executed" % (ts, ti))
    print("## by stall samples")
    for key, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print("%5.1f%% smp %5.1f%% ins  %s:%d  %s" % (100.0 * a[0] / ts, 100.0 * a[1] / ti, key[0], key[1], text(key)))
    print("## by instructions executed")
    for key, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print("%5.1f%% ins %5.1f%% smp  %s:%d  %s" % (100.0 * a[1] / ti, 100.0 * a[0] / ts, key[0], key[1], text(key)))


if __name__ == "__main__":
    main()
