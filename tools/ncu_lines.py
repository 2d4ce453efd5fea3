#!/usr/bin/env python
"""Per-source-line share of executed warp instructions and of warp-stall samples of one kernel of an ``ncu --set full
--import-source on`` report (``ncu -i x.ncu-rep --page source --csv --print-source cuda,sass``, SASS rows summed under the CUDA
line they belong to):

  python tools/ncu_lines.py gpurun_out/x.ncu-rep [top]
"""
import collections
import csv
import io
import subprocess
import sys


def main():
    path, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    cur, fname, txt = None, "", {}
    inst, samp = collections.Counter(), collections.Counter()
    for r in csv.reader(io.StringIO(out)):
        if not r:
            continue
        if r[0] == "File Path":
            fname = r[1].rsplit("/", 1)[-1]
        elif r[0] == "Function Name":
            print("== %s" % r[1])
        elif r[0] not in ("", "Line No"):
            cur = (fname, r[0])
            txt[cur] = " ".join(x.strip() for x in r[1:5])[:120]
        elif r[0] == "" and len(r) > 7:
            try:
                inst[cur] += int(r[7])
                samp[cur] += int(r[4])
            except ValueError:
                pass
    ti, ts = max(sum(inst.values()), 1), max(sum(samp.values()), 1)
    print("warp instructions executed %d, stall samples %d" % (ti, ts))
    print("%-22s %8s %8s  source" % ("file:line", "inst %", "stall %"))
    for k, c in sorted(inst.items(), key=lambda kv: -kv[1])[:top]:
        print("%-22s %8.2f %8.2f  %s" % ("%s:%s" % k, 100.0 * c / ti, 100.0 * samp[k] / ts, txt.get(k, "")))


if __name__ == "__main__":
    main()
