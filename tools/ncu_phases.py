"""Splits the SASS page of an .ncu-rep at BAR.SYNC instructions (the phases of a phase-program kernel) and prints
executed warp-instructions, stall samples and the opcode mix per phase.
usage: python tools/ncu_phases.py gpurun_out/x.ncu-rep"""
import collections
import csv
import subprocess
import sys


def main(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hi]
    si, ii, sa = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    wf, wfi = hdr.index("L1 Wavefronts Shared"), hdr.index("L1 Wavefronts Shared Ideal")
    phases, cur = [], {"n": 0, "s": 0, "ops": collections.Counter(), "wf": 0, "wfi": 0}
    for r in rows[hi + 1:]:
        if len(r) <= ii:
            continue
        n, s = int(r[ii]), int(r[sa])
        op = r[si].split()
        op = [t for t in op if not t.startswith("@")][0].rstrip(";")
        cur["n"] += n; cur["s"] += s; cur["ops"][op.split(".")[0]] += n
        cur["wf"] += int(r[wf] or 0); cur["wfi"] += int(r[wfi] or 0)
        if op.startswith("BAR"):
            phases.append(cur)
            cur = {"n": 0, "s": 0, "ops": collections.Counter(), "wf": 0, "wfi": 0}
    phases.append(cur)
    tn, ts = sum(p["n"] for p in phases), sum(p["s"] for p in phases)
    print(f"total warp-instructions {tn}, samples {ts}")
    for i, p in enumerate(phases):
        if p["n"] == 0:
            continue
        top = ", ".join(f"{k} {100 * v / p['n']:.0f}%" for k, v in p["ops"].most_common(7))
        print(f"phase {i:2d}: {100 * p['n'] / tn:5.1f}% inst  {100 * p['s'] / max(ts, 1):5.1f}% samples  smem wavefronts {p['wf']:>9d} (ideal {p['wfi']})  | {top}")


if __name__ == "__main__":
    main(sys.argv[1])
