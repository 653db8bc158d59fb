"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel.
usage: python tools/summarize_launches.py profiles/<file>.csv [top_n]"""
import collections
import csv
import re
import sys


def main(path, top=30):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        full = row["Kernel Name"]
        m = re.search(r"ffc_kernel(?:_coop)?<(\w+)(<[^>]*>)?", full)
        if m:
            name = "ffc_b200:" + m.group(1) + (m.group(2) or "")
        elif re.match(r"(void )?(conv_v5|wgrad_v5|pack_v5|conv_small\w*|wgrad_small\w*|conv1x1_narrow|fu3_mix|fu3_wgrad|fu3_pack|fu4|fu4_bwd|ffc_zero)_kernel|(void )?fu3_wgrad_reduce", full):
            name = "ffc_b200:" + re.sub(r"\(.*", "", full).replace("void ", "").replace("fu4::", "")
        else:
            name = re.sub(r"<.*", "", full)[:70]
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        v = v / 1000 if unit.startswith("n") else (v * 1000 if unit.startswith("m") else v)
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    ours = sum(v[1] for k, v in agg.items() if k.startswith("ffc_b200:"))
    print(f"{sum(v[0] for v in agg.values())} launches, {tot:.0f} us total; ffc_b200 kernels {100 * ours / tot:.1f}% of device time")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"{v[1]:10.1f} us {100 * v[1] / tot:5.1f}%  n={v[0]:4d}  avg {v[1] / v[0]:8.1f} us  {k}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30)
