"""Prints the last N launches of an `ncu --csv --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum`
launch list as one line per kernel.   usage: python tools/launch_table.py launches.csv [N]"""
import csv
import sys


def main(path, last=12):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]
    ki, mi, vi, ii = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
    d, order = {}, []
    for r in rows[1:]:
        k = (r[ii], r[ki])
        if k not in d:
            d[k] = {}
            order.append(k)
        d[k][r[mi]] = float(r[vi].replace(",", ""))
    for k in order[-last:]:
        m = d[k]
        print(k[0].rjust(4), k[1][:64].ljust(64), f"{m.get('gpu__time_duration.sum', 0) / 1000:8.1f} us  rd {m.get('dram__bytes_read.sum', 0) / 1e6:7.1f} MB"
              f"  wr {m.get('dram__bytes_write.sum', 0) / 1e6:7.1f} MB  inst {m.get('smsp__inst_executed.sum', 0) / 1e6:6.1f} M")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 12)
