"""Aggregate an `ncu --page source --print-source cuda,sass --csv` export by CUDA source line."""
import csv
import sys


def num(x):
    try:
        return int(x.replace(",", ""))
    except ValueError:
        return 0


def main(path, top=45):
    rows = list(csv.reader(open(path)))
    cur, out = None, []
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur = r[1].split("/")[-1]
            continue
        if r[0] in ("Line No", "Function Name"):
            continue
        if r[0] != "":
            out.append((cur, num(r[0]), r[1][:100], num(r[4]), num(r[7])))
    ts = sum(o[3] for o in out) or 1
    ti = sum(o[4] for o in out) or 1
    print("total samples", ts, "total warp-level instructions", ti)
    for o in sorted(out, key=lambda o: -o[3])[:top]:
        print(f"{100 * o[3] / ts:5.1f}% smp {100 * o[4] / ti:5.1f}% inst  {o[0]}:{o[1]}  {o[2]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 45)
