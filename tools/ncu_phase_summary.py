"""Summarise an `ncu --page source --print-source cuda,sass --csv` export of policy_tc.cu by kernel phase
(line ranges looked up from marker comments in the source) and by stall reason."""
import csv
import re
import sys


def num(x):
    try:
        return int(x.replace(",", ""))
    except ValueError:
        return 0


def main(path, src):
    lines = open(src).read().split("\n")
    marks = []
    pats = [("mma/loadw", r"void load_w2?\("),
            ("agg", r"void agg_nbh_tc2?\("), ("epilogue", r"^// y = GELU|^// ---- layer epilogue"),
            ("matvec", r"^// -* ?decoder mat-vec"), 
            ("kernel head/carve", r"^__global__ void"), ("prologue", r"// -+ prologue"), ("h0", r"// ---- h_0 ="),
            ("layer loop", r"for \(int l = 0; l < [67]"), 
            ("ne epilogue", r"// -+ node embedding rows"),
            ("dec b/a te,stats", r"// -+ decoder$"), ("dec c ctxt", r"// \(c\)"), ("dec d qp", r"// \(d\)"),
            ("dec e scores", r"// \(e\)"), ("dec f softmax/pn", r"// \(f\)"), ("dec g gbar", r"// \(g\)"),
            ("dec h g", r"// \(h\)"), ("dec i lp", r"// \(i\)"), ("dec j logits", r"// \(j\)"),
            ("select", r"// ---- log-softmax"), ("host", r"^static size_t p[t2]_smem_bytes")]
    for name, pat in pats:
        for i, l in enumerate(lines):
            if re.search(pat, l):
                marks.append((i + 1, name))
                break
    marks.sort()
    rows = list(csv.reader(open(path)))
    hdr = [r for r in rows if r and r[0] == "Line No"][0]
    idx = {h: i for i, h in enumerate(hdr)}
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    cur, agg, st_tot = None, {}, {s: 0 for s in stalls}
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur = r[1].split("/")[-1]
            continue
        if r[0] in ("Line No", "Function Name"):
            continue
        if r[0] == "":
            if len(r) >= len(hdr):
                for s in stalls:
                    st_tot[s] += num(r[idx[s]])
            continue
        key = cur
        if cur == src.split("/")[-1]:
            ln = num(r[0])
            key = "?"
            for a, name in marks:
                if ln >= a:
                    key = name
        d = agg.setdefault(key, [0, 0])
        d[0] += num(r[4])
        d[1] += num(r[7])
    ts = sum(v[0] for v in agg.values()) or 1
    ti = sum(v[1] for v in agg.values()) or 1
    print(f"total samples {ts}, warp-level instructions {ti}")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        print(f"{k:24s} samples {100 * v[0] / ts:5.1f}%   inst {100 * v[1] / ti:5.1f}%")
    T = sum(st_tot.values()) or 1
    print("stall reasons: " + ", ".join(f"{s[6:]} {100 * v / T:.1f}%" for s, v in sorted(st_tot.items(), key=lambda kv: -kv[1])[:9]))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "jampr_plus_b200/csrc/policy_tc.cu")
