#!/usr/bin/env python
"""Turn the ncu captures brought back in gpurun_out/ into the text summaries committed under profiles/ (run locally; ncu
reads .ncu-rep files without a GPU):

    python tools/make_profiles.py r2          # expects gpurun_out/r2_decoder_tc.ncu-rep, r2_policy_tc2_split.ncu-rep,
                                              # r2_final_launches.csv (ncu --metrics gpu__time_duration.sum launch list)
Also writes the SASS opcode table of the built library (cuobjdump) and profiles/traffic_policy_step.json."""
import collections
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
KEEP = re.compile(r"^(gpu__time_duration\.sum|dram__bytes_(read|write)\.sum$|dram__throughput\.avg\.pct|lts__throughput\.avg\.pct|"
                  r"lts__t_bytes\.sum$|l1tex__throughput\.avg\.pct|l1tex__data_pipe_lsu_wavefronts_mem_shared\.sum$|"
                  r"l1tex__data_bank_conflicts_pipe_lsu_mem_shared\.sum$|sm__throughput\.avg\.pct|sm__inst_executed\.sum$|"
                  r"sm__inst_executed\.avg\.per_cycle_active|smsp__issue_active\.avg\.pct|sm__warps_active\.avg\.pct|"
                  r"sm__pipe_tensor_cycles_active\.avg\.pct_of_peak_sustained_active|sm__mem_tensor_cycles_active\.avg\.pct_of_peak_sustained_active|"
                  r"sm__pipe_fma_cycles_active\.avg\.pct|sm__pipe_xu_cycles_active\.avg\.pct|sm__pipe_alu_cycles_active\.avg\.pct|"
                  r"smsp__average_warps_issue_stalled_\w+_per_issue_active\.ratio|launch__(grid_size|block_size|registers_per_thread|shared_mem_per_block_dynamic|occupancy_limit_(registers|shared_mem|warps)|waves_per_multiprocessor)$|smsp__cycles_active\.avg$)")


def ncu_csv(rep, page, extra=()):
    r = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv", *extra], capture_output=True, text=True)
    return list(csv.reader(io.StringIO(r.stdout)))


def raw_summary(rep, dst, header):
    rows = ncu_csv(rep, "raw")
    if len(rows) < 3:
        print("no data in", rep)
        return {}
    names, units, vals = rows[0], rows[1], rows[2]
    out, d = [header], {}
    for n, u, v in zip(names, units, vals):
        d[n] = (v, u)
        if KEEP.match(n):
            out.append(f"{n:95s} {v} {u}")
    open(dst, "w").write("\n".join(out) + "\n")
    print("wrote", dst)
    return d


def launches(csv_path, dst, header):
    lines = [l for l in open(csv_path) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for x in csv.DictReader(lines):
        try:
            v = float(x["Metric Value"].replace(",", ""))
        except (ValueError, KeyError):
            continue
        u = x["Metric Unit"]
        v = v / 1e6 if u in ("nsecond", "ns") else v / 1e3 if u in ("usecond", "us") else v
        agg.setdefault(x["Kernel Name"][:70], []).append(v)
    tot = sum(sum(v) for v in agg.values())
    out = [header, f"# {sum(len(v) for v in agg.values())} launches, {tot:.1f} ms in total (cold-cache, serialised under ncu: shares, not bench values)"]
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        out.append(f"{k:72s} n={len(v):4d}  mean {sum(v) / len(v):9.4f} ms  total {sum(v):9.2f} ms  {100 * sum(v) / tot:5.1f} %")
    open(dst, "w").write("\n".join(out) + "\n")
    print("wrote", dst)


def sass_table(dst):
    lib = os.path.join(ROOT, "jampr_plus_b200", "lib", "libjampr_b200.so")
    txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    ops = ["UTCHMMA", "LDTM", "STTM", "UBLKCP", "UTMALDG", "UTMASTG", "SYNCS", "UTCBAR", "HMMA", "LDSM", "LDGSTS", "FFMA2", "FMUL2",
           "FADD2", "HMNMX2", "HMUL2", "MUFU", "LDS", "STS", "LDG", "STG", "SHFL", "HADD2.F32"]
    out = ["# cuobjdump -sass jampr_plus_b200/lib/libjampr_b200.so | per-kernel opcode counts (static instructions).",
           "# UTCHMMA = tcgen05.mma kind::f16, LDTM / STTM = tcgen05.ld / st (TMEM), UBLKCP = cp.async.bulk (copy engine, 1-D),",
           "# UTMALDG / UTMASTG = tensor-map TMA (not used: operand tiles are pre-swizzled images moved by 1-D bulk copies or",
           "# written in place by the threads), SYNCS = mbarrier ops, HMMA = mma.sync (per-rollout phases of the decoder), LDSM = ldmatrix,",
           "# LDGSTS = cp.async (16-byte asynchronous row staging), FFMA2 / FMUL2 / FADD2 = packed fp32 pipe.",
           "%-64s" % "kernel" + " ".join("%9s" % o for o in ops)]
    for f in re.split(r"\n\s*Function : ", txt)[1:]:
        name = f.split("\n", 1)[0].strip()
        cnt = [len(re.findall(re.escape(o) + r"\b", f)) for o in ops]
        out.append("%-64s" % name[:64] + " ".join("%9d" % c for c in cnt))
    open(dst, "w").write("\n".join(out) + "\n")
    print("wrote", dst)


def main(tag):
    g = os.path.join(ROOT, "gpurun_out")
    cmd = ("ncu --set full --clock-control none --import-source on -k regex:%s -s 20 -c 1 python tools/profile_step.py 30   "
           "(eager full-size steps of BASELINE config 2: 65,536 rollouts, fp16; one mid-episode launch; times under ncu are "
           "not bench values)")
    tr = {}
    for rep, kern, name in ((f"{tag}_policy_tc2_split", "policy_step_tc2", f"{tag}_policy_step_tc2_split_ncu_raw.txt"),
                            (f"{tag}_decoder_tc", "decoder_tc_kernel", f"{tag}_decoder_tc_ncu_raw.txt")):
        p = os.path.join(g, rep + ".ncu-rep")
        if os.path.exists(p):
            d = raw_summary(p, os.path.join(OUT, name), "# round 2: " + cmd % kern)
            for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                if k in d:
                    v, u = d[k]
                    mult = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[u]
                    tr.setdefault(kern, {})[k] = float(v.replace(",", "")) * mult
    if tr:
        rd = sum(v.get("dram__bytes_read.sum", 0) for v in tr.values())
        wr = sum(v.get("dram__bytes_write.sum", 0) for v in tr.values())
        json.dump({"dram_bytes_read": rd, "dram_bytes_write": wr, "per_kernel": tr,
                   "source": f"profiles/{tag}_policy_step_tc2_split_ncu_raw.txt + profiles/{tag}_decoder_tc_ncu_raw.txt: one ncu --set "
                             "full capture of each of the two launches of a policy step at 65,536 rollouts (fp16)"},
                  open(os.path.join(OUT, "traffic_policy_step.json"), "w"), indent=1)
        print("wrote traffic_policy_step.json", rd + wr)
    lc = os.path.join(g, f"{tag}_final_launches.csv")
    if os.path.exists(lc):
        launches(lc, os.path.join(OUT, f"{tag}_launches_summary.txt"),
                 f"# round 2: ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv python bench.py --steps 1 --warmup 3 "
                 f"--extras none --no-cpu-baseline   (launch list of the default bench command, first 600 launches)")
    sass_table(os.path.join(OUT, f"{tag}_sass_counts.txt"))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "r2")
