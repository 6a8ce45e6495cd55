"""Regenerate the tracked profile summaries from a gpurun_out/ capture (run here after tools/final_capture.sh ran on the box).
Reads gpurun_out/prof_leaf*.ncu-rep (ncu --set full of k_leaf_hash) and gpurun_out/launches.csv (launch list of a bench
step) and writes profiles/r01_ncu_raw_k_leaf_hash.csv, profiles/r01_launches_bench_cfg2.csv, profiles/r01_traffic.json."""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")


def raw_page(rep, out):
    with open(out, "w") as f:
        subprocess.check_call(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=f, stderr=subprocess.DEVNULL)
    rows = list(csv.reader(open(out)))
    return dict(zip(rows[0], rows[2])), dict(zip(rows[0], rows[1]))


def main(rep_name="prof_leaf_final.ncu-rep", leaves_log2=19, width=234):
    rep = os.path.join(G, rep_name)
    vals, units = raw_page(rep, os.path.join(P, "r01_ncu_raw_k_leaf_hash.csv"))
    to_bytes = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
    rd = float(vals["dram__bytes_read.sum"]) * to_bytes[units["dram__bytes_read.sum"]]
    wr = float(vals["dram__bytes_write.sum"]) * to_bytes[units["dram__bytes_write.sum"]]
    n_leaves = 1 << leaves_log2
    alg = 8.0 * n_leaves * width + 32.0 * n_leaves
    perms = n_leaves * ((width + 7) // 8)
    winst = float(vals["smsp__inst_executed.sum"])
    tj = os.path.join(P, "r01_traffic.json")
    tr = json.load(open(tj))
    tr["k_leaf_hash"] = {
        "dram_bytes_per_algorithmic_byte": round((rd + wr) / alg, 4),
        "source": "ncu --set full, 2^%d leaves x %d: dram read %.3f GB + write %.3f GB vs %.3f GB algorithmic "
                  "(profiles/r01_ncu_raw_k_leaf_hash.csv)" % (leaves_log2, width, rd / 1e9, wr / 1e9, alg / 1e9),
        "thread_instructions_per_permutation": int(round(winst * 32 / perms, -1)),
        "instr_source": "ncu smsp__inst_executed.sum = %.2f G warp-inst for %.2f M permutations "
                        "(profiles/r01_ncu_raw_k_leaf_hash.csv)" % (winst / 1e9, perms / 1e6),
        "duration_ms_under_ncu": float(vals["gpu__time_duration.sum"]),
        "pipes_pct": {k: float(vals[m]) for k, m in {
            "issue_active": "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
            "alu": "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
            "fmaheavy": "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
            "fp64": "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
            "xu": "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"}.items() if m in vals},
    }
    json.dump(tr, open(tj, "w"), indent=1)
    print(json.dumps(tr["k_leaf_hash"], indent=1))
    src = os.path.join(G, "launches.csv")
    if os.path.exists(src):
        import shutil
        shutil.copy(src, os.path.join(P, "r01_launches_bench_cfg2.csv"))
        # per-kernel totals
        rows = [r for r in csv.reader(open(src)) if len(r) > 5]
        hdr = rows[0]
        ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
        tot = {}
        for r in rows[1:]:
            try:
                v = float(r[vi].replace(",", ""))
            except ValueError:
                continue
            k = r[ki].split("(")[0]
            c = tot.setdefault(k, [0, 0.0])
            c[0] += 1
            c[1] += v
        s = sum(v for _, v in tot.values())
        for k, (c, v) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
            print("%-40s %5d launches %12.0f (%5.1f %%)" % (k, c, v, 100 * v / s))


if __name__ == "__main__":
    main(*sys.argv[1:2])
