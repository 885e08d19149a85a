#!/usr/bin/env python
"""Turn the scratch artefacts of a GPU run (gpurun_out/) into the tracked evidence under profiles/:
  profiles/rNN_launches.csv          ncu launch list (gpu__time_duration per launch) + share table
  profiles/rNN_fused_ncu.txt         key metrics of the `ncu --set full` capture of the fused kernel + stall summary
  profiles/fused_traffic.json        dram bytes per fused launch (bench.py reads it for roofline.traffic)
  profiles/sass/*.sass               cuobjdump -sass of every kernel in libeffex_fx.so
usage: python tools/make_profiles.py r01 [tag]"""
import csv
import io
import json
import os
import re
import subprocess
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rnd = sys.argv[1] if len(sys.argv) > 1 else "r01"
tag = ("_" + sys.argv[2]) if len(sys.argv) > 2 else ""
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")
os.makedirs(os.path.join(P, "sass"), exist_ok=True)

# ---- launch list ------------------------------------------------------------
lc = os.path.join(G, "launches.csv")
step_dram = None
if os.path.exists(lc):
    lines = [l for l in open(lc) if not l.startswith("==")]
    rows = list(csv.reader(lines))
    hdr = rows[0]
    ki, vi, idi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("ID")
    mi, ui = hdr.index("Metric Name"), hdr.index("Metric Unit")
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9, "nsecond": 1, "usecond": 1e3, "msecond": 1e6}
    launches = {}                     # id -> {name, time_ns, dram}
    for r in rows[1:]:
        if len(r) <= vi:
            continue
        L = launches.setdefault(r[idi], {"name": r[ki], "time": 0.0, "dram": 0.0})
        v = float(r[vi].replace(",", "")) * scale.get(r[ui], 1)
        if r[mi].startswith("gpu__time_duration"):
            L["time"] = v
        elif r[mi].startswith("dram__bytes"):
            L["dram"] += v
    d = defaultdict(list)
    for L in launches.values():
        d[L["name"]].append(L)
    tot = sum(L["time"] for L in launches.values())
    # one step of bench.py --profile = byte sums + fused + finalize: traffic of a step = sum of the means
    step_names = [k for k in d if any(t in k for t in ("block_sums_kernel", "fused_kernel_stag", "finalize_rows"))]
    if step_names:
        step_dram = sum(sum(L["dram"] for L in d[k]) / len(d[k]) for k in step_names)
    with open(os.path.join(P, f"{rnd}{tag}_launches.csv"), "w") as fh:
        fh.write("# ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none python bench.py --steps 2 --warmup 3 --profile\n")
        fh.write("# per-launch times are cold-cache and serialised: compare SHARES, not absolutes\n")
        if step_dram:
            fh.write(f"# dram bytes of one step (byte sums + fused + finalize, means): {step_dram:.0f} = {step_dram / 594739200:.3f} x the algorithmic 594739200\n")
        fh.write("kernel,launches,mean_us,total_us,share,mean_dram_MB\n")
        for k, v in sorted(d.items(), key=lambda kv: -sum(L["time"] for L in kv[1])):
            t = [L["time"] for L in v]
            fh.write(f"\"{k}\",{len(v)},{sum(t) / len(t) / 1e3:.1f},{sum(t) / 1e3:.1f},{sum(t) / tot:.4f},{sum(L['dram'] for L in v) / len(v) / 1e6:.1f}\n")
        fh.write("# ---- raw list ----\n")
        fh.writelines(lines)

# ---- full capture -----------------------------------------------------------
rep = os.path.join(G, "prof_fused.ncu-rep")
if os.path.exists(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    m = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
    keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "dram__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
            "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic" if "launch__shared_mem_per_block_dynamic" in m else "launch__shared_mem_per_block",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_issued.avg.per_cycle_active",
            "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_tma.avg.pct_of_peak_sustained_active",
            "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
            "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.avg", "gpc__cycles_elapsed.max"]
    keys += sorted(k for k in m if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("per_issue_active.ratio"))

    def num(k):
        v, u = m[k]
        f = float(v.replace(",", ""))
        return f * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}.get(u, 1)
    with open(os.path.join(P, f"{rnd}{tag}_fused_ncu.txt"), "w") as fh:
        fh.write("# ncu --set full --clock-control none --import-source on -k regex:fused_kernel -s 3 -c 1 "
                 "python bench.py --steps 1 --warmup 3 --profile\n# kernel: fx::fused4096::fused_kernel_stag<0, false>, "
                 "550 block pairs (S=262144, N=4096, T=4)\n")
        for k in keys:
            if k in m:
                fh.write(f"{k:92s} {m[k][0]:>18s} {m[k][1]}\n")
        src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
        summ = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_src_summary.py"), "30"],
                              input=src, capture_output=True, text=True).stdout
        fh.write("\n# ---- warp-stall sampling (SASS view) ----\n" + summ)
    traffic = num("dram__bytes_read.sum") + num("dram__bytes_write.sum")
    facts = {"dram_bytes_per_launch": traffic, "dram_bytes_read": num("dram__bytes_read.sum"),
             "dram_bytes_write": num("dram__bytes_write.sum"), "source": f"profiles/{rnd}{tag}_fused_ncu.txt",
             "kernel": "fx::fused4096::fused_kernel_stag<0, false>", "units_per_launch": "550 block pairs",
             "fma_pipe_pct": float(m["sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"][0].replace(",", "")),
             "issue_active_pct": float(m["smsp__issue_active.avg.pct_of_peak_sustained_active"][0].replace(",", "")),
             "registers": int(float(m["launch__registers_per_thread"][0].replace(",", "")))}
    if step_dram:
        facts["step_dram_bytes"] = step_dram
        facts["step_dram_source"] = f"profiles/{rnd}{tag}_launches.csv (byte sums + fused + finalize)"
    json.dump(facts, open(os.path.join(P, "fused_traffic.json"), "w"), indent=1)

# ---- SASS listings -------------------------------------------------------------
lib = os.path.join(ROOT, "effex_b200", "libeffex_fx.so")
if os.path.exists(lib):
    sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    cur, name = [], None
    out = {}
    for line in sass.splitlines():
        mm = re.match(r"\s*Function : (\S+)", line)
        if mm:
            if name:
                out[name] = cur
            name, cur = mm.group(1), []
        if name:
            cur.append(line)
    if name:
        out[name] = cur
    dem = subprocess.run(["c++filt"] + list(out), capture_output=True, text=True).stdout.splitlines()
    for (mangled, body), d in zip(out.items(), dem):
        short = re.sub(r"\(.*", "", d).replace("fx::", "").replace("::", "_").replace("<", "_").replace(">", "").replace(" ", "")
        short = re.sub(r"[^A-Za-z0-9_]", "_", short)
        with open(os.path.join(P, "sass", short + ".sass"), "w") as fh:
            fh.write(f"// {d}\n// cuobjdump -sass effex_b200/libeffex_fx.so (sm_100a)\n")
            fh.write("\n".join(body) + "\n")
    print("sass:", len(out), "kernels")
    # static FMA-pipe instruction counts of the fused kernel (bench.py's fp32_pipe view): packed = FFMA2/FADD2/FMUL2
    # (2 issue cycles of the pipe per warp), scalar = FFMA/FADD/FMUL.  Whole kernel: the loop body executes
    # once per frame and thread, the prologue/epilogue instructions (a few per cent) once per segment.
    counts = {}
    for (mangled, body), dname in zip(out.items(), dem):
        for key, pat in (("no_autos", "fused_kernel_stag<0, false>"), ("autos", "fused_kernel_stag<0, true>")):
            if pat in dname.replace("(bool)0", "false").replace("(bool)1", "true").replace("(int)0", "0"):
                text = "\n".join(body)
                counts[key] = {"packed": len(re.findall(r"\b(FFMA2|FADD2|FMUL2)\b", text)),
                               "scalar": len(re.findall(r"\b(FFMA|FADD|FMUL)\b", text)),
                               "sttm": len(re.findall(r"\bSTTM\b", text)), "ldtm": len(re.findall(r"\bLDTM\b", text)),
                               "ublkcp": len(re.findall(r"\bUBLKCP\b", text))}
    fj = os.path.join(P, "fused_traffic.json")
    if counts and os.path.exists(fj):
        facts = json.load(open(fj))
        facts["sass_static_counts"] = counts
        json.dump(facts, open(fj, "w"), indent=1)
        print("sass counts:", counts)
print("profiles written")
