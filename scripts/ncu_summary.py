#!/usr/bin/env python
"""Condense an ncu report (.ncu-rep of the dominant kernel, `ncu --set full --import-source on`) into the small
tracked artefacts under profiles/: <out>_summary.json (raw-page metrics incl. DRAM bytes per launch, stall
sampling, top source lines) and <out>_details.txt (details page).  Usage:
    python scripts/ncu_summary.py gpurun_out/prof5.ncu-rep profiles/round1_ncu "description" [instances_per_launch] [kernel_tag]
kernel_tag: what bench.py's config.kernel_tag says for the run the profile belongs to (bench.py only quotes `traffic` from a
summary whose tag matches the kernel it launched)."""
import collections
import csv
import json
import subprocess
import sys


def raw_metrics(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, d = rows[0], rows[1], rows[2]
    return {h: (d[i], units[i]) for i, h in enumerate(hdr)}


def source_page(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[2]
    ci, cs = hdr.index("Instructions Executed"), hdr.index("# Samples")
    stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    cur = curfile = None
    seen = set()
    agg = collections.defaultdict(lambda: [0, 0])
    src = {}
    stalls = collections.Counter()
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            curfile = r[1]; continue
        if r[0] in ("Function Name", "Line No"):
            continue
        if r[0] != "":
            try:
                cur = (curfile.split("/")[-1], int(r[0])); src[cur] = r[1]
            except Exception:
                cur = None
            continue
        if cur is None or len(r) <= ci or not r[ci].isdigit() or r[2] in seen:
            continue
        seen.add(r[2])
        agg[cur][0] += int(r[ci]); agg[cur][1] += int(r[cs]) if r[cs].isdigit() else 0
        for i in stall_cols:
            if i < len(r) and r[i].isdigit():
                stalls[hdr[i][6:]] += int(r[i])
    return agg, src, stalls, len(seen)


def main():
    rep, outp, what = sys.argv[1], sys.argv[2], sys.argv[3]
    n_inst = int(sys.argv[4]) if len(sys.argv) > 4 else 4096
    tag = sys.argv[5] if len(sys.argv) > 5 else None
    m = raw_metrics(rep)
    mul = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}

    def f(name):
        return float(m[name][0])

    rd = f("dram__bytes_read.sum") * mul[m["dram__bytes_read.sum"][1]]
    wr = f("dram__bytes_write.sum") * mul[m["dram__bytes_write.sum"][1]]
    agg, src, stalls, n_sass = source_page(rep)
    tot_i = sum(v[0] for v in agg.values()); tot_s = sum(v[1] for v in agg.values()); st = sum(stalls.values())
    top = sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]
    summ = {
        "what": what, "kernel_tag": tag,
        "gpu_time_us": f("gpu__time_duration.sum") * (1e3 if m["gpu__time_duration.sum"][1] == "ms" else 1.0),
        "dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes_per_launch": rd + wr,
        "registers_per_thread": f("launch__registers_per_thread"),
        "grid": f("launch__grid_size"), "block": f("launch__block_size"),
        "occupancy_limit_blocks": {"registers": f("launch__occupancy_limit_registers"), "shared_mem": f("launch__occupancy_limit_shared_mem")},
        "warps_active_pct_of_peak": f("sm__warps_active.avg.pct_of_peak_sustained_active"),
        "issue_active_pct": f("smsp__issue_active.avg.pct_of_peak_sustained_active"),
        "sm_throughput_pct": f("sm__throughput.avg.pct_of_peak_sustained_elapsed"),
        "pipe_pct": {"lsu": f("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
                     "fma": f("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
                     "alu": f("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
                     "fp64": f("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active")},
        "warp_instructions": f("smsp__inst_executed.sum"),
        "warp_instructions_per_instance": f("smsp__inst_executed.sum") / n_inst,
        "avg_active_threads_per_warp_inst": f("smsp__thread_inst_executed_per_inst_executed.ratio"),
        "cycles_elapsed": f("sm__cycles_elapsed.max"),
        "unique_sass_instructions_executed_kernel": n_sass,
        "stall_sampling_pct": {k: round(100.0 * v / st, 1) for k, v in stalls.most_common(8)},
        "top_source_lines_by_samples": [{"line": f"{k[0]}:{k[1]}", "samples_pct": round(100.0 * v[1] / tot_s, 1),
                                         "inst_pct": round(100.0 * v[0] / tot_i, 1), "source": src[k].strip()[:100]} for k, v in top],
    }
    summ["all_source_lines"] = [[f"{k[0]}:{k[1]}", round(100.0 * v[0] / tot_i, 2), round(100.0 * v[1] / tot_s, 2)]
                                for k, v in sorted(agg.items()) if v[1] * 1000 >= tot_s or v[0] * 1000 >= tot_i]     # [line, inst %, samples %]
    json.dump(summ, open(outp + "_summary.json", "w"), indent=1)
    det = subprocess.run(["ncu", "-i", rep, "--page", "details"], capture_output=True, text=True).stdout
    open(outp + "_details.txt", "w").write("\n".join(det.splitlines()[:170]) + "\n")
    print(json.dumps({k: summ[k] for k in ("gpu_time_us", "dram_bytes_per_launch", "issue_active_pct", "warp_instructions_per_instance", "stall_sampling_pct")}))


if __name__ == "__main__":
    main()
