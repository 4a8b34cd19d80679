#!/usr/bin/env python
"""One `ncu --set full` capture (.ncu-rep) -> a short markdown summary for profiles/: headline metrics from the raw
page, warp-stall reasons and the hottest instructions from the source page.
Usage: summarize_ncu_full.py prof.ncu-rep "what was profiled" [algorithmic_flops] [algorithmic_bytes] > out.md"""
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("sm__cycles_elapsed.avg", "SM cycles elapsed"),
    ("sm__cycles_elapsed.avg.per_second", "SM clock"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem / block"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor pipe active (% of elapsed)"),
    ("sm__inst_executed_pipe_uniform.sum", "uniform-pipe instructions"),
    ("smsp__inst_executed.sum", "warp instructions executed"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy (%)"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy (%)"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput (% of peak)"),
    ("lts__t_bytes.sum", "L2 bytes (all)"),
    ("lts__t_sectors_srcunit_tex.sum", "L2 sectors from SMs"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput (% of peak)"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "LSU shared-memory wavefronts"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1/TEX throughput (% of peak)"),
]


def page(rep, name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep, what = sys.argv[1], sys.argv[2]
    flops = float(sys.argv[3]) if len(sys.argv) > 3 else 0.0
    abytes = float(sys.argv[4]) if len(sys.argv) > 4 else 0.0
    raw = page(rep, "raw")
    hdr, units = raw[0], raw[1]
    src = page(rep, "source")
    # a report may hold several launches (-c N): one raw row each; the source page repeats its header per launch
    tables, cur = [], None
    for r in src:
        if "Source" in r and "# Samples" in r:
            cur = {"h": r, "rows": []}
            tables.append(cur)
        elif cur is not None and len(r) == len(cur["h"]):
            cur["rows"].append(r)
    print(f"# ncu --set full: {what}\n")
    for li, vals in enumerate(raw[2:]):
        m = {h: (u, v) for h, u, v in zip(hdr, units, vals)}
        if len(raw) > 3:
            print(f"\n## launch {li + 1}\n")
        print(f"`{m.get('Kernel Name', ('', '?'))[1][:150]}`\n")
        print("| metric | value |\n|---|---:|")
        for k, label in KEYS:
            if k in m:
                u, v = m[k]
                print(f"| {label} (`{k}`) | {v} {u} |")
        try:
            us = float(m["gpu__time_duration.sum"][1].replace(",", ""))
            unit = m["gpu__time_duration.sum"][0]
            sec = us * {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}.get(unit, 1e-6)
            scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            dram = sum(float(m[k][1].replace(",", "")) * scale.get(m[k][0], 1.0) for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
            print(f"| DRAM traffic per launch | {dram / 1e6:.1f} MB ({dram / sec / 1e9:.0f} GB/s under the profiler) |")
            if flops and li == 0:
                print(f"| algorithmic FLOPs per launch | {flops / 1e9:.1f} GFLOP ({flops / sec / 1e12:.0f} TFLOP/s under the profiler) |")
            if abytes and li == 0:
                print(f"| algorithmic bytes per launch | {abytes / 1e6:.1f} MB (traffic / algorithmic = {dram / abytes:.2f}) |")
        except (KeyError, ValueError):
            pass
        if li < len(tables):
            h, rows = tables[li]["h"], tables[li]["rows"]
            ix = {c: i for i, c in enumerate(h)}
            stalls = [c for c in h if c.startswith("stall_") and "Not Issued" not in c]

            def num(x):
                try:
                    return int(x or 0)
                except ValueError:
                    return 0

            tot = {s_: sum(num(r[ix[s_]]) for r in rows) for s_ in stalls}
            n = sum(num(r[ix["# Samples"]]) for r in rows) or 1
            print("\n### Warp stall samples (all warps of the CTA)\n")
            print("| reason | share |\n|---|---:|")
            for s_, v in sorted(tot.items(), key=lambda x: -x[1])[:8]:
                print(f"| {s_[6:]} | {100 * v / n:.1f} % |")
            print("\n### Hottest instructions\n")
            print("| samples | executed | SASS | top stall |\n|---:|---:|---|---|")
            for r in sorted(rows, key=lambda r: -num(r[ix["# Samples"]]))[:12]:
                st = sorted(((num(r[ix[s_]]), s_[6:]) for s_ in stalls), reverse=True)[0]
                print(f"| {r[ix['# Samples']]} | {r[ix['Instructions Executed']]} | `{r[ix['Source']].strip()[:70]}` | {st[1]} |")


if __name__ == "__main__":
    main()
