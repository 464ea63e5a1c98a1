#!/usr/bin/env python
"""Turn a gpurun evidence directory (tools/evidence_r02.sh) into the tracked summaries under profiles/.

usage: tools/summarize_profile.py gpurun_out/<tag> <round-name> <workload>
writes profiles/<round-name>_<workload>_launches.csv   (the ncu per-launch durations, kernel names shortened)
       profiles/<round-name>_<workload>_ncu.md         (the `--set full` metrics that the roofline line quotes)
       profiles/traffic.json                           (dram bytes per launch of the dominant kernel, read by bench.py)
"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__cycles_elapsed.max", "lts__t_bytes.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
]
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}


def main():
    src, rnd, wl = sys.argv[1], sys.argv[2], sys.argv[3]
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    lpath = os.path.join(src, f"launches_{wl}.csv")
    if os.path.exists(lpath):
        lines = [l for l in open(lpath) if l.startswith('"')]
        rows = list(csv.reader(lines))
        hdr, body = rows[0], rows[1:]
        ki, vi, gi, bi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size"), hdr.index("Block Size")
        out = os.path.join(ROOT, "profiles", f"{rnd}_{wl}_launches.csv")
        tot = {}
        with open(out, "w") as f:
            f.write("# ncu --metrics gpu__time_duration.sum --clock-control none; command: python bench.py "
                    f"--workload {wl} --steps 2 --warmup 3 --no-cpu-baseline --no-e2e\n")
            f.write("# cold-cache, serialised launches: compare shares, not absolutes. at:: kernels = torch generating the "
                    "synthetic input / gathering the parity sample, outside the timed region\n")
            f.write("id,kernel,grid,block,duration_ns\n")
            for r in body:
                name = r[ki]
                short = name.split("(")[0].replace("void ", "")
                if len(short) > 90:
                    short = short[:87] + "..."
                short = short.replace(",", ";")
                f.write(f"{r[0]},{short},{r[gi].replace(',', 'x')},{r[bi].replace(',', 'x')},{r[vi]}\n")
                key = "nds" if ("nds::" in name or "unnamed>::" in name) else "other"
                tot[key] = tot.get(key, 0) + float(r[vi].replace(",", ""))
            f.write("# totals_ns: " + json.dumps(tot) + "\n")
        print("wrote", out, tot)
    rep = os.path.join(src, f"prof_{wl}.ncu-rep")
    raw = os.path.join(src, f"prof_{wl}.raw.csv")   # `ncu -i X.ncu-rep --page raw --csv` taken on the GPU box
    if os.path.exists(rep) or os.path.exists(raw):
        txt = open(raw).read() if os.path.exists(raw) else \
            subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(txt)))
        hdr, units = rows[0], rows[1]
        out = os.path.join(ROOT, "profiles", f"{rnd}_{wl}_ncu.md")
        traffic_path = os.path.join(ROOT, "profiles", "traffic.json")
        traffic = json.load(open(traffic_path)) if os.path.exists(traffic_path) else {}
        with open(out, "w") as f:
            f.write(f"# ncu --set full --clock-control none, workload {wl} ({rnd})\n\n")
            f.write(f"source report: gpurun_out/{os.path.basename(src)}/prof_{wl}.ncu-rep (scratch, not tracked); "
                    "command: `python bench.py --workload %s --steps 2 --warmup 3 --no-cpu-baseline --no-e2e`\n\n" % wl)
            for r in rows[2:]:
                kname = r[hdr.index("Kernel Name")]
                f.write(f"## {kname}\n\n| metric | value | unit |\n|---|---|---|\n")
                rd = wr = 0.0
                for w in WANT:
                    if w in hdr:
                        i = hdr.index(w)
                        f.write(f"| {w} | {r[i]} | {units[i]} |\n")
                        if w.startswith("dram__bytes_read"):
                            rd = float(r[i].replace(",", "")) * UNIT.get(units[i], 1.0)
                        if w.startswith("dram__bytes_write"):
                            wr = float(r[i].replace(",", "")) * UNIT.get(units[i], 1.0)
                f.write(f"\nDRAM traffic per launch (read + write): {(rd + wr) / 1e9:.4f} GB\n\n")
                short = kname.split("(")[0].split("::")[-1].split("<")[0]
                traffic[wl] = {"kernel": short, "dram_bytes_per_launch": rd + wr, "round": rnd}
        json.dump(traffic, open(traffic_path, "w"), indent=1, sort_keys=True)
        print("wrote", out, traffic)
    for name in (f"bench_{wl}.json", f"bench_{wl}_reference.json"):
        p = os.path.join(src, name)
        if os.path.exists(p) and os.path.getsize(p):
            open(os.path.join(ROOT, "profiles", f"{rnd}_{name}"), "w").write(open(p).read())


if __name__ == "__main__":
    main()
