"""Summarise an `ncu --set full` capture (.ncu-rep) for profiles/: one CSV row per kernel launch with
the metrics DESIGN.md and bench.py refer to, and (with --traffic CONFIG N) the profiles/r02_traffic.json
entry bench.py reads `roofline.traffic` from.

usage: python tools/ncu_summary.py gpurun_out/x.ncu-rep profiles/r02_x_ncu_full.csv [--traffic 2 10000000]"""
import csv
import json
import os
import subprocess
import sys

WANT = ["Kernel Name", "Block Size", "Grid Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__waves_per_multiprocessor", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps", "sm__cycles_active.avg",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    cols = [hdr.index(w) for w in WANT if w in hdr]
    stall = [i for i, h in enumerate(hdr) if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued")]
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow([f"{hdr[i]} [{units[i]}]" if units[i] else hdr[i] for i in cols] + ["top stall reasons (% of samples)"])
        for r in rows[2:]:
            st = {hdr[i].replace("smsp__pcsamp_warps_issue_stalled_", ""): float(r[i]) for i in stall if r[i] not in ("", "n/a")}
            tot = sum(st.values()) or 1.0
            top = "; ".join(f"{k} {v / tot * 100:.0f}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:5])
            w.writerow([r[i] for i in cols] + [top])
    if "--traffic" in sys.argv:
        k = sys.argv.index("--traffic")
        config, n = int(sys.argv[k + 1]), int(sys.argv[k + 2])
        ki, ri, wi = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        for r in rows[2:]:
            if "sweep_tick_kernel" in r[ki]:
                b = float(r[ri]) * scale[units[ri]] + float(r[wi]) * scale[units[wi]]
                path = os.path.join(os.path.dirname(out) or ".", "r02_traffic.json")
                try:
                    doc = json.load(open(path))
                except Exception:
                    doc = {"what": "dram__bytes_read.sum + dram__bytes_write.sum of ONE sweep_tick_kernel launch "
                                   "(ncu --set full --clock-control none); read by bench.py (roofline.traffic)", "captures": []}
                doc["captures"] = [c for c in doc["captures"] if not (c["config"] == config and c["n"] == n)]
                doc["captures"].append({"config": config, "n": n, "dram_bytes": int(b), "kernel": r[ki],
                                        "source": os.path.relpath(out, os.path.dirname(os.path.dirname(os.path.abspath(out)))) })
                json.dump(doc, open(path, "w"), indent=1)
                break


if __name__ == "__main__":
    main()
