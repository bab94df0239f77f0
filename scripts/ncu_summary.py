"""Turn an Nsight Compute report (ncu --set full ... -o gpurun_out/X) into the committed evidence under profiles/:
a per-kernel text summary of the counters DESIGN.md argues from, and profiles/r02_ncu_traffic.json — the DRAM bytes per
launch that bench.py cites in `roofline.traffic` (with report name, kernel and capture time, never a literal).

    python scripts/ncu_summary.py gpurun_out/r02k_prof.ncu-rep profiles/r02_ncu_field_kernels_summary.txt \
        [--traffic field_bwd_fine=field_bwd4:12582912 --traffic field_fwd_fine=mlp_tc_fwd:12582912]
"""
import argparse
import csv
import datetime
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEEP = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_red.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld_lookup_hit.sum",
    "smsp__inst_executed_op_global_red.sum", "l1tex__m_l1tex2xbar_write_sectors_mem_global_op_red.sum",
    "lts__d_atomic_input_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_requests_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_local_op_st.sum",
]


def load(report):
    out = subprocess.run(["ncu", "-i", report, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[2:]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("out")
    ap.add_argument("--traffic", action="append", default=[], help="key=kernel_substring:points_per_launch")
    a = ap.parse_args()
    h, rows = load(a.report)
    ki = h.index("Kernel Name")
    stalls = [x for x in h if "issue_stalled" in x and "per_issue_active" in x]
    lines = ["# %s  (ncu --set full --clock-control none; per-launch counters, cold-cache replays: compare shares, not absolutes)"
             % os.path.basename(a.report), ""]
    for r in rows:
        lines.append("Kernel  " + r[ki])
        for k in KEEP:
            if k in h and r[h.index(k)] not in ("", "n/a"):
                lines.append("  %-78s %s" % (k, r[h.index(k)]))
        st = sorted(((float(r[h.index(s)] or 0), s) for s in stalls), reverse=True)[:7]
        lines.append("  stall cycles per issue: " + ", ".join("%s %.2f" % (s.split("stalled_")[1].split("_per")[0], v) for v, s in st))
        lines.append("")
    with open(a.out, "w") as f:
        f.write("\n".join(lines))
    print("wrote", a.out)
    if a.traffic:
        p = os.path.join(ROOT, "profiles", "r02_ncu_traffic.json")
        d = json.load(open(p)) if os.path.isfile(p) else {}
        when = datetime.datetime.utcfromtimestamp(os.path.getmtime(a.report)).strftime("%Y-%m-%dT%H:%M:%SZ")
        for spec in a.traffic:
            key, rest = spec.split("=")
            sub, pts = rest.split(":")
            for r in rows:
                if sub in r[ki]:
                    def num(name):
                        v, u = float(r[h.index(name)].replace(",", "")), rows and name
                        return v
                    # ncu prints dram bytes in the unit of the header row below the names; read it from the units row
                    units = list(csv.reader(io.StringIO(subprocess.run(["ncu", "-i", a.report, "--page", "raw", "--csv"],
                                                                        capture_output=True, text=True).stdout)))[1]
                    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
                    rd = num("dram__bytes_read.sum") * scale[units[h.index("dram__bytes_read.sum")]]
                    wr = num("dram__bytes_write.sum") * scale[units[h.index("dram__bytes_write.sum")]]
                    d[key] = {"dram_bytes": rd + wr, "dram_bytes_read": rd, "dram_bytes_write": wr, "points": int(pts),
                              "kernel": r[ki][:120], "report": os.path.basename(a.report), "captured": when,
                              "summary": os.path.relpath(a.out, ROOT)}
                    break
        json.dump(d, open(p, "w"), indent=1)
        print("updated", p)


if __name__ == "__main__":
    main()
