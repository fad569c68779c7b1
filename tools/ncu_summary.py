"""Read an `ncu --set full` report (.ncu-rep) here (no GPU needed) and print the metrics the profiles/ notes quote as a
markdown table; for the count kernel also write profiles/count_kernel_traffic.json, which bench.py's roofline.traffic
reads (DRAM bytes per launch, tied to the sha256 of the kernel's sources so that a stale capture is never reported).

usage: python tools/ncu_summary.py gpurun_out/r02_prof_count.ncu-rep [--traffic-json "capture description"]"""
import csv
import hashlib
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "registers / thread"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem / block"), ("launch__shared_mem_per_block_static", "static smem / block"),
    ("launch__occupancy_limit_registers", "CTAs / SM allowed by registers"), ("launch__occupancy_limit_shared_mem", "CTAs / SM allowed by smem"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe, % of peak (active cycles)"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "FP64 pipe"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots used"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active, % of 64"),
    ("smsp__warps_eligible.avg.per_cycle_active", "eligible warps / cycle"),
    ("sm__cycles_active.avg", "SM cycles active"), ("sm__cycles_elapsed.avg", "SM cycles elapsed"),
    ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM written"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput, % of peak"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput, % of peak"),
    ("lts__t_sectors_srcunit_tex_op_red.sum", "L2 reduction sectors"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput, % of peak"),
]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        name = vals[hdr.index("Kernel Name")]
        print(f"### `{name}`\n\n| metric | value |\n|---|---:|")
        got = {}
        for key, label in WANT:
            if key in hdr:
                i = hdr.index(key)
                got[key] = (vals[i], units[i])
                print(f"| {label} (`{key}`) | {vals[i]} {units[i]} |")
        stalls = []
        for i, h in enumerate(hdr):
            if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued"):
                try:
                    stalls.append((float(vals[i]), h.replace("smsp__pcsamp_warps_issue_stalled_", "")))
                except ValueError:
                    pass
        tot = sum(v for v, _ in stalls) or 1.0
        print("\nstall reasons (PC samples, share of all samples): " +
              ", ".join(f"{n} {v / tot * 100:.1f} %" for v, n in sorted(stalls, reverse=True)[:8]))
        if "--traffic-json" in sys.argv and "count_kernel" in name:
            def mb(key):
                v, u = got[key]
                f = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
                return float(v) * f
            h = hashlib.sha256()
            for src in ("nk_count.cu", "nk_device.cuh"):
                with open(os.path.join(ROOT, "neurokmer_b200", "csrc", src), "rb") as f:
                    h.update(f.read())
            rec = {"kernel": name, "dram_bytes_read": mb("dram__bytes_read.sum"), "dram_bytes_written": mb("dram__bytes_write.sum"),
                   "dram_bytes_per_launch": mb("dram__bytes_read.sum") + mb("dram__bytes_write.sum"),
                   "capture": sys.argv[sys.argv.index("--traffic-json") + 1], "source_sha256": h.hexdigest()}
            with open(os.path.join(ROOT, "profiles", "count_kernel_traffic.json"), "w") as f:
                json.dump(rec, f, indent=1)
            print(f"\n(wrote profiles/count_kernel_traffic.json: {rec['dram_bytes_per_launch'] / 1e6:.1f} MB per launch)")


if __name__ == "__main__":
    main()
