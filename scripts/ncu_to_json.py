"""Extract the per-launch figures bench.py quotes from an `ncu --set full` capture into profiles/<workload>_ncu.json.

    python scripts/ncu_to_json.py gpurun_out/s6_symik.ncu-rep symik k_symik_solve r1_s6
"""
import csv
import io
import json
import os
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = {
    "gpu__time_duration.sum": "duration_us_under_ncu",
    "dram__bytes_read.sum": "dram_bytes_read",
    "dram__bytes_write.sum": "dram_bytes_write",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active": "fp64_pipe_pct_of_peak",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "launch__registers_per_thread": "registers_per_thread",
    "smsp__inst_executed.sum": "warp_instructions",
    "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum": "thread_dfma",
    "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum": "thread_dmul",
    "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum": "thread_dadd",
}
UNIT_SCALE = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "ms": 1e3, "us": 1.0, "ns": 1e-3, "s": 1e6}


def main():
    rep, wl, kernel, tag = sys.argv[1:5]
    sum_all = len(sys.argv) > 5 and sys.argv[5] == "sum"      # a step of several kernels (K3): byte / instruction counts are summed
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    out = {"source": f"ncu --set full --clock-control none, {os.path.basename(rep)} ({tag}); per launch; under the profiler "
                     "(use for traffic / ratios, not as a timing)", "kernel": kernel}
    additive = {"duration_us_under_ncu", "dram_bytes_read", "dram_bytes_write", "warp_instructions", "thread_dfma", "thread_dmul", "thread_dadd"}
    seen = []
    parts = []      # sum mode: the per-kernel figures, for the duration-weighted means of the percentages
    for r in rows[2:]:
        name_k = r[hdr.index("Kernel Name")]
        if kernel not in name_k:
            continue
        short = name_k.split("(")[0]
        if short in seen:
            continue
        seen.append(short)
        part = {"kernel": short}
        for k, name in WANT.items():
            if k in hdr:
                part[name] = float(r[hdr.index(k)].replace(",", "")) * UNIT_SCALE.get(units[hdr.index(k)], 1.0)
        parts.append(part)
        for k, name in WANT.items():
            if k in hdr:
                v = float(r[hdr.index(k)].replace(",", ""))
                v *= UNIT_SCALE.get(units[hdr.index(k)], 1.0)
                if sum_all and name in additive:
                    out[name] = out.get(name, 0.0) + v
                elif name not in out:
                    out[name] = v
        if not sum_all:
            break
    if sum_all:
        out["kernels_summed"] = seen
        total = sum(p_["duration_us_under_ncu"] for p_ in parts)
        for name in ("fp64_pipe_pct_of_peak", "issue_active_pct", "warps_active_pct"):      # duration-weighted over the pass
            out[name] = sum(p_.get(name, 0.0) * p_["duration_us_under_ncu"] for p_ in parts) / total
        out["registers_per_thread"] = max(p_.get("registers_per_thread", 0.0) for p_ in parts)
        out["per_kernel"] = [{k: p_.get(k) for k in ("kernel", "duration_us_under_ncu", "warp_instructions", "issue_active_pct",
                                                      "fp64_pipe_pct_of_peak", "warps_active_pct", "registers_per_thread",
                                                      "dram_bytes_read", "dram_bytes_write")} for p_ in parts]
    sha = os.path.join(os.path.dirname(rep), f"{tag}_csrc_sha16.txt")
    if os.path.exists(sha):
        out["csrc_sha16"] = open(sha).read().strip()
    out["dram_bytes_per_launch"] = out.get("dram_bytes_read", 0.0) + out.get("dram_bytes_write", 0.0)
    if "thread_dfma" in out:
        out["executed_fp64_flop"] = 2 * out["thread_dfma"] + out.get("thread_dmul", 0.0) + out.get("thread_dadd", 0.0)
    with open(os.path.join(REPO, "profiles", f"{wl}_ncu.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
