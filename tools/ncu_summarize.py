"""Summarise an ncu report (`ncu --set full ...`) into profiles/: one record per distinct kernel (first launch
seen) with the metrics the roofline discussion uses, and the per-launch DRAM traffic table bench.py reads.

usage: python tools/ncu_summarize.py gpurun_out/prof.ncu-rep profiles/rN_ncu_full_summary.json [profiles/ncu_traffic.json] [source note]
"""
import csv
import io
import json
import re
import subprocess
import sys

KEEP = ["launch__grid_size", "launch__block_size", "launch__registers_per_thread", "gpu__time_duration.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "l1tex__m_xbar2l1tex_read_sectors_mem_global_op_tma_ld.sum",
        "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__throughput.avg.pct_of_peak_sustained_active", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio"]
# profiler names used by the library's CUDA-event profiler (bench.py roofline keys) per kernel-name fragment
NAMES = [("k_orth", None), ("k_csr_spmv_bulk", "csr_spmv"), ("k_csr_spmv_stream", "csr_spmv"), ("k_csr_spmv", "csr_spmv"),
         ("k_vq_tma", "vq_tma"), ("k_vq_mma", "vq_mma"), ("k_start_step", "start_step"), ("k_spmv_rows", "gram_spmv")]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    traffic_out = sys.argv[3] if len(sys.argv) > 3 else None
    note = sys.argv[4] if len(sys.argv) > 4 else ""
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    seen, kernels = {}, []
    for r in rows[2:]:
        name = r[ki]
        key = name.split(">(")[0] + ">"
        seen[key] = seen.get(key, 0) + 1
        if seen[key] > 1:
            continue
        rec = {"Kernel Name": name}
        for h, u, v in zip(hdr, units, r):
            if h in KEEP and v != "":
                rec[f"{h} [{u}]" if u else h] = v
        kernels.append(rec)
    json.dump({"source": note, "launches_per_kernel": seen, "kernels": kernels}, open(out, "w"), indent=1)
    if traffic_out:
        def to_bytes(v, u):
            mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
            return float(v) * mult
        ri, wi = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
        acc = {}
        for r in rows[2:]:
            name = r[ki]
            tag = None
            if "k_upd<" in name:
                tag = "update_spec_tma" if re.search(r"k_upd<[^,]+, (?:\([^)]*\))?(1|true)", name) else "reorth_tma"
            elif "k_start_step" in name:
                tag = "start_step"
            elif "k_orth" in name:
                m = re.search(r"k_orth<[^,]+, (?:\([^)]*\))?(\d)>", name)
                tag = {"0": "dots_tma", "1": "reorth_tma", "2": "update_spec_tma"}.get(m.group(1) if m else "", None)
            else:
                for frag, t in NAMES:
                    if frag in name and t:
                        tag = t
                        break
            if not tag:
                continue
            b = to_bytes(r[ri], units[ri]) + to_bytes(r[wi], units[wi])
            a = acc.setdefault(tag, [0.0, 0])
            a[0] += b
            a[1] += 1
        json.dump({k: {"dram_bytes_per_launch": v[0] / v[1], "launches_captured": v[1], "note": note}
                   for k, v in acc.items()}, open(traffic_out, "w"), indent=1)
    print("kernels:", {k: v for k, v in seen.items()})


if __name__ == "__main__":
    main()
