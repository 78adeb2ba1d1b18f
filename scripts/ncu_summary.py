"""Extracts the metrics the profiles/ summaries quote from an .ncu-rep (ncu -i ... --page raw --csv)."""
import csv
import json
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_shared_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__warps_eligible.avg.per_cycle_active", "smsp__average_warp_latency_per_inst_issued.ratio"]


def stamp(files):
    """SHA-1 (16 hex) of the kernel sources the capture was taken from -- what bench.py compares before quoting traffic."""
    import hashlib
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, "clustering-hidden-markov-models-with-variational-bayesian-hierarchical-em_b200", "csrc")
    h = hashlib.sha1()
    for f in files:
        with open(os.path.join(pkg, f), "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()[:16]


def main(path, sources=()):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    res = []
    for vals in rows[2:]:
        d = {}
        for h, u, v in zip(hdr, units, vals):
            if h in KEYS or h == "Kernel Name" or (h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")):
                try:
                    fv = float(v.replace(",", ""))
                except ValueError:
                    fv = v
                if h.startswith("smsp__average_warps_issue_stalled") and isinstance(fv, float) and fv < 0.05:
                    continue
                d[h] = {"unit": u, "value": fv} if u else fv
        res.append(d)
    if sources and len(res) == 1:
        res[0]["source_sha1"] = stamp(sources)
        res[0]["source_files"] = list(sources)
    print(json.dumps(res[0] if len(res) == 1 else res, indent=1))


if __name__ == "__main__":
    main(sys.argv[1], tuple(sys.argv[2:]))     # optional: csrc file names whose SHA-1 is stamped into the summary
