"""Summarise an .ncu-rep (ncu --set full) into the small JSON committed under profiles/.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep "description of the command" [workload] > profiles/x.json
"""
import csv
import json
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct",
    "l1tex__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__cycles_elapsed.avg", "l1tex__m_xbar2l1tex_read_bytes.sum", "smsp__inst_executed.sum",
    "sm__inst_executed_pipe_tensor.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warp_latency_per_inst_issued.ratio", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
]
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}


def main(path, source):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    launches = []
    for vals in rows[2:]:
        rec = {"kernel": vals[hdr.index("Kernel Name")]}
        for h, u, v in zip(hdr, units, vals):
            if h in KEEP:
                rec[h] = {"value": v, "unit": u}
        launches.append(rec)

    def nbytes(rec, key):
        m = rec.get(key)
        return float(m["value"].replace(",", "")) * UNIT.get(m["unit"], 1.0) if m else None
    res = {"source": source, "workload": sys.argv[3] if len(sys.argv) > 3 else "cbg", "launches": launches}
    if launches:
        l0 = launches[0]
        rd, wr = nbytes(l0, "dram__bytes_read.sum"), nbytes(l0, "dram__bytes_write.sum")
        if rd is not None and wr is not None:
            res["dram_bytes_per_launch"] = rd + wr
        res["xbar2l1_read_bytes_per_launch"] = nbytes(l0, "l1tex__m_xbar2l1tex_read_bytes.sum")
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "")
