"""Summarises an Nsight Compute report (ncu -i REP --page raw --csv) into the JSON kept under profiles/.

  python profiles/ncu_summary.py gpurun_out/prof.ncu-rep "what was captured" > profiles/r02/ncu_xxx.json
"""
import csv
import io
import json
import subprocess
import sys

KEYS = {
    "gpu_time_ms": ("gpu__time_duration.sum", 1.0),
    "registers": ("launch__registers_per_thread", 1.0),
    "grid": ("launch__grid_size", 1.0),
    "block": ("launch__block_size", 1.0),
    "warps_active_pct": ("sm__warps_active.avg.pct_of_peak_sustained_active", 1.0),
    "issue_active_pct": ("smsp__issue_active.avg.pct_of_peak_sustained_active", 1.0),
    "inst_executed": ("smsp__inst_executed.sum", 1.0),
    "lanes_per_instruction": ("smsp__thread_inst_executed_per_inst_executed.ratio", 1.0),
    "stall_no_instruction_per_issue": ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", 1.0),
    "stall_long_scoreboard_per_issue": ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", 1.0),
    "stall_wait_per_issue": ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", 1.0),
    "stall_lg_throttle_per_issue": ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", 1.0),
    "l1_data_pipe_pct": ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", 1.0),
    "l1_hit_pct": ("l1tex__t_sector_hit_rate.pct", 1.0),
    "l2_hit_pct": ("lts__t_sector_hit_rate.pct", 1.0),
    "sm_throughput_pct": ("sm__throughput.avg.pct_of_peak_sustained_elapsed", 1.0),
    "dram_bytes_read": ("dram__bytes_read.sum", None),
    "dram_bytes_write": ("dram__bytes_write.sum", None),
    "local_load_sectors": ("l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum", 1.0),
    "local_store_sectors": ("l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum", 1.0),
}
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def main():
    rep, what = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    res = {"report": rep, "what": what, "launches": []}
    for r in rows[2:]:
        d = {"kernel": r[idx["Kernel Name"]]}
        for k, (m, scale) in KEYS.items():
            if m not in idx:
                continue
            try:
                v = float(r[idx[m]].replace(",", ""))
            except ValueError:
                continue
            if scale is None:
                v *= UNIT.get(units[idx[m]], 1.0)
            elif k == "gpu_time_ms":
                v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(units[idx[m]], 1.0)
            d[k] = v
        if "dram_bytes_read" in d and "gpu_time_ms" in d:
            d["dram_GBps"] = (d["dram_bytes_read"] + d.get("dram_bytes_write", 0.0)) / (d["gpu_time_ms"] * 1e-3) / 1e9
        res["launches"].append(d)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
