"""Summarise an .ncu-rep (one kernel launch): headline metrics, stall reasons per issue, top stall sites.
    python tools/ncu_summary.py gpurun_out/x.ncu-rep [out.json]"""
import csv
import io
import json
import subprocess
import sys


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return dict(zip(rows[0], rows[2])), dict(zip(rows[0], rows[1]))


def source(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[1]
    return hdr, rows[2:]


KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"]


def main():
    rep = sys.argv[1]
    vals, units = raw(rep)
    out = {"report": rep, "metrics": {}, "stall_cycles_per_issue": {}}
    for k in KEYS:
        if k in vals:
            out["metrics"][k] = {"value": vals[k], "unit": units[k]}
    for k, v in vals.items():
        if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio") and "not_issued" not in k:
            try:
                if float(v) >= 0.2:
                    out["stall_cycles_per_issue"][k[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]] = float(v)
            except ValueError:
                pass
    hdr, data = source(rep)
    ix = {h: i for i, h in enumerate(hdr)}
    ts = sum(int(r[ix["# Samples"]]) for r in data) or 1
    ti = sum(int(r[ix["Instructions Executed"]]) for r in data) or 1
    top = sorted(range(len(data)), key=lambda i: -int(data[i][ix["# Samples"]]))[:14]
    out["top_stall_sites"] = [{"sass": data[i][1].strip()[:70], "pct_samples": round(100.0 * int(data[i][ix["# Samples"]]) / ts, 1),
                               "main_reason": max(((k[6:], int(data[i][ix[k]])) for k in hdr if k.startswith("stall_") and "Not Issued" not in k),
                                                  key=lambda kv: kv[1])[0]} for i in sorted(top)]
    out["warp_instructions"] = ti
    js = json.dumps(out, indent=1)
    if len(sys.argv) > 2:
        open(sys.argv[2], "w").write(js)
    print(js)


if __name__ == "__main__":
    main()
