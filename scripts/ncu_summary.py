"""Summarise an .ncu-rep (read here, no GPU needed) into profiles/<tag>_summary.{json,md} and, when a third argument names the bench key
(k_project_cast_hist_bytes_per_launch / k_hist_u8_cols_bytes_per_launch), record the launch's DRAM bytes in profiles/traffic.json."""
import csv, io, json, subprocess, sys
from pathlib import Path

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes.sum.per_second",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.per_cycle_active",
    "smsp__warps_active.avg.per_cycle_active", "smsp__warps_eligible.avg.per_cycle_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor", "smsp__inst_executed.sum",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
]

def main(rep, tag, kernel_key):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    summ = []
    for vals in rows[2:]:
        d = dict(zip(hdr, vals))
        u = dict(zip(hdr, units))
        rec = {"kernel": d.get("Kernel Name"), "grid": d.get("Grid Size"), "block": d.get("Block Size")}
        for k in KEYS:
            if k in d and d[k] != "":
                rec[k] = {"value": d[k], "unit": u[k]}
        stalls = {h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""): float(d[h])
                  for h in hdr if "issue_stalled" in h and h.endswith("per_issue_active.ratio") and d[h] not in ("", "n/a")}
        rec["stall_cycles_per_issue"] = dict(sorted(stalls.items(), key=lambda kv: -kv[1])[:8])
        summ.append(rec)
    P = Path("profiles"); P.mkdir(exist_ok=True)
    (P / f"{tag}_summary.json").write_text(json.dumps(summ, indent=1))
    def scale(v, u):
        v = float(v)
        return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "Tbyte": 1e12}[u]
    r = summ[0]
    rd = scale(r["dram__bytes_read.sum"]["value"], r["dram__bytes_read.sum"]["unit"])
    wr = scale(r["dram__bytes_write.sum"]["value"], r["dram__bytes_write.sum"]["unit"])
    if kernel_key:                       # only when the caller names the bench key this capture is the full-size launch of
        tr_path = P / "traffic.json"
        tr = json.loads(tr_path.read_text()) if tr_path.exists() else {}
        tr[kernel_key] = rd + wr
        tr[kernel_key + "_source"] = f"profiles/{tag}_summary.json (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum, one launch)"
        tr_path.write_text(json.dumps(tr, indent=1))
    lines = [f"# ncu --set full summary: {tag}", "", f"kernel: `{r['kernel']}`  grid {r['grid']} block {r['block']}", ""]
    for k in KEYS:
        if k in r:
            lines.append(f"- {k}: {r[k]['value']} {r[k]['unit']}")
    lines.append(f"- DRAM traffic per launch: {(rd + wr) / 1e9:.3f} GB (read {rd / 1e9:.3f} + write {wr / 1e9:.3f})")
    lines.append("- top stall reasons (warp-cycles per issued instruction): " +
                 ", ".join(f"{k} {v:.2f}" for k, v in r["stall_cycles_per_issue"].items()))
    (P / f"{tag}_summary.md").write_text("\n".join(lines) + "\n")
    print("\n".join(lines))

if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else None)
