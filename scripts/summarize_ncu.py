"""Turn an .ncu-rep into the small text/JSON summaries committed under profiles/.
usage: python scripts/summarize_ncu.py <report.ncu-rep> <out_prefix> [workload_key]"""
import csv, json, subprocess, sys, io

KEYS = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__cluster_size", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio"]

def main():
    rep, out = sys.argv[1], sys.argv[2]
    key = sys.argv[3] if len(sys.argv) > 3 else None
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    lines, traffic = [], {}
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")].split("(")[0]
        lines.append(f"== {name}")
        vals = {}
        for i, h in enumerate(hdr):
            if h in KEYS or h.startswith("smsp__pcsamp_warps_issue_stalled_") and "not_issued" not in h:
                if r[i] not in ("", "n/a"):
                    lines.append(f"  {h:85s} {r[i]:>18s} {units[i]}")
                    vals[h] = (r[i], units[i])
        def to_bytes(k):
            v, u = vals.get(k, ("0", "byte"))
            f = float(v.replace(",", ""))
            return f * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(u, 1)
        if key and ("tc_search" in name or "tail" in name):
            traffic[key + ("_tail" if "tail" in name else "")] = {"dram_bytes_per_launch": to_bytes("dram__bytes_read.sum") + to_bytes("dram__bytes_write.sum"),
                            "dram_read": to_bytes("dram__bytes_read.sum"), "dram_write": to_bytes("dram__bytes_write.sum"),
                            "source": "ncu --set full --clock-control none, one launch, " + rep.split("/")[-1]}
    open(out + ".txt", "w").write("\n".join(lines) + "\n")
    if traffic:
        path = "profiles/tc_search_traffic.json"
        try:
            cur = json.load(open(path))
        except Exception:
            cur = {}
        cur.update(traffic)
        json.dump(cur, open(path, "w"), indent=1)
    print("\n".join(lines[:60]))

if __name__ == "__main__":
    main()
