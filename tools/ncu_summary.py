"""Summarise an ncu report of the stiffness kernel into a markdown table (run here, no GPU needed):
    python tools/ncu_summary.py gpurun_out/r2n_full.ncu-rep [label] >> profiles/r2_ncu_stiff_brick_kernel.md
Also prints the per-instruction shared-memory wavefront excess and the top stall reasons from the
source page, and writes profiles/traffic.json (DRAM bytes per launch) when --traffic is given."""
import csv
import io
import json
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__shared_mem_config_size", "launch__grid_size", "launch__block_size",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__throughput.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__average_warp_latency_per_inst_issued.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "lts__t_sector_hit_rate.pct",
]


def page(rep, name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep = sys.argv[1]
    label = sys.argv[2] if len(sys.argv) > 2 and not sys.argv[2].startswith("--") else rep
    raw = page(rep, "raw")
    hdr, units, vals = raw[0], raw[1], raw[2]
    got = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
    print(f"### {label}\n")
    print(f"kernel: `{got.get('Kernel Name', ('?', ''))[0]}`\n")
    print("| metric | value |\n|---|---|")
    for m in METRICS:
        if m in got:
            print(f"| `{m}` | {got[m][0]} {got[m][1]} |")
    src = page(rep, "source")
    h2 = src[1]
    ci = {h: i for i, h in enumerate(h2)}
    data = src[2:]
    stalls = [h for h in h2 if h.startswith("stall_") and "Not Issued" not in h]
    agg = {s: sum(int(r[ci[s]] or 0) for r in data) for s in stalls}
    tot = sum(agg.values())
    top = sorted(agg.items(), key=lambda kv: -kv[1])[:6]
    print("\nwarp-state samples (source page): " + ", ".join(f"{k[6:]} {100 * v / tot:.0f} %" for k, v in top))
    w = sum(int(r[ci["L1 Wavefronts Shared"]] or 0) for r in data)
    wi = sum(int(r[ci["L1 Wavefronts Shared Ideal"]] or 0) for r in data)
    print(f"\nshared-memory wavefronts {w} (ideal {wi}, excess {w - wi}) per launch")
    ops = {}
    for r in data:
        op = r[1].split()[0] if not r[1].strip().startswith("@") else r[1].split()[1]
        op = op.split(".")[0]
        ops[op] = ops.get(op, 0) + 1
    keys = ["LDGSTS", "UBLKPF", "UBLKCP", "SYNCS", "LDG", "STG", "LDS", "STS", "DFMA", "ATOM", "RED", "BAR", "LDL", "STL"]
    print("\nSASS instruction census of the captured kernel: " + ", ".join(f"{k} {ops.get(k, 0)}" for k in keys))
    if "--traffic" in sys.argv:
        rd = float(got["dram__bytes_read.sum"][0]) * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}[got["dram__bytes_read.sum"][1]]
        wr = float(got["dram__bytes_write.sum"][0]) * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}[got["dram__bytes_write.sum"][1]]
        json.dump({"dram_bytes_per_launch": rd + wr, "read": rd, "write": wr, "report": rep,
                   "note": "dram__bytes_read.sum + dram__bytes_write.sum of one colour launch (512 CTAs) of stiff_brick_kernel, ncu --set full"},
                  open("profiles/traffic.json", "w"), indent=1)


if __name__ == "__main__":
    main()
