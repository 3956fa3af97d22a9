"""Summarise an .ncu-rep (ncu --set full) into the few numbers DESIGN.md / bench.py quote: per launch duration, DRAM bytes
read / written, L2 -> SM bytes, tensor-pipe and DRAM utilisation, achieved occupancy, top stall reason.

    python tools/ncu_summary.py gpurun_out/r2_prof_nt.ncu-rep > profiles/r2_ncu_nt_summary.txt
"""
import csv
import io
import subprocess
import sys

WANT = [("gpu__time_duration.sum", "duration"),
        ("dram__bytes_read.sum", "dram_read"), ("dram__bytes_write.sum", "dram_write"),
        ("lts__t_bytes.sum", "l2_bytes"), ("lts__t_sectors_srcunit_tex.sum", "l2_sectors_from_sm"),
        ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
        ("sm__inst_executed_pipe_tensor.sum", "tensor_inst"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pipe_pct"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct"),
        ("sm__warps_active.avg.per_cycle_active", "warps_active"),
        ("smsp__cycles_active.avg", "cycles_active"),
        ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__cluster_dim_x", "cluster_x"),
        ("launch__registers_per_thread", "regs"), ("launch__shared_mem_per_block_dynamic", "smem_dyn")]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[0]
    units = rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    print("# %s" % path)
    for r in rows[2:]:
        if len(r) < len(hdr):
            continue
        name = r[idx["Kernel Name"]][:90]
        print("\n[%s] %s  grid %s block %s" % (r[idx["ID"]], name, r[idx.get("Grid Size", 0)], r[idx.get("Block Size", 0)]))
        for key, label in WANT:
            if key in idx:
                print("    %-22s %s %s" % (label, r[idx[key]], units[idx[key]]))
        stalls = [(h, r[i]) for h, i in idx.items() if h.startswith("smsp__average_warp_latency_issue_stalled") or
                  h.startswith("smsp__average_warps_issue_stalled")]
        try:
            stalls = sorted(((h, float(v.replace(",", ""))) for h, v in stalls if v), key=lambda t: -t[1])[:3]
            for h, v in stalls:
                print("    stall %-60s %.2f" % (h.replace("smsp__average_warps_issue_stalled_", "").replace("smsp__average_warp_latency_issue_stalled_", ""), v))
        except ValueError:
            pass


if __name__ == "__main__":
    main(sys.argv[1])
