"""Summarise `ncu --set full` reports (gpurun_out/full_*.ncu-rep) into a markdown table for profiles/.

    python tools/ncu_summary.py gpurun_out/full_hist_c2.ncu-rep ... > profiles/r02_ncu_full.md

One row per captured launch: duration, DRAM bytes, achieved DRAM GB/s, issue-slot utilisation, LSU pipe utilisation,
occupancy, registers, and the three largest warp-stall reasons (stalled warps per issued instruction).
"""
import csv
import io
import subprocess
import sys

COLS = {
    "dur": "gpu__time_duration.sum",
    "rd": "dram__bytes_read.sum",
    "wr": "dram__bytes_write.sum",
    "issue": "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "lsu": "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "warps": "sm__warps_active.avg.pct_of_peak_sustained_active",
    "regs": "launch__registers_per_thread",
    "grid": "launch__grid_size",
    "block": "launch__block_size",
    "inst": "smsp__inst_executed.sum",
    "l2hit": "lts__t_sector_hit_rate.pct",
    "smem_wave": "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "smem_conf": "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
}
STALL = "smsp__average_warps_issue_stalled_%s_per_issue_active.ratio"
STALLS = ["long_scoreboard", "short_scoreboard", "mio_throttle", "lg_throttle", "barrier", "wait", "math_pipe_throttle",
          "dispatch_stall", "branch_resolving", "no_instruction", "membar", "sleeping", "drain", "imc_miss", "not_selected"]
SCALE = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def num(v, unit):
    try:
        x = float(v.replace(",", ""))
    except ValueError:
        return None
    return x * SCALE.get(unit, 1.0)


def main():
    print("| report | kernel | grid x block | regs | time us | DRAM rd+wr MB | DRAM GB/s | issue % | LSU pipe % | warps active % | L2 hit % | smem wavefronts (conflicts) | top stalls (warps per issue) |")
    print("|---|---|---|---|---|---|---|---|---|---|---|---|---|")
    for path in sys.argv[1:]:
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        if len(rows) < 3:
            continue
        hdr, units = rows[0], rows[1]
        idx = {h: i for i, h in enumerate(hdr)}
        for r in rows[2:]:
            g = {k: (num(r[idx[c]], units[idx[c]]) if c in idx else None) for k, c in COLS.items()}
            name = r[idx["Kernel Name"]].split("(")[0].replace("void ", "")
            st = []
            for s in STALLS:
                c = STALL % s
                if c in idx:
                    v = num(r[idx[c]], "")
                    if v is not None:
                        st.append((v, s))
            st.sort(reverse=True)
            mb = ((g["rd"] or 0) + (g["wr"] or 0)) / 1e6
            gbs = mb / 1e3 / (g["dur"] * 1e-6) if g["dur"] else 0.0
            print("| %s | %s | %d x %d | %d | %.1f | %.1f | %.0f | %.1f | %.1f | %.1f | %.1f | %.3g (%.3g) | %s |" % (
                path.split("/")[-1].replace(".ncu-rep", ""), name, g["grid"] or 0, g["block"] or 0, g["regs"] or 0, g["dur"] or 0, mb, gbs,
                g["issue"] or 0, g["lsu"] or 0, g["warps"] or 0, g["l2hit"] or 0, g["smem_wave"] or 0, g["smem_conf"] or 0,
                ", ".join("%s %.2f" % (s, v) for v, s in st[:3])))


if __name__ == "__main__":
    main()
