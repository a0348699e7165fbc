"""Reduce an `ncu --set full` report to the metrics the roofline discussion uses.
   python scripts/ncu_summary.py gpurun_out/r2_prof.ncu-rep profiles/r2_ncu_raw.csv profiles/r2_ncu_full.md
Writes a CSV (header row, unit row, one row per launch; read by bench.py for `roofline.traffic`) and a markdown table."""
import csv
import subprocess
import sys

KEEP = ["ID", "Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__inst_executed.sum", "launch__cluster_size"]


def main():
    rep, out_csv, out_md = sys.argv[1:4]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr = rows[0]
    idx = [hdr.index(k) for k in KEEP if k in hdr]
    names = [hdr[i] for i in idx]
    table = [[r[i] for i in idx] for r in rows[1:]]
    with open(out_csv, "w", newline="") as fh:
        w = csv.writer(fh)
        w.writerow(names)
        w.writerows(table)
    short = {"gpu__time_duration.sum": "time", "dram__bytes_read.sum": "dram rd", "dram__bytes_write.sum": "dram wr",
             "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram %", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor % (active)",
             "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed": "tensor % (elapsed)", "launch__registers_per_thread": "regs",
             "sm__warps_active.avg.pct_of_peak_sustained_active": "warps %", "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm %",
             "lts__t_sector_hit_rate.pct": "L2 hit %", "lts__t_bytes.sum": "L2 bytes", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum": "smem wavefronts",
             "smsp__inst_executed.sum": "warp insts", "launch__cluster_size": "cluster"}
    with open(out_md, "a") as fh:
        fh.write("| " + " | ".join(short.get(n, n) for n in names) + " |\n")
        fh.write("|" + "---|" * len(names) + "\n")
        for r in table:
            cells = []
            for n, v in zip(names, r):
                if n == "Kernel Name":
                    v = v.split("(")[0].replace("myr::", "")
                else:
                    try:
                        v = ("%.2f" % float(v)) if "." in v else v
                    except ValueError:
                        pass
                cells.append(v)
            fh.write("| " + " | ".join(cells) + " |\n")


if __name__ == "__main__":
    main()
