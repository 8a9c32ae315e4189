"""Turn an .ncu-rep (brought back in gpurun_out/) into the small CSV summaries committed under profiles/.
  python tools/ncu_to_profile.py single <rep> <out.csv> "<header comment>"      one kernel: metric,unit,value rows
  python tools/ncu_to_profile.py table  <rep> <out.csv> "<header comment>"      many kernels: one row per launch
Runs where ncu is installed (the build container has ncu but no GPU: it can read reports)."""
import csv, io, subprocess, sys

KEEP = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
        "dram__bytes_write.sum.per_second", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "lts__t_bytes.sum", "lts__t_sectors_srcunit_tex_op_read.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__cycles_active.avg", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__cluster_size", "launch__shared_mem_per_block_dynamic",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__cycles_active.avg", "sm__cycles_elapsed.avg.per_second", "dram__cycles_elapsed.avg.per_second")


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]


def main():
    mode, rep, dst, comment = sys.argv[1:5]
    hdr, units, rows = raw(rep)
    with open(dst, "w") as f:
        for ln in comment.split("\\n"):
            f.write("# " + ln + "\n")
        if mode == "single":
            f.write("metric,unit,value\n")
            r = rows[0]
            for i, h in enumerate(hdr):
                if h == "Kernel Name" or h in KEEP:
                    f.write(f'"{h}",{units[i]},{r[i]}\n')
        else:
            cols = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
                    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
                    "launch__grid_size"]
            idx = [hdr.index(c) for c in cols if c in hdr]
            f.write(",".join(f"{hdr[i]}[{units[i]}]" for i in idx) + "\n")
            for r in rows:
                f.write(",".join('"' + r[i] + '"' if "," in r[i] or "<" in r[i] else r[i] for i in idx) + "\n")


if __name__ == "__main__":
    main()
