"""Summarise an ncu report (`ncu --set full`) as a small table: one row per captured launch.
usage: python tools/ncu_summary.py report.ncu-rep > profiles/xxx.md"""
import csv, io, subprocess, sys

METRICS = [
    ("time_us", "gpu__time_duration.sum"),
    ("dram_read_MB", "dram__bytes_read.sum"),
    ("dram_write_MB", "dram__bytes_write.sum"),
    ("dram_pct", "dram__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("sm_pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("issue_pct", "sm__issue_active.avg.pct_of_peak_sustained_elapsed"),
    ("alu_pct", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
    ("fma_pct", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
    ("lsu_pct", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
    ("xu_pct", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
    ("tensor_pct", "sm__pipe_tensor_subpipe_op_cycles_active.avg.pct_of_peak_sustained_active"),
    ("warps_active_pct", "sm__warps_active.avg.pct_of_peak_sustained_active"),
    ("regs", "launch__registers_per_thread"),
    ("grid", "launch__grid_size"),
    ("block", "launch__block_size"),
    ("smem_dyn_KB", "launch__shared_mem_per_block_dynamic"),
    ("cycles", "gpc__cycles_elapsed.max"),
]
UNIT_SCALE = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}


def main(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    col = {name: i for i, name in enumerate(hdr)}
    tensor_alt = [n for n in hdr if "pipe_tensor" in n and n.endswith("pct_of_peak_sustained_active")]
    print(f"ncu --set full summary of `{path.split('/')[-1]}` (one row per captured launch; clocks not locked)\n")
    names = [m[0] for m in METRICS]
    print("| kernel | " + " | ".join(names) + " |")
    print("|---|" + "---|" * len(names))
    for r in rows[2:]:
        if len(r) < len(hdr):
            continue
        k = r[col["Kernel Name"]].split("(")[0]
        vals = []
        for short, name in METRICS:
            if name not in col and short == "tensor_pct" and tensor_alt:
                name = tensor_alt[0]
            if name not in col:
                vals.append("n/a"); continue
            v, u = r[col[name]].replace(",", ""), units[col[name]]
            try:
                x = float(v)
            except ValueError:
                vals.append(v); continue
            if short in ("time_us", "dram_read_MB", "dram_write_MB"):
                x *= UNIT_SCALE.get(u, 1.0)
            if short == "smem_dyn_KB":
                x *= {"byte": 1 / 1024, "Kbyte": 1.0, "Mbyte": 1024.0}.get(u, 1.0)
            vals.append(f"{x:.1f}" if abs(x) < 1e6 and x != int(x) else f"{int(x)}")
        print(f"| {k} | " + " | ".join(vals) + " |")


if __name__ == "__main__":
    main(sys.argv[1])
