"""Turn ncu reports (gpurun_out/*.ncu-rep) into the compact CSV summaries committed under profiles/.

    python profiles/summarize.py gpurun_out/r1_gemm_cfg6_8192x12288x4096.ncu-rep [...]

For every kernel in a report: duration, SM clock, tensor-pipe utilisation
(sm__mem_tensor_cycles_active: the metric that tracks tcgen05 MMA busy time on sm_100 -- it reads
87.7 % for the cuBLASLt INT8 kernel whose cycles/ideal ratio is 87.6 %), shared-memory / L2 / DRAM
traffic, registers, launch shape.
"""
import csv
import subprocess
import sys
from pathlib import Path

KEYS = [
    ("duration_us", "gpu__time_duration.sum"),
    ("sm_clock_ghz", "sm__cycles_elapsed.avg.per_second"),
    ("tensor_pipe_pct", "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"),
    ("sm_throughput_pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("issue_active_pct", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
    ("tc_smem_read_pct", "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"),
    ("tma_fill_pct", "l1tex__m_xbar2l1tex_read_bytes.sum.pct_of_peak_sustained_elapsed"),
    ("l2_throughput_pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("dram_read_mb", "dram__bytes_read.sum"),
    ("dram_write_mb", "dram__bytes_write.sum"),
    ("dram_read_gbs", "dram__bytes_read.sum.per_second"),
    ("dram_write_gbs", "dram__bytes_write.sum.per_second"),
    ("regs", "launch__registers_per_thread"),
    ("dyn_smem_kb", "launch__shared_mem_per_block_dynamic"),
    ("grid", "Grid Size"),
    ("block", "Block Size"),
    ("cluster_x", "launch__cluster_dim_x"),
]


def rows_of(rep: Path):
    out = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        d = dict(zip(hdr, vals))
        u = dict(zip(hdr, units))
        yield d, u


def main():
    w = csv.writer(sys.stdout)
    w.writerow(["report", "kernel"] + [k for k, _ in KEYS])
    for rep in sys.argv[1:]:
        for d, u in rows_of(Path(rep)):
            name = d.get("Kernel Name", "")[:90].replace("\n", " ")
            vals = []
            for k, m in KEYS:
                v = d.get(m, "")
                unit = u.get(m, "")
                try:
                    f = float(v)
                    if unit in ("byte", "Kbyte", "Mbyte", "Gbyte") and k.endswith("_mb"):
                        f *= {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}[unit]
                    if k == "duration_us" and unit in ("ns", "ms", "s"):
                        f *= {"ns": 1e-3, "ms": 1e3, "s": 1e6}[unit]
                    v = f"{f:.4g}"
                except ValueError:
                    pass
                vals.append(v)
            w.writerow([Path(rep).name, name] + vals)


if __name__ == "__main__":
    main()
