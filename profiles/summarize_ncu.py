"""Compact markdown summary of an `ncu --set full` report (per kernel launch).

    python profiles/summarize_ncu.py gpurun_out/report.ncu-rep > profiles/rNN_name.md
"""
import csv
import subprocess
import sys

METRICS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "FMA pipe cycles active %"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe %"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU (MUFU) pipe %"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe %"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe cycles active % (of active cycles)"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor pipe cycles active % (of elapsed)"),
    ("sm__ops_path_tensor_op_utchmma_src_tf32_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed", "UTCHMMA tf32 ops % of peak"),
    ("smsp__sass_inst_executed_op_utcmma.sum", "tcgen05.mma instructions"),
    ("smsp__sass_inst_executed_op_tmem_ldt.sum", "tcgen05.ld instructions"),
    ("smsp__sass_inst_executed_op_tmem_stt.sum", "tcgen05.st instructions"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "shared-memory wavefronts"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "shared-memory bank conflicts"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem / CTA"),
]

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
kn, gs, bs = hdr.index("Kernel Name"), hdr.index("Grid Size"), hdr.index("Block Size")
print(f"# ncu --set full --clock-control none: `{rep.split('/')[-1]}` (one row block per captured launch; cold caches, serialised)\n")
for r in rows[2:]:
    name = r[kn].split("(")[0].replace("void ", "")
    print(f"## `{name}`  grid {r[gs]} block {r[bs]}\n")
    print("| metric | value |")
    print("|---|---|")
    for m, label in METRICS:
        if m in hdr:
            i = hdr.index(m)
            v = r[i]
            if v in ("", "0", "0.000000") and ("tensor" in m or "utc" in m or "tmem" in m):
                continue
            try:
                v = f"{float(v.replace(',', '')):,.2f}".rstrip("0").rstrip(".")
            except ValueError:
                pass
            print(f"| {label} (`{m}`) | {v} {units[i]} |")
    print()
