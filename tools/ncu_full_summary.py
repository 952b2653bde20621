"""Selected `ncu --set full` metrics of every kernel in a report, one column per kernel (evidence tool, no GPU needed).
    python tools/ncu_full_summary.py gpurun_out/prof.ncu-rep > profiles/rN_ncu_full_hot_kernels.txt"""
import csv
import io
import re
import subprocess
import sys

KEEP = re.compile(
    r"^(dram__bytes_(read|write)\.sum(\.per_second|\.pct_of_peak_sustained_elapsed)?|gpu__dram_throughput\.avg\.pct_of_peak_sustained_elapsed|"
    r"gpu__time_duration\.sum|l1tex__data_bank_conflicts_pipe_lsu_mem_shared\.sum|launch__(block_size|grid_size|cluster_size|registers_per_thread)|"
    r"lts__t_sectors_srcunit_tex_op_read\.sum(\.per_second)?|lts__t_sectors_srcunit_tex_op_write\.sum|lts__throughput\.avg\.pct_of_peak_sustained_elapsed|"
    r"sm__cycles_elapsed\.max(\.per_second)?|sm__pipe_tensor_cycles_active\.avg\.pct_of_peak_sustained_elapsed|"
    r"sm__throughput\.avg\.pct_of_peak_sustained_elapsed|sm__warps_active\.avg\.pct_of_peak_sustained_active|smsp__inst_executed\.sum|"
    r"smsp__issue_active\.avg\.pct_of_peak_sustained_active|sm__inst_executed_pipe_xu\.avg\.pct_of_peak_sustained_active|"
    r"sm__pipe_fma_cycles_active\.avg\.pct_of_peak_sustained_active|sm__pipe_alu_cycles_active\.avg\.pct_of_peak_sustained_active)$")

out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units, data = rows[0], rows[1], rows[2:]
ik = hdr.index("Kernel Name")
names = [re.sub(r"\(.*", "", r[ik]).replace("void ", "").replace("tsasr::", "") for r in data]
print("columns: " + " | ".join(names))
for i, h in enumerate(hdr):
    if KEEP.match(h):
        print(f"{h:84s} {units[i]:12s} " + " | ".join(r[i] for r in data))
