"""Print selected raw metrics of every launch in an ncu report: python tools/ncu_raw.py <rep> [substr ...]"""
import csv, subprocess, sys, io
rep = sys.argv[1]
extra = sys.argv[2:]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum", "l1tex__m_l1tex2xbar_write_bytes_mem_global_op_tma_st.sum",
        "l1tex__data_bank_conflicts_pipe_lsu.sum", "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "launch__registers_per_thread", "l1tex__tmain_requests.sum.pct_of_peak_sustained_elapsed"]
for r in rows[2:]:
    print("----", r[hdr.index("Kernel Name")][:60], "id", r[0])
    for i, h in enumerate(hdr):
        if h in want or any(e in h for e in extra):
            print(f"   {h:75s} {r[i]:>16s} {units[i]}")
