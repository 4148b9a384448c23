"""Print the handful of ncu raw-page metrics we read when tuning (usage: ncu_keys.py raw.csv)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
want = ["Kernel Name", "gpu__time_duration.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts.avg",
        "l1tex__data_pipe_lsu_wavefronts_mem_lgds.avg", "l1tex__data_pipe_lsu_wavefronts_mem_shared.avg",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second", "launch__grid_size", "launch__registers_per_thread",
        "smsp__cycles_active.avg", "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__m_l1tex2xbar_write_bytes.sum",
        "lts__t_sectors_srcunit_tex_op_red.sum", "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active"]
for i, h in enumerate(hdr):
    name = h.split(".Triage")[0] if False else h
    if any(name == w or name.endswith("." + w) for w in want):
        print(f"{h} [{units[i]}]:", [r[i] for r in rows[2:]])
