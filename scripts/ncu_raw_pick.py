"""Picks the metrics the round's summaries quote out of an `ncu -i <rep> --page raw --csv` dump."""
import csv, sys
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__cycles_elapsed.avg", "sm__cycles_elapsed.max",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_tensor.sum", "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_tc_wavefronts_mem_shared.sum",
        "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "launch__registers_per_thread", "launch__grid_size", "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__cycles_active.avg", "launch__shared_mem_per_block_dynamic",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "lts__t_sectors_srcunit_tex_op_read.sum", "sm__warps_active.avg.pct_of_peak_sustained_active"]
rows = list(csv.reader(l for l in open(sys.argv[1]) if not l.startswith("==")))
hdr, units, vals = rows[0], rows[1], rows[2]
pat = sys.argv[2] if len(sys.argv) > 2 else None
for h, u, v in zip(hdr, units, vals):
    if (pat and pat in h) or (not pat and (h in WANT or h == "Kernel Name")):
        print(f"{h} [{u}] = {v}")
