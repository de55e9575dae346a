"""Print the handful of ncu metrics that decide what bounds a kernel: python tools/ncu_brief.py report.ncu-rep"""
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_reads.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_writes.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum",
        "l1tex__m_l1tex2xbar_write_bytes.sum", "l1tex__m_l1tex2xbar_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__cycles_active.avg", "sm__cycles_elapsed.max"]


def main():
    out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h = rows[0]
    for r in rows[2:]:
        d = dict(zip(h, r))
        print("==", d.get("Kernel Name", "")[:100])
        for k in KEYS:
            if k in d:
                print("  %-80s %s" % (k, d[k]))
        for k in h:
            if k.startswith("smsp__pcsamp_warps_issue_stalled") and "not_issued" not in k and d.get(k, "0") not in ("0", ""):
                print("  %-80s %s" % (k, d[k]))


if __name__ == "__main__":
    main()
