"""Key counters of one kernel of an ncu report: ncu -i X.ncu-rep --page raw --csv | python tools/ncu_summary.py"""
import csv, sys
rows = list(csv.reader(sys.stdin))
hdr = rows[0]
for vals in rows[2:]:
    d = dict(zip(hdr, vals))
    print("kernel:", d.get("Kernel Name", "?")[:90])
    for k in ("gpu__time_duration.sum", "sm__cycles_active.avg", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
              "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
              "launch__registers_per_thread", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
              "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "lts__t_sectors_srcunit_tex_op_read.sum",
              "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
              "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed"):
        if k in d:
            print(f"  {k} = {d[k]}")
    st = {k.replace("smsp__pcsamp_warps_issue_stalled_", ""): int(float(v)) for k, v in d.items()
          if k.startswith("smsp__pcsamp_warps_issue_stalled_") and not k.endswith("_not_issued") and v}
    tot = sum(st.values()) or 1
    print("  stall samples:", ", ".join(f"{k} {100 * v / tot:.1f}%" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:8]))
