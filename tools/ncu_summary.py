"""Key metrics of an `ncu --set full` report as text (for profiles/).  python tools/ncu_summary.py rep.ncu-rep > out.txt"""
import csv, subprocess, sys, io
KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.avg", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
for rep in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        d = dict(zip(hdr, zip(vals, units)))
        print(f"# {rep}: {d.get('Kernel Name', ('?',))[0][:110]}")
        for k in KEYS:
            if k in d:
                print(f"{k:75s} {d[k][0]:>16s} {d[k][1]}")
        st = sorted(((float(v[0]), k) for k, v in d.items() if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("_per_issue_active.ratio")), reverse=True)[:6]
        for v, k in st:
            print(f"{k:75s} {v:16.3f}")
        print()
