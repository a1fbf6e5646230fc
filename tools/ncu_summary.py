#!/usr/bin/env python
"""Summarise an .ncu-rep (ncu --set full) into the text kept under profiles/: per kernel launch the duration, DRAM
bytes, L2 / DRAM / tensor-pipe utilisation, occupancy and the top warp-stall reasons.
    python tools/ncu_summary.py gpurun_out/r1_conv.ncu-rep > profiles/r1_conv_ncu.txt"""
import csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}
want = [("gpu__time_duration.sum", "duration"), ("sm__cycles_elapsed.avg.per_second", "sm clock"),
        ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "regs/thread"),
        ("launch__shared_mem_per_block_dynamic", "dyn smem/block"),
        ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram write"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram throughput %"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"),
        ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
        ("l1tex__throughput.avg.pct_of_peak_sustained_active", "L1/TEX throughput %"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active % (of active cycles)"),
        ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma pipe %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
        ("smsp__cycles_active.avg", "smsp active cycles"), ("sm__cycles_elapsed.avg", "elapsed cycles")]
stalls = [h for h in hdr if h.startswith("smsp__average_warp") and "issue_stalled" in h and h.endswith("_per_warp_active.pct")] or \
         [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith(".ratio")]
print(f"# ncu --set full --clock-control none summary of {rep.split('/')[-1]} (one block per profiled launch)")
for r in rows[2:]:
    name = r[col["Kernel Name"]]
    print("\n== " + name[:160])
    for key, label in want:
        if key in col:
            print(f"   {label:42s} {r[col[key]]:>16s} {units[col[key]]}")
    if "dram__bytes_read.sum" in col:
        def tob(v, u):
            v = float(v.replace(",", ""))
            return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
        t = tob(r[col["dram__bytes_read.sum"]], units[col["dram__bytes_read.sum"]]) + tob(r[col["dram__bytes_write.sum"]], units[col["dram__bytes_write.sum"]])
        print(f"   {'dram traffic (read + write)':42s} {t / 1e6:16.3f} MB")
    st = []
    for h in stalls:
        try:
            st.append((float(r[col[h]].replace(",", "")), h))
        except ValueError:
            pass
    st.sort(reverse=True)
    for v, h in st[:5]:
        print(f"   stall {h.replace('smsp__average_warps_issue_stalled_', '').replace('smsp__average_warp_latency_issue_stalled_', '')[:50]:44s} {v:10.3f}")
