#!/bin/bash
# Runs on the GPU box after the captures: turns every gpurun_out/*.ncu-rep into small text summaries (the reports themselves
# exceed the 64 MiB that gpurun copies back) and deletes the reports.
cd gpurun_out || exit 1
M='gpu__time_duration.sum|dram__bytes_read.sum |dram__bytes_write.sum |dram__throughput.avg.pct|sm__throughput.avg.pct|sm__pipe_tc_cycles_active|sm__pipe_tensor|sm__inst_executed_pipe_xu|sm__inst_executed_pipe_fma|sm__inst_executed_pipe_alu|sm__inst_executed_pipe_lsu|smsp__issue_active.avg.pct|l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct|lts__throughput.avg.pct|launch__grid_size|launch__cluster|launch__registers_per_thread|smsp__inst_executed.sum |sm__cycles_elapsed.max|sm__warps_active.avg.pct|lts__t_bytes.sum |l1tex__t_bytes.sum '
for r in *.ncu-rep; do
  b=${r%.ncu-rep}
  ncu -i "$r" --page raw --csv > /tmp/$b.raw.csv 2>/dev/null
  python - "$b" <<'PY'
import csv, re, sys
b = sys.argv[1]
rows = list(csv.reader(open("/tmp/%s.raw.csv" % b)))
hdr = rows[0]
units = rows[1] if len(rows) > 1 else []
keep = re.compile(r"gpu__time_duration.sum|dram__bytes_(read|write).sum$|dram__throughput.avg.pct_of_peak_sustained_elapsed|sm__throughput.avg.pct|sm__pipe_tc_cycles_active|sm__inst_executed_pipe_(xu|fma|alu|lsu).*pct_of_peak_sustained_active|smsp__issue_active.avg.pct|l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct|lts__throughput.avg.pct|launch__(grid_size|cluster_size|registers_per_thread)|smsp__inst_executed.sum$|sm__cycles_elapsed.max|sm__warps_active.avg.pct|lts__t_bytes.sum$")
idx = [i for i, h in enumerate(hdr) if keep.search(h)]
ki = hdr.index("Kernel Name") if "Kernel Name" in hdr else 4
with open("%s_summary.txt" % b, "w") as f:
    for r in rows[2:]:
        if len(r) <= ki:
            continue
        f.write("== %s\n" % r[ki][:150])
        for i in idx:
            f.write("  %-80s %s %s\n" % (hdr[i], r[i], units[i] if i < len(units) else ""))
PY
done
# per-source-line stall picture of the fused edge kernel (first captured launch)
if [ -f r2_edge.ncu-rep ]; then
  ncu -i r2_edge.ncu-rep --page source --csv --print-source sass 2>/dev/null | head -c 30000000 > /tmp/edge_src.csv
  python - <<'PY'
import csv, collections
rows = list(csv.reader(open("/tmp/edge_src.csv", errors="ignore")))
# find header containing 'Source' and sampling columns
h = None
for i, r in enumerate(rows[:50]):
    if "Source" in r or "# Samples" in " ".join(r):
        h = i; break
if h is not None:
    hdr = rows[h]
    with open("r2_edge_source_top.txt", "w") as f:
        f.write(",".join(hdr) + "\n")
        def samples(r):
            for name in ("Warp Stall Sampling (All Samples)", "# Samples", "Warp Stall Sampling (All Cycles)"):
                if name in hdr:
                    try: return float(r[hdr.index(name)])
                    except Exception: return 0.0
            return 0.0
        body = [r for r in rows[h + 1:] if len(r) == len(hdr)]
        tot = sum(samples(r) for r in body) or 1.0
        f.write("total samples %.0f over %d instructions\n" % (tot, len(body)))
        for r in sorted(body, key=samples, reverse=True)[:150]:
            f.write("%6.2f%% | %s\n" % (100 * samples(r) / tot, " | ".join(x[:70] for x in r[:6])))
PY
fi
rm -f *.ncu-rep
du -sh .
