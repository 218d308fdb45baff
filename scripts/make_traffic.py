"""profiles/traffic.json from the ncu capture of the bench's own 16 Mi-ray launch (scripts/gpu_full.sh, last step):
the two kernels of one trace launch -- the walk (tracePoolKernel) and the record expansion (expandRecordsKernel).
usage: make_traffic.py <tag>"""
import csv, json, sys
tag = sys.argv[1]
rows = [r for r in csv.reader(open(f"gpurun_out/{tag}_traffic.csv")) if len(r) > 14 and r[0].isdigit()]
per = {}
for r in rows:
    name = "tracePoolKernel" if "tracePool" in r[4] else "expandRecordsKernel" if "expandRecords" in r[4] else None
    if name is None or (name in per and r[12] in per[name]):
        continue                                   # other kernels, or a second launch of the same kernel
    per.setdefault(name, {"grid": r[8], "block": r[7]})[r[12]] = float(r[14])
tot_r = sum(k["dram__bytes_read.sum"] for k in per.values())
tot_w = sum(k["dram__bytes_write.sum"] for k in per.values())
out = {"kernels": per, "dram_bytes_read": tot_r, "dram_bytes_write": tot_w, "dram_bytes_per_launch": tot_r + tot_w,
       "gpu_time_ns_under_ncu": sum(k["gpu__time_duration.sum"] for k in per.values()), "traces_per_launch": 1 << 24,
       "command": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none "
                  "-k regex:'tracePool|expandRecords' -s 6 -c 2 python bench.py --steps 1 --warmup 3 --no-cpu-baseline",
       "source": f"profiles/{tag}_traffic.csv"}
# warp execution efficiency and issue utilisation of the walk kernel from the same set's ncu --set full capture
try:
    import datetime
    k = json.load(open(f"profiles/{tag}_trace_kernel_ncu.json"))["kernels"][0]
    out["lanes_per_instruction"] = float(k["smsp__thread_inst_executed_per_inst_executed.ratio"]["value"])
    out["issue_active_pct"] = float(k["smsp__issue_active.avg.pct_of_peak_sustained_active"]["value"])
    out["full_source"] = f"profiles/{tag}_trace_kernel_ncu.json"
    out["captured"] = f"{datetime.date.today().isoformat()} (set {tag})"
except Exception as ex:
    print("no full capture for this set:", ex)
json.dump(out, open("profiles/traffic.json", "w"), indent=1)
print(out)
