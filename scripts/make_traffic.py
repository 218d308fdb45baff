"""profiles/traffic.json from the ncu capture of the bench's own 16 Mi-ray launch (scripts/gpu_full.sh, last step).
usage: make_traffic.py <tag>"""
import csv, json, sys
tag = sys.argv[1]
rows = [r for r in csv.reader(open(f"gpurun_out/{tag}_traffic.csv")) if len(r) > 14 and r[0].isdigit()]
d = {r[12]: float(r[14]) for r in rows}
out = {"kernel": rows[0][4], "grid": rows[0][8], "block": rows[0][7],
       "dram_bytes_read": d["dram__bytes_read.sum"], "dram_bytes_write": d["dram__bytes_write.sum"],
       "dram_bytes_per_launch": d["dram__bytes_read.sum"] + d["dram__bytes_write.sum"],
       "gpu_time_ns_under_ncu": d["gpu__time_duration.sum"], "traces_per_launch": 1 << 24,
       "command": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none "
                  "-k regex:tracePool -s 3 -c 1 python bench.py --steps 1 --warmup 3 --no-cpu-baseline",
       "source": f"profiles/{tag}_traffic.csv"}
json.dump(out, open("profiles/traffic.json", "w"), indent=1)
print(out)
