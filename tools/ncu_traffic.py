"""Extract the persistent decoder's DRAM traffic and headline counters from an `ncu --set full` report (raw page CSV).
usage: ncu -i rep.ncu-rep --page raw --csv > raw.csv ; python tools/ncu_traffic.py raw.csv profiles/r1_ncu_traffic.json profiles/r1_ncu_full_decoder_bf16.csv"""
import csv
import json
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units, vals = rows[0], rows[1], rows[2]
d = {h: (u, v) for h, u, v in zip(hdr, units, vals)}


def num(k):
    u, v = d[k]
    x = float(v.replace(",", ""))
    return x * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}.get(u, 1.0)


import hashlib
import os
import subprocess
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
try:
    commit = subprocess.check_output(["git", "-C", root, "rev-parse", "HEAD"], text=True).strip()
except Exception:
    commit = os.environ.get("GSTK_COMMIT", "unknown")   # the GPU box has no .git: pass the hash in
ksrc = os.path.join(root, "gst_tacotron_b200", "csrc", "decoder_bf16.cuh")
out = {"commit": commit, "kernel_source_sha1": hashlib.sha1(open(ksrc, "rb").read()).hexdigest(),
       "dram_bytes_per_launch": num("dram__bytes_read.sum") + num("dram__bytes_write.sum"),
       "dram_bytes_read": num("dram__bytes_read.sum"), "dram_bytes_write": num("dram__bytes_write.sum"),
       "l2_to_sm_read_bytes_per_launch": num("lts__t_sectors_srcunit_tex_op_read.sum") * 32.0 if "lts__t_sectors_srcunit_tex_op_read.sum" in d else None,
       "kernel": d["Kernel Name"][1], "gpu_time_ms_under_ncu": float(d["gpu__time_duration.sum"][1].replace(",", "")),
       "source": "ncu --set full --clock-control none -k regex:decoder_bf16_kernel, bench.py workload (B=256, T_v=150, 1000 steps)"}
json.dump(out, open(sys.argv[2], "w"), indent=1)
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "lts__t_sectors_srcunit_tex_op_read.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "sm__cycles_active.avg"]
with open(sys.argv[3], "w") as f:
    f.write("metric,unit,value\n")
    for k in want:
        if k in d:
            f.write("{},{},{}\n".format(k, d[k][0], d[k][1]))
print(json.dumps(out))
