"""Measured DRAM bytes / tensor-pipe activity of the kernels bench.py reports, for THIS build -> profiles/r2_traffic.json
(bench.py's `roofline.traffic`, `*.dram_bytes_measured`; stamped with the digest of the CUDA sources so that a number taken
from another build is never reported).  Run on the GPU box:

    python tools/make_traffic_stamp.py          # runs `ncu ... python bench.py --steps 1 --warmup 3 --no-configs --no-cpu-baseline`

Also writes the launch list of that bench run (per-launch gpu__time_duration) to gpurun_out/r2_launches_bench.csv.
"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rrnco_b200.build import build_digest  # noqa: E402

OUT = os.path.join(ROOT, "gpurun_out")
os.makedirs(OUT, exist_ok=True)
log = os.path.join(OUT, "r2_launches_bench.csv")
metrics = "gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"
cmd = ["ncu", "--metrics", metrics, "--clock-control", "none", "-c", "4000", "--csv", "--log-file", log,
       sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "3", "--no-configs", "--no-cpu-baseline"]
subprocess.run(cmd, check=True, cwd=ROOT, stdout=subprocess.DEVNULL)

rows = list(csv.DictReader(l for l in open(log) if not l.startswith("==")))
per = {}
for r in rows:  # one row per (launch id, metric)
    d = per.setdefault(r["ID"], {"name": r["Kernel Name"]})
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"].lower()
    scale = {"kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "byte": 1.0, "usecond": 1e-3, "msecond": 1.0, "nsecond": 1e-6, "second": 1e3, "us": 1e-3, "ms": 1.0, "ns": 1e-6, "s": 1e3}.get(unit, 1.0)
    d[r["Metric Name"]] = v * scale
launches = [per[k] for k in sorted(per, key=int)]


def last(sub):
    c = [l for l in launches if sub in l["name"]]
    return c[-1] if c else None


stamp = {"build_digest": build_digest(), "source": "tools/make_traffic_stamp.py: ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum "
         "--clock-control none over one bench.py run, last launch of each kernel"}
for key, sub in (("rollout_c2_dram_bytes", "rollout_lean_kernel"), ("rcvrp_step_dram_bytes", "rcvrp_step_vec_kernel"),
                 ("atsp_step_dram_bytes", "atsp_step_vec_kernel"), ("rcvrptw_step_dram_bytes", "rmtvrp_step_kernel"),
                 ("gather_dram_bytes", "gather_submatrix_kernel<float"), ("nab_dram_bytes", "nab_gating_table_kernel")):
    l = last(sub)
    if l is None:
        continue
    stamp[key] = l["dram__bytes_read.sum"] + l["dram__bytes_write.sum"]
    stamp[key.replace("_dram_bytes", "_ms_under_ncu")] = l["gpu__time_duration.sum"]
l = last("rollout_lean_kernel")
if l is not None:
    stamp["rollout_c2_source"] = "ncu dram__bytes_read.sum + dram__bytes_write.sum of one full-size launch (8192 tiles), this build"
    stamp["rollout_tensor_pipe_active_pct"] = l["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
tot = sum(l["gpu__time_duration.sum"] for l in launches)
share = {}
for l in launches:
    n = l["name"].split("(")[0][-60:]
    share[n] = share.get(n, 0.0) + l["gpu__time_duration.sum"]
stamp["launch_share_pct"] = {k: round(100 * v / tot, 2) for k, v in sorted(share.items(), key=lambda kv: -kv[1])[:8]}
for d in (os.path.join(ROOT, "profiles"), OUT):  # (gpurun brings back gpurun_out/ only: copy it into profiles/ afterwards)
    json.dump(stamp, open(os.path.join(d, "r2_traffic.json"), "w"), indent=1)
print(json.dumps(stamp, indent=1))
