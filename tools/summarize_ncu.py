"""Summarises an ncu launch list (--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv) of
`bench.py --steps 1 --warmup 2`: isolates ONE steady-state training step (between two input-pack launches), and writes
per-kernel-class time shares and DRAM bytes.

    python tools/summarize_ncu.py gpurun_out/launches.csv profiles/r01b
"""
import csv, json, sys
from collections import OrderedDict

src, prefix = sys.argv[1], sys.argv[2]
rows = []
with open(src) as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.DictReader(lines)
for r in rd:
    rows.append(r)
# long format: one row per (launch ID, metric)
launches = OrderedDict()
for r in rows:
    i = int(r["ID"])
    d = launches.setdefault(i, {"name": r["Kernel Name"], "id": i})
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    m = r["Metric Name"]
    if m == "gpu__time_duration.sum":
        scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit.replace("second", "s").replace("nsecond", "ns"), None)
        if scale is None:
            scale = {"nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3, "second": 1e6}[unit]
        d["us"] = v * scale
    else:
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
        d[m] = v * scale
L = list(launches.values())
starts = [k for k, d in enumerate(L) if "input_to_nhwc32" in d["name"] or "input_u8_nhwc32" in d["name"]]
if len(starts) < 3:
    raise SystemExit("need at least 3 steps in the capture, found %d" % len(starts))
a, b = starts[-2], starts[-1]        # the last complete step of the run (steady state: after the warm-up steps)
step = L[a - 1:b - 1]                 # the batched weight-pack launch precedes the input repack


def cls(name):
    for key, c in (("conv3x3_wgrad", "conv_wgrad"), ("wgrad_reduce", "conv_wgrad"), ("conv3x3_igemm", "conv_fwd+dgrad"), ("conv3x3_c32", "conv_fwd+dgrad"), ("grad_router", "router"),
                   ("up_transpose", "up_transpose"), ("upsample2x", "upsample2x"), ("activate_", "activate"), ("optim_", "optimizer"), ("bn_bwd_apply", "bn_bwd_apply"), ("bn_apply", "bn_apply"),
                   ("bn_", "bn_small"), ("head_", "head"), ("loss_", "loss"), ("iou_", "loss"), ("pack_weights", "pack"), ("input_to", "input_pack"), ("input_u8", "input_pack"),
                   ("maxpool", "maxpool"), ("nccl", "nccl")):
        if key in name:
            return c
    return "torch (optimizer, loss glue)"


agg = OrderedDict()
for d in step:
    c = cls(d["name"])
    g = agg.setdefault(c, {"launches": 0, "us": 0.0, "dram_read": 0.0, "dram_write": 0.0})
    g["launches"] += 1
    g["us"] += d.get("us", 0.0)
    g["dram_read"] += d.get("dram__bytes_read.sum", 0.0)
    g["dram_write"] += d.get("dram__bytes_write.sum", 0.0)
tot_us = sum(g["us"] for g in agg.values())
tot_b = sum(g["dram_read"] + g["dram_write"] for g in agg.values())
with open(prefix + "_ncu_launch_shares.csv", "w") as f:
    f.write("class,launches,time_us,time_share,dram_read_MB,dram_write_MB\n")
    for c, g in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
        f.write("%s,%d,%.1f,%.4f,%.1f,%.1f\n" % (c, g["launches"], g["us"], g["us"] / tot_us, g["dram_read"] / 1e6, g["dram_write"] / 1e6))
    f.write("TOTAL,%d,%.1f,1.0,%.1f,%.1f\n" % (len(step), tot_us, sum(g["dram_read"] for g in agg.values()) / 1e6, sum(g["dram_write"] for g in agg.values()) / 1e6))
with open(prefix + "_ncu_launches_step.csv", "w") as f:
    f.write("id,kernel,time_us,dram_read_B,dram_write_B\n")
    for d in step:
        f.write("%d,\"%s\",%.2f,%.0f,%.0f\n" % (d["id"], d["name"][:80], d.get("us", 0.0), d.get("dram__bytes_read.sum", 0.0), d.get("dram__bytes_write.sum", 0.0)))
conv = sum(g["dram_read"] + g["dram_write"] for c, g in agg.items() if c.startswith("conv"))
json.dump({"batch": 256, "size": 96, "steps_in_window": 1.0, "conv_dram_bytes_per_step": conv, "all_dram_bytes_per_step": tot_b,
           "kernel_time_us_per_step_serialized": tot_us, "launches_per_step": len(step),
           "source": "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none over bench.py --steps 1 --warmup 2; "
                     "last complete step of the run (tools/summarize_ncu.py)"}, open(prefix + "_traffic.json", "w"), indent=1)
print(open(prefix + "_ncu_launch_shares.csv").read())
