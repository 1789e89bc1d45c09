#!/usr/bin/env python
"""ncu `--set full` capture -> the per-kernel counters bench.py quotes (profiles/ncu_counters.json).

    ncu -i gpurun_out/X.ncu-rep --page raw --csv > /tmp/raw.csv
    python tools/ncu_counters.py /tmp/raw.csv --workload c2 --commit $(git rev-parse --short HEAD) --samples 16777216 \
        --source profiles/X_ncu_summary.txt -o profiles/ncu_counters.json

Per kernel (first launch of each name in the capture): executed FP32 thread instructions by opcode class (FFMA / FADD / FMUL:
rate per cycle x elapsed cycles), FLOPs executed = 2 FFMA + FADD + FMUL, all thread / warp instructions, issue-active fraction,
DRAM bytes read + written, duration under ncu.  bench.py divides the per-launch figures by the samples of the captured workload
and multiplies by the samples of the run it is timing: the FLOP figure it reports is a measured instruction count, not an estimate
(SURVEY §8d asked for exactly that replacement)."""
import argparse
import csv
import json
import re


def num(x):
    try:
        return float(x.replace(",", ""))
    except ValueError:
        return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("raw_csv"); ap.add_argument("-o", "--out", required=True)
    ap.add_argument("--workload", required=True); ap.add_argument("--commit", required=True)
    ap.add_argument("--samples", type=int, required=True, help="samples (pixels x spp) one launch of the captured workload shades")
    ap.add_argument("--source", default="")
    a = ap.parse_args()
    rows = list(csv.reader(open(a.raw_csv)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    unit_scale = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0, "usecond": 1e-6, "msecond": 1e-3, "nsecond": 1e-9, "second": 1.0}

    def get(r, name, scaled=False):
        i = col.get(name)
        if i is None:
            return None
        v = num(r[i])
        if v is None:
            return None
        return v * unit_scale.get(units[i], 1.0) if scaled else v

    out = {"commit": a.commit, "workload": a.workload, "samples_per_launch": a.samples, "source": a.source,
           "how": "ncu --set full --clock-control none; first captured launch of each kernel; FLOPs = 2 FFMA + FADD + FMUL thread instructions (pred on)",
           "kernels": {}}
    for r in rows[2:]:
        name = r[col["Kernel Name"]]
        m = re.search(r"(\w+_kernel)", name)
        short = (m.group(1) if m else name).replace("_kernel", "")
        if short in out["kernels"]:
            continue
        cyc = get(r, "sm__cycles_elapsed.max")
        op = {k: (get(r, f"smsp__sass_thread_inst_executed_op_{k}_pred_on.sum.per_cycle_elapsed") or 0.0) * cyc for k in ("ffma", "fadd", "fmul")}
        flops = 2 * op["ffma"] + op["fadd"] + op["fmul"]
        out["kernels"][short] = {
            "kernel": name[:160], "duration_s_under_ncu": get(r, "gpu__time_duration.sum", True), "cycles": cyc,
            "thread_inst": get(r, "thread_inst_executed"), "warp_inst": get(r, "smsp__inst_executed.sum"),
            "ffma": op["ffma"], "fadd": op["fadd"], "fmul": op["fmul"], "flops_executed": flops,
            "flops_per_sample": flops / a.samples, "thread_inst_per_sample": (get(r, "thread_inst_executed") or 0) / a.samples,
            "issue_active_frac": (get(r, "sm__issue_active.avg.pct_of_peak_sustained_elapsed") or 0) / 100.0,
            "fp32_pipe_inst_frac": (op["ffma"] + op["fadd"] + op["fmul"]) / cyc / (128.0 * 148.0),
            "dram_bytes": (get(r, "dram__bytes_read.sum", True) or 0) + (get(r, "dram__bytes_write.sum", True) or 0),
            "registers": get(r, "launch__registers_per_thread"), "waves_per_sm": get(r, "launch__waves_per_multiprocessor"),
            "warps_active_frac": (get(r, "sm__warps_active.avg.pct_of_peak_sustained_active") or 0) / 100.0,
        }
    json.dump(out, open(a.out, "w"), indent=1)
    for k, v in out["kernels"].items():
        print(k, {x: (round(y, 4) if isinstance(y, float) else y) for x, y in v.items() if x != "kernel"})


if __name__ == "__main__":
    main()
