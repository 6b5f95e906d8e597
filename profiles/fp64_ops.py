"""Sums the FP64 work of one config-3 step from an ncu metric log:

    ncu --metrics smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,\
smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,gpu__time_duration.sum --clock-control none -k regex:dubins \
        --csv --log-file gpurun_out/edge_fp64.csv python benchmarks/edge_once.py
    python profiles/fp64_ops.py gpurun_out/edge_fp64.csv > profiles/r02_edge_fp64_ops.json

flop = DADD + DMUL + 2 x DFMA thread-level instructions (predicated-on), summed over the edge-check
kernels of the step; bench.py divides it by the step time it measures and by the measured FP64 peak.
"""
import collections
import csv
import json
import sys


def main(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, data = rows[hi], rows[hi + 1:]
    ki, mi, vi = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
    ui = hdr.index("Metric Unit")
    per = collections.OrderedDict()
    for r in data:
        if len(r) <= vi:
            continue
        k = r[ki].split("(")[0][:60]
        v = float(r[vi].replace(",", ""))
        if r[mi] == "gpu__time_duration.sum":
            v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[ui], 1.0)
        d = per.setdefault(k, collections.defaultdict(float))
        d[r[mi]] += v
        if r[mi] == "gpu__time_duration.sum":
            d["launches"] += 1
    out = {"source": "ncu thread-level FP64 instruction counts (op_dadd / op_dmul / op_dfma, predicated on) of one "
                     "config-3 step: python benchmarks/edge_once.py, kernels matching 'dubins'", "kernels": {}}
    total = 0.0
    for k, d in per.items():
        add = d.get("smsp__sass_thread_inst_executed_op_dadd_pred_on.sum", 0.0)
        mul = d.get("smsp__sass_thread_inst_executed_op_dmul_pred_on.sum", 0.0)
        fma = d.get("smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", 0.0)
        flop = add + mul + 2.0 * fma
        total += flop
        out["kernels"][k] = {"launches": int(d["launches"]), "ms_under_ncu": d.get("gpu__time_duration.sum", 0.0),
                             "dadd": add, "dmul": mul, "dfma": fma, "fp64_flop": flop}
    out["fp64_flop_per_step"] = total
    out["fp64_flop_per_edge"] = total / 1e7
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main(sys.argv[1])
