#!/usr/bin/env python
"""profiles/rNN_ncu_traffic.json from an ncu summary (tools/ncu_summary.py output, one JSON object per launch):
DRAM bytes per launch of every hand-written kernel, under the names bench.py's `kernels_ms_per_step` uses
(the two size classes of the fused trainer and the two refine kernels are one logical launch each).

    python tools/make_traffic.py gpurun_out/r02_ncu_default.jsonl default > profiles/r02_ncu_traffic.json"""
import collections
import json
import sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
ALIAS = {"kmeans_fused_kernel<256>": "kmeans_fused", "kmeans_fused_kernel<512>": "kmeans_fused",
         "refine_kernel": "refine_block", "gather_kernel<unsigned int>": "gather", "gather_kernel<double>": "gather"}


def to_bytes(text):
    v, u = text.split()
    return float(v) * UNIT[u]


def main():
    path, workload = sys.argv[1], sys.argv[2]
    launches = collections.OrderedDict()
    for line in open(path):
        d = json.loads(line)
        name = d["kernel"].replace("void ", "").replace("flc::", "")
        if name.startswith("at::") or "dram_read" not in d:
            continue
        name = ALIAS.get(name, name[: -len("_kernel")] if name.endswith("_kernel") else name)
        launches.setdefault(name, []).append(to_bytes(d["dram_read"]) + to_bytes(d["dram_write"]))
    # launches that make up one logical launch are summed; repeated launches of a kernel (DBSCAN sweeps) are averaged
    merged = {"kmeans_fused": 2, "refine_block": 2, "gather": 2}
    out = {}
    for name, vals in launches.items():
        per = merged.get(name, 1)
        n_logical = max(1, len(vals) // per)
        out[name] = {"workload": workload, "dram_bytes_per_launch": sum(vals) / n_logical, "launches_in_capture": len(vals),
                     "source": f"ncu --set full --clock-control none, one steady-state step ({path.split('/')[-1]}): "
                               "dram__bytes_read.sum + dram__bytes_write.sum"}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
