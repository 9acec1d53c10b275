"""Per-kernel-family totals of an `ncu --csv` launch list (gpu__time_duration.sum [+ dram__bytes_read/write.sum]).

    python tools/ncu_family_summary.py profiles/r01_ncu_launches_one_forward_v5.csv [--last-forward N_LAUNCHES] > out.json

`--last-forward N` keeps only the last N launches (the warm second forward of tools/profile_forward.py)."""
import collections
import csv
import io
import json
import re
import sys

FAMILIES = [("conv_tc", r"conv_tc"), ("mbconv_t", r"mbconv_t_kernel"), ("expand_sums", r"expand_sums"),
            ("mbconv_fused", r"mbconv_fused_kernel"), ("mbconv_noexpand_fused", r"mbconv1_fused"),
            ("dwconv_tma", r"dwconv_tma"), ("stem_tc", r"stem_tc"), ("gate_fc", r"gate_fc"), ("scale_act", r"scale_act"),
            ("scale_weights", r"scale_weights"), ("upsample_logits_nchw", r"upsample_logits_nchw"),
            ("upsample_argmax", r"upsample_argmax"), ("bilinear_nhwc", r"bilinear"), ("channel_sum", r"channel_sum"),
            ("psp_pool", r"psp_pool"), ("psp_concat", r"psp_concat"), ("attention_tc", r"attn|transpose_v|attention"),
            ("cab_combine", r"cab_combine"), ("memset/other", r".")]


def main():
    path = sys.argv[1]
    last = int(sys.argv[sys.argv.index("--last-forward") + 1]) if "--last-forward" in sys.argv else 0
    lines = open(path).read().splitlines()
    start = next(i for i, l in enumerate(lines) if l.startswith('"ID"'))
    launches = collections.OrderedDict()
    for r in csv.DictReader(io.StringIO("\n".join(lines[start:]))):
        d = launches.setdefault(int(r["ID"]), {"name": r["Kernel Name"]})
        d[r["Metric Name"]] = float(r["Metric Value"].replace(",", ""))
    rows = list(launches.values())
    if last:
        rows = rows[-last:]
    fam = collections.OrderedDict()
    for r in rows:
        name = next(f for f, pat in FAMILIES if re.search(pat, r["name"]))
        f = fam.setdefault(name, {"launches": 0, "time_ns": 0.0, "dram_read_bytes": 0.0, "dram_write_bytes": 0.0,
                                  "per_launch_dram_bytes": []})  # in launch order (= schedule order of the forward)
        f["launches"] += 1
        f["per_launch_dram_bytes"].append(r.get("dram__bytes_read.sum", 0.0) + r.get("dram__bytes_write.sum", 0.0))
        f["time_ns"] += r.get("gpu__time_duration.sum", 0.0)
        f["dram_read_bytes"] += r.get("dram__bytes_read.sum", 0.0)
        f["dram_write_bytes"] += r.get("dram__bytes_write.sum", 0.0)
    total = sum(f["time_ns"] for f in fam.values())
    for f in fam.values():
        f["time_share"] = f["time_ns"] / total
    out = {"source": path, "launches": len(rows), "total_time_ms": total / 1e6,
           "total_dram_bytes": sum(f["dram_read_bytes"] + f["dram_write_bytes"] for f in fam.values()),
           "families": dict(sorted(fam.items(), key=lambda kv: -kv[1]["time_ns"]))}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
