"""Aggregate an ncu metrics pass over ONE encode+decode step into per-kernel-class DRAM traffic.

    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
        --launch-skip <launches of the warm-up step> -c <launches of one step> --csv --log-file X.csv \
        python tools/profile_step.py 36 1
    python tools/ncu_traffic.py X.csv profiles/<round>_traffic.json

The classes are the ones bench.py reports (libescb200's escb_profile_end names); a LayerNorm statistics launch is
booked to the GEMM it precedes, exactly as the per-op CUDA-event timing does.  bench.py reads the JSON to fill
roofline.traffic (measured DRAM bytes per launch of the dominant kernel class, averaged over the step's launches the
same way roofline.achieved averages the algorithmic bytes)."""
import collections
import csv
import json
import sys

RULES = [   # (substring of the demangled kernel name, class); first match wins
    ("mlp_fused_kernel", "mlp_fused"),
    ("pvq_stream_kernel", "pvq_stream_fused"),
    ("rvq_chain_kernel", "codebook_argmin"),
    ("rvq_gather_kernel", "layout"),
    ("code_histogram_kernel", "layout"),
    ("EpiAttn<", "qkv_attention_fused"),
    ("AWindow, EpiRows", "qkv_gemm"),
    ("window_attn_kernel", "window_attention"),
    ("EpiWindow", "proj_gemm"),
    ("EpiRows<1, 0>", "mlp1_gemm"),
    ("EpiRows<0, 1>", "mlp2_gemm"),
    ("AMerge", "merge_gemm"),
    ("EpiSplit", "split_gemm"),
    ("AFrame", "pvq_down_gemm"),
    ("ACodes", "pvq_up_gemm"),
    ("codebook_argmin_kernel", "codebook_argmin"),
    ("AIm2col", "deembed_conv5x5_gemm"),
    ("conv3x3_out", "deembed_conv3x3"),
    ("AStftFrames", "stft_gemm"),
    ("AIstft", "istft_gemm"),
    ("patch_embed", "patch_embed"),
    ("vq_loss_kernel", "vq_loss"),
    ("transpose_kernel", "layout"),
    ("repitch_kernel", "layout"),
]


def classify(name):
    for sub, cls in RULES:
        if sub in name:
            return cls
    return None


def main(src, dst):
    per_launch = collections.OrderedDict()            # launch id -> {name, metric: value}
    for r in csv.reader(open(src, errors="replace")):
        if len(r) < 15 or not r[0].isdigit():
            continue
        d = per_launch.setdefault(int(r[0]), {"name": r[4]})
        v = float(r[14].replace(",", ""))
        unit = r[13]
        if r[12].startswith("dram__bytes"):
            v *= {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
        elif r[12].startswith("gpu__time"):
            v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(unit, 1.0)
        d[r[12]] = v
    # the capture may hold warm-up steps: keep the LAST step (every encode starts with the STFT GEMM)
    ids = list(per_launch)
    starts = [i for i in ids if "AStftFrames" in per_launch[i]["name"]]
    if len(starts) > 1:
        for i in ids:
            if i < starts[-1]:
                del per_launch[i]
    out = collections.OrderedDict()
    pending = None                                     # LayerNorm statistics launch waiting for its GEMM
    for lid, d in per_launch.items():
        cls = classify(d["name"])
        rd, wr, us = d.get("dram__bytes_read.sum", 0.0), d.get("dram__bytes_write.sum", 0.0), d.get("gpu__time_duration.sum", 0.0)
        if "ln_stats_kernel" in d["name"]:
            pending = (rd, wr, us)
            continue
        if cls is None:
            print("unclassified:", d["name"][:100], file=sys.stderr)
            continue
        e = out.setdefault(cls, {"launches": 0, "dram_read_bytes": 0.0, "dram_write_bytes": 0.0, "time_us": 0.0})
        e["launches"] += 1
        e["dram_read_bytes"] += rd
        e["dram_write_bytes"] += wr
        e["time_us"] += us
        if pending:
            e["dram_read_bytes"] += pending[0]
            e["dram_write_bytes"] += pending[1]
            e["time_us"] += pending[2]
            pending = None
    total_bytes = sum(e["dram_read_bytes"] + e["dram_write_bytes"] for e in out.values())
    for e in out.values():
        e["dram_bytes_per_launch"] = (e["dram_read_bytes"] + e["dram_write_bytes"]) / max(e["launches"], 1)
    tot = sum(e["time_us"] for e in out.values()) or 1.0
    for e in out.values():
        e["share_of_step"] = round(e["time_us"] / tot, 4)
    json.dump({"source": src, "note": "one encode+decode step, ESC-Base, 36 x 3 s clips; ncu serialises launches and "
               "runs them cold-cache, so shares (not absolute times) are comparable with bench.py",
               "launches_in_step": sum(e["launches"] for e in out.values()), "dram_bytes_per_step": total_bytes, "classes": out},
              open(dst, "w"), indent=1)
    for k, e in sorted(out.items(), key=lambda kv: -kv[1]["time_us"]):
        print(f"{k:24s} {e['launches']:4d} launches {e['time_us']:9.1f} us  share {e['share_of_step']:.3f}  "
              f"{e['dram_bytes_per_launch'] / 1e6:8.1f} MB/launch")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
