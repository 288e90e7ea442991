"""``python -m scripts.test`` of the reference (scripts/test.py:9-82): the evaluation sweep over all bitrates, same CLI
and same ``perf_stats.json``, with the codec and the code histogram running on libescb200.

    python -m scripts.test --eval_folder_path ../evaluation_set/test --batch_size 12 --model_path ./esc9kbps --device cuda

``eval_epoch`` calls the model exactly like the reference does - ``model(x=x, x_feat=None, num_streams=s)`` in eval
mode (scripts/test.py:37) - which is ONE fused encode+decode pass of the native library per batch (escb_forward), and
feeds ``outputs["codes"]`` (still on the GPU) to ``EntropyCounter.update`` (one histogram launch).  Differences from the
reference, all at the edges: ``PESQ`` is left out when the ``pesq`` package is not installed (it is CPU-side quality
scoring, not codec work); wav files are read with ``scripts.utils.load_wav`` (torchaudio.load needs torchcodec here);
the model is left in eval mode (there is no training path to return to).
"""
import argparse
import json

import numpy as np
import torch
from torch.utils.data import DataLoader, default_collate

from esc.models import make_model
from .metrics import PESQ, SISDR, EntropyCounter, MelSpectrogramDistance
from .utils import EvalSet, read_yaml


def parse_args(argv=None):
    parser = argparse.ArgumentParser()
    parser.add_argument("--eval_folder_path", type=str, required=True)
    parser.add_argument("--batch_size", type=int, default=1)
    parser.add_argument("--model_path", type=str, required=True, help="folder contains model configuration and checkpoint")
    parser.add_argument("--save_path", type=str, default=None, help="folder to save test statistics")
    parser.add_argument("--device", type=str, default="cpu")
    return parser.parse_args(argv)


@torch.no_grad()
def eval_epoch(model, eval_loader: DataLoader, metric_funcs: dict, e_counter: EntropyCounter, device: str,
               bps_per_stream: float, num_streams=None, verbose: bool = True):
    """One pass over ``eval_loader`` per bitrate; returns {metric: [value per bitrate], "utilization": [...]}."""
    model.eval()
    all_perf = {k: [] for k in metric_funcs}
    all_perf["utilization"] = []
    streams = range(num_streams, num_streams + 1) if num_streams is not None else range(1, model.max_streams + 1)
    for s in streams:
        perf = {k: [] for k in metric_funcs}
        e_counter.reset_stats(num_streams=s)
        for x in eval_loader:
            x = x.to(device)
            outputs = model(**dict(x=x, x_feat=None, num_streams=s))
            recon_x, codes = outputs["recon_audio"], outputs["codes"]
            for k, func in metric_funcs.items():
                perf[k].extend(func(x, recon_x).tolist())
            e_counter.update(codes)
        for k, v in perf.items():
            all_perf[k].append(round(float(np.mean(v)), 4))
        rate, _ = e_counter.compute_utilization()
        perf["utilization"] = [rate]
        all_perf["utilization"].append(rate)
        if verbose:
            print(f"Test Metrics at {s * bps_per_stream:.2f}kbps: " +
                  " | ".join(f"{k}: {np.mean(v):.4f}" for k, v in perf.items()))
    return all_perf


def make_metrics(device):
    funcs = {}
    try:
        funcs["PESQ"] = PESQ()
    except ImportError:
        print("scripts.test: the `pesq` package is not installed; PESQ is left out of perf_stats.json")
    funcs["MelDistance"] = MelSpectrogramDistance().to(device)
    funcs["SISDR"] = SISDR().to(device)
    return funcs


def run(args):
    eval_set = EvalSet(args.eval_folder_path)
    eval_loader = DataLoader(eval_set, batch_size=args.batch_size, shuffle=False, collate_fn=default_collate)
    metric_funcs = make_metrics(args.device)

    cfg = read_yaml(f"{args.model_path}/config.yaml")
    model = make_model(cfg["model"], cfg["model_name"])
    model.load_state_dict(torch.load(f"{args.model_path}/model.pth", map_location="cpu")["model_state_dict"])
    model = model.to(args.device)
    e_counter = EntropyCounter(cfg["model"]["codebook_size"], num_streams=cfg["model"]["max_streams"],
                               num_groups=cfg["model"]["group_size"], device=args.device)

    performances = eval_epoch(model, eval_loader, metric_funcs, e_counter, args.device,
                              num_streams=None, verbose=True, bps_per_stream=1.5)   # all bitrates
    save_path = args.model_path if args.save_path is None else args.save_path
    json.dump(performances, open(f"{save_path}/perf_stats.json", "w"), indent=2)
    print(f"Test statistics saved into {save_path}/perf_stats.json")
    return performances


if __name__ == "__main__":
    run(parse_args())
