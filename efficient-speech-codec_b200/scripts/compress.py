"""``python -m scripts.compress`` of the reference (scripts/compress.py:6-40), unchanged CLI, B200-native codec.

    python -m scripts.compress --input ./audio.wav --save_path ./output --model_path ./esc9kbps \\
        --num_streams 6 --device cuda

The argparse block and the body of ``main`` deliberately track the reference's CLI line for line (same flags, same help
strings, same output names): this file IS the API surface the north star says to keep.  Deviations, all at the edges:
``make_model`` is given ``model_name`` (the reference's one-argument call raises TypeError as shipped,
scripts/compress.py:22 vs esc/models/codecs.py:190), the model is put in ``.eval()`` (there is no training path),
wav I/O goes through scipy (torchaudio.load/save need torchcodec, absent here; the decoded wav is 32-bit float like
``torchaudio.save`` of a float tensor), and ``--bitstream`` is an addition.

Outputs are the reference's: ``decoded_{kbps}kbps_{name}`` (wav) and ``encoded_{kbps}kbps_{stem}.pth`` (the int64 code
tensor, ``torch.save``).  ``--device cpu`` keeps the tensors on the host and stages them through the GPU (there is
no CPU compute path)."""
import argparse
import os
import warnings

import torch

from esc.models import make_model
from .utils import load_wav, read_yaml, save_wav

warnings.filterwarnings("ignore")


def parse_args(argv=None):
    parser = argparse.ArgumentParser()
    parser.add_argument("--input", type=str, required=True, help="input 16kHz mono audio file to encode")
    parser.add_argument("--save_path", type=str, default="./output", help="folder to save codes and reconstructed audio")
    parser.add_argument("--model_path", type=str, required=True, help="folder contains model configuration and checkpoint")
    parser.add_argument("--num_streams", type=int, default=6, help="number of transmitted streams in encoding")
    parser.add_argument("--device", type=str, default="cpu")
    parser.add_argument("--bitstream", action="store_true",
                        help="(esc-b200 addition) also write encoded_*.escb: the codes packed at 10 bits per index")
    return parser.parse_args(argv)


def main(args):
    x, sr = load_wav(f"{args.input}")
    x = x.to(args.device)

    cfg = read_yaml(f"{args.model_path}/config.yaml")
    model = make_model(cfg["model"], cfg.get("model_name", "csvq+swinT"))
    model.load_state_dict(
        torch.load(f"{args.model_path}/model.pth", map_location="cpu")["model_state_dict"],
    )
    model = model.to(args.device).eval()

    codes, size = model.encode(x, num_streams=args.num_streams)
    recon_x = model.decode(codes, size)

    fname = args.input.split("/")[-1]
    if not os.path.exists(args.save_path):
        os.makedirs(args.save_path)
    save_wav(f"{args.save_path}/decoded_{args.num_streams*1.5}kbps_{fname}", recon_x, sr)
    torch.save(codes, f"{args.save_path}/encoded_{args.num_streams*1.5}kbps_{fname.split('.')[0]}.pth")
    if getattr(args, "bitstream", False):
        from escb200.bitstream import save_codes
        save_codes(f"{args.save_path}/encoded_{args.num_streams*1.5}kbps_{fname.split('.')[0]}.escb", codes,
                   cfg["model"].get("codebook_size", 1024))
    print(f"compression outputs saved into {args.save_path}")
    return codes, recon_x


if __name__ == "__main__":
    main(parse_args())
