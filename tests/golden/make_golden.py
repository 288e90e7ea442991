#!/usr/bin/env python
"""Generate the golden vectors under tests/golden/ by running the REAL reference.

Runs only in the build container (needs /root/reference, read-only).  The GPU
box never runs this; it only reads the committed ``*.npz`` / ``*.json`` files.

    python tests/golden/make_golden.py

The reference imports two packages that are not installed here at module scope
(SURVEY.md §8c): ``timm`` (attention.py:6, two helpers) and ``audiotools``
(discriminator.py:8-10, pulled in by esc/models/__init__.py:2).  Both are
shimmed with the minimum surface before ``import esc``.  Nothing from the
reference is copied into this repository: only its *outputs* on seeded inputs.

Weights are the deterministic synthetic ones of ``escb200.synthetic`` —
written over the reference model's own ``state_dict`` (``strict=True``), which
also proves our key/shape manifest equals the reference's.
"""
import collections.abc
import itertools
import json
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"


def _install_shims():
    timm = types.ModuleType("timm")
    models = types.ModuleType("timm.models")
    layers = types.ModuleType("timm.models.layers")
    layers.trunc_normal_ = torch.nn.init.trunc_normal_

    def to_2tuple(x):
        if isinstance(x, collections.abc.Iterable) and not isinstance(x, str):
            return tuple(x)
        return tuple(itertools.repeat(x, 2))
    layers.to_2tuple = to_2tuple
    timm.models, models.layers = models, layers
    sys.modules.update({"timm": timm, "timm.models": models, "timm.models.layers": layers})

    at = types.ModuleType("audiotools")
    at.AudioSignal = type("AudioSignal", (), {})
    at.STFTParams = type("STFTParams", (), {})
    ml = types.ModuleType("audiotools.ml")
    ml.BaseModel = torch.nn.Module
    at.ml = ml
    sys.modules.update({"audiotools": at, "audiotools.ml": ml})


def main():
    _install_shims()
    sys.path.insert(0, REF)                                          # reference `esc`
    sys.path.append(os.path.join(ROOT, "efficient-speech-codec_b200"))  # our `escb200`
    from esc.models import make_model                                # the reference
    import esc as ref_esc
    assert ref_esc.__file__.startswith(REF), ref_esc.__file__
    import yaml
    from escb200.spec import CodecSpec
    from escb200.synthetic import synth_state_dict, synth_audio

    torch.manual_seed(0)
    torch.set_num_threads(8)

    def build(cfg_model, seed):
        model = make_model(cfg_model, "csvq+swinT").eval()
        spec = CodecSpec.from_kwargs(**cfg_model)
        sd = synth_state_dict(spec, seed)
        ref_sd = model.state_dict()
        assert set(sd) == set(ref_sd), (set(sd) ^ set(ref_sd))
        for k in ref_sd:
            assert tuple(ref_sd[k].shape) == tuple(sd[k].shape), k
            assert ref_sd[k].dtype == sd[k].dtype, k
        # buffers we regenerate analytically must equal the reference ctor's
        for k, v in ref_sd.items():
            if k.endswith("relative_position_index") or k.endswith(".window"):
                assert torch.equal(v, sd[k]), k
        model.load_state_dict(sd, strict=True)
        return model, spec, sd

    def yaml_model(name):
        return yaml.safe_load(open(f"{REF}/configs/{name}"))["model"]

    base_cfg = yaml_model("9kbps_esc_base.yaml")
    large_cfg = yaml_model("9kbps_esc_large.yaml")
    adv_cfg = yaml_model("9kbps_esc_base_adv.yaml")

    out = {}
    manifests = {}

    # ---------------------------------------------------------------- case A: Base, 3 s, B=2, all bitrates
    model, spec, sd = build(base_cfg, seed=0)
    manifests["base"] = [[e.key, list(e.shape), e.dtype] for e in spec.manifest()]
    assert float(model.max_bps) == 9.0 == float(spec.max_bps)
    x = synth_audio(2, 48000, seed=1)
    with torch.no_grad():
        codes6, fs = model.encode(x, 6)
        audio6 = model.decode(codes6, fs)
    assert fs == (2, 300) and tuple(audio6.shape) == (2, 47920)
    out["A_codes"] = codes6.numpy().astype(np.int16)
    out["A_audio"] = audio6.numpy()
    for s in range(1, 6):
        with torch.no_grad():
            cs, _ = model.encode(x[:1], s)
            au = model.decode(cs, fs)
        assert torch.equal(cs, codes6[:1, :s])      # the chain is causal in the stream index
        out[f"A_audio_s{s}_head"] = au[0, :4000].numpy()
        out[f"A_audio_s{s}_sums"] = np.array([au.double().sum().item(), au.double().abs().sum().item()])
    # forward(eval) == decode(encode(x)) on the 47 920-sample trimmed clip (scripts/utils.py:40)
    xt = x[:1, :-80]
    with torch.no_grad():
        fo = model(x=xt, x_feat=None, num_streams=6)
        ce, fse = model.encode(xt, 6)
        ae = model.decode(ce, fse)
    assert torch.equal(fo["codes"], ce) and torch.equal(fo["recon_audio"], ae)
    out["A_fwd_cm_loss"] = fo["cm_loss"].numpy()
    out["A_fwd_cb_loss"] = fo["cb_loss"].numpy()
    out["A_fwd_codes"] = fo["codes"].numpy().astype(np.int16)
    out["A_fwd_audio_head"] = fo["recon_audio"][0, :4000].numpy()

    # ---------------------------------------------------------------- case B: Large (depth 4, d=8), 3 s, B=1
    model, spec, sd = build(large_cfg, seed=2)
    manifests["large"] = [[e.key, list(e.shape), e.dtype] for e in spec.manifest()]
    x = synth_audio(1, 48000, seed=3)
    with torch.no_grad():
        c, fs = model.encode(x, 6)
        a = model.decode(c, fs)
    out["B_codes"] = c.numpy().astype(np.int16)
    out["B_audio"] = a.numpy()

    # ---------------------------------------------------------------- case C: Base with adv codebook dims, 1 s
    model, spec, sd = build(adv_cfg, seed=4)
    x = synth_audio(2, 16000, seed=5)          # 201 frames -> W=100 -> 50 VQ frames
    with torch.no_grad():
        c, fs = model.encode(x, 6)
        a = model.decode(c, fs)
    assert fs == (2, 100)
    out["C_codes"] = c.numpy().astype(np.int16)
    out["C_audio"] = a.numpy()

    # ---------------------------------------------------------------- case D: Base, ragged W=10 (pads to 12), B=3, full taps
    model, spec, sd = build(base_cfg, seed=6)
    x = synth_audio(3, 1680, seed=7)           # 22 frames -> W=11?  see assert below
    T = 1 + 1680 // 80
    W = T // 2
    assert W == 11 or W == 10
    if W % 2:
        x = synth_audio(3, 1600, seed=7)       # 21 frames -> W=10
    taps = {}

    def hook(name):
        def f(_m, _i, o):
            taps[name] = (o[0] if isinstance(o, tuple) else o).detach().clone()
        return f
    hs = [model.encoder.patch_embed.register_forward_hook(hook("patch_embed")),
          model.encoder.pre_nn.register_forward_hook(hook("enc_hs.0")),
          model.encoder.pre_nn.swint_blocks[0].register_forward_hook(hook("pre_nn.block0")),
          model.encoder.pre_nn.swint_blocks[1].register_forward_hook(hook("pre_nn.block1")),
          model.decoder.post_nn.register_forward_hook(hook("post_nn")),
          model.decoder.patch_deembed.register_forward_hook(hook("recon_feat"))]
    for i, blk in enumerate(model.encoder.blocks):
        hs.append(blk.register_forward_hook(hook(f"enc_hs.{i + 1}")))
    for i, blk in enumerate(model.decoder.blocks):
        hs.append(blk.register_forward_hook(hook(f"dec_hs.{i + 1}")))
    with torch.no_grad():
        planes = model.spec_transform(x)
        c, fs = model.encode(x, 6)
        enc_taps = dict(taps)
        taps.clear()
        a = model.decode(c, fs)
    for h in hs:
        h.remove()
    assert fs == (2, 10), fs
    out["D_x"] = x.numpy()
    out["D_stft"] = planes.contiguous().numpy()
    out["D_codes"] = c.numpy().astype(np.int16)
    out["D_audio"] = a.numpy()
    for k, v in enc_taps.items():
        if k.startswith(("patch_embed", "enc_hs", "pre_nn")):
            out[f"D_tap_{k}"] = v.numpy()
    for k, v in taps.items():
        out[f"D_tap_{k}"] = v.numpy()
    # eval-mode forward for S=3 and S=6 (loss values + short-circuit of untransmitted streams, csrvq.py:35-36)
    xt = x[:, :-80] if (1 + (x.shape[1] - 80) // 80) // 2 % 2 == 0 else x
    for s in (3, 6):
        with torch.no_grad():
            fo = model(x=x, x_feat=None, num_streams=s)
        out[f"D_fwd{s}_cm_loss"] = fo["cm_loss"].numpy()
        out[f"D_fwd{s}_codes"] = fo["codes"].numpy().astype(np.int16)
        out[f"D_fwd{s}_audio"] = fo["recon_audio"].numpy()
        out[f"D_fwd{s}_recon_feat"] = fo["recon_feat"].numpy()
    # per-op pins for the VQ layer on an adversarial input: exact ties -> first index wins
    q = model.quantizers[3]
    z = torch.from_numpy(np.random.RandomState(11).standard_normal((2, 8 * 10, 144)).astype(np.float32))
    with torch.no_grad():
        out["D_pvq3_in"] = z.numpy()
        cq = q.encode(z)
        out["D_pvq3_codes"] = cq.numpy().astype(np.int16)
        out["D_pvq3_dec"] = q.decode(cq, dims=3).numpy()
    vq = q.vqs[0]
    tie_table = vq.embedding.weight.detach().clone()
    tie_table[700] = tie_table[5]          # duplicate rows: argmin must return the lower index
    tie_table[9] = 2.0 * tie_table[3]      # same direction, different norm: ties after l2-normalisation
    zt = torch.cat([tie_table[5:6], tie_table[3:4], -tie_table[3:4], torch.zeros(1, tie_table.shape[1])])[None]
    saved = vq.embedding.weight.data.clone()
    vq.embedding.weight.data.copy_(tie_table)
    with torch.no_grad():
        out["D_tie_table"] = tie_table.numpy()
        out["D_tie_z"] = zt.numpy()
        out["D_tie_codes"] = vq.encode(zt).numpy().astype(np.int16)
    vq.embedding.weight.data.copy_(saved)

    np.savez_compressed(os.path.join(HERE, "golden_v1.npz"), **out)
    json.dump(manifests, open(os.path.join(HERE, "state_dict_manifest.json"), "w"))
    meta = {
        "reference_commit": json.load(open(f"{REF}/.SUBMODULES.json")) if os.path.exists(f"{REF}/.SUBMODULES.json") else None,
        "torch": torch.__version__, "threads": torch.get_num_threads(),
        "cases": {
            "A": "configs/9kbps_esc_base.yaml, weights seed 0, clips synth_audio(2,48000,seed=1), S=1..6",
            "B": "configs/9kbps_esc_large.yaml, weights seed 2, clips synth_audio(1,48000,seed=3), S=6",
            "C": "configs/9kbps_esc_base_adv.yaml, weights seed 4, clips synth_audio(2,16000,seed=5), S=6",
            "D": f"configs/9kbps_esc_base.yaml, weights seed 6, clips synth_audio(3,{x.shape[1]},seed=7), all taps",
        },
        "keys": sorted(out.keys()),
    }
    json.dump(meta, open(os.path.join(HERE, "golden_v1.meta.json"), "w"), indent=1)
    print("wrote", len(out), "arrays;", os.path.getsize(os.path.join(HERE, "golden_v1.npz")) / 1e6, "MB")


if __name__ == "__main__":
    main()
