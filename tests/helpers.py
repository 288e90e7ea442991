"""Shared fixtures for the parity tests: the reference's three shipped model configs
(/root/reference/configs/9kbps_esc_{base,base_adv,large}.yaml, `model:` block) and builders."""
import numpy as np
import torch

BASE = dict(backbone="transformer", in_dim=2, in_freq=192, h_dims=[45, 72, 96, 144, 192, 384], max_streams=6,
            win_len=20, hop_len=5, sr=16000, patch_size=[3, 2], swin_heads=[3, 6, 12, 24, 24], swin_depth=2,
            window_size=4, mlp_ratio=4.0, overlap=2, group_size=3, codebook_size=1024,
            codebook_dims=[32, 32, 16, 12, 8, 6], l2norm=True)
ADV = dict(BASE, codebook_dims=[8] * 6)
LARGE = dict(BASE, swin_depth=4, codebook_dims=[8] * 6)


def make_oracle(cfg, seed):
    from escb200.models.spec import CodecSpec
    from escb200.utils.synthetic import synth_state_dict
    from oracle.esc_oracle import EscOracle
    sd = synth_state_dict(CodecSpec.from_kwargs(**cfg), seed)
    return EscOracle(cfg, sd), sd


def i64(a):
    return torch.from_numpy(np.asarray(a).astype(np.int64))
