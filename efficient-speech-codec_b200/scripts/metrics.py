"""Evaluation metrics of the reference's ``scripts/metrics.py`` that the eval sweep (scripts/test.py) needs.

``EntropyCounter`` (scripts/metrics.py:12-77 of the reference) keeps the reference's public surface - constructor
arguments, ``reset_stats``, ``update``, ``compute_utilization``, the ``codebook_counts`` / ``total_counts`` / ``dist``
/ ``entropy`` attributes - but its ``update`` is ONE launch of libescb200's histogram kernel over the code tensor the
codec just produced on the GPU (the reference builds 18 one-hot ``[B*T, 1024]`` tensors per batch).  CUDA codes with
no CUDA library is an error, not a fallback; host code tensors (e.g. a loaded ``encoded_*.pth``) are counted with
``torch.bincount`` since no device is involved.

``SISDR`` and ``MelSpectrogramDistance`` (scripts/metrics.py:95-171) are restated on stock torch / torchaudio ops -
they are quality metrics on the output audio, not part of the accelerated path.  ``PESQ`` wraps the third-party
``pesq`` package exactly like the reference and raises ``ImportError`` at construction when it is absent (it is absent
in this image); ``scripts.test`` then leaves it out of the table.
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

MEL_WINDOWS = [32, 64, 128, 256, 512, 1024, 2048]
MEL_BINS = [5, 10, 20, 40, 80, 160, 320]
SR = 16000


class EntropyCounter:
    """Codebook utilisation (bitrate efficiency) over a held-out set."""

    def __init__(self, codebook_size=1024, num_streams=6, num_groups=3, device="cuda"):
        self.num_groups = num_groups
        self.codebook_size = codebook_size
        self.device = device
        self.reset_stats(num_streams)

    def reset_stats(self, num_streams):
        self.num_streams = num_streams
        # one [S*G, K] matrix; the reference's per-codebook dict is a view of its rows
        self._counts = torch.zeros(num_streams * self.num_groups, self.codebook_size, device=self.device)
        self.codebook_counts = {f"stream_{s}_group_{g + 1}": self._counts[s * self.num_groups + g]
                                for s in range(num_streams) for g in range(self.num_groups)}
        self.total_counts = 0
        self.dist = None
        self.entropy = None
        self.max_entropy_per_book = math.log2(self.codebook_size)
        self.max_total_entropy = num_streams * self.num_groups * self.max_entropy_per_book

    def update(self, codes):
        """codes: (B, num_streams, group_size, T) int64."""
        assert codes.size(1) == self.num_streams and codes.size(2) == self.num_groups, "code indices size not match"
        B, S, G, T = codes.shape
        self.total_counts += B * T
        self.dist = self.entropy = None
        if codes.is_cuda:
            from escb200 import native
            if self._counts.device != codes.device:
                raise ValueError(f"EntropyCounter lives on {self._counts.device} but the codes are on {codes.device}")
            c = codes.contiguous().to(torch.int64)
            with torch.cuda.device(codes.device):
                native.check(native.lib().escb_code_histogram(
                    native.ptr(c), B, S, G, T, self.codebook_size, native.ptr(self._counts),
                    torch.cuda.current_stream(codes.device).cuda_stream))
        else:
            flat = codes.permute(1, 2, 0, 3).reshape(S * G, B * T)
            offs = torch.arange(S * G).unsqueeze(1) * self.codebook_size
            binc = torch.bincount((flat + offs).reshape(-1), minlength=S * G * self.codebook_size)
            self._counts += binc.view(S * G, self.codebook_size).to(self._counts)

    def _form_distribution(self):
        assert self.total_counts > 0, "No data collected, please update on a specific dataset"
        self.dist = {k: v / float(self.total_counts) for k, v in self.codebook_counts.items()}

    def _form_entropy(self):
        assert self.dist is not None, "Please compute posterior distribution first using self._form_distribution()"
        self.entropy = {k: (-torch.sum(p * torch.log2(p + 1e-10))).item() for k, p in self.dist.items()}

    def compute_utilization(self):
        if self.dist is None:
            self._form_distribution()
        if self.entropy is None:
            self._form_entropy()
        utilization = {k: round(e / self.max_entropy_per_book, 4) for k, e in self.entropy.items()}
        return round(sum(self.entropy.values()) / self.max_total_entropy, 4), utilization


class PESQ:
    """Wide-band PESQ per clip through the third-party ``pesq`` package (CPU)."""

    def __init__(self):
        from pesq import pesq          # noqa: F401  (ImportError here: the package is not installed)
        self._pesq = pesq

    def __call__(self, x, y):
        return torch.tensor([self._pesq(SR, x[b].cpu().numpy(), y[b].cpu().numpy(), "wb") for b in range(x.size(0))])


class MelSpectrogramDistance(nn.Module):
    """Sum over 7 resolutions of the mean L1 distance between log10 power mel spectrograms, per clip."""

    def __init__(self, win_lengths=MEL_WINDOWS, n_mels=MEL_BINS, clamp_eps=1e-5):
        super().__init__()
        import torchaudio.transforms as T
        self.mel_transf = nn.ModuleList([
            T.MelSpectrogram(sample_rate=SR, n_fft=w, win_length=w, hop_length=w // 4, n_mels=m, power=1)
            for w, m in zip(win_lengths, n_mels)])
        self.clamp_eps = clamp_eps

    def forward(self, raw_audio, recon_audio):
        total = 0.0
        for mel in self.mel_transf:
            a = mel(raw_audio).clamp(self.clamp_eps).pow(2).log10()
            b = mel(recon_audio).clamp(self.clamp_eps).pow(2).log10()
            total = total + F.l1_loss(a, b, reduction="none").mean(dim=[1, 2])
        return total


class SISDR(nn.Module):
    """Scale-invariant signal-to-distortion ratio in dB per clip (zero-mean, optimal scaling)."""

    def __init__(self, scaling=True, reduction="none", zero_mean=True):
        super().__init__()
        self.scaling, self.reduction, self.zero_mean = scaling, reduction, zero_mean

    def forward(self, x, y):
        eps = 1e-8
        ref = x.reshape(x.shape[0], -1)
        est = y.reshape(y.shape[0], -1)
        if self.zero_mean:
            ref = ref - ref.mean(dim=1, keepdim=True)
            est = est - est.mean(dim=1, keepdim=True)
        if self.scaling:
            alpha = ((est * ref).sum(dim=1, keepdim=True) + eps) / ((ref ** 2).sum(dim=1, keepdim=True) + eps)
        else:
            alpha = 1.0
        target = alpha * ref
        noise = est - target
        return 10 * torch.log10((target ** 2).sum(dim=1) / (noise ** 2).sum(dim=1) + eps)
