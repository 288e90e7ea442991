"""The helpers of the reference's scripts/utils.py the inference drivers need (read_yaml :87-90, EvalSet :28-40, wav I/O).

torchaudio.load/save need torchcodec in torchaudio >= 2.9 (absent in this image), so wav files go through
``scipy.io.wavfile`` with torchaudio's conventions: float32 in [-1, 1], shape (channels, samples)."""
import numpy as np
import torch
import yaml


def read_yaml(pth):
    with open(pth, "r") as f:
        return yaml.safe_load(f)


def load_wav(path):
    """-> (float32 tensor [channels, samples], sample_rate), like ``torchaudio.load``."""
    from scipy.io import wavfile
    sr, data = wavfile.read(path)
    if data.dtype == np.int16:
        x = data.astype(np.float32) / 32768.0
    elif data.dtype == np.int32:
        x = data.astype(np.float32) / 2147483648.0
    elif data.dtype == np.uint8:
        x = (data.astype(np.float32) - 128.0) / 128.0
    else:
        x = data.astype(np.float32)
    if x.ndim == 1:
        x = x[None, :]
    else:
        x = x.T
    return torch.from_numpy(np.ascontiguousarray(x)), int(sr)


def save_wav(path, x, sr):
    """float32 [channels, samples] -> 32-bit float wav, like ``torchaudio.save`` on a float tensor."""
    from scipy.io import wavfile
    a = x.detach().cpu().float().numpy()
    wavfile.write(path, int(sr), a.T if a.shape[0] > 1 else a[0])


class EvalSet(torch.utils.data.Dataset):
    """Evaluation clips of a folder (``*.wav``, else ``*/*.wav``), each returned as ``x[0, :-80]`` - the reference's
    EvalSet (scripts/utils.py:28-40): 3 s test clips lose one hop so that 600 STFT frames come back out."""

    def __init__(self, eval_folder_path) -> None:
        super().__init__()
        import glob
        files = sorted(glob.glob(f"{eval_folder_path}/*.wav")) or sorted(glob.glob(f"{eval_folder_path}/*/*.wav"))
        self.testset_files = files[:180000]

    def __len__(self):
        return len(self.testset_files)

    def __getitem__(self, i):
        x, _ = load_wav(self.testset_files[i])
        return x[0, :-80]
