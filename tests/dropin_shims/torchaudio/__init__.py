"""`torchaudio.load` for mono PCM16 wav files via the standard library (TorchCodec is absent from this image)."""
import wave

import numpy as np
import torch


def load(path, *a, **k):
    with wave.open(str(path), "rb") as w:
        assert w.getnchannels() == 1 and w.getsampwidth() == 2, "mono PCM16 only"
        sr, raw = w.getframerate(), w.readframes(w.getnframes())
    pcm = np.frombuffer(raw, dtype="<i2")
    return torch.from_numpy(pcm.astype(np.float32) / 32768.0)[None], sr
