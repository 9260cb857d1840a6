"""`soundfile.read / write` for mono PCM16 wav files via the standard library."""
import wave

import numpy as np


def write(path, data, samplerate, *a, **k):
    pcm = np.clip(np.rint(np.asarray(data, dtype=np.float64).reshape(-1) * 32767.0), -32768, 32767).astype("<i2")
    with wave.open(str(path), "wb") as w:
        w.setnchannels(1)
        w.setsampwidth(2)
        w.setframerate(int(samplerate))
        w.writeframes(pcm.tobytes())


def read(path, *a, **k):
    with wave.open(str(path), "rb") as w:
        sr, raw = w.getframerate(), w.readframes(w.getnframes())
    return np.frombuffer(raw, dtype="<i2").astype(np.float32) / 32768.0, sr
