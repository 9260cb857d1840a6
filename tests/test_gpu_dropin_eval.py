"""Drop-in at the reference's own call site: the reference's UNMODIFIED `eval.py` script (byte copy under
baseline/_ref, skipped where absent) runs against this repo's `sgmse` mirror -- PYTHONPATH points at
`snr_aligned_diffse_b200/` where the reference had `sgmse-bbed/`.

What the script exercises (eval.py:94-140,156-165): `ScoreModel.load_from_checkpoint(ckpt, base_dir="", batch_size=16,
num_workers=0, kwargs=dict(gpu=False))` on a Lightning-format checkpoint with an EMA section, `model.eval(no_ema=False)`,
`model.cpu()`, `model.sde.__class__.__name__` / `model.sde._T = ...`, `model.enhance(x, y, sampler_type=..., predictor=...,
corrector=..., corrector_steps=..., N=..., snr=..., atol=..., rtol=..., timestep_type=..., correct_stepsize=..., oracle=...,
clean_rms=1, noise_rms=1)` with the import-time SNR estimator checkpoint at the reference's relative path, and a
1-D float32 numpy result written with `soundfile.write`.  Input: the reference's own valid/ folder (p232_001.wav).
Only I/O packages absent from this image are stood in for (tests/dropin_shims).
"""
import os
import shutil
import subprocess
import sys
import wave

import numpy as np
import pytest
import torch

from oracle.topology import NCSNppConfig, param_specs, snrnet_param_specs
from snr_aligned_diffse_b200.synth import synth_state_dict

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")


@pytest.mark.skipif(not os.path.isfile(os.path.join(REF, "sgmse-bbed", "eval.py")), reason="baseline/_ref not present")
def test_reference_eval_script_runs_on_the_mirror(tmp_path):
    run = tmp_path / "run"
    out = tmp_path / "out"
    (run / "sgmse-bbed" / "sgmse").mkdir(parents=True)
    (out / "all").mkdir(parents=True)
    for f in ("eval.py", "utils.py"):                      # the scripts only: the `sgmse` package must come from the mirror
        shutil.copy(os.path.join(REF, "sgmse-bbed", f), run / f)
    sd = synth_state_dict(param_specs(NCSNppConfig()), seed=0)
    names = [n for n in sd if n != "dnn.all_modules.0.W"]
    hp = dict(backbone="ncsnpp", sde="ouve", model_type="sebridge_v3", snr_conditioned="true", fixed_snr=0.17783,
              theta=1.5, sigma_min=0.05, sigma_max=1.0, base_dir="/data/was/elsewhere")
    torch.save({"state_dict": {k: v * 1.5 for k, v in sd.items()}, "hyper_parameters": hp,   # raw weights differ from EMA:
                "ema": {"decay": 0.999, "num_updates": 10, "shadow_params": [sd[n] for n in names],   # eval() must pick EMA
                        "collected_params": None}}, run / "model.ckpt")
    torch.save({"state_dict": synth_state_dict(snrnet_param_specs(), seed=1),
                "hyper_parameters": {"backbone": "snrnet", "base_dir": ""}}, run / "sgmse-bbed" / "sgmse" / "snr_estimator.ckpt")
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([os.path.join(ROOT, "snr_aligned_diffse_b200"), os.path.join(ROOT, "tests", "dropin_shims"), ROOT])
    r = subprocess.run([sys.executable, "eval.py", "--test_dir", os.path.join(REF, "dataset", "VBD_SNR-5", "valid"),
                        "--ckpt", "model.ckpt", "--destination_folder", str(out) + os.sep], cwd=run, env=env,
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    wav = out / "all" / "p232_001.wav"
    assert wav.is_file() and (out / "_results.csv").is_file() and (out / "_avg_results.txt").is_file()
    with wave.open(str(wav), "rb") as w:
        assert w.getframerate() == 16000 and w.getnframes() == 27861            # same length as the noisy input
        pcm = np.frombuffer(w.readframes(w.getnframes()), dtype="<i2")
    assert np.abs(pcm).max() > 0
    rows = open(out / "_results.csv").read().strip().splitlines()
    assert rows[0].startswith("filename") and rows[1].startswith("p232_001.wav")
