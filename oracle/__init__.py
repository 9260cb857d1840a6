"""CPU oracle for the SNR-aligned diffusion enhancement hot path.

TEST INFRASTRUCTURE ONLY.  This package is a CPU (PyTorch fp32 / numpy fp64) restatement of the
reference algorithm (yh-jun/SNR-Aligned_diffSE, `sgmse-bbed/sgmse/...`).  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference` legs may import it,
and there only as the checker or as the timed CPU baseline -- never as part of the product path.

Parity status: PINNED.  `oracle/make_golden.py` imports the unmodified reference from
`/root/reference` (through the dependency stand-ins in `oracle/ref_shims/`), runs both the reference
and this restatement on the same seeded weights / inputs / noise, asserts they agree, and writes the
reference's outputs to `tests/golden/`.  `tests/test_oracle_golden.py` re-checks this restatement
against those committed fixtures on every CPU test run.

Third-party arithmetic the reference delegates to (and this oracle delegates to as well):
`torch.stft/istft`, `torch.nn.functional.conv2d/group_norm`, `torch.nn.LSTM`, `scipy.special.expi`.
"""
