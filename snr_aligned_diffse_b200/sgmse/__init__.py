"""Drop-in mirror of the reference's `sgmse` package API for the enhancement hot path.

Put `<repo>/snr_aligned_diffse_b200` on PYTHONPATH in place of the reference's `sgmse-bbed` directory
and `from sgmse.model import ScoreModel` resolves here; or import it as
`snr_aligned_diffse_b200.sgmse`.  Same class / function names, argument meaning and error behaviour
as the reference (sgmse-bbed/sgmse/), with every tensor computation running in the sm_100a library.
"""
import importlib
import os
import sys

_SUBMODULES = ("util", "util.registry", "util.other", "data_module", "sdes", "sampling", "sampling.predictors",
               "sampling.correctors", "backbones", "backbones.shared", "backbones.ncsnpp", "backbones.snrnet",
               "backbones.ncsnpp_utils", "backbones.ncsnpp_utils.op", "backbones.ncsnpp_utils.op.upfirdn2d",
               "backbones.ncsnpp_utils.up_or_down_sampling",
               "snr_estimator", "model")


def install_alias():
    """Make `import sgmse...` (and pickled references to `sgmse.data_module.SpecsDataModule` inside
    Lightning checkpoints, model.py:93) resolve to this package."""
    real = importlib.import_module("snr_aligned_diffse_b200.sgmse")
    sys.modules["sgmse"] = real
    for sub in _SUBMODULES:
        sys.modules["sgmse." + sub] = importlib.import_module("snr_aligned_diffse_b200.sgmse." + sub)
    return real


if __name__ == "sgmse":
    # Imported as a top-level package (its parent directory is on sys.path, like `sgmse-bbed` was):
    # re-home under the real parent package so the relative imports of the sub-modules work.
    _root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    if _root not in sys.path:
        sys.path.insert(0, _root)
    install_alias()
