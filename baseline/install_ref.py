#!/usr/bin/env python
"""Place an UNMODIFIED copy of the reference under baseline/_ref/ (git-ignored; it still travels to the
GPU box with the gpurun snapshot) so that `bench.py --impl reference`, the `gpu_eager_baseline` record and
the on-box parity tests can run the reference's own code there.

    python baseline/install_ref.py [--prebuild]

The reference has no setup.py / pyproject.toml (`pip install /root/reference` has nothing to build), so the
"install" is a byte copy of `sgmse-bbed/` (sources) and `dataset/` (the wav fixtures + active_rms.txt).
No reference source enters the git history.  --prebuild additionally JIT-builds the reference's two native
ops (upfirdn2d, fused) for sm_100a into baseline/_ref/torch_extensions so a GPU box does not spend a
minute compiling them.  Runs only where /root/reference exists; a no-op elsewhere.
"""
import filecmp
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference"
DST = os.path.join(HERE, "_ref")


def _same_tree(a, b):
    if not os.path.isdir(b):
        return False
    c = filecmp.dircmp(a, b, ignore=["__pycache__"])
    if c.left_only or c.diff_files or c.funny_files:
        return False
    return all(_same_tree(os.path.join(a, d), os.path.join(b, d)) for d in c.common_dirs)


def install(prebuild=False, verbose=True):
    if not os.path.isdir(os.path.join(SRC, "sgmse-bbed")):
        if verbose:
            print("install_ref: /root/reference absent, nothing to do")
        return os.path.isdir(os.path.join(DST, "sgmse-bbed"))
    os.makedirs(DST, exist_ok=True)
    for sub in ("sgmse-bbed", "dataset"):
        s, d = os.path.join(SRC, sub), os.path.join(DST, sub)
        if _same_tree(s, d):
            continue
        if os.path.isdir(d):
            shutil.rmtree(d)
        shutil.copytree(s, d, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
        if verbose:
            print(f"install_ref: copied {s} -> {d}")
    if prebuild:
        env = dict(os.environ)
        env["TORCH_CUDA_ARCH_LIST"] = "10.0a"
        r = subprocess.run([sys.executable, os.path.join(HERE, "ref_runner.py"), "--task", "import"], env=env,
                           capture_output=True, text=True)
        if verbose:
            print("install_ref: prebuild", "ok" if r.returncode == 0 else "FAILED\n" + r.stderr[-2000:])
    return True


if __name__ == "__main__":
    install(prebuild="--prebuild" in sys.argv)
