"""Installs the UNMODIFIED reference (apple/ml-neuman) under baseline/_ref/ so that it travels to the GPU box with the
gpurun snapshot (baseline/_ref is git-ignored, not gpurun-ignored):

    python tools/install_reference.py            # build container only (/root/reference present)

The reference is not a Python package (no setup.py / pyproject: `pip install --target baseline/_ref /root/reference`
fails with "does not appear to be a Python project"), so its Python sources are copied file by file, untouched.
Used by: bench.py --impl reference / cpu_baseline (times the reference's own CPU functions on the box's host cores) and
tests/test_gpu_dropin.py (neuman_b200.install() under the reference's own callers on the B200).
"""
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.environ.get("NEUMAN_REFERENCE_SRC", "/root/reference")
DST = os.path.join(ROOT, "baseline", "_ref")
PACKAGES = ["cameras", "data_io", "datasets", "geometry", "models", "options", "trainers", "utils"]


def install(verbose=True):
    if not os.path.isdir(os.path.join(SRC, "utils")):
        if verbose:
            print(f"{SRC} not present: keeping {DST} as it is ({'present' if os.path.isdir(DST) else 'absent'})")
        return os.path.isdir(os.path.join(DST, "utils"))
    os.makedirs(DST, exist_ok=True)
    n = 0
    for pkg in PACKAGES:
        for dp, dn, fn in os.walk(os.path.join(SRC, pkg)):
            rel = os.path.relpath(dp, SRC)
            os.makedirs(os.path.join(DST, rel), exist_ok=True)
            for f in fn:
                if f.endswith(".py"):
                    shutil.copy2(os.path.join(dp, f), os.path.join(DST, rel, f))
                    n += 1
    for f in os.listdir(SRC):
        if f.endswith(".py") or f in ("LICENSE", "ACKNOWLEDGMENTS"):
            shutil.copy2(os.path.join(SRC, f), os.path.join(DST, f))
            n += 1
    try:
        head = subprocess.run(["git", "-C", SRC, "rev-parse", "HEAD"], capture_output=True, text=True).stdout.strip()
    except Exception:
        head = ""
    with open(os.path.join(DST, "INSTALLED_FROM"), "w") as fp:
        fp.write(f"{SRC} {head}\n{n} files copied unmodified by tools/install_reference.py\n")
    if verbose:
        print(f"installed {n} files into {DST}")
    return True


if __name__ == "__main__":
    sys.exit(0 if install() else 1)
