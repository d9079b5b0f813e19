"""Stages the UNMODIFIED reference for the CPU arm of bench.py:

    python baseline/make_ref.py            # build container only: needs /root/reference

copies /root/reference/scripts verbatim into baseline/_ref/scripts (git-ignored -- reference sources never enter the
history -- but NOT gpurun-ignored, so the copy travels to the GPU box like the built .so).  The reference is a flat
directory of scripts with no setup.py / pyproject.toml, so `pip install --target baseline/_ref /root/reference` has nothing
to install; a plain copy is the install.  `baseline/ref_step.py` imports it through three stub modules for the packages the
reference imports at module top but this image lacks (albumentations, tensorboardX, torchsummary; none of them is touched by
the G+D loop body).
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/scripts"
DST = os.path.join(HERE, "_ref", "scripts")


def make(verbose=True):
    if not os.path.isdir(SRC):
        if verbose:
            print("baseline/make_ref.py: %s not present (GPU box); keeping whatever baseline/_ref holds" % SRC)
        return os.path.isdir(DST)
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    shutil.copytree(SRC, DST, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    n = sum(len(fs) for _, _, fs in os.walk(DST))
    with open(os.path.join(HERE, "_ref", "PROVENANCE.txt"), "w") as f:
        f.write("verbatim copy of %s (%d files), made by baseline/make_ref.py; git-ignored\n" % (SRC, n))
    if verbose:
        print("baseline/_ref/scripts: %d files copied from %s" % (n, SRC))
    return True


if __name__ == "__main__":
    sys.exit(0 if make() else 1)
