"""Build recipe of `oracle/_ref/`: the UNMODIFIED reference package, for timing and live parity.

TEST / BENCH INFRASTRUCTURE ONLY.  The reference (regeirk/pycwt) is pure Python: four source
files, no build step.  Where its checkout exists (the build container: /root/reference) this
script copies `pycwt/{__init__,wavelet,mothers,helpers}.py` byte for byte into the git-ignored
directory `oracle/_ref/pycwt/` and records their SHA-256 in `oracle/_ref/MANIFEST.json`.  The
directory is git-ignored (no reference source enters the history) but not gpurun-ignored, so it
travels to the GPU box like the built `.so` files, where

  * `bench.py --impl reference` times the stock `pycwt.cwt` (reference wavelet.py:13-124) through
    its stock single-threaded scipy.fftpack path on the box's host cores, and
  * `bench.py`'s `cpu_baseline` leg does the same on a bounded sample.

Nothing under `pycwt_b200/` imports it.  `__graft_entry__.build()` runs this recipe when
/root/reference is present; on a box without the checkout the prebuilt copy is used as is.
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.environ.get("PYCWT_REFERENCE", "/root/reference")
DST = os.path.join(HERE, "_ref")
FILES = ("__init__.py", "wavelet.py", "mothers.py", "helpers.py")


def make(verbose=False):
    """Copy the reference package into oracle/_ref/pycwt.  Returns the destination directory,
    or None if the reference checkout does not exist here."""
    src = os.path.join(REF_ROOT, "pycwt")
    if not all(os.path.isfile(os.path.join(src, f)) for f in FILES):
        return None
    pkg = os.path.join(DST, "pycwt")
    os.makedirs(pkg, exist_ok=True)
    manifest = {"source": src, "files": {}}
    for f in FILES:
        shutil.copyfile(os.path.join(src, f), os.path.join(pkg, f))
        with open(os.path.join(pkg, f), "rb") as fh:
            manifest["files"][f] = hashlib.sha256(fh.read()).hexdigest()
    with open(os.path.join(DST, "MANIFEST.json"), "w") as fh:
        json.dump(manifest, fh, indent=1, sort_keys=True)
    if verbose:
        print("oracle/_ref: reference package copied from", src)
    return DST


def available():
    return all(os.path.isfile(os.path.join(DST, "pycwt", f)) for f in FILES)


def load():
    """Import the stock reference package from oracle/_ref (raises ImportError if absent).
    It is imported under its own name `pycwt`; the product package is `pycwt_b200`."""
    if not available():
        raise ImportError("oracle/_ref/pycwt is missing: run `python oracle/make_ref.py` where "
                          "the reference checkout exists")
    import importlib
    import warnings
    if DST not in sys.path:
        sys.path.insert(0, DST)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")     # SyntaxWarning / DeprecationWarning of the reference
        return importlib.import_module("pycwt")


if __name__ == "__main__":
    out = make(verbose=True)
    print(out or "reference checkout not found at %s" % REF_ROOT)
