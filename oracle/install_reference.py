"""Per-pod install of the UNMODIFIED reference into the git-ignored ``baseline/_ref/`` -- TEST INFRASTRUCTURE ONLY.

    python oracle/install_reference.py            # run in the build container, where /root/reference is mounted

``/root/reference`` does not exist on the GPU box, but ``baseline/_ref/`` travels there with the repository snapshot
(it is git-ignored, never committed, not gpurun-ignored).  With it in place ``bench.py --impl reference`` and the
``cpu_baseline`` leg time the reference's own ``ConformerEncoder`` (``kind: "reference"``) instead of the oracle port.

Order of attempts (the outcome is written to ``baseline/_ref/INSTALL_LOG.txt``):
  1. the documented offline install
        python -m pip install --no-index --no-build-isolation --find-links /opt/wheelhouse --no-deps \
               --target baseline/_ref <copy of /root/reference under /tmp>
     In this image it fails: the reference's setup.py lists ``pytest-runner`` in ``setup_requires`` and no wheel of it
     is in /opt/wheelhouse.
  2. what that install would have produced for a pure-Python package: the ``nemo`` package tree (``*.py`` only) copied
     verbatim to ``baseline/_ref/nemo``.  The package is then loaded by ``oracle/reference_loader.py`` exactly as from
     ``/root/reference`` (stub parents for the absent hydra / pytorch_lightning / sox imports; the five hot-path files
     execute unmodified).
Nothing under ``baseline/_ref`` is product source and nothing in the product imports it.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.environ.get("CONFORMER_REF_SRC", "/root/reference")
DST = os.path.join(ROOT, "baseline", "_ref")
MARKER = "nemo/collections/asr/modules/conformer_encoder.py"


def installed() -> bool:
    return os.path.isfile(os.path.join(DST, MARKER))


def install(force: bool = False) -> str:
    """Returns a one-line description of what is under baseline/_ref afterwards ("" if nothing could be installed)."""
    if installed() and not force:
        return "already installed"
    if not os.path.isfile(os.path.join(SRC, MARKER)):
        return ""
    os.makedirs(DST, exist_ok=True)
    log = [f"{time.strftime('%Y-%m-%dT%H:%M:%SZ', time.gmtime())} install of {SRC} into {DST}"]
    how = ""
    with tempfile.TemporaryDirectory() as tmp:
        work = os.path.join(tmp, "reference")
        shutil.copytree(SRC, work, ignore=shutil.ignore_patterns(".git"))  # the build writes into the source tree
        cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--find-links",
               "/opt/wheelhouse", "--no-deps", "--target", DST, work]
        try:
            r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
            log.append("$ " + " ".join(cmd))
            log.append(f"rc={r.returncode}")
            log.append((r.stdout + r.stderr)[-1500:])
            if r.returncode == 0 and installed():
                how = "pip install --target (offline)"
        except Exception as e:  # pragma: no cover
            log.append(f"pip raised {e!r}")
    if not how:
        # pure-Python package: the install is its package tree
        dst_pkg = os.path.join(DST, "nemo")
        if os.path.isdir(dst_pkg):
            shutil.rmtree(dst_pkg)
        shutil.copytree(os.path.join(SRC, "nemo"), dst_pkg,
                        ignore=lambda d, names: [n for n in names
                                                 if not (n.endswith(".py") or os.path.isdir(os.path.join(d, n)))])
        how = "pip failed (setup_requires pytest-runner has no offline wheel); package tree nemo/ copied verbatim"
        log.append(how)
    with open(os.path.join(DST, "INSTALL_LOG.txt"), "w") as f:
        f.write("\n".join(log) + "\n")
    return how


if __name__ == "__main__":
    print(install(force="--force" in sys.argv) or f"no reference tree at {SRC}")
