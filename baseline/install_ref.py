# -*- coding: utf-8 -*-
"""Put the UNMODIFIED reference package under git-ignored ``baseline/_ref/`` so that ``bench.py --impl reference``
(and ``cpu_baseline``) time the reference's own code on the GPU box, where ``/root/reference`` does not exist.

    python baseline/install_ref.py

``pip install --no-index --no-build-isolation --target baseline/_ref <copy of /root/reference>`` is tried first; it
fails in this image (the reference's setup.py imports ``distutils`` / ``pytest-runner``, gone from Python 3.12 and
the wheelhouse), so the pure-Python package directory ``wavenet_vocoder/nets`` -- the only part that carries hot-path
arithmetic and imports with numpy + torch alone -- is installed by a plain file copy, byte for byte.
``baseline/_ref`` is listed in .gitignore (no reference sources in the history) but NOT in .gpurunignore (it travels).
"""
import filecmp
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")
SRC = "/root/reference"


def install(verbose=True):
    if not os.path.isdir(os.path.join(SRC, "wavenet_vocoder", "nets")):
        return os.path.isdir(os.path.join(DST, "wavenet_vocoder", "nets"))   # GPU box: use what travelled
    tmp = tempfile.mkdtemp()
    try:
        cp = os.path.join(tmp, "ref")
        shutil.copytree(SRC, cp)
        rc = subprocess.run([sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps",
                             "--find-links", "/opt/wheelhouse", "--target", DST, cp],
                            stdout=subprocess.PIPE, stderr=subprocess.STDOUT).returncode
    except Exception:
        rc = 1
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    if rc != 0:
        pkg = os.path.join(DST, "wavenet_vocoder")
        os.makedirs(pkg, exist_ok=True)
        open(os.path.join(pkg, "__init__.py"), "w").close()       # the reference's own __init__.py is empty too
        dn = os.path.join(pkg, "nets")
        if os.path.isdir(dn):
            shutil.rmtree(dn)
        shutil.copytree(os.path.join(SRC, "wavenet_vocoder", "nets"), dn)
        same = filecmp.cmp(os.path.join(SRC, "wavenet_vocoder", "nets", "wavenet.py"), os.path.join(dn, "wavenet.py"),
                           shallow=False)
        assert same
        if verbose:
            print("baseline/_ref: pip install failed (rc=%d); copied wavenet_vocoder/nets unmodified" % rc)
    return True


def import_ref():
    """The reference ``wavenet_vocoder.nets`` module from baseline/_ref (None if it is not there)."""
    if not os.path.isdir(os.path.join(DST, "wavenet_vocoder", "nets")):
        return None
    import importlib.util
    spec = importlib.util.spec_from_file_location("_ref_wavenet", os.path.join(DST, "wavenet_vocoder", "nets", "wavenet.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    print(install())
