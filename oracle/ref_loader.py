"""Load the UNMODIFIED reference sources from /root/reference for oracle validation.

TEST INFRASTRUCTURE ONLY.  The reference package cannot be imported normally in the
build container: ``regularizepsf/__init__.py:9`` needs installed dist metadata, and
h5py / astropy / matplotlib / sep / scikit-image are absent.  The hot path
(``transform.py``, ``psf.py``, ``util.py``, ``exceptions.py``) needs only numpy and
scipy, so we register empty stand-in modules for the absent imports and exec the four
source files where they lie.  Nothing is copied into this repository.

``/root/reference`` does not exist on the GPU box.  ``__graft_entry__.build()`` therefore pip-installs
the unmodified reference (``--no-deps``; it is pure Python) into ``baseline/_ref/`` — git-ignored, but it
travels with gpurun — and this loader falls back to that copy, so the oracle can be re-pinned and the
reference itself timed on the GPU box.  With neither present ``available()`` is False and callers skip.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_CANDIDATES = [os.environ.get("RPSF_REFERENCE_ROOT"), "/root/reference", os.path.join(_REPO, "baseline", "_ref")]
REF_ROOT = next((c for c in _CANDIDATES if c and os.path.isfile(os.path.join(c, "regularizepsf", "transform.py"))),
                "/root/reference")
_REF_PKG = os.path.join(REF_ROOT, "regularizepsf")
_loaded = None


def available() -> bool:
    return os.path.isfile(os.path.join(_REF_PKG, "transform.py"))


def install_copy(force: bool = False) -> str | None:
    """pip-install the unmodified reference from /root/reference into baseline/_ref (build container only)."""
    import shutil
    import subprocess
    import tempfile

    src = "/root/reference"
    dst = os.path.join(_REPO, "baseline", "_ref")
    if not os.path.isfile(os.path.join(src, "pyproject.toml")):
        return dst if os.path.isdir(os.path.join(dst, "regularizepsf")) else None
    if os.path.isfile(os.path.join(dst, "regularizepsf", "transform.py")) and not force:
        return dst
    with tempfile.TemporaryDirectory() as tmp:
        work = os.path.join(tmp, "reference")          # /root/reference is read-only; the build writes egg-info
        shutil.copytree(src, work, ignore=shutil.ignore_patterns(".git"))
        cmd = [sys.executable, "-m", "pip", "install", "--quiet", "--no-index", "--no-build-isolation", "--no-deps",
               "--find-links", "/opt/wheelhouse", "--upgrade", "--target", dst, work]
        proc = subprocess.run(cmd, capture_output=True, text=True)
        if proc.returncode != 0:
            raise RuntimeError("reference install failed:\n" + proc.stdout + proc.stderr)
    return dst


def _stub(name: str, **attrs) -> types.ModuleType:
    mod = types.ModuleType(name)
    mod.__dict__.update(attrs)
    sys.modules[name] = mod
    return mod


def load():
    """Return a namespace with the reference's ``exceptions``, ``util``, ``psf`` and ``transform`` modules."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError(f"reference sources not found under {REF_ROOT}")
    saved = {k: sys.modules.get(k) for k in list(sys.modules)
             if k == "regularizepsf" or k.startswith("regularizepsf.")}
    for name in ("h5py", "astropy", "astropy.io", "astropy.io.fits"):
        if name not in sys.modules:
            _stub(name)
    sys.modules["astropy.io"].fits = sys.modules["astropy.io.fits"]
    if "matplotlib" not in sys.modules:
        mpl = _stub("matplotlib")
        mpl.colors = _stub("matplotlib.colors")
        mpl.pyplot = _stub("matplotlib.pyplot")
    _stub("regularizepsf.visualize", KERNEL_IMSHOW_ARGS_DEFAULT={}, PSF_IMSHOW_ARGS_DEFAULT={},
          visualize_grid=None)
    pkg = types.ModuleType("regularizepsf")
    pkg.__path__ = [_REF_PKG]
    sys.modules["regularizepsf"] = pkg
    mods = {}
    for name in ("exceptions", "util", "psf", "transform"):
        spec = importlib.util.spec_from_file_location(f"regularizepsf.{name}", os.path.join(_REF_PKG, f"{name}.py"))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[spec.name] = mod
        spec.loader.exec_module(mod)
        mods[name] = mod
    _loaded = types.SimpleNamespace(**mods)
    del saved
    return _loaded


_builder = None


def load_builder():
    """The reference's ``builder`` and ``image_processing`` modules, unmodified, with ``sep`` replaced by
    ``oracle/fake_sep.py`` and an empty ``skimage.transform`` (only used at ``interpolation_scale != 1``)."""
    global _builder
    if _builder is not None:
        return _builder
    ref = load()
    from oracle import fake_sep

    fake_sep.install()
    if "skimage" not in sys.modules:
        _stub("skimage")
        sys.modules["skimage"].transform = _stub("skimage.transform", downscale_local_mean=None)
    mods = {}
    for name in ("image_processing", "builder"):
        spec = importlib.util.spec_from_file_location(f"regularizepsf.{name}", os.path.join(_REF_PKG, f"{name}.py"))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[spec.name] = mod
        spec.loader.exec_module(mod)
        mods[name] = mod
    _builder = types.SimpleNamespace(**vars(ref), **mods)
    return _builder
