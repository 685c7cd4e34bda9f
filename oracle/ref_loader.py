"""Imports the UNMODIFIED reference modules (Networks.py / Losses.py of Baverne/VAE-CYCLEGAN-Implementation).

TEST / BENCH INFRASTRUCTURE ONLY.  Source of the modules, in this order: /root/reference (build container) or the
archive oracle/build_ref.py wrote into oracle/_ref/ (GPU box).  The repo root also holds drop-in modules called
Networks / Losses (the product's shims); the reference's ``from Losses import *`` must not resolve to those, so the
reference pair is loaded under private names with sys.modules patched during the import only."""
import importlib.util
import os
import sys
import tarfile
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ARCHIVE = os.path.join(HERE, "_ref", "reference_py.tar.gz")
_CACHE = {}


def _source_dir():
    if os.path.exists("/root/reference/Networks.py"):
        return "/root/reference", "checkout"
    if os.path.exists(ARCHIVE):
        d = tempfile.mkdtemp(prefix="vcg_reference_")
        with tarfile.open(ARCHIVE, "r:gz") as tar:
            for m in tar.getmembers():
                if m.name in ("Networks.py", "Losses.py") and m.isfile():
                    with open(os.path.join(d, m.name), "wb") as fh:
                        fh.write(tar.extractfile(m).read())
        return d, "oracle/_ref archive"
    return None, None


def load():
    """-> (reference Networks module, origin string) or (None, reason)."""
    if "mod" in _CACHE:
        return _CACHE["mod"], _CACHE["origin"]
    d, origin = _source_dir()
    if d is None:
        return None, "no /root/reference and no oracle/_ref/reference_py.tar.gz (run oracle/build_ref.py in the build container)"
    saved = {k: sys.modules.get(k) for k in ("Losses", "Networks")}
    try:
        def imp(name, alias):
            spec = importlib.util.spec_from_file_location(alias, os.path.join(d, name + ".py"))
            mod = importlib.util.module_from_spec(spec)
            sys.modules[name] = mod                 # what `from Losses import *` inside the reference resolves to
            spec.loader.exec_module(mod)
            return mod
        imp("Losses", "_reference_Losses")
        net = imp("Networks", "_reference_Networks")
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    _CACHE["mod"], _CACHE["origin"] = net, origin
    return net, origin
