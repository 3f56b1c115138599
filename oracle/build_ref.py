#!/usr/bin/env python
"""Make the UNMODIFIED reference available to bench.py's reference arm on the GPU box.

The reference (jgslunde/Lanczos) is a tree of plain Python scripts with no build system.  When
/root/reference (or $LANCZOS_REF) is present - in the authoring container; the GPU box only sees
what this recipe produced - the three files the hot path lives in are copied byte for byte into
oracle/_ref/ :

    Python/Regular/Lanczos.py  Python/Irregular/IrrLanczos.py  Python/Regular/Hamiltonian.py

oracle/_ref/ is listed in .gitignore (reference sources never enter this repository's history) and
not in .gpurunignore (so it travels to the GPU box like the built .so files).  A MANIFEST.json with
the sha256 of every file is written next to them.  `load()` imports the copies with cupy / cupyx /
matplotlib stubbed (they are imported at module top, Lanczos.py:3-5, and are not installed here);
only the use_cuda=False branch can run.

Test infrastructure: only tests/, __graft_entry__ and bench.py's CPU-baseline / reference legs use this.
"""
import hashlib
import json
import os
import shutil
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
FILES = ["Python/Regular/Lanczos.py", "Python/Irregular/IrrLanczos.py", "Python/Regular/Hamiltonian.py"]


def reference_root():
    for cand in (os.environ.get("LANCZOS_REF"), "/root/reference"):
        if cand and os.path.isdir(os.path.join(cand, "Python", "Regular")):
            return cand
    return None


def build(verbose=True):
    """Copy the reference files into oracle/_ref/ when the reference tree is present.  Returns the
    manifest (dict) or None when neither the tree nor an earlier copy exists."""
    root = reference_root()
    if root is None:
        return manifest()
    os.makedirs(DEST, exist_ok=True)
    man = {"source": root, "files": {}}
    for rel in FILES:
        src = os.path.join(root, rel)
        dst = os.path.join(DEST, os.path.basename(rel))
        shutil.copyfile(src, dst)
        man["files"][os.path.basename(rel)] = {"from": rel, "sha256": hashlib.sha256(open(dst, "rb").read()).hexdigest()}
    with open(os.path.join(DEST, "MANIFEST.json"), "w") as f:
        json.dump(man, f, indent=1)
    if verbose:
        print("oracle/_ref:", ", ".join(sorted(man["files"])))
    return man


def manifest():
    try:
        with open(os.path.join(DEST, "MANIFEST.json")) as f:
            return json.load(f)
    except Exception:
        return None


def available():
    man = manifest()
    return bool(man) and all(os.path.exists(os.path.join(DEST, n)) for n in man["files"])


def _stub_modules():
    cupy = types.ModuleType("cupy")
    cupy.ndarray = type("ndarray", (), {})
    cupyx = types.ModuleType("cupyx")
    cupyx_scipy = types.ModuleType("cupyx.scipy")
    cupyx_sparse = types.ModuleType("cupyx.scipy.sparse")
    cupyx.scipy = cupyx_scipy
    cupyx_scipy.sparse = cupyx_sparse
    mpl = types.ModuleType("matplotlib")
    plt = types.ModuleType("matplotlib.pyplot")
    mpl.pyplot = plt
    for name, mod in [("cupy", cupy), ("cupyx", cupyx), ("cupyx.scipy", cupyx_scipy),
                      ("cupyx.scipy.sparse", cupyx_sparse), ("matplotlib", mpl), ("matplotlib.pyplot", plt)]:
        sys.modules.setdefault(name, mod)


def load():
    """(Lanczos module, IrrLanczos module) of the vendored reference, or None when it is absent.
    Imported under private names so that the repo's own Python/Regular/Lanczos.py shim is untouched."""
    if not available():
        return None
    import importlib.util
    _stub_modules()
    mods = []
    for name in ("Lanczos", "IrrLanczos"):
        spec = importlib.util.spec_from_file_location("_lanczos_ref_" + name, os.path.join(DEST, name + ".py"))
        m = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(m)
        mods.append(m)
    return tuple(mods)


if __name__ == "__main__":
    build()
