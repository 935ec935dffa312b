"""Staging of the UNMODIFIED reference Python package next to a chosen shared library (shared by the drop-in tests).
The package is taken from /root/reference (build container) or from baseline/_ref/python-package (the copy __graft_entry__.build()
places there — git-ignored, it travels to the GPU box like oracle/_ref); nothing of it is part of the product."""
import json
import os
import subprocess
import sys
import tempfile
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CANDIDATES = ["/root/reference/python-package/gpboost", os.path.join(ROOT, "baseline", "_ref", "python-package", "gpboost")]


def ref_package_dir():
    for c in CANDIDATES:
        if os.path.isdir(c):
            return c
    return None


def stage(lib_path):
    """scratch dir with gpboost/ = the reference's module files + lib_gpboost.so -> lib_path (libpath.py:36 finds it there)"""
    src = ref_package_dir()
    tmp = tempfile.mkdtemp()
    pkg = os.path.join(tmp, "gpboost")
    os.makedirs(pkg)
    for f in os.listdir(src):
        if f.endswith(".py") or f == "VERSION.txt":
            os.symlink(os.path.join(src, f), os.path.join(pkg, f))
    os.symlink(lib_path, os.path.join(pkg, "lib_gpboost.so"))
    return tmp


def run_with(lib_path, body, timeout=900):
    """runs `body` (python source that fills a dict `out`) with `import gpboost as gpb` bound to lib_path; returns out"""
    tmp = stage(lib_path)
    code = textwrap.dedent("""
        import json, os, sys, types
        sys.modules.setdefault("optuna", types.ModuleType("optuna"))   # hard import of the package, not installed here
        sys.path.insert(0, %r)
        import numpy as np
        import gpboost as gpb
        assert gpb.basic._LIB._name.endswith("lib_gpboost.so")
        out = {}
        """ % tmp) + textwrap.dedent(body) + "\nprint('RESULT' + json.dumps(out))\n"
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=timeout)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    line = [ln for ln in r.stdout.split("\n") if ln.startswith("RESULT")][-1]
    return json.loads(line[len("RESULT"):])
