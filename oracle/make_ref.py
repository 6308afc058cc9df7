"""Build a runnable copy of the *literal* reference under ``oracle/_ref/pyvb``.

TEST INFRASTRUCTURE ONLY.  Nothing in ``pyvb_b200/`` imports this.

The reference (``/root/reference/src/pyvb``) is Python-2 only.  This script
copies its six importable files and applies the purely syntactic py3 patch
listed in SURVEY.md Appendix D -- relative imports, ``print`` calls, ``raise``
syntax and one tuple in a comprehension.  No arithmetic is touched.  The output
directory is git-ignored (never committed) but *does* travel to the GPU box
with the gpurun snapshot, where it serves as the ``--impl reference`` arm of
``bench.py``.

Patched sites (reference file:line):
  src/pyvb/__init__.py:2-3            implicit relative imports
  src/pyvb/network.py:2,43,51,54,96   import + print statements
  src/pyvb/nodes/__init__.py:2-5      implicit relative imports
  src/pyvb/nodes/gaussian.py:4-5,52,58,61
  src/pyvb/nodes/nodes_todo.py:5
"""
import os
import re
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEFAULT_SRC = "/root/reference/src/pyvb"
DEFAULT_DST = os.path.join(HERE, "_ref", "pyvb")

FILES = [
    "__init__.py",
    "network.py",
    os.path.join("nodes", "__init__.py"),
    os.path.join("nodes", "node.py"),
    os.path.join("nodes", "gaussian.py"),
    os.path.join("nodes", "nodes_todo.py"),
]


def _patch(rel, text):
    if rel == "__init__.py":
        text = text.replace("import nodes", "from . import nodes")
        text = text.replace("from network import Network", "from .network import Network")
    elif rel == "network.py":
        text = text.replace("from nodes import *", "from .nodes import *")
        # the four py2 print statements
        text = re.sub(r"^(\s*)print (.+)$", r"\1print(\2)", text, flags=re.M)
    elif rel == os.path.join("nodes", "__init__.py"):
        text = text.replace("from gaussian import", "from .gaussian import")
        text = text.replace("from node import", "from .node import")
        text = text.replace("from nodes_todo import", "from .nodes_todo import")
    elif rel == os.path.join("nodes", "gaussian.py"):
        text = text.replace("from node import *", "from .node import *")
        text = text.replace("from nodes_todo import *", "from .nodes_todo import *")
        text = re.sub(r"raise ConjugacyError,(.+)$", r"raise ConjugacyError(\1)", text, flags=re.M)
        text = text.replace(
            "for e in Gamma,DiagonalGamma,Wishart,Constant]",
            "for e in (Gamma,DiagonalGamma,Wishart,Constant)]",
        )
    elif rel == os.path.join("nodes", "nodes_todo.py"):
        text = text.replace("import node\n", "from . import node\n")
    return text


def make_ref(src=DEFAULT_SRC, dst=DEFAULT_DST, quiet=False):
    """Returns True when ``dst`` holds an importable translated reference."""
    if not os.path.isdir(src):
        if not quiet:
            print("make_ref: %s not present; keeping existing %s" % (src, dst))
        return os.path.isfile(os.path.join(dst, "__init__.py"))
    if os.path.isdir(dst):
        shutil.rmtree(dst)
    os.makedirs(os.path.join(dst, "nodes"))
    for rel in FILES:
        with open(os.path.join(src, rel), "r", encoding="utf-8") as f:
            text = f.read()
        with open(os.path.join(dst, rel), "w", encoding="utf-8") as f:
            f.write(_patch(rel, text))
    if not quiet:
        print("make_ref: wrote %d files to %s" % (len(FILES), dst))
    return True


REF_MODULE = "pyvb_literal_reference"


def import_ref():
    """Import the translated reference (or return None).  It is loaded under the module name ``pyvb_literal_reference`` --
    NOT ``pyvb``: the repo's own top-level ``pyvb`` package is the drop-in alias of pyvb_b200, and the two must be able to
    live in one process (the parity tests build the same graph with both)."""
    pkg = os.path.join(HERE, "_ref", "pyvb")
    if not os.path.isfile(os.path.join(pkg, "__init__.py")):
        return None
    if REF_MODULE in sys.modules:
        return sys.modules[REF_MODULE]
    import importlib.util
    import warnings

    warnings.simplefilter("ignore")
    spec = importlib.util.spec_from_file_location(REF_MODULE, os.path.join(pkg, "__init__.py"),
                                                  submodule_search_locations=[pkg])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[REF_MODULE] = mod
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    ok = make_ref()
    sys.exit(0 if ok else 1)
