"""Import the reference's hot-path modules *verbatim* from /root/reference.

Used only by ``make_golden.py`` (and by ``tests/test_reference_live.py`` when the
reference tree is present).  Nothing that runs on the GPU box imports this.

``import laml_libs`` fails as shipped because ``laml_libs/__init__.py`` pulls in
jax / scikit-bio / cvxpy through ``runner.py``.  The hot-path files themselves
(``ML_solver.py``, ``EM_solver.py``, ``sequence_lib.py``, ``lca_lib.py``) are pure
Python; they load unmodified once
  * ``laml_libs`` is registered as an empty namespace pointing at the reference
    directory (skipping its ``__init__``),
  * ``treeswift`` resolves to ``laml_b200.tree`` (same API subset), and
  * ``cvxpy`` resolves to an empty stub (only ``EM_solver.Mstep`` touches it).
"""
import os
import sys
import types

REF = os.environ.get("LAML_REFERENCE", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REF, "laml_libs", "EM_solver.py"))


def load():
    if not available():
        raise RuntimeError("reference tree not found at %s" % REF)
    repo = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    if repo not in sys.path:
        sys.path.insert(0, repo)
    import laml_b200.tree as tree_mod
    if "treeswift" not in sys.modules:
        ts = types.ModuleType("treeswift")
        for name in ("Node", "Tree", "read_tree_newick"):
            setattr(ts, name, getattr(tree_mod, name))
        ts.__all__ = ["Node", "Tree", "read_tree_newick"]
        sys.modules["treeswift"] = ts
    if "cvxpy" not in sys.modules:
        sys.modules["cvxpy"] = types.ModuleType("cvxpy")
    if "laml_libs" not in sys.modules:
        pkg = types.ModuleType("laml_libs")
        pkg.__path__ = [os.path.join(REF, "laml_libs")]
        sys.modules["laml_libs"] = pkg
    from laml_libs.EM_solver import EM_solver
    from laml_libs.ML_solver import ML_solver
    from laml_libs import sequence_lib
    return types.SimpleNamespace(EM_solver=EM_solver, ML_solver=ML_solver, sequence_lib=sequence_lib, root=REF)


def enable_cvxpy_shim():
    """Give the ``cvxpy`` module object the reference imported (an empty stub, see :func:`load`) the few names
    ``EM_solver.Mstep`` uses, backed by scipy (tests/golden/cvxpy_shim.py).  Works whether or not the reference
    has been imported already: ``EM_solver``'s ``cp`` is this same module object."""
    load()
    import cvxpy_shim
    mod = sys.modules["cvxpy"]
    for name in cvxpy_shim.API:
        setattr(mod, name, getattr(cvxpy_shim, name))
    return mod
