"""Golden posterior-annotation files produced by the REFERENCE's own code.

The annotation writer lives inside ``run_from_namespace`` (``laml_libs/runner.py:367-435``), which
cannot be imported here (jax / scikit-bio).  This script reads that block of source lines from
/root/reference at generation time, wraps it in a function and executes it on the reference's
``EM_solver`` (imported verbatim through ref_loader.py) -- nothing of it is copied into the repo.
Outputs: ``annot_<case>.txt`` next to this file; inputs are the committed ``ref_*.npz`` cases.

    python tests/golden/make_annotation_golden.py
"""
import os
import re
import sys
import textwrap

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import ref_loader  # noqa: E402
from helpers import load_npz_case  # noqa: E402

CASES = {
    # name -> (npz case, strip internal labels?)
    "small": ("ref_rand_small", False),
    "ragged_unlabelled": ("ref_rand_ragged", True),
    "forest": ("ref_rand_forest", False),
    "forest_unlabelled": ("ref_rand_forest", True),
    "zeroprior": ("ref_rand_zeroprior", True),
    "nu0": ("ref_rand_nu0", True),
    "test4_interior": ("ref_test4_interior", False),
}
POLYTOMY = {
    "trees": ["((a:0.3,b:0.2,c:0.4,d:0.1):0.2,(e:0.1,f:0.3):0.1,g:0.5):0.05;"],
    "msa": {"a": [1, 0, 3], "b": [1, 2, "?"], "c": [0, 2, 3], "d": [1, "?", -1], "e": [0, 0, 4], "f": [2, 0, 4],
            "g": ["?", 5, 0]},
    "Q": [{0: 0.0, 1: 2.0, 2: 1.0}, {2: 0.5, 5: 0.25, 7: 0.25}, {3: 1.0, 4: 3.0}],
    "phi": 0.1, "nu": 0.2,
}


def strip_internal_labels(newick):
    return re.sub(r"\)[^:;,()\[\]]+", ")", newick)


def reference_writer(ref):
    """the annotation block of runner.py as a callable"""
    src = open(os.path.join(ref.root, "laml_libs", "runner.py")).read().splitlines()
    lo = next(i for i, l in enumerate(src) if l.strip().startswith("# annotations"))
    hi = next(i for i, l in enumerate(src) if l.strip().startswith("# params file"))
    body = textwrap.dedent("\n".join(src[lo:hi]))
    fn = "def _annotate(EM_solver, opt_trees, msa, Q, opt_params, out_annotate):\n" + textwrap.indent(body, "    ")
    ns = {}
    exec(compile(fn, "runner.py[%d:%d]" % (lo + 1, hi), "exec"), ns)
    return lambda *a: ns["_annotate"](ref.EM_solver, *a)


def main():
    ref = ref_loader.load()
    write = reference_writer(ref)
    for name, (npz, strip) in CASES.items():
        case, z = load_npz_case(npz)
        trees = [strip_internal_labels(t) if strip else t for t in case.newicks]
        out = os.path.join(HERE, "annot_%s.txt" % name)
        write(trees, case.msa, case.Q, {"phi": case.phi, "nu": case.nu}, out)
        print(name, os.path.getsize(out), "bytes")
    out = os.path.join(HERE, "annot_polytomy.txt")
    write(POLYTOMY["trees"], POLYTOMY["msa"], POLYTOMY["Q"], {"phi": POLYTOMY["phi"], "nu": POLYTOMY["nu"]}, out)
    print("polytomy", os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
