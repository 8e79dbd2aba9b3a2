"""BASELINE config 1 end to end: the reference's OWN runner (``laml_libs/runner.py: run_from_namespace``, what
``run_laml`` and ``api.run_laml_infer`` call) with the patch INTEGRATION.md section 1 describes, on
``examples/example1`` with the command line of the published run (``examples/out_example1/example1.log``).

The runner's source is read from /root/reference at test time, the INTEGRATION.md hunks are applied to the text in
memory (nothing of it is written anywhere), and the result is executed with the modules the runner imports but this
path never calls (jax, the JAX solver, the NJ starting tree, the multiprocessing search) stubbed out.  Skips where
the reference tree is absent (the GPU box).  CPU only: the engine is the oracle-backed stand-in."""
import os
import re
import sys
import types

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import ref_loader  # noqa: E402

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="reference tree not present")

import laml_b200.solver as solver_mod  # noqa: E402
from oracle_engine import OracleEngine  # noqa: E402


def integration_patch(src: str, search: str = "reference") -> str:
    """the hunks of INTEGRATION.md section 1, as text substitutions that fail loudly if the runner changes;
    ``search``: keep the reference's Topology_search class ("reference"), or the optional hunk that swaps in
    GPU_Topology_search ("batched" / "batched_ultra")"""
    def sub(old, new, count=1):
        nonlocal src
        assert src.count(old) >= 1, "runner.py no longer contains: " + old
        src = src.replace(old, new, count)

    # hunk 1 (runner.py:161): one more solver name
    sub('if solver_str not in {"fastem-cpu", "fastem-gpu", "em", "scipy"}:',
        'if solver_str not in {"fastem-cpu", "fastem-gpu", "em", "scipy", "em-b200"}:')
    # hunk 2 (runner.py:213): pick the class
    sub('    elif solver_str != "em":\n        selected_solver = ML_solver\n        em_selected = False',
        '    elif solver_str == "em-b200":\n        from laml_b200.solver import GPU_EM_solver\n'
        '        selected_solver = GPU_EM_solver\n        em_selected = True\n'
        '    elif solver_str != "em":\n        selected_solver = ML_solver\n        em_selected = False')
    # hunk 3 (runner.py:391-435): the annotation block becomes one call ...
    a = src.index("    my_solver = EM_solver(opt_trees, {'charMtrx': msa}")
    b = src.index("    # params file")
    src = src[:a] + ("    from laml_b200.annotate import annotate\n"
                     "    solved_trees = annotate(opt_trees, msa, Q, opt_params, out_annotate)\n\n") + src[b:]
    # ... which also returns what the runner read off my_solver.trees (runner.py:450)
    sub("    solved_trees = [t.newick() for t in my_solver.trees]\n", "")
    # hunk 4 (runner.py:99-119): the MOSEK licence is needed by the solvers that call MOSEK
    a = src.index("    # 2) MOSEK license checks")
    b = src.index("    # 3) Inputs and basic validation")
    block = src[a:b]
    guarded = '    if str(canon["solver"]).lower() != "em-b200":\n' + "".join(
        ("    " + line if line.strip() else line) for line in block.splitlines(keepends=True))
    src = src[:a] + guarded + src[b:]
    if search != "reference":
        # optional hunk (runner.py:246-247): the search class
        sub("    myTopoSearch = Topology_search(input_trees, selected_solver, data=data, prior=prior, params=params)\n",
            "    from laml_b200.topology import GPU_Topology_search\n"
            "    myTopoSearch = GPU_Topology_search(input_trees, selected_solver, data=data, prior=prior, params=params%s)\n"
            % (", batch_ultra=True" if search == "batched_ultra" else ""))
    return src


@pytest.fixture
def runner(monkeypatch, request):
    search = getattr(request, "param", "reference")
    R = ref_loader.load()
    monkeypatch.setattr(solver_mod, "Engine", OracleEngine)
    pkg = sys.modules["laml_libs"]
    init_src = open(os.path.join(R.root, "laml_libs", "__init__.py")).read()
    for name in ("PROGRAM_NAME", "PROGRAM_VERSION"):
        monkeypatch.setattr(pkg, name, re.search(name + r'\s*=\s*"([^"]*)"', init_src).group(1), raising=False)
    from laml_libs import core_constants
    monkeypatch.setattr(pkg, "core_constants", core_constants, raising=False)
    # imported by the runner, never called on this path
    stubs = {"jax": dict(__version__="stub"),
             "laml_libs.fastEM_solver": dict(fastEM_solver=None, parse_data=None, parse_tree=None),
             "laml_libs.Topology_search_parallel": dict(Topology_search_parallel=None),
             "laml_libs.starting_tree": dict(build_starting_tree=None)}
    for name, attrs in stubs.items():
        if name not in sys.modules:
            mod = types.ModuleType(name)
            for k, v in attrs.items():
                setattr(mod, k, v)
            monkeypatch.setitem(sys.modules, name, mod)
    src = integration_patch(open(os.path.join(R.root, "laml_libs", "runner.py")).read(), search)
    mod = types.ModuleType("laml_libs_runner_patched")
    exec(compile(src, "runner.py+INTEGRATION.md", "exec"), mod.__dict__)
    monkeypatch.delenv("MOSEKLM_LICENSE_FILE", raising=False)
    monkeypatch.setitem(sys.modules, "laml_libs.runner", mod)       # what laml_libs/api.py imports
    monkeypatch.delitem(sys.modules, "laml_libs.api", raising=False)
    stdout = sys.stdout
    yield mod, R
    sys.modules.pop("laml_libs.api", None)
    sys.stdout = stdout                       # the runner swaps stdout for its tee logger


@pytest.mark.parametrize("runner", ["reference", "batched"], indirect=True)
def test_example1_through_the_reference_runner(runner, tmp_path):
    mod, R = runner
    ex = os.path.join(R.root, "examples", "example1")
    pub = os.path.join(R.root, "examples", "out_example1")
    prefix = str(tmp_path / "example1")
    stdout = sys.stdout
    try:
        tree, params, ann_path, _, log_path = mod.run_from_namespace(dict(
            character_matrix=os.path.join(ex, "character_matrix.csv"), tree_topology=os.path.join(ex, "starting.tree"),
            priors=os.path.join(ex, "priors.csv"), output=prefix, nInitials=1, timescale=10, solver="em-b200",
            randseeds="9697", verbose=False))
    finally:
        if hasattr(sys.stdout, "log"):
            sys.stdout.log.close()
        sys.stdout = stdout
    log = open(prefix + ".log").read()
    # the published run's seed: the same initial phi and nu, digit for digit (its initial likelihood is that of
    # LAML 0.0.2's branch-length start; the one printed here is the live reference's, v1.0.5, from the same seed)
    assert "Initial point 1. Random seed: 9697" in log
    assert "Initial phi: 0.6552157923672384. Initial nu: 0.5217815084485568. Initial nllh: " in log
    from random import seed
    msa, site_names = R.sequence_lib.read_sequences(os.path.join(ex, "character_matrix.csv"), filetype="charMtrx",
                                                    delimiter=",", masked_symbol="?")
    Q = R.sequence_lib.read_priors(os.path.join(ex, "priors.csv"), msa, site_names=site_names)
    ref = R.EM_solver([l.strip() for l in open(os.path.join(ex, "starting.tree"))], {'charMtrx': msa}, {'Q': Q},
                      {'nu': 1e-10, 'phi': 1e-10})
    seed(a=9697)
    ref.x2params(ref.ini_all())
    ini = float(re.search(r"Initial nllh: ([0-9.e+-]+)", log).group(1))
    assert abs(ini - ref.negative_llh()) <= 1e-12 * ini
    assert "Optimization by EM algorithm" in log and "Numerical optimization finished successfully" in log
    # the published optimum (MOSEK M-steps, 40 EM iterations)
    pub_params = dict(line.strip().split(": ", 1) for line in open(os.path.join(pub, "example1_params.txt")))
    got_params = dict(line.strip().split(": ", 1) for line in open(prefix + "_params.txt"))
    pub_nllh = float(pub_params["Negative-llh"].strip("()").split(",")[0])
    got_nllh = float(got_params["Negative-llh"].strip("()").split(",")[0])
    assert "'optimal'" in got_params["Negative-llh"]
    assert abs(got_nllh - pub_nllh) <= 1e-6 * pub_nllh
    # two EM runs from different starts (0.0.2's and 1.0.5's initial lengths) stop 1e-8-relative likelihood steps
    # apart on a flat optimum: parameters agree to ~1e-4 (observed: phi 2.5e-5, nu 9.4e-5, rate 1.1e-4 relative),
    # and the likelihood reached here is the lower one by 2.5e-4
    assert abs(float(got_params["Dropout rate"]) - float(pub_params["Dropout rate"])) < 2e-4
    assert abs(float(got_params["Silencing rate"]) - float(pub_params["Silencing rate"])) < 2e-4
    assert abs(float(got_params["Mutation rate"]) - float(pub_params["Mutation rate"])) < 3e-4 * float(pub_params["Mutation rate"])
    assert abs(params["nll"][0] - pub_nllh) <= 1e-6 * pub_nllh and abs(params["lambda"] - float(got_params["Mutation rate"])) < 1e-12
    # output trees: same topology and leaf set as the published file, total height = the requested timescale
    from laml_b200.tree import read_tree_newick
    got_t = read_tree_newick(open(prefix + "_trees.nwk").readline().strip())
    pub_t = read_tree_newick(open(os.path.join(pub, "example1_trees.nwk")).readline().strip())

    def clades(t):
        out, below = set(), {}
        for nd in t.traverse_postorder():
            below[nd] = frozenset([nd.label]) if nd.is_leaf() else frozenset().union(*(below[c] for c in nd.children))
            out.add(below[nd])
        return out
    assert clades(got_t) == clades(pub_t)
    assert abs(got_t.height(weighted=True) - 10.0) < 1e-9
    d_got = {frozenset(k): v for k, v in _lengths_by_clade(got_t).items()}
    d_pub = {frozenset(k): v for k, v in _lengths_by_clade(pub_t).items()}
    diff = np.array([abs(d_got[k] - d_pub[k]) for k in d_pub])
    print("branch lengths vs published (10-unit tree): mean |diff| %.2e, worst %.2e" % (diff.mean(), diff.max()))
    assert diff.mean() < 5e-3 and diff.max() < 0.1        # flat directions of the optimum (see above)
    # annotation file: one tree line + one line per node, and the returned newick is its first line
    ann = open(str(ann_path)).read().splitlines()
    pub_ann = open(os.path.join(pub, "example1_annotations.txt")).read().splitlines()
    assert len(ann) == len(pub_ann) and ann[0] == tree
    assert [l.split(",", 1)[0] for l in ann[1:]] == [l.split(",", 1)[0] for l in pub_ann[1:]]
    same = sum(a == b for a, b in zip(ann[1:], pub_ann[1:]))
    assert same >= 0.9 * (len(ann) - 1)                   # posteriors printed to 2 places at a slightly different optimum
    assert os.path.isfile(prefix + "_trees.collapsed.nwk") and str(log_path) == prefix + ".log"


def _lengths_by_clade(t):
    out, below = {}, {}
    for nd in t.traverse_postorder():
        below[nd] = frozenset([nd.label]) if nd.is_leaf() else frozenset().union(*(below[c] for c in nd.children))
        if nd.edge_length is not None:
            out[below[nd]] = nd.edge_length
    return out


_EXAMPLE2_INITIAL = []


@pytest.mark.parametrize("runner", ["reference", "batched", "batched_ultra"], indirect=True)
def test_example2_topology_search_through_the_reference_runner(runner, tmp_path):
    """the published example2 command line (``--topology_search``, 64 cells of which 2 are duplicates, ultrametric
    constraint on) cut to two NNI iterations: with the reference's own search class driving GPU_EM_solver, and with
    the search-class hunk (candidate by candidate, and neighbourhoods batched under the constraint)"""
    from random import seed
    mod, R = runner
    ex = os.path.join(R.root, "examples", "example2")
    prefix = str(tmp_path / "example2")
    stdout = sys.stdout
    seed(a=4242)
    try:
        tree, params, ann_path, _, log_path = mod.run_from_namespace(dict(
            character_matrix=os.path.join(ex, "character_matrix.csv"), tree_topology=os.path.join(ex, "starting.tree"),
            priors=os.path.join(ex, "priors.csv"), output=prefix, nInitials=1, randomreps=1, topology_search=True,
            timescale=10, solver="em-b200", maxIters=2, verbose=True))
    finally:
        if hasattr(sys.stdout, "log"):
            sys.stdout.log.close()
        sys.stdout = stdout
    log = open(prefix + ".log").read()
    assert "Started with 64 sequences. Deduplication removes 2 sequences." in log and "Starting topology search" in log
    # the score of the starting topology after numerical optimisation under the constraint (the published log's
    # -1126.92 is LAML 1.0.0's unconstrained one; 1.0.5's runner forces ultra_constr, runner.py:91): the same in all
    # three modes, and above the unconstrained optimum
    ini = float(re.search(r"Initial score: (-[0-9.]+)", log).group(1))
    _EXAMPLE2_INITIAL.append(ini)
    assert abs(ini - _EXAMPLE2_INITIAL[0]) <= 1e-6 * abs(ini) and -ini > 1126.9195350702505
    nll = params["nll"]
    assert np.isfinite(nll) and nll <= -ini + 1e-3            # the search never returns worse than its start
    assert "Optimal topology found. Re-optimizing other parameters" in log or "Optimal score for this search" in log
    from laml_b200.tree import read_tree_newick
    out_t = read_tree_newick(open(prefix + "_trees.nwk").readline().strip())
    assert len(list(out_t.traverse_leaves())) == 64           # duplicates are put back (add_dup)
    assert abs(out_t.height(weighted=True) - 10.0) < 1e-9
    ann = open(str(ann_path)).read().splitlines()
    assert ann[0] == tree and len(ann) > 100
    for f in ("_params.txt", "_trees.collapsed.nwk"):
        assert os.path.isfile(prefix + f)


def test_api_run_laml_infer(runner, tmp_path):
    """``laml_libs.api.run_laml_infer`` (the programmatic entry point, api.py:39-131; imported verbatim) on top of
    the patched runner: example1, published command line"""
    mod, R = runner
    from laml_libs.api import LAMLOutput, run_laml_infer
    ex = os.path.join(R.root, "examples", "example1")
    stdout = sys.stdout
    from random import seed
    seed(a=7)
    try:
        out = run_laml_infer(character_matrix=os.path.join(ex, "character_matrix.csv"),
                             tree_topology=os.path.join(ex, "starting.tree"), priors=os.path.join(ex, "priors.csv"),
                             solver="em-b200", nInitials=1, randomreps=1, timescale=10, output_prefix=str(tmp_path / "api"))
    finally:
        if hasattr(sys.stdout, "log"):
            sys.stdout.log.close()
        sys.stdout = stdout
    assert isinstance(out, LAMLOutput) and out.tree_newick.endswith(";")
    # another random start, the same optimum (published: phi 0.14378501771164973, nu 0.13359737967697907, rate 0.1034192687776938)
    assert abs(out.params.phi_dropout - 0.14378501771164973) < 3e-4
    assert abs(out.params.nu_silence - 0.13359737967697907) < 3e-4
    assert abs(out.params.lambda_edit - 0.1034192687776938) < 5e-4 * 0.1034192687776938
    assert out.annotations_path.is_file() and out.log_path.is_file()
    nllh = float(open(str(tmp_path / "api_params.txt")).read().split("Negative-llh: (")[1].split(",")[0])
    assert abs(nllh - 6987.352449704049) < 1e-6 * 6987.35
