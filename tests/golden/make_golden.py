#!/usr/bin/env python
"""Generate the golden vectors under tests/golden/ from the REFERENCE itself.

Run in the container that has /root/reference:

    python tests/golden/make_golden.py

It imports the reference's ``EM_solver`` verbatim (``ref_loader.py``), runs it on
 (a) every known-answer / property test input in
     laml_unit_tests/unit_tests_EMSolver.py (tests 1-46; extracted with ``ast``),
 (b) the reference's shipped fixtures (test_EM/test1, test4; examples/example1 at
     its published optimum; case_tests/old/test1 stored optimum),
 (c) seeded random inputs that exercise ragged alphabets, '?' and literal -1
     entries, nu = 0, phi = 0, zero prior mass, forests of several trees,
and stores inputs + the reference's outputs.  The outputs are what
tests/test_oracle_golden.py pins oracle/pmm_oracle.c against, and what the GPU
parity tests compare with.  The files are small; regenerate only when the
reference version changes.
"""
import ast
import json
import os
import random
import sys
from math import exp

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import ref_loader  # noqa: E402

R = ref_loader.load()
from laml_b200.tree import read_tree_newick  # noqa: E402

NEG = float("-inf")


# ----------------------------------------------------------------------------
def label_internal(newick):
    t = read_tree_newick(newick)
    i = 0
    for nd in t.traverse_preorder():
        if nd.label is None or nd.label == "":
            nd.label = "I%d" % i
            i += 1
    return t.newick()


def run_reference(trees, msa, Q, phi, nu, full=True):
    """Reference EM_solver: az_partition + Estep; returns dict of outputs keyed by
    node label (per tree index)."""
    solver = R.EM_solver(list(trees), {"charMtrx": msa}, {"Q": Q}, {"phi": phi, "nu": nu})
    # the constructor suppresses unifurcations and resolves polytomies
    # (ML_solver.py:30, EM_solver.py:26-43); label whatever is still unlabelled so
    # rows can be matched by name.  Labels do not enter any computation.
    i = 0
    for tree in solver.trees:
        for nd in tree.traverse_preorder():
            if nd.label is None or nd.label == "":
                nd.label = "J%d" % i
                i += 1
    nllh = solver.negative_llh()
    solver.Estep()
    K = solver.numsites
    out = {"nllh": nllh, "labels": [], "tree_of": []}
    names = ["L0", "L1", "out0", "out1", "post0", "post1", "S0", "S1", "S2", "S3", "S4"]
    cols = {k: [] for k in names}
    tags = []
    for ti, tree in enumerate(solver.trees):
        for nd in tree.traverse_postorder():
            out["labels"].append(nd.label)
            out["tree_of"].append(ti)
            for k in names:
                cols[k].append([float(x) for x in getattr(nd, k)])
            tags.append([-1 if a == "z" else (-2 if a == "?" else int(a)) for a in nd.alpha])
    for k in names:
        out[k] = np.array(cols[k], np.float64).reshape(len(out["labels"]), K)
    out["tag"] = np.array(tags, np.int64).reshape(len(out["labels"]), K)
    out["site_llh"] = np.sum([np.array(t.root.L0, np.float64) for t in solver.trees], axis=0)
    # Mstep's phi statistics, EM_solver.py:318-324
    Rn, Rt = 0.0, 0.0
    for tree in solver.trees:
        for v in tree.traverse_leaves():
            Rn += sum([x != "?" for x in solver.charMtrx[v.label]])
            Rt += sum([1 - exp(p) for (j, p) in enumerate(v.post1) if solver.charMtrx[v.label][j] == "?"])
    out["R"], out["Rtilde"] = Rn, Rt
    out["newick_out"] = solver.get_tree_newick()
    return out


def msa_to_array(msa, cells):
    K = len(msa[cells[0]])
    a = np.zeros((len(cells), K), np.int64)
    for r, c in enumerate(cells):
        for s, x in enumerate(msa[c]):
            a[r, s] = -1 if x == "?" else (-2 if x == -1 else int(x))
    return a


def q_to_csr(Q):
    qptr, st, val = [0], [], []
    for q in Q:
        for k in sorted(q):
            st.append(int(k)); val.append(float(q[k]))
        qptr.append(len(st))
    return np.array(qptr, np.int64), np.array(st, np.int64), np.array(val, np.float64)


def save_case(name, trees, msa, Q, phi, nu, full=True):
    ref = run_reference(trees, msa, Q, phi, nu)
    trees = ref["newick_out"]           # as the solver sees them (binary, labelled)
    cells = list(msa.keys())
    qptr, qst, qraw = q_to_csr(Q)
    payload = dict(
        trees=np.array(trees), cells=np.array(cells), matrix=msa_to_array(msa, cells),
        qptr=qptr, qstate=qst, qraw=qraw, phi=phi, nu=nu, nllh=ref["nllh"],
        labels=np.array(ref["labels"]), tree_of=np.array(ref["tree_of"]), site_llh=ref["site_llh"],
        R=ref["R"], Rtilde=ref["Rtilde"],
        Ssum=np.stack([ref["S%d" % k].sum(axis=1) for k in range(5)]),
    )
    if full:
        for k in ["L0", "L1", "out0", "out1", "post0", "post1", "S0", "S1", "S2", "S3", "S4", "tag"]:
            payload[k] = ref[k]
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **payload)
    print("%-28s nllh=%.12f nodes=%d sites=%d" % (name, ref["nllh"], len(ref["labels"]), payload["matrix"].shape[1]))
    return ref


# ----------------------------------------------------------------------------
def extract_unit_tests():
    """Pull the literal inputs of tests 1-46 out of the reference's test file."""
    path = os.path.join(R.root, "laml_unit_tests", "unit_tests_EMSolver.py")
    src = open(path).read()
    mod = ast.parse(src)
    cases = []
    for cls in mod.body:
        if not isinstance(cls, ast.ClassDef):
            continue
        for fn in cls.body:
            if not (isinstance(fn, ast.FunctionDef) and fn.name.startswith("test_")):
                continue
            no = int(fn.name.split("_")[1])
            if no > 46 or no == 45:
                continue
            env = {}
            for st in fn.body:
                if isinstance(st, ast.Assign) and len(st.targets) == 1 and isinstance(st.targets[0], ast.Name):
                    nm = st.targets[0].id
                    if nm in ("Q", "T", "msa", "true_nllh", "phi", "nu"):
                        env[nm] = ast.literal_eval(st.value)
                    elif nm == "mySolver":          # EM_solver([T],{...},{...},{'phi':..,'nu':..})
                        params = ast.literal_eval(st.value.args[3])
                        env["phi"], env["nu"] = params["phi"], params["nu"]
            env["no"] = no
            env["line"] = fn.lineno
            cases.append(env)
    return sorted(cases, key=lambda c: c["no"])


def jsonable_msa(msa):
    return {c: [x for x in row] for c, row in msa.items()}


def jsonable_Q(Q):
    return [{str(k): v for k, v in q.items()} for q in Q]


def unit_test_goldens():
    kat, prop = [], []
    for c in extract_unit_tests():
        rec = dict(no=c["no"], ref_line=c["line"], T=c["T"], Q=jsonable_Q(c["Q"]), msa=jsonable_msa(c["msa"]),
                   phi=c["phi"], nu=c["nu"])
        if "true_nllh" in c:
            s = R.EM_solver([c["T"]], {"charMtrx": c["msa"]}, {"Q": c["Q"]}, {"phi": c["phi"], "nu": c["nu"]})
            rec["true_nllh"] = c["true_nllh"]
            rec["ref_nllh"] = s.negative_llh()
            assert abs(rec["ref_nllh"] - rec["true_nllh"]) < 5e-6, rec
            kat.append(rec)
        else:
            Tl = label_internal(c["T"])
            ref = run_reference([Tl], c["msa"], c["Q"], c["phi"], c["nu"])
            rec["T"] = Tl
            rec["labels"] = ref["labels"]
            for k in ["L0", "L1", "out0", "out1", "post0", "post1", "S0", "S1", "S2", "S3", "S4"]:
                rec[k] = [[x if x != NEG else "-inf" for x in row] for row in ref[k].tolist()]
            rec["nllh"] = ref["nllh"]
            prop.append(rec)
    json.dump(kat, open(os.path.join(HERE, "kat_likelihood.json"), "w"), indent=1)
    json.dump(prop, open(os.path.join(HERE, "outside_cases.json"), "w"), indent=1)
    print("unit-test goldens: %d likelihood KATs, %d outside/posterior cases" % (len(kat), len(prop)))


# ----------------------------------------------------------------------------
def random_tree(n_leaves, rng, zero_len_prob=0.0):
    """Random binary tree by splitting a uniformly chosen current leaf."""
    from laml_b200.tree import Node, Tree
    t = Tree()
    t.root.label = "r"
    leaves = [t.root]
    k = 0
    while len(leaves) < n_leaves:
        nd = leaves.pop(rng.randrange(len(leaves)))
        for _ in range(2):
            c = Node(label="n%d" % k); k += 1
            nd.add_child(c)
            leaves.append(c)
    for i, nd in enumerate(t.traverse_postorder()):
        nd.edge_length = 0.0 if rng.random() < zero_len_prob else round(rng.uniform(0.01, 1.5), 6)
        if nd.is_leaf():
            nd.label = "c%d" % i
    return t


def random_case(seed, n_leaves, K, max_A, phi, nu, p_missing=0.2, p_known_silenced=0.05, n_trees=1,
                zero_prior=False):
    rng = random.Random(seed)
    Q = []
    for s in range(K):
        A = rng.randint(1, max_A)
        w = [rng.random() + 0.01 for _ in range(A)]
        if zero_prior and A > 2 and s % 3 == 0:
            w[rng.randrange(A)] = 0.0
        keys = rng.sample(range(1, 3 * max_A), A)
        q = {k: x for k, x in zip(keys, w)}
        if s % 2 == 0:
            q[0] = 0
        Q.append(q)
    trees = [random_tree(n_leaves, rng) for _ in range(n_trees)]
    # all trees must carry the same leaf set
    names = [nd.label for nd in trees[0].traverse_leaves()]
    for t in trees[1:]:
        for nd, nm in zip(t.traverse_leaves(), names):
            nd.label = nm
    # crude forward simulation on the first tree so that clades share states
    seq = {}
    for nd in trees[0].traverse_preorder():
        par = seq[nd.parent] if nd.parent is not None else [0] * K
        cur = []
        for s in range(K):
            c = par[s]
            if c == 0 and rng.random() < 0.35:
                pos = [k for k in Q[s] if k != 0 and Q[s][k] > 0]
                c = rng.choice(pos)
            cur.append(c)
        seq[nd] = cur
    msa = {}
    for nd in trees[0].traverse_leaves():
        row = []
        for s in range(K):
            u = rng.random()
            if u < p_missing and (phi > 0 or nu > 0):
                row.append("?")
            elif u < p_missing + p_known_silenced and nu > 0:
                row.append(-1)
            else:
                row.append(seq[nd][s])
        msa[nd.label] = row
    return [t.newick() for t in trees], msa, Q


def fixture_cases():
    sl = R.sequence_lib
    base = os.path.join(R.root, "laml_unit_tests", "test_data", "test_EM")
    # test_45's input (unit_tests_EMSolver.py:506-520)
    msa, _ = sl.read_sequences(os.path.join(base, "test1_charMtrx.txt"), filetype="charMtrx", delimiter=",",
                               masked_symbol="-", suppress_warnings=True)
    Q = []
    for i in range(60):
        M_i = set(msa[x][i] for x in msa if msa[x][i] not in [0, "?"])
        q = {x: 1 / len(M_i) for x in M_i}
        q[0] = 0
        Q.append(q)
    T = open(os.path.join(base, "test1.tre")).read().strip()
    save_case("ref_test45_test1", [T], msa, Q, 0.05231954386883335, 0.15877477685098262, full=True)

    # test_47/48's input (unit_tests_EMSolver.py:535-602) at an interior point
    msa, _ = sl.read_sequences(os.path.join(base, "test4_charMtrx.txt"), filetype="charMtrx", delimiter=",",
                               masked_symbol="?")
    Q = sl.read_priors(os.path.join(base, "test4_prior.csv"), msa)
    T = open(os.path.join(base, "test4.tre")).read().strip()
    save_case("ref_test4_interior", [T], msa, Q, 0.07, 0.11, full=True)
    save_case("ref_test4_zero", [T], msa, Q, 0.0, 0.0, full=True)

    # example1 at its published optimum (examples/out_example1/example1_params.txt)
    ex = os.path.join(R.root, "examples")
    msa, _ = sl.read_sequences(os.path.join(ex, "example1", "character_matrix.csv"), filetype="charMtrx",
                               delimiter=",", masked_symbol="?")
    Q = sl.read_priors(os.path.join(ex, "example1", "priors.csv"), msa)
    lam, phi, nu = 0.1034192687776938, 0.14378501771164973, 0.13359737967697907
    tree = read_tree_newick(open(os.path.join(ex, "out_example1", "example1_trees.nwk")).read().strip())
    for nd in tree.traverse_preorder():
        if nd.edge_length is not None:
            nd.edge_length = nd.edge_length * lam
    ref = save_case("ref_example1_optimum", [tree.newick()], msa, Q, phi, nu, full=False)
    published = 6987.352449704049
    assert abs(ref["nllh"] - published) / published < 1e-12, ref["nllh"]
    # the starting tree of example1 (config 1 of BASELINE.json) for EM runs
    # (its branches carry no lengths: give every branch 0.1 so the point is defined)
    t0 = read_tree_newick(open(os.path.join(ex, "example1", "starting.tree")).read().strip())
    for nd in t0.traverse_preorder():
        if nd.edge_length is None:
            nd.edge_length = 0.1
    save_case("ref_example1_start", [t0.newick()], msa, Q, 0.1, 0.1, full=False)


def random_cases():
    specs = [
        ("ref_rand_small", dict(seed=11, n_leaves=9, K=7, max_A=5, phi=0.17, nu=0.23)),
        ("ref_rand_ragged", dict(seed=12, n_leaves=40, K=24, max_A=40, phi=0.17, nu=0.23)),
        ("ref_rand_nu0", dict(seed=13, n_leaves=25, K=12, max_A=8, phi=0.2, nu=0.0)),
        ("ref_rand_phi0", dict(seed=14, n_leaves=25, K=12, max_A=8, phi=0.0, nu=0.3)),
        ("ref_rand_nomissing", dict(seed=15, n_leaves=30, K=10, max_A=6, phi=0.0, nu=0.0)),
        ("ref_rand_zeroprior", dict(seed=16, n_leaves=30, K=12, max_A=9, phi=0.1, nu=0.15, zero_prior=True)),
        ("ref_rand_forest", dict(seed=17, n_leaves=12, K=9, max_A=6, phi=0.12, nu=0.4, n_trees=3)),
        ("ref_rand_deep", dict(seed=18, n_leaves=120, K=16, max_A=30, phi=0.3, nu=0.3, p_missing=0.4)),
    ]
    for name, kw in specs:
        trees, msa, Q = random_case(**kw)
        save_case(name, trees, msa, Q, kw["phi"], kw["nu"], full=True)


def stored_optima():
    """Converged values the reference's own tests / logs pin across the MOSEK
    boundary; inputs for these live in the npz fixtures above."""
    rec = {
        "test_47": dict(source="laml_unit_tests/unit_tests_EMSolver.py:556-566", fixture="ref_test4_zero",
                        nllh=1007.009158767873, phi=0.0, nu=0.0, rel_tol_nllh=1e-4, abs_tol_params=1e-4,
                        randseed=1221),
        "example1": dict(source="examples/out_example1/example1_params.txt:1-4", fixture="ref_example1_optimum",
                         nllh=6987.352449704049, phi=0.14378501771164973, nu=0.13359737967697907,
                         mutation_rate=0.1034192687776938),
    }
    json.dump(rec, open(os.path.join(HERE, "stored_optima.json"), "w"), indent=1)


if __name__ == "__main__":
    unit_test_goldens()
    fixture_cases()
    random_cases()
    stored_optima()
