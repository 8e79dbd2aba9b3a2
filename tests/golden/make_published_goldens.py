"""Goldens from the reference's PUBLISHED example2 run (``examples/out_example2``: LAML 1.0.0, MOSEK, treeswift,
topology search, de-duplication): the optimised trees, phi, nu, the mutation rate, the negative log-likelihood
1124.4446364799137 and the posterior annotation file the real pipeline wrote.

* ``ref_example2_published.npz`` -- the de-duplicated problem the published likelihood refers to (62 cells; the
  duplicates and the dummy nodes ``add_dup`` re-inserted, ``sequence_lib.py:54-84``, removed again), same layout
  as the other ``ref_*.npz`` with statistics dumped from the reference's EM_solver;
* ``example2_annotation_case.json`` -- the inputs of the annotation step (``runner.py:391-435``): the published
  trees in mutation units (time-scaled lengths x mutation rate), the full character matrix, the raw prior;
* ``annot_example2_published.txt`` -- the published ``example2_annotations.txt``, verbatim;
* ``annot_example1_published.txt`` -- likewise for ``examples/out_example1`` (LAML 0.0.2; inputs: ``ref_example1_optimum.npz``).

    python tests/golden/make_published_goldens.py
"""
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as G  # noqa: E402

R = G.R
PUBLISHED = dict(nllh=1124.4446364799137, phi=0.008240523490840344, nu=0.009490727450532633,
                 mutation_rate=0.23302478040012137)          # examples/out_example2/example2_params.txt


# the other published example2 runs (examples/<dir>/<stem>_params.txt): (fixture name, dir/stem, nllh, phi, nu, mutation rate)
OTHER_RUNS = [
    ("ref_example2_api", "out_example2_api/example2_LAMLapi", 1168.3841303822578, 0.008091323558598398,
     0.010748193893885536, 0.10737189112228884),                      # run_laml_api.py, ultrametric (default), MOSEK
    ("ref_example2_fastem", "out_example2_fastEM_cpu/example2_fastEM_cpu", 1168.312849021511, 0.008109886705275154,
     0.010683326841117283, 0.10736286736761827),                      # --solver fastEM-cpu, ultrametric
    ("ref_example2_fastem_nostart", "out_example2_fastEM_cpu_nostart/example2_fastEM_cpu_nostart", 1124.4461061331122,
     0.008241383472221163, 0.009487736331438177, 0.23302487862478943),   # fastEM-cpu with topology search from an NJ tree
]


def undo_add_dup(tree):
    """drop the duplicates ``add_dup`` re-inserted (sequence_lib.py:54-84); the representative gets its branch back"""
    dummies = [nd for nd in tree.traverse_preorder() if nd.label and str(nd.label).startswith("node_dummy_")]
    drop = []
    for d in dummies:
        rep = str(d.label).split("node_dummy_")[1]
        drop += [c.label for c in d.children if str(c.label) != rep]
    pruned = tree.extract_tree_without(drop) if drop else tree
    for d in dummies:
        rep = str(d.label).split("node_dummy_")[1]
        pruned.label_to_node()[rep].edge_length -= 0.005
    return pruned, drop


def other_runs(msa, Q):
    from laml_b200.tree import read_tree_newick
    ex = os.path.join(R.root, "examples")
    rec = {}
    for name, stem, nllh, phi, nu, lam in OTHER_RUNS:
        tree = read_tree_newick(open(os.path.join(ex, stem + "_trees.nwk")).read().strip())
        for nd in tree.traverse_preorder():
            if nd.edge_length is not None:
                nd.edge_length *= lam
        annot_newick = tree.newick()
        pruned, drop = undo_add_dup(tree)
        msa2 = {k: v for k, v in msa.items() if k not in drop}
        ref = G.save_case(name, [pruned.newick()], msa2, Q, phi, nu, full=False)
        assert abs(ref["nllh"] - nllh) <= 1e-12 * nllh, (name, ref["nllh"], nllh)
        shutil.copyfile(os.path.join(ex, stem + "_annotations.txt"), os.path.join(HERE, "annot_%s_published.txt" % name[4:]))
        rec[name] = dict(source="examples/" + stem + "_params.txt", nllh=nllh, phi=phi, nu=nu, mutation_rate=lam,
                         annotation_trees=[annot_newick], annotation="annot_%s_published.txt" % name[4:])
    json.dump(rec, open(os.path.join(HERE, "published_runs.json"), "w"), indent=1)


def main():
    from laml_b200.tree import read_tree_newick
    sl = R.sequence_lib
    ex = os.path.join(R.root, "examples")
    msa, _ = sl.read_sequences(os.path.join(ex, "example2", "character_matrix.csv"), filetype="charMtrx", delimiter=",",
                               masked_symbol="?")
    Q = sl.read_priors(os.path.join(ex, "example2", "priors.csv"), msa)
    tree = read_tree_newick(open(os.path.join(ex, "out_example2", "example2_trees.nwk")).read().strip())
    for nd in tree.traverse_preorder():
        if nd.edge_length is not None:
            nd.edge_length *= PUBLISHED["mutation_rate"]
    annot_newick = tree.newick()
    # undo add_dup: drop the re-inserted duplicates; the representative gets its branch back
    dummies = [nd for nd in tree.traverse_preorder() if nd.label and str(nd.label).startswith("node_dummy_")]
    drop = []
    for d in dummies:
        rep = str(d.label).split("node_dummy_")[1]
        drop += [c.label for c in d.children if str(c.label) != rep]
    pruned = tree.extract_tree_without(drop)
    for d in dummies:
        rep = str(d.label).split("node_dummy_")[1]
        pruned.label_to_node()[rep].edge_length -= 0.005          # add_dup hung the representative at dmin below the dummy
    msa2 = {k: v for k, v in msa.items() if k not in drop}
    ref = G.save_case("ref_example2_published", [pruned.newick()], msa2, Q, PUBLISHED["phi"], PUBLISHED["nu"], full=False)
    assert abs(ref["nllh"] - PUBLISHED["nllh"]) <= 1e-12 * PUBLISHED["nllh"], ref["nllh"]
    json.dump(dict(source="examples/out_example2 (LAML 1.0.0)", published=PUBLISHED, trees=[annot_newick],
                   msa={k: [x for x in v] for k, v in msa.items()}, Q=[{str(k): v for k, v in q.items()} for q in Q],
                   dropped_duplicates=drop),
              open(os.path.join(HERE, "example2_annotation_case.json"), "w"))
    shutil.copyfile(os.path.join(ex, "out_example2", "example2_annotations.txt"), os.path.join(HERE, "annot_example2_published.txt"))
    # example1 (LAML 0.0.2): its inputs are ref_example1_optimum.npz (make_golden.py); the published file, verbatim
    shutil.copyfile(os.path.join(ex, "out_example1", "example1_annotations.txt"), os.path.join(HERE, "annot_example1_published.txt"))
    other_runs(msa, Q)
    print("ok", ref["nllh"], "dropped", drop)


if __name__ == "__main__":
    main()
