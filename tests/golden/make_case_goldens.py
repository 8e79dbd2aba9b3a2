"""Known answers the reference ships with its case tests (``case_tests/old/test{1,2,3,6}/sample_output.txt``:
final tree, dropout and silencing rate, negative log-likelihood of a stored run), turned into fixtures:
``ref_case_test<k>.npz`` (same layout as the other ``ref_*.npz``; statistics dumped from the reference's
``EM_solver`` imported verbatim) and ``case_tests.json`` (the stored numbers next to what the reference's
``ML_solver`` and ``EM_solver`` give at the stored point here).  SURVEY.md 8c lists these as golden vectors.

    python tests/golden/make_case_goldens.py
"""
import json
import os
import re
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as G  # noqa: E402  (loads the reference through ref_loader)

R = G.R
BASE = os.path.join(R.root, "case_tests", "old")
CASES = {
    1: ("n64_d0_r1_character_matrix.csv", "n64_d0_priors.csv"),
    2: ("n64_d0s32_r1_character_matrix.csv", "n64_d0s32_priors.csv"),
    3: ("n64_d0s32_r1_character_matrix.csv", "n64_d0s32_priors.csv"),
    6: ("n64_d0s32_r1_character_matrix.csv", "n64_d0s32_priors.csv"),
    4: ("n64_d0_r1_character_matrix.csv", "n64_d0_priors.csv"),                   # --resolve_search
    5: ("n64_d0s32_r1_character_matrix.csv", "n64_d0s32_priors.csv"),             # topology search on a polytomous start
    7: ("n64_d0s32_r1_character_matrix.csv", "n64_d0s32_priors.csv"),
    9: ("bar10_character_matrix_without_cluster.csv", "bar10_mutation_prior.csv"),   # real data, polytomies
    10: ("3724_NT_All_dedup_mtx.txt", "3724_NT_All_priors_norm.pkl"),             # real data, 3 724 cells, --ultrametric
    "8": ("character_matrix.csv", "prior_k30_r01.csv", "problin_output.txt"),       # 250 cells, topology search
    "8u": ("character_matrix.csv", "prior_k30_r01.csv", "problin_output_ultr.txt"),  # ... --ultrametric --parallel
}


def parse_output(path):
    text = open(path).read()
    m = re.search(r"Newick tree: (.*)", text) or re.search(r"Newick tree \(s\):\s*\n(.*)", text)
    tree = m.group(1).strip()
    num = r"([-+0-9.eE]+)"
    nllh = float(re.search(r"[Nn]egative-llh: \(?" + num, text).group(1))
    phi = float(re.search(r"[Dd]ropout rate: " + num, text).group(1))
    nu = float(re.search(r"[Ss]ilencing rate: " + num, text).group(1))
    return tree, nllh, phi, nu


def main():
    sl = R.sequence_lib
    rec = {}
    for k, spec in CASES.items():
        mtx, pri = spec[0], spec[1]
        output = spec[2] if len(spec) > 2 else "sample_output.txt"
        d = os.path.join(BASE, "test%s" % str(k).rstrip("u"))
        msa, _ = sl.read_sequences(os.path.join(d, mtx), filetype="charMtrx", delimiter=",", masked_symbol="?")
        Q = sl.read_priors(os.path.join(d, pri), msa)
        tree, nllh, phi, nu = parse_output(os.path.join(d, output))
        ml = R.ML_solver([tree], {"charMtrx": msa}, {"Q": Q}, {"phi": phi, "nu": nu}).negative_llh() if k != 10 else None      # test10: ML_solver's Python recursion is slow at 2 919 nodes; the stored value is EM_solver's
        ref = G.save_case("ref_case_test%s" % k, [tree], msa, Q, phi, nu, full=False)
        rec["test%s" % k] = dict(source="case_tests/old/test%s/%s" % (str(k).rstrip("u"), output), fixture="ref_case_test%s" % k,
                                 stored_nllh=nllh, phi=phi, nu=nu, ml_solver_nllh=ml, em_solver_nllh=ref["nllh"])
        print(k, "stored", nllh, "ML_solver", ml, "EM_solver", ref["nllh"])
    json.dump(rec, open(os.path.join(HERE, "case_tests.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
