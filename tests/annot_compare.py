"""Compare two annotation files (``runner.py:391-435`` format).

``exact=True``: byte equality (the CPU tests: oracle statistics agree with the reference's to ~1e-12, no
rounding boundary is hit in the committed cases).  ``exact=False``: the structure (lines, labels, tree
shape, states listed) must agree and every printed number within one unit of its last printed digit plus
``tol`` -- for the GPU run, whose posteriors differ from the reference's in the 12th digit and may
therefore round a printed 2-digit probability the other way at a tie."""
import re


def _entry(text):
    """'0' -> {'0': 1.0};  '0:0.3/-1:0.5/7:0.2' -> {'0': .3, '-1': .5, '7': .2}"""
    if ':' not in text:
        return {text: 1.0}
    out = {}
    for item in text.split('/'):
        if item == '':
            continue
        k, v = item.rsplit(':', 1)
        out[k] = float(v)
    return out


def _newick_numbers(nwk):
    nums = [float(x) for x in re.findall(r":([0-9.eE+-]+)", nwk)]
    return re.sub(r":[0-9.eE+-]+", ":#", nwk), nums


def assert_annotations_match(got, ref, exact=True, tol=0.0101):
    if exact:
        if got != ref:
            g, r = got.splitlines(), ref.splitlines()
            assert len(g) == len(r), "line counts differ: %d vs %d" % (len(g), len(r))
            for i, (a, b) in enumerate(zip(g, r)):
                assert a == b, "line %d differs:\n got %s\n ref %s" % (i + 1, a[:400], b[:400])
        return
    g, r = got.splitlines(), ref.splitlines()
    assert len(g) == len(r), "line counts differ: %d vs %d" % (len(g), len(r))
    for i, (a, b) in enumerate(zip(g, r)):
        if a == b:
            continue
        if b.lstrip().startswith(("(", "[")):
            sa, na = _newick_numbers(a); sb, nb = _newick_numbers(b)
            assert sa == sb, "tree line %d: structure / labels differ" % (i + 1)
            assert all(abs(x - y) <= 0.0011 for x, y in zip(na, nb)), "tree line %d: lengths differ" % (i + 1)
            continue
        fa, fb = a.split(','), b.split(',')
        assert fa[0] == fb[0] and len(fa) == len(fb), "line %d: label / site count differ" % (i + 1)
        for j, (x, y) in enumerate(zip(fa[1:], fb[1:])):
            if x == y:
                continue
            ex, ey = _entry(x), _entry(y)
            for k in set(ex) | set(ey):
                assert abs(ex.get(k, 0.0) - ey.get(k, 0.0)) <= tol, \
                    "line %d site %d state %s: %r vs %r" % (i + 1, j, k, x, y)
