#!/usr/bin/env python
"""Compare, kernel by kernel, the SASS of two builds of laml_b200/csrc/pmm_engine.cu (no GPU needed).

    python scripts/sass_diff.py <git-rev> [<git-rev2> | WORKTREE]

Each revision's csrc/ + include/ is extracted to a scratch directory and compiled to a cubin for sm_100a; kernels are
matched by demangled name (template arguments stripped as a fallback, so that a kernel that became a template
instance is still paired) and their instruction streams compared after removing addresses and encodings.  Used to
show that code written after a round's GPU budget is spent leaves the GPU-verified kernels as they were measured."""
import difflib
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NVCC = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-cubin"]


def cubin_of(rev, tmp):
    out = os.path.join(tmp, rev.replace("/", "_") + ".cubin")
    if rev == "WORKTREE":
        src = os.path.join(ROOT, "laml_b200", "csrc")
    else:
        top = os.path.join(tmp, "src_" + rev)
        os.makedirs(top)
        tar = subprocess.run(["git", "-C", ROOT, "archive", rev, "laml_b200/csrc", "include"], check=True, capture_output=True).stdout
        subprocess.run(["tar", "-x", "-C", top], input=tar, check=True)
        src = os.path.join(top, "laml_b200", "csrc")
    subprocess.run(NVCC + ["-o", out, "pmm_engine.cu"], cwd=src, check=True)
    return out


def strip_params(name):
    end = name.rfind(")")
    if end < 0:
        return name
    depth = 0
    for i in range(end, -1, -1):
        depth += (name[i] == ")") - (name[i] == "(")
        if depth == 0:
            return name[:i]
    return name


def kernels_of(cubin):
    text = subprocess.run(["cuobjdump", "-sass", cubin], check=True, capture_output=True, text=True).stdout
    fns, cur = {}, None
    for line in text.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = subprocess.run(["cu++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            cur = strip_params(cur)                            # name (+ template arguments), no parameter list
            fns[cur] = []
        elif cur and re.match(r"\s+/\*[0-9a-f]{4,}\*/", line):
            ins = re.sub(r"/\*.*?\*/", "", line).strip()
            fns[cur].append(re.sub(r"0x[0-9a-f]+", "ADDR", ins) if re.search(r"\b(CALL|BRA|BSSY|BSYNC|JMP)\b", ins) else ins)
    return fns


def main():
    a_rev = sys.argv[1]
    b_rev = sys.argv[2] if len(sys.argv) > 2 else "WORKTREE"
    with tempfile.TemporaryDirectory() as tmp:
        a, b = kernels_of(cubin_of(a_rev, tmp)), kernels_of(cubin_of(b_rev, tmp))
    base = lambda k: re.sub(r"<.*", "", re.sub(r"^void ", "", k))
    notes = []
    print("| kernel (%s -> %s) | instructions | result |" % (a_rev, b_rev))
    print("|---|---|---|")
    for k in sorted(set(a) | set(b)):
        if k in a and k in b:
            ka = kb = k
        elif k in b and k not in a:
            cands = [x for x in a if x not in b and base(x) == base(k)]
            if not cands:
                print("| `%s` | %d | new |" % (k, len(b[k])))
                continue
            ka, kb = cands[0], k
        else:
            if not any(base(x) == base(k) for x in b if x not in a):
                print("| `%s` | %d | removed |" % (k, len(a[k])))
            continue
        if a[ka] == b[kb]:
            res = "identical"
        else:
            d = [l for l in difflib.unified_diff(a[ka], b[kb], lineterm="", n=0) if l[:1] in "+-" and l[:3] not in ("+++", "---")]
            regs = lambda l: re.sub(r"\b(U?R|U?P)\d+\b", r"\1", l[1:])
            minus, plus = sorted(regs(l) for l in d if l[0] == "-"), sorted(regs(l) for l in d if l[0] == "+")
            res = ("same instructions, %d with other register names" % len(minus)) if minus == plus else "%d lines differ" % len(d)
            if minus != plus and len(d) <= 24:
                notes.append("`%s`:\n\n```\n%s\n```" % (kb, "\n".join(d)))
        name = kb if ka == kb else "%s (was `%s`)" % (kb, ka)
        print("| `%s` | %d | %s |" % (name, len(b[kb]), res))
    for note in notes:
        print("\n" + note)


if __name__ == "__main__":
    main()
