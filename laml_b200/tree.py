"""Minimal rooted-tree container for the PMM engine's host side.

The reference builds its solvers on the third-party ``treeswift`` package
(``laml_libs/ML_solver.py:1,28-32``), which is not installed in this image.
The hot path only needs a small subset of that API: a newick reader, the three
traversal orders whose *order* fixes the branch<->vector index mapping
(``laml_libs/EM_solver.py:45-51`` uses postorder; ``:150`` preorder), and the
two structural edits the solver constructor performs
(``suppress_unifurcations`` at ``ML_solver.py:30`` and ``resolve_polytomies``
at ``EM_solver.py:35``).  This module provides exactly that subset with the
same method names so host code written against the reference reads the same.
It is written from the documented behaviour, not from
the treeswift sources.
"""
from __future__ import annotations

import os
from collections import deque
from typing import Dict, Iterator, List, Optional


class Node:
    __slots__ = ("label", "edge_length", "children", "parent", "__dict__")

    def __init__(self, label=None, edge_length=None):
        self.label = label
        self.edge_length = edge_length
        self.children: List["Node"] = []
        self.parent: Optional["Node"] = None

    # -- structure ---------------------------------------------------------
    def add_child(self, child: "Node") -> None:
        self.children.append(child)
        child.parent = self

    def remove_child(self, child: "Node") -> None:
        self.children.remove(child)
        child.parent = None

    def child_nodes(self) -> List["Node"]:
        return list(self.children)

    def num_children(self) -> int:
        return len(self.children)

    def is_leaf(self) -> bool:
        return len(self.children) == 0

    def is_root(self) -> bool:
        return self.parent is None

    # -- accessors (treeswift spelling) ------------------------------------
    def get_parent(self):
        return self.parent

    def set_parent(self, parent):
        self.parent = parent

    def get_label(self):
        return self.label

    def set_label(self, label):
        self.label = label

    def get_edge_length(self):
        return self.edge_length

    def set_edge_length(self, length):
        self.edge_length = length

    # -- traversals --------------------------------------------------------
    def traverse_preorder(self, leaves=True, internal=True) -> Iterator["Node"]:
        """Single-stack preorder: a node, then its children last-to-first."""
        stack = [self]
        while stack:
            n = stack.pop()
            if (leaves and not n.children) or (internal and n.children):
                yield n
            stack.extend(n.children)

    def traverse_postorder(self, leaves=True, internal=True) -> Iterator["Node"]:
        """Children left-to-right, then the node."""
        s1 = [self]
        s2 = []
        while s1:
            n = s1.pop()
            s2.append(n)
            s1.extend(n.children)
        while s2:
            n = s2.pop()
            if (leaves and not n.children) or (internal and n.children):
                yield n

    def traverse_leaves(self) -> Iterator["Node"]:
        return self.traverse_preorder(leaves=True, internal=False)

    def traverse_levelorder(self) -> Iterator["Node"]:
        q = deque([self])
        while q:
            n = q.popleft()
            yield n
            q.extend(n.children)

    def traverse_ancestors(self, include_self=True) -> Iterator["Node"]:
        n = self if include_self else self.parent
        while n is not None:
            yield n
            n = n.parent

    def __str__(self):
        return "" if self.label is None else str(self.label)

    # -- serialisation -----------------------------------------------------
    def newick(self) -> str:
        out: List[str] = []
        # iterative to survive caterpillar trees deeper than the recursion limit
        stack = [(self, 0)]
        while stack:
            node, state = stack.pop()
            if state == 0:
                if node.children:
                    out.append("(")
                    stack.append((node, 1))
                    for i, c in enumerate(reversed(node.children)):
                        stack.append((c, 0))
                        if i != len(node.children) - 1:
                            stack.append((None, 2))
                else:
                    out.append(_fmt_label(node.label))
                    if node is not self:
                        out.append(_fmt_len(node.edge_length))
            elif state == 1:
                out.append(")")
                out.append(_fmt_label(node.label))
                if node is not self:
                    out.append(_fmt_len(node.edge_length))
            else:
                out.append(",")
        return "".join(out)


def _fmt_label(label) -> str:
    return "" if label is None else str(label)


def _fmt_len(length) -> str:
    if length is None:
        return ""
    if isinstance(length, int):
        return ":%d" % length
    f = float(length)
    if f.is_integer() and abs(f) < 1e15:
        return ":%d" % int(f)
    return ":" + repr(f)


class Tree:
    def __init__(self, is_rooted=True):
        self.root = Node()
        self.is_rooted = is_rooted

    # -- traversals --------------------------------------------------------
    def traverse_preorder(self, leaves=True, internal=True):
        return self.root.traverse_preorder(leaves=leaves, internal=internal)

    def traverse_postorder(self, leaves=True, internal=True):
        return self.root.traverse_postorder(leaves=leaves, internal=internal)

    def traverse_leaves(self):
        return self.root.traverse_leaves()

    def traverse_levelorder(self):
        return self.root.traverse_levelorder()

    def traverse_internal(self):
        return self.root.traverse_preorder(leaves=False, internal=True)

    def labels(self, leaves=True, internal=True):
        for n in self.traverse_preorder(leaves=leaves, internal=internal):
            if n.label is not None:
                yield n.label

    def num_nodes(self, leaves=True, internal=True) -> int:
        return sum(1 for _ in self.traverse_preorder(leaves=leaves, internal=internal))

    def label_to_node(self, selection="leaves") -> Dict[str, Node]:
        if selection == "all":
            it = self.traverse_preorder()
        elif selection == "leaves":
            it = self.traverse_leaves()
        elif selection == "internal":
            it = self.traverse_internal()
        else:
            wanted = set(selection)
            it = (n for n in self.traverse_preorder() if n.label in wanted)
        return {n.label: n for n in it if n.label is not None}

    # -- serialisation -----------------------------------------------------
    def newick(self) -> str:
        suffix = _fmt_len(self.root.edge_length) + ";"
        body = self.root.newick()
        return ("[&R] " if self.is_rooted else "") + body + suffix

    def __str__(self):
        return self.newick()

    # -- metrics -----------------------------------------------------------
    def height(self, weighted=True) -> float:
        """Largest root-to-node distance; the root edge counts (as in the
        reference's time-scale computation, ``laml_libs/runner.py:336``)."""
        best = 0.0
        dist = {}
        for n in self.traverse_preorder():
            up = dist[n.parent] if n.parent is not None else 0.0
            step = (n.edge_length if n.edge_length is not None else 0.0) if weighted else 1.0
            dist[n] = up + step
            if dist[n] > best:
                best = dist[n]
        return best

    # -- structural edits --------------------------------------------------
    def suppress_unifurcations(self) -> None:
        """Remove out-degree-1 nodes; the removed node's edge length is added
        to its only child (``laml_libs/ML_solver.py:30``)."""
        q = deque([self.root])
        while q:
            node = q.popleft()
            while len(node.children) == 1:
                child = node.children[0]
                if node.edge_length is not None or child.edge_length is not None:
                    child.edge_length = (child.edge_length or 0) + (node.edge_length or 0)
                if child.label is None and node.label is not None:
                    child.label = node.label
                par = node.parent
                if par is None:
                    child.parent = None
                    self.root = child
                else:
                    idx = par.children.index(node)
                    par.children[idx] = child
                    child.parent = par
                node = child
            q.extend(node.children)

    def resolve_polytomies(self) -> None:
        """Binarise: while a node has more than two children, its last two
        children are re-attached under a fresh zero-length node.  Fresh nodes
        carry no user attributes, which is how the reference recognises them
        (``laml_libs/EM_solver.py:35-43``)."""
        q = deque([self.root])
        while q:
            node = q.popleft()
            while len(node.children) > 2:
                c1 = node.children.pop()
                c2 = node.children.pop()
                nn = Node(edge_length=0)
                node.add_child(nn)
                nn.add_child(c1)
                nn.add_child(c2)
            q.extend(node.children)

    def collapse_short_branches(self, threshold) -> None:
        """Contract internal branches of length <= threshold."""
        q = deque([self.root])
        while q:
            node = q.popleft()
            if node.is_root() or node.is_leaf():
                q.extend(node.children)
                continue
            if node.edge_length is None or node.edge_length <= threshold:
                par = node.parent
                par.remove_child(node)
                for c in node.children:
                    par.add_child(c)
                    q.append(c)
                node.children = []
            else:
                q.extend(node.children)

    def extract_tree_without(self, labels, suppress_unifurcations=True) -> "Tree":
        drop = set(labels)
        out = read_tree_newick(self.newick())
        out.is_rooted = True          # treeswift builds the result in a fresh Tree(), whose default is rooted
        for leaf in list(out.traverse_leaves()):
            if leaf.label in drop:
                n = leaf
                while n.parent is not None and not n.children:
                    p = n.parent
                    p.remove_child(n)
                    n = p
        if suppress_unifurcations:
            out.suppress_unifurcations()
        return out


def read_tree_newick(newick) -> Tree:
    """Parse a newick string, a path to a newick file, or anything whose
    ``str()`` is a newick string (the reference's tests hand an already parsed
    tree to the solver constructor, ``laml_unit_tests/unit_tests_EMSolver.py:508-523``).
    Root edge lengths ``(...):1;`` are kept on ``tree.root.edge_length``."""
    if not isinstance(newick, str):
        newick = str(newick)
    text = newick.strip()
    if text and not text.startswith(("(", "[")) and os.path.isfile(os.path.expanduser(text)):
        with open(os.path.expanduser(text)) as fh:
            text = fh.read().strip()
    tree = Tree()
    # treeswift keeps the rooting comment of the input: a string without ``[&R]`` is written back without it
    # (the reference's published trees, examples/out_example1/example1_trees.nwk, and the string slicing in
    # laml_unit_tests/unit_tests_TopoSearch.py:47 rely on that)
    tree.is_rooted = False
    if text.startswith("[&"):
        close = text.index("]")
        tree.is_rooted = text[2:close].strip().upper() == "R"
        text = text[close + 1:].strip()
    node = tree.root
    i = 0
    n = len(text)
    while i < n:
        ch = text[i]
        if ch == ";":
            break
        if ch == "(":
            child = Node()
            node.add_child(child)
            node = child
            i += 1
        elif ch == ")":
            node = node.parent
            i += 1
        elif ch == ",":
            sib = Node()
            node.parent.add_child(sib)
            node = sib
            i += 1
        elif ch == ":":
            j = i + 1
            while j < n and text[j] not in ",();[":
                j += 1
            tok = text[i + 1:j].strip()
            if tok:
                val = float(tok)
                node.edge_length = int(val) if val.is_integer() and "." not in tok and "e" not in tok.lower() else val
            i = j
        elif ch == "[":
            i = text.index("]", i) + 1  # comment
        elif ch in " \t\r\n":
            i += 1
        elif ch in "'\"":
            j = text.index(ch, i + 1)
            node.label = text[i + 1:j]
            i = j + 1
        else:
            j = i
            while j < n and text[j] not in ":,();[":
                j += 1
            node.label = text[i:j].strip()
            i = j
    if node is not tree.root:
        raise ValueError("unbalanced newick string")
    return tree
