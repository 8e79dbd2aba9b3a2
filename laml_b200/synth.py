"""Synthetic PMM workloads (trees, priors, character matrices) as arrays.

The reference's generators cannot produce the benchmark inputs here:
``simulate_tree.py`` needs Cassiopeia, which ``laml_libs/sim_lib.py:8,17`` does
not even import, and ``simulate_seqs`` (``sim_lib.py:64-105``) is a per-entry
``random.choices`` loop (~10^7 calls at 10k x 500).  This module draws from the
*same transition law* (``sim_lib.py:77-93``) with vectorised torch ops so the
100k x 2k workload is generated in seconds on the GPU (or on the CPU for tests):

  parent 0      -> 0 w.p. p^(1+nu); alpha w.p. q_alpha p^nu (1-p); silenced w.p. 1-p^nu
  parent alpha  -> alpha w.p. p^nu;  silenced w.p. 1-p^nu
  silenced      -> silenced
  leaves: silenced -> '?', anything else -> '?' w.p. dropout (sim_lib.py:91-93,103)

Trees: repeatedly split a uniformly chosen current leaf (seeded); branch
lengths i.i.d. lognormal as in ``simTree_lnorm`` (``sim_lib.py:16-19``), then
rescaled so the mean root-to-tip depth is ``depth`` (``simulate_tree.py:7-23``).
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np
import torch

from .flatten import FlatForest, forest_from_parent


def random_tree(n_leaves: int, seed: int = 1, caterpillar_bias: float = 0.0) -> FlatForest:
    """Random rooted binary tree; leaves get ids 0..N-1, the root is the last
    internal node created first (ids are then leaves-first).  With
    ``caterpillar_bias`` in (0,1) the most recently created leaf is split with
    that probability, which produces deep, narrow lineages."""
    rng = np.random.default_rng(seed)
    N = int(n_leaves)
    n = 2 * N - 1
    # build with temporary ids: node 0 is the root
    left = np.full(n, -1, np.int64)
    right = np.full(n, -1, np.int64)
    parent = np.full(n, -1, np.int64)
    leaves = [0]
    nxt = 1
    picks = rng.random(N)
    idxs = rng.random(N)
    for i in range(N - 1):
        if caterpillar_bias > 0 and picks[i] < caterpillar_bias:
            j = len(leaves) - 1
        else:
            j = int(idxs[i] * len(leaves))
        v = leaves[j]
        leaves[j] = leaves[-1]
        leaves.pop()
        left[v], right[v] = nxt, nxt + 1
        parent[nxt] = parent[nxt + 1] = v
        leaves.extend((nxt, nxt + 1))
        nxt += 2
    is_leaf = left < 0
    new_id = np.empty(n, np.int64)
    new_id[is_leaf] = np.arange(N)
    new_id[~is_leaf] = N + np.arange(N - 1)
    p2 = np.full(n, -1, np.int32); l2 = np.full(n, -1, np.int32); r2 = np.full(n, -1, np.int32)
    p2[new_id] = np.where(parent >= 0, new_id[np.maximum(parent, 0)], -1)
    l2[new_id] = np.where(left >= 0, new_id[np.maximum(left, 0)], -1)
    r2[new_id] = np.where(right >= 0, new_id[np.maximum(right, 0)], -1)
    return forest_from_parent(p2, l2, r2)


def lognormal_branch_lengths(forest: FlatForest, seed: int = 2, std: float = 0.1, depth: float = 1.0) -> np.ndarray:
    rng = np.random.default_rng(seed)
    n = forest.n_nodes
    log_sd = math.sqrt(math.log((1 + math.sqrt(1 + 4 * std * std)) / 2))     # sim_lib.py:16
    d = rng.lognormal(0.0, log_sd, n)
    # mean root-to-tip depth (root edge included) -> `depth`
    dist = np.zeros(n)
    for lvl in range(int(forest.depth.max()) + 1):
        idx = np.nonzero(forest.depth == lvl)[0]
        par = forest.parent[idx]
        dist[idx] = d[idx] + np.where(par >= 0, dist[np.maximum(par, 0)], 0.0)
    mean_tip = dist[: forest.n_leaves].mean()
    return d * (depth / mean_tip)


def random_priors(n_sites: int, a_lo: int, a_hi: int, seed: int = 5, uniform: bool = False):
    """CSR prior: A_s ~ U{a_lo..a_hi}, weights Dirichlet(1) (or uniform as ``sim_Q``, sim_lib.py:121-131)."""
    rng = np.random.default_rng(seed)
    A = rng.integers(a_lo, a_hi + 1, n_sites)
    qptr = np.zeros(n_sites + 1, np.int32)
    qptr[1:] = np.cumsum(A)
    qval = np.empty(qptr[-1], np.float64)
    for s in range(n_sites):
        w = np.full(A[s], 1.0 / A[s]) if uniform else rng.dirichlet(np.ones(A[s]))
        qval[qptr[s]:qptr[s + 1]] = w
    return qptr, qval


@torch.no_grad()
def simulate_codes(forest: FlatForest, d: np.ndarray, qptr: np.ndarray, qval: np.ndarray, nu: float, phi: float,
                   seed: int = 7, device: str = "cpu", site_block: int = 256, sites=None) -> np.ndarray:
    """Leaf codes (n_leaves, K') int32 in the engine's coding (0, 1..A_s, -1 = '?')
    for the sites ``[lo, hi)`` (default all).  Every block of ``site_block`` sites
    has its own seeded stream, so a rank that simulates only its shard gets the
    same columns as a rank that simulates everything."""
    K = len(qptr) - 1
    lo, hi = (0, K) if sites is None else sites
    n, N = forest.n_nodes, forest.n_leaves
    dev = torch.device(device)
    depth = torch.as_tensor(forest.depth, device=dev)
    parent = torch.as_tensor(forest.parent.astype(np.int64), device=dev)
    dt = torch.as_tensor(d, dtype=torch.float64, device=dev)
    p = torch.exp(-dt)
    stay0 = p ** (1.0 + nu)
    keep = p ** nu                       # not silenced on this branch
    levels = [torch.nonzero(depth == l).flatten() for l in range(int(forest.depth.max()) + 1)]
    levels = [(idx, parent[idx]) for idx in levels if idx.numel()]
    out = np.empty((N, hi - lo), np.int32)
    Amax = int(np.max(np.diff(qptr))) if K else 0
    SIL = -3
    sil = torch.full((), SIL, dtype=torch.int16, device=dev)
    zero = torch.zeros((), dtype=torch.int16, device=dev)
    for blk in range(lo // site_block, (hi + site_block - 1) // site_block):
        b0, b1 = blk * site_block, min(K, (blk + 1) * site_block)
        kb = b1 - b0
        gen = torch.Generator(device=dev)
        gen.manual_seed(seed * 1000003 + blk)
        cdf = np.ones((kb, Amax), np.float64)
        for j, s in enumerate(range(b0, b1)):
            c = np.cumsum(qval[qptr[s]:qptr[s + 1]])
            cdf[j, : c.shape[0]] = c / c[-1]
        cdf = torch.as_tensor(cdf, device=dev)
        A_blk = torch.as_tensor(np.diff(qptr)[b0:b1].astype(np.int64), device=dev)
        state = torch.zeros((n, kb), dtype=torch.int16, device=dev)
        for idx, par in levels:
            m = idx.numel()
            pst = torch.where((par >= 0)[:, None], state[par.clamp(min=0)], zero)
            u = torch.rand((m, kb), dtype=torch.float64, device=dev, generator=gen)
            k_ = keep[idx][:, None]
            s0 = stay0[idx][:, None]
            silenced = (pst == SIL) | (u >= k_)                     # u in [p^nu, 1): silenced on this branch
            mutate = (pst == 0) & (u >= s0) & (u < k_)              # u in [p^(1+nu), p^nu): 0 -> alpha
            r = torch.rand((m, kb), dtype=torch.float64, device=dev, generator=gen)
            drawn = torch.searchsorted(cdf, r.t().contiguous()).t() + 1     # (m,kb) in 1..A
            drawn = torch.minimum(drawn, A_blk[None, :]).to(torch.int16)
            new = torch.where(mutate, drawn, pst)
            state[idx] = torch.where(silenced, sil, new)
        leaf = state[:N].to(torch.int32)
        drop = torch.rand((N, kb), dtype=torch.float64, device=dev, generator=gen) < phi
        leaf = torch.where((leaf == SIL) | drop, torch.full((), -1, dtype=torch.int32, device=dev), leaf)
        c0, c1 = max(b0, lo), min(b1, hi)
        out[:, c0 - lo:c1 - lo] = leaf[:, c0 - b0:c1 - b0].cpu().numpy()
    return out


@dataclass
class Workload:
    forest: FlatForest
    d: np.ndarray
    qptr: np.ndarray
    qval: np.ndarray
    codes: np.ndarray
    nu: float
    phi: float
    name: str = ""

    @property
    def W(self) -> int:
        return int(self.qptr[-1]) + 2 * (len(self.qptr) - 1)


def make_workload(n_leaves: int, n_sites: int, a_lo: int, a_hi: int, nu: float, phi: float, seed: int = 1,
                  uniform_prior: bool = False, device: str = "cpu", name: str = "", caterpillar_bias: float = 0.0,
                  sites=None) -> Workload:
    """``sites=(lo,hi)`` restricts the simulated character matrix (and the returned
    prior) to a contiguous site range of the same workload (site sharding)."""
    forest = random_tree(n_leaves, seed=seed, caterpillar_bias=caterpillar_bias)
    d = lognormal_branch_lengths(forest, seed=seed + 1)
    qptr, qval = random_priors(n_sites, a_lo, a_hi, seed=seed + 4, uniform=uniform_prior)
    codes = simulate_codes(forest, d, qptr, qval, nu, phi, seed=seed + 6, device=device, sites=sites)
    if sites is not None:
        lo, hi = sites
        qval = qval[qptr[lo]:qptr[hi]].copy()
        qptr = (qptr[lo:hi + 1] - qptr[lo]).astype(np.int32)
    return Workload(forest, d, qptr, qval, codes, nu, phi, name)


#: BASELINE.json configs 2-5
CONFIGS = {
    "config2_1k_x_30": dict(n_leaves=1000, n_sites=30, a_lo=30, a_hi=30, nu=0.1, phi=0.1, uniform_prior=True),
    "config3_10k_x_500": dict(n_leaves=10000, n_sites=500, a_lo=2, a_hi=100, nu=0.3, phi=0.3),
    "config4_100k_x_2k": dict(n_leaves=100000, n_sites=2000, a_lo=2, a_hi=100, nu=0.3, phi=0.3),
    "config5_5k_x_300": dict(n_leaves=5000, n_sites=300, a_lo=2, a_hi=30, nu=0.1, phi=0.1),
}


def reference_inputs(w: Workload):
    """A synthetic workload in the form the reference's solver constructor takes (``ML_solver.py:19-26``):
    ``(newick strings, {cell: [states]}, [ {state: prior} per site ])`` -- leaf ``i`` is cell ``"c<i>"``,
    state ids are the 1-based ranks, ``'?'`` for missing entries, -1 for known-silenced ones."""
    f = w.forest
    text = [None] * f.n_nodes
    for v in f.order:                                   # children first
        v = int(v)
        length = ":%r" % float(w.d[v])
        if f.left[v] < 0:
            text[v] = "c%d%s" % (v, length)
        else:
            l, r = int(f.left[v]), int(f.right[v])
            text[v] = "(" + text[l] + "," + text[r] + ")" + length
            text[l] = text[r] = None
    newicks = [text[int(r)] + ";" for r in f.roots]
    K = w.codes.shape[1]
    msa = {}
    for i in range(f.n_leaves):
        msa["c%d" % i] = ['?' if x == -1 else (-1 if x == -2 else x) for x in w.codes[i].tolist()]
    Q = [{a + 1: float(w.qval[w.qptr[s] + a]) for a in range(int(w.qptr[s + 1] - w.qptr[s]))} for s in range(K)]
    return newicks, msa, Q
