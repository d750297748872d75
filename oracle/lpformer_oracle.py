"""CPU oracle for LPFormer's per-link pairwise-encoding path.

TEST INFRASTRUCTURE ONLY.  Only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` may import this module;
nothing under `lpformer_b200/` does.  It is the checker, never the product.

Parity status: PINNED.  Every function here is checked (tests/test_oracle.py)
against golden vectors in `tests/golden/*.npz` that were produced by running the
UNMODIFIED reference model files (`/root/reference/src/models/link_transformer.py`,
`modules/layers.py`, `modules/node_encoder.py`, `models/other_models.py`) on CPU
through the third-party shims in `oracle/shims/` (generator:
`tests/golden/make_golden.py`).  The reference itself ships no tests/fixtures
(SURVEY.md §4), so those generated outputs are the pin.

This file restates the reference *semantics* with plain numpy sets and dense
algebra.  `oracle/ref_port.py` restates the
reference's own *algorithm* (torch sparse-COO algebra) and is what the CPU
baseline times.  All `file:line` citations are relative to /root/reference/src.
"""
from __future__ import annotations

import numpy as np

F32 = np.float32


# --------------------------------------------------------------------------- #
# CSR helpers
# --------------------------------------------------------------------------- #
class CSR:
    """Sorted-column CSR.  `val` is None for the 0/1 adjacency mask."""

    def __init__(self, indptr, indices, val=None, n=None):
        self.indptr = np.asarray(indptr, dtype=np.int64)
        self.indices = np.asarray(indices, dtype=np.int64)
        self.val = None if val is None else np.asarray(val, dtype=F32)
        self.n = int(n if n is not None else len(self.indptr) - 1)

    @staticmethod
    def from_coo(row, col, val, n):
        row = np.asarray(row, dtype=np.int64)
        col = np.asarray(col, dtype=np.int64)
        order = np.lexsort((col, row))
        row, col = row[order], col[order]
        indptr = np.zeros(n + 1, dtype=np.int64)
        np.add.at(indptr, row + 1, 1)
        indptr = np.cumsum(indptr)
        return CSR(indptr, col, None if val is None else np.asarray(val)[order], n)

    def row(self, i):
        s, e = self.indptr[i], self.indptr[i + 1]
        return (self.indices[s:e], None if self.val is None else self.val[s:e])


def quantise(p):
    """q(p) = fl32(p + 1) - 1, the value the reference actually thresholds and
    feeds to the RPE MLP.  models/link_transformer.py:290-291 computes
    `ppr*t + t` on the pair pattern (t in {1,2}), :316-317 undo it with
    `(v - t)/t`; :464-476 do `p + sign(other) - 1`.  All reduce to q."""
    p = np.asarray(p, dtype=F32)
    return (p + F32(1.0)) - F32(1.0)


def model_mode(th_1hop, th_non1hop):
    """models/link_transformer.py:39-44."""
    if th_non1hop == 1 and th_1hop == 1:
        return "cn"
    if th_non1hop == 1 and th_1hop < 1:
        return "1-hop"
    return "all"


# --------------------------------------------------------------------------- #
# Node selection (compute_node_mask / get_ppr_vals / get_non_1hop_ppr)
# --------------------------------------------------------------------------- #
def _lookup(cols, vals, u):
    """(present mask, q(P(x,u)) or 0) for sorted `cols`."""
    if len(cols) == 0:
        return np.zeros(len(u), bool), np.zeros(len(u), F32)
    pos = np.searchsorted(cols, u)
    pos_c = np.minimum(pos, len(cols) - 1)
    hit = cols[pos_c] == u
    q = np.where(hit, quantise(vals[pos_c]), F32(0.0)).astype(F32)
    return hit, q


def select_link(adj: CSR, ppr: CSR, a, b, th_cn, th_1hop, th_non1hop, mode):
    """Selected node sets of ONE link (a,b): returns dict type -> (nodes, qa, qb),
    nodes ascending.  models/link_transformer.py:214-276 (CN / 1-hop split and
    thresholds :241-250, :263-268), :279-319 (values), :434-481 (>1-hop)."""
    th_cn, th_1hop, th_non1hop = F32(th_cn), F32(th_1hop), F32(th_non1hop)  # compare in fp32
    Aa, _ = adj.row(a)
    Ab, _ = adj.row(b)
    Pa_c, Pa_v = ppr.row(a)
    Pb_c, Pb_v = ppr.row(b)
    out = {}

    cn = np.intersect1d(Aa, Ab)                       # pair_adj == 2 (:237)  /  src*tgt (:234)
    _, qa = _lookup(Pa_c, Pa_v, cn)
    _, qb = _lookup(Pb_c, Pb_v, cn)
    keep = (qa >= th_cn) & (qb >= th_cn)              # :241
    out["cn"] = (cn[keep], qa[keep], qb[keep])
    if mode == "cn":
        return out

    oh = np.setxor1d(Aa, Ab)                          # pair_adj == 1
    _, qa = _lookup(Pa_c, Pa_v, oh)
    _, qb = _lookup(Pb_c, Pb_v, oh)
    keep = (qa >= th_1hop) & (qb >= th_1hop)          # :242
    out["1hop"] = (oh[keep], qa[keep], qb[keep])
    if mode == "1-hop":
        return out

    # >1-hop (:443-481): pattern = P(a,.) U P(b,.); entries on A(a) U A(b) are zeroed
    # (:452-460); value = p' + sign(other') - 1 (:464-476); keep both >= th (:478).
    cand = np.union1d(Pa_c, Pb_c)
    in_adj = np.isin(cand, Aa) | np.isin(cand, Ab)
    ha, _ = _lookup(Pa_c, Pa_v, cand)
    hb, _ = _lookup(Pb_c, Pb_v, cand)
    pa = np.zeros(len(cand), F32)
    pb = np.zeros(len(cand), F32)
    if len(Pa_c):
        pa[ha] = Pa_v[np.searchsorted(Pa_c, cand[ha])]
    if len(Pb_c):
        pb[hb] = Pb_v[np.searchsorted(Pb_c, cand[hb])]
    pa[in_adj] = 0
    pb[in_adj] = 0
    sv = (pa + np.sign(pb).astype(F32)) - F32(1.0)
    tv = (pb + np.sign(pa).astype(F32)) - F32(1.0)
    keep = (sv >= th_non1hop) & (tv >= th_non1hop)
    out["non1hop"] = (cand[keep], sv[keep].astype(F32), tv[keep].astype(F32))
    return out


def select_sets(adj: CSR, ppr: CSR, links, th_cn, th_1hop, th_non1hop, adj_far: CSR = None):
    """All links of a batch.  Returns (mode, {type: (link_idx int64[S_t], node int64[S_t],
    src_ppr f32[S_t], tgt_ppr f32[S_t])}) sorted by (link, node) within a type, exactly
    the tuples `compute_node_mask` returns (ix[0]=position in batch, ix[1]=node id).

    adj_far: the reference takes the >1-hop set from the STORED adjacency even when the caller hands
    compute_node_mask another one for the CN / 1-hop sets (models/link_transformer.py:226-254 uses `adj`,
    get_non_1hop_ppr :443-447 calls get_adj itself); pass the stored table here in that case."""
    links = np.asarray(links, dtype=np.int64)
    mode = model_mode(th_1hop, th_non1hop)
    types = {"cn": ["cn"], "1-hop": ["cn", "1hop"], "all": ["cn", "1hop", "non1hop"]}[mode]
    acc = {t: ([], [], [], []) for t in types}
    for i in range(links.shape[1]):
        res = select_link(adj, ppr, links[0, i], links[1, i], th_cn, th_1hop, th_non1hop, mode)
        if adj_far is not None and mode == "all":
            res = dict(res, non1hop=select_link(adj_far, ppr, links[0, i], links[1, i], th_cn, th_1hop, th_non1hop, mode)["non1hop"])
        for t in types:
            n, qa, qb = res[t]
            acc[t][0].append(np.full(len(n), i, np.int64))
            acc[t][1].append(n)
            acc[t][2].append(qa)
            acc[t][3].append(qb)
    out = {}
    for t in types:
        out[t] = (np.concatenate(acc[t][0]) if acc[t][0] else np.zeros(0, np.int64),
                  np.concatenate(acc[t][1]).astype(np.int64) if acc[t][1] else np.zeros(0, np.int64),
                  np.concatenate(acc[t][2]).astype(F32) if acc[t][2] else np.zeros(0, F32),
                  np.concatenate(acc[t][3]).astype(F32) if acc[t][3] else np.zeros(0, F32))
    return mode, out


def structure_counts(sets, mode, num_links):
    """get_structure_cnts / get_count (models/link_transformer.py:340-386): post-filter
    set sizes; num_neighbors = |CN| + |1hop| (:348-350 recount with thresh=0 over the
    already-filtered set).  Column order as concatenated at :155, :173, :175."""
    def cnt(t):
        return np.bincount(sets[t][0], minlength=num_links).astype(F32)
    if mode == "cn":
        return np.stack([cnt("cn")], 1)
    if mode == "1-hop":
        return np.stack([cnt("cn"), cnt("1hop"), cnt("cn") + cnt("1hop")], 1)
    return np.stack([cnt("cn"), cnt("1hop"), cnt("non1hop"), cnt("cn") + cnt("1hop")], 1)


# --------------------------------------------------------------------------- #
# Dense pieces
# --------------------------------------------------------------------------- #
def layer_norm(x, w, b, eps=1e-5):
    mu = x.mean(-1, keepdims=True)
    var = ((x - mu) ** 2).mean(-1, keepdims=True)
    return (x - mu) / np.sqrt(var + eps) * w + b


def mlp2(x, P, prefix):
    """models/other_models.py:125-138 with num_layers=2: Linear -> LayerNorm -> ReLU -> Linear."""
    h = x @ P[f"{prefix}.linears.0.weight"].T + P[f"{prefix}.linears.0.bias"]
    h = layer_norm(h, P[f"{prefix}.norm.weight"], P[f"{prefix}.norm.bias"])
    h = np.maximum(h, 0)
    return h @ P[f"{prefix}.linears.1.weight"].T + P[f"{prefix}.linears.1.bias"]


def gcn_norm(adj_w: CSR):
    """PyG 2.2.0 gcn_norm on a SparseTensor: diagonal SET to 1,
    deg = row sum, A_hat = D^-1/2 A D^-1/2.  Returns COO (row, col, val) float64."""
    n = adj_w.n
    row = np.repeat(np.arange(n), np.diff(adj_w.indptr))
    col = adj_w.indices
    val = np.ones(len(col)) if adj_w.val is None else adj_w.val.astype(np.float64)
    keep = row != col
    row = np.concatenate([row[keep], np.arange(n)])
    col = np.concatenate([col[keep], np.arange(n)])
    val = np.concatenate([val[keep], np.ones(n)])
    deg = np.bincount(row, weights=val, minlength=n)
    with np.errstate(divide="ignore"):
        dis = np.where(deg > 0, deg ** -0.5, 0.0)
    return row, col, val * dis[row] * dis[col]


def propagate(x, adj_w: CSR, P, cfg):
    """LinkTransformer.propagate (models/link_transformer.py:110-129) ->
    NodeEncoder.forward (modules/node_encoder.py:35-44, eval: no dropout) ->
    GCN.forward (models/other_models.py:61-76) -> gnn_norm."""
    row, col, val = gcn_norm(adj_w)
    n = adj_w.n
    x = x.astype(np.float64)
    for i in range(cfg["gnn_layers"]):
        pre = f"node_encoder.gnn_encoder.convs.{i}"
        xw = x @ P[f"{pre}.lin.weight"].T
        xi = np.zeros((n, xw.shape[1]))
        np.add.at(xi, row, val[:, None] * xw[col])
        xi = xi + P[f"{pre}.bias"]
        if cfg["layer_norm"]:
            ln = f"node_encoder.gnn_encoder.lns.{i}"
            xi = layer_norm(xi, P[f"{ln}.weight"], P[f"{ln}.bias"])
        if cfg["relu"]:
            xi = np.maximum(xi, 0)
        x = x + xi if (cfg["residual"] and x.shape[-1] == xi.shape[-1]) else xi
    return layer_norm(x, P["gnn_norm.weight"], P["gnn_norm.bias"])


_ENC = {"cn": "ppr_encoder_cn", "1hop": "ppr_encoder_onehop", "non1hop": "ppr_encoder_non1hop"}


def pos_encodings(sets, mode, P):
    """get_pos_encodings (models/link_transformer.py:182-211): per type
    MLP([pa,pb]) + MLP([pb,pa]); concatenated CN | 1-hop | >1-hop."""
    out = []
    for t in sets:
        _, _, pa, pb = sets[t]
        ab = np.stack([pa, pb], 1).astype(np.float64)
        ba = np.stack([pb, pa], 1).astype(np.float64)
        out.append(mlp2(ab, P, _ENC[t]) + mlp2(ba, P, _ENC[t]))
    return np.concatenate(out, 0)


def attention_layer(link_idx, node_idx, e1, e2, node_x, pe, P, l, heads, num_links):
    """LinkTransformerLayer.forward (modules/layers.py:39-82) + LinkAttention
    (:161-224): v = lin_r([X[u] | pe]); e = lin_l(e1)+lin_l(e2); s = sum_c att*leaky_relu(v*e);
    alpha = segment softmax (max-subtracted, denominator + 1e-16); out = sum alpha*v + bias;
    LayerNorm."""
    pre = f"att_layers.{l}.att"
    Wl, bl = P[f"{pre}.lin_l.weight"], P[f"{pre}.lin_l.bias"]
    Wr, br = P[f"{pre}.lin_r.weight"], P[f"{pre}.lin_r.bias"]
    att = P[f"{pre}.att"].reshape(heads, -1)
    C = att.shape[1]
    v = (np.concatenate([node_x[node_idx], pe], 1) @ Wr.T + br).reshape(-1, heads, C)
    e = ((e1 @ Wl.T + bl) + (e2 @ Wl.T + bl)).reshape(-1, heads, C)
    x = v * e[link_idx]
    x = np.where(x > 0, x, 0.2 * x)
    s = (x * att[None]).sum(-1)                                    # [S, H]
    smax = np.full((num_links, heads), -np.inf)
    np.maximum.at(smax, link_idx, s)
    ex = np.exp(s - smax[link_idx])
    den = np.zeros((num_links, heads))
    np.add.at(den, link_idx, ex)
    alpha = ex / (den[link_idx] + 1e-16)
    out = np.zeros((num_links, heads, C))
    np.add.at(out, link_idx, v * alpha[:, :, None])
    out = out.reshape(num_links, heads * C) + P[f"{pre}.bias"]
    out = layer_norm(out, P[f"att_layers.{l}.post_att_norm.weight"], P[f"att_layers.{l}.post_att_norm.bias"])
    return out, alpha


def calc_pairwise(links, X, adj: CSR, ppr: CSR, P, cfg):
    """LinkTransformer.calc_pairwise (models/link_transformer.py:132-178), eval mode.
    Returns (pairwise_feats [BS,d], sets, counts, alpha of last layer)."""
    links = np.asarray(links, dtype=np.int64)
    bs = links.shape[1]
    mode, sets = select_sets(adj, ppr, links, cfg["thresh_cn"], cfg["thresh_1hop"], cfg["thresh_non1hop"])
    counts = structure_counts(sets, mode, bs)
    link_idx = np.concatenate([sets[t][0] for t in sets])
    node_idx = np.concatenate([sets[t][1] for t in sets])
    pe = pos_encodings(sets, mode, P)
    X = X.astype(np.float64)
    feats = np.concatenate([X[links[0]], X[links[1]]], 1)
    alpha = None
    for l in range(cfg["trans_layers"]):
        half = feats.shape[1] // 2
        feats, alpha = attention_layer(link_idx, node_idx, feats[:, :half], feats[:, half:], X, pe, P, l,
                                       cfg["num_heads"], bs)
    feats = np.concatenate([feats, counts.astype(np.float64)], 1)
    return mlp2(feats, P, "pairwise_lin"), (mode, sets), counts, alpha


def link_features(links, X, adj, ppr, P, cfg):
    """Body of the eval loop, train/testing.py:29-31 / :113-115: [elementwise | pairwise]."""
    links = np.asarray(links, dtype=np.int64)
    X64 = X.astype(np.float64)
    el = mlp2(X64[links[0]] * X64[links[1]], P, "elementwise_lin")
    pw, sel, counts, alpha = calc_pairwise(links, X, adj, ppr, P, cfg)
    return np.concatenate([el, pw], 1), sel, counts, alpha


def mlp_score(feats, S):
    """models/other_models.py:173-179: Linear/ReLU ... Linear -> sigmoid.  Returns (logit, prob)."""
    n = len([k for k in S if k.endswith(".weight")])
    x = feats
    for i in range(n - 1):
        x = np.maximum(x @ S[f"lins.{i}.weight"].T + S[f"lins.{i}.bias"], 0)
    logit = (x @ S[f"lins.{n-1}.weight"].T + S[f"lins.{n-1}.bias"]).squeeze(-1)
    return logit, 1.0 / (1.0 + np.exp(-logit))


# --------------------------------------------------------------------------- #
# PPR push (util/calc_ppr_scores.py:137-192), pure Python: small graphs only.
# larger ones take the host port of the same kernel (lpformer_b200/csrc/ppr_push.cpp, pinned by tests/test_host.py).
# --------------------------------------------------------------------------- #
def ppr_push(indptr, indices, alpha, eps):
    """Andersen push with a LIFO queue, float64 arithmetic, exactly the op order of
    calc_ppr (:155-190); then create_sparse_ppr_matrix (:221-241) casts to fp32 and
    sorts by column.  Returns a CSR with fp32 values."""
    n = len(indptr) - 1
    deg = np.diff(indptr)
    alpha_eps = alpha * eps
    rows, cols, vals = [], [], []
    for s in range(n):
        p = {s: 0.0}
        r = {s: alpha}
        q = [s]
        while q:
            u = q.pop()
            res = r[u] if u in r else 0
            p[u] = p.get(u, 0.0) + res
            r[u] = 0
            for v in indices[indptr[u]:indptr[u + 1]]:
                v = int(v)
                _val = (1 - alpha) * res / deg[u]
                r[v] = r[v] + _val if v in r else _val
                if r[v] >= alpha_eps * deg[v] and v not in q:
                    q.append(v)
        ks = sorted(p)
        rows += [s] * len(ks)
        cols += ks
        vals += [p[k] for k in ks]
    return CSR.from_coo(rows, cols, np.asarray(vals, dtype=np.float64).astype(F32), n)
