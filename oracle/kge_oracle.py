"""CPU oracle for the KGEModel hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A numpy restatement of the reference algorithm (``/root/reference/codes/model.py`` and the
filter semantics of ``codes/dataloader.py``).  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import this module; nothing under
``knowledgegraphembedding_b200/`` does.

Parity pin: the reference ships no tests or golden vectors ("parity unpinned"
by the reference itself), so this oracle is pinned against *outputs of the reference run in the
build container*: ``tests/golden/make_golden.py`` imports the unmodified reference, runs it on
seeded inputs and commits the results under ``tests/golden/*.npz``; ``tests/test_oracle_golden.py``
checks every function below against those vectors.

All arithmetic is done in ``dtype`` (float32 restates the reference bit-for-bit per element up to
libm sin/cos and reduction order; float64 is used for error attribution).
Every function cites the reference lines it follows.
"""
from __future__ import annotations

import math

import numpy as np

MODELS = ("TransE", "DistMult", "ComplEx", "RotatE", "pRotatE")
MODES = ("single", "head-batch", "tail-batch")

PI_ROTATE = 3.14159265358979323846    # model.py:202
PI_PROTATE = 3.14159262358979323846   # model.py:232 (sic: the reference's constant differs at 1e-8)
EPSILON = 2.0                         # model.py:30


# --------------------------------------------------------------------------------------------------
# state (model.py:23-70)
# --------------------------------------------------------------------------------------------------
def embedding_range(gamma: float, hidden_dim: int) -> float:
    """model.py:32-40: both scalars live in fp32 parameters and are read back with .item()."""
    g = float(np.float32(gamma))
    return float(np.float32((g + EPSILON) / hidden_dim))


def dims(model_name, hidden_dim, double_entity_embedding=False, double_relation_embedding=False):
    """model.py:42-43."""
    return (hidden_dim * 2 if double_entity_embedding else hidden_dim,
            hidden_dim * 2 if double_relation_embedding else hidden_dim)


def check_config(model_name, double_entity_embedding, double_relation_embedding):
    """model.py:63-70 (same messages)."""
    if model_name not in MODELS:
        raise ValueError('model %s not supported' % model_name)
    if model_name == 'RotatE' and (not double_entity_embedding or double_relation_embedding):
        raise ValueError('RotatE should use --double_entity_embedding')
    if model_name == 'ComplEx' and (not double_entity_embedding or not double_relation_embedding):
        raise ValueError('ComplEx should use --double_entity_embedding and --double_relation_embedding')


def init_tables(model_name, nentity, nrelation, hidden_dim, gamma, de=False, dr=False, seed=0):
    """Synthetic U(-rho, rho) tables (same law as model.py:45-57; numpy RNG so that the stream is
    portable between the build container and the GPU box)."""
    check_config(model_name, de, dr)
    rho = embedding_range(gamma, hidden_dim)
    D_e, D_r = dims(model_name, hidden_dim, de, dr)
    rng = np.random.RandomState(seed)
    E = rng.uniform(-rho, rho, size=(nentity, D_e)).astype(np.float32)
    R = rng.uniform(-rho, rho, size=(nrelation, D_r)).astype(np.float32)
    state = {"entity_embedding": E, "relation_embedding": R}
    if model_name == 'pRotatE':
        state["modulus"] = np.array([[0.5 * rho]], dtype=np.float32)   # model.py:59-60
    return state


# --------------------------------------------------------------------------------------------------
# score functions on gathered rows (model.py:166-249).  head/relation/tail: [B, 1|N, D]
# --------------------------------------------------------------------------------------------------
def _split(x):
    d = x.shape[-1] // 2
    return x[..., :d], x[..., d:]          # torch.chunk(x, 2, dim=2): first half real, second imag


def score_rows(model_name, head, relation, tail, mode, gamma, rho, modulus=None, dtype=np.float32):
    f = dtype
    head, relation, tail = head.astype(f), relation.astype(f), tail.astype(f)
    g = f(np.float32(gamma))
    if model_name == 'TransE':                                   # model.py:166-173
        x = head + (relation - tail) if mode == 'head-batch' else (head + relation) - tail
        return g - np.abs(x).sum(axis=2, dtype=f)
    if model_name == 'DistMult':                                 # model.py:175-182
        x = head * (relation * tail) if mode == 'head-batch' else (head * relation) * tail
        return x.sum(axis=2, dtype=f)
    if model_name == 'ComplEx':                                  # model.py:184-199
        hr, hi = _split(head); rr, ri = _split(relation); tr, ti = _split(tail)
        if mode == 'head-batch':
            re = rr * tr + ri * ti
            im = rr * ti - ri * tr
            x = hr * re + hi * im
        else:
            re = hr * rr - hi * ri
            im = hr * ri + hi * rr
            x = re * tr + im * ti
        return x.sum(axis=2, dtype=f)
    if model_name == 'RotatE':                                   # model.py:201-229
        hr, hi = _split(head); tr, ti = _split(tail)
        phase = relation / f(rho / PI_ROTATE)                    # model.py:209
        c, s = np.cos(phase), np.sin(phase)
        if mode == 'head-batch':
            re = c * tr + s * ti
            im = c * ti - s * tr
            re = re - hr
            im = im - hi
        else:
            re = hr * c - hi * s
            im = hr * s + hi * c
            re = re - tr
            im = im - ti
        m = np.sqrt(re * re + im * im)                           # stack + norm(dim=0), model.py:225-226
        return g - m.sum(axis=2, dtype=f)
    if model_name == 'pRotatE':                                  # model.py:231-249
        k = f(rho / PI_PROTATE)
        ph, pr, pt = head / k, relation / k, tail / k
        x = ph + (pr - pt) if mode == 'head-batch' else (ph + pr) - pt
        x = np.abs(np.sin(x))
        return g - x.sum(axis=2, dtype=f) * f(np.asarray(modulus).reshape(())[()])
    raise ValueError('model %s not supported' % model_name)      # model.py:162


def gather(state, sample, mode):
    """model.py:83-146: index_select of head / relation / tail rows."""
    E, R = state["entity_embedding"], state["relation_embedding"]
    if mode == 'single':
        s = np.asarray(sample)
        return E[s[:, 0]][:, None, :], R[s[:, 1]][:, None, :], E[s[:, 2]][:, None, :]
    pos, neg = sample
    pos, neg = np.asarray(pos), np.asarray(neg)
    if mode == 'head-batch':
        return E[neg.reshape(-1)].reshape(neg.shape[0], neg.shape[1], -1), \
            R[pos[:, 1]][:, None, :], E[pos[:, 2]][:, None, :]
    if mode == 'tail-batch':
        return E[pos[:, 0]][:, None, :], R[pos[:, 1]][:, None, :], \
            E[neg.reshape(-1)].reshape(neg.shape[0], neg.shape[1], -1)
    raise ValueError('mode %s not supported' % mode)             # model.py:149


def forward(model_name, state, sample, mode, gamma, hidden_dim, dtype=np.float32):
    """KGEModel.forward (model.py:72-164) -> [B, N] scores."""
    h, r, t = gather(state, sample, mode)
    rho = embedding_range(gamma, hidden_dim)
    return score_rows(model_name, h, r, t, mode, gamma, rho, state.get("modulus"), dtype)


# --------------------------------------------------------------------------------------------------
# loss (model.py:268-297) and its closed-form backward (autograd of model.py:301)
# --------------------------------------------------------------------------------------------------
def _logsigmoid(x):
    return np.minimum(x, 0) - np.log1p(np.exp(-np.abs(x)))


def _sigmoid(x):
    e = np.exp(-np.abs(x))
    return np.where(x >= 0, 1 / (1 + e), e / (1 + e))


def loss_and_dscore(neg_score, pos_score, weight, adversarial, alpha, uni_weight, dtype=np.float32):
    """model.py:270-288.  Returns (positive_sample_loss, negative_sample_loss, loss,
    dL/dneg_score [B,N], dL/dpos_score [B])."""
    f = dtype
    s = neg_score.astype(f)
    p = pos_score.astype(f).reshape(-1)
    B, N = s.shape
    if adversarial:                                              # model.py:272-273
        z = s * f(alpha)
        z = z - z.max(axis=1, keepdims=True)
        e = np.exp(z)
        w = e / e.sum(axis=1, keepdims=True, dtype=f)            # softmax(...).detach()
    else:                                                        # model.py:275: mean over N
        w = np.full((B, N), 1.0 / N, dtype=f)
    neg_row = (w * _logsigmoid(-s)).sum(axis=1, dtype=f)
    pos_row = _logsigmoid(p)                                     # model.py:279
    if uni_weight:                                               # model.py:281-283
        u = np.full((B,), 1.0 / B, dtype=f)
        pos_loss = -pos_row.mean(dtype=f)
        neg_loss = -neg_row.mean(dtype=f)
    else:                                                        # model.py:285-286
        wt = weight.astype(f)
        wsum = wt.sum(dtype=f)
        u = wt / wsum
        pos_loss = -(wt * pos_row).sum(dtype=f) / wsum
        neg_loss = -(wt * neg_row).sum(dtype=f) / wsum
    loss = (pos_loss + neg_loss) / f(2)                          # model.py:288
    # d(-sum_i u_i sum_j w_ij logsig(-s_ij))/ds_ij / 2 = u_i w_ij sigmoid(s_ij) / 2
    dneg = (f(0.5) * u[:, None] * w * _sigmoid(s)).astype(f)
    dpos = (-f(0.5) * u * _sigmoid(-p)).astype(f)
    return f(pos_loss), f(neg_loss), f(loss), dneg, dpos


def l3_regularization(state, reg, dtype=np.float32):
    """model.py:290-297: reg * (||E||_3^3 + ||R||_3^3); value and dense gradients 3*reg*x*|x|."""
    f = dtype
    E = state["entity_embedding"].astype(f)
    R = state["relation_embedding"].astype(f)
    val = f(reg) * ((np.abs(E) ** 3).sum(dtype=np.float64) + (np.abs(R) ** 3).sum(dtype=np.float64))
    return f(val), (f(3 * reg) * E * np.abs(E)), (f(3 * reg) * R * np.abs(R))


def score_backward(model_name, state, sample, mode, dscore, gamma, hidden_dim, dtype=np.float64):
    """Dense gradients of sum(dscore * score) w.r.t. E, R (and modulus): what autograd produces for
    model.py:86-146 (index_select backward = scatter-add) composed with model.py:166-249."""
    f = dtype
    E, R = state["entity_embedding"], state["relation_embedding"]
    h, r, t = (x.astype(f) for x in gather(state, sample, mode))
    rho = embedding_range(gamma, hidden_dim)
    g = np.asarray(dscore, dtype=f)
    if g.ndim == 1:
        g = g[:, None]
    g = g[:, :, None]
    dmod = None
    hb = mode == 'head-batch'
    if model_name == 'TransE':
        x = h + (r - t) if hb else (h + r) - t
        dx = -g * np.sign(x)
        dh, dr, dt = dx, dx, -dx
    elif model_name == 'DistMult':
        dh, dr, dt = g * (r * t), g * (h * t), g * (h * r)
    elif model_name == 'ComplEx':
        hr, hi = _split(h); rr, ri = _split(r); tr, ti = _split(t)
        # score = sum Re(h * r * conj(t))
        dh = np.concatenate([g * (rr * tr + ri * ti), g * (rr * ti - ri * tr)], axis=-1)
        dr = np.concatenate([g * (hr * tr + hi * ti), g * (hr * ti - hi * tr)], axis=-1)
        dt = np.concatenate([g * (hr * rr - hi * ri), g * (hr * ri + hi * rr)], axis=-1)
    elif model_name == 'RotatE':
        hr, hi = _split(h); tr, ti = _split(t)
        scale = f(np.float32(rho / PI_ROTATE))
        ph = r / scale
        c, s = np.cos(ph), np.sin(ph)
        if hb:
            a = c * tr + s * ti - hr
            b = c * ti - s * tr - hi
        else:
            a = hr * c - hi * s - tr
            b = hr * s + hi * c - ti
        m = np.sqrt(a * a + b * b)
        with np.errstate(divide='ignore', invalid='ignore'):
            da = np.where(m > 0, -g * a / m, 0)                   # norm subgradient 0 at m == 0
            db = np.where(m > 0, -g * b / m, 0)
        if hb:
            dh = np.concatenate([-da, -db], axis=-1)
            dt = np.concatenate([da * c - db * s, da * s + db * c], axis=-1)
            dth = -(da * tr + db * ti) * s + (da * ti - db * tr) * c
        else:
            dt = np.concatenate([-da, -db], axis=-1)
            dh = np.concatenate([da * c + db * s, -da * s + db * c], axis=-1)
            dth = -(da * hr + db * hi) * s + (-da * hi + db * hr) * c
        dr = dth / scale
    elif model_name == 'pRotatE':
        k = f(np.float32(rho / PI_PROTATE))
        mod = f(state["modulus"].reshape(())[()])
        ph, pr, pt = h / k, r / k, t / k
        x = ph + (pr - pt) if hb else (ph + pr) - pt
        sx = np.sin(x)
        dx = -g * mod * np.sign(sx) * np.cos(x)
        dh, dr, dt = dx / k, dx / k, -dx / k
        dmod = -(g[:, :, 0] * np.abs(sx).sum(axis=2)).sum().reshape(1, 1)
    else:
        raise ValueError('model %s not supported' % model_name)

    def red(x, like):      # broadcast-gradient: sum over N when the operand was [B,1,D]
        x = np.broadcast_to(x, np.broadcast_shapes(x.shape, (g.shape[0], g.shape[1], 1)))
        return x.sum(axis=1) if like.shape[1] == 1 else x.reshape(-1, x.shape[-1])

    gE = np.zeros(E.shape, dtype=f)
    gR = np.zeros(R.shape, dtype=f)
    if mode == 'single':
        s_ = np.asarray(sample)
        hi_, ri_, ti_ = s_[:, 0], s_[:, 1], s_[:, 2]
    else:
        pos, neg = (np.asarray(x) for x in sample)
        ri_ = pos[:, 1]
        hi_ = neg.reshape(-1) if hb else pos[:, 0]
        ti_ = pos[:, 2] if hb else neg.reshape(-1)
    np.add.at(gE, hi_, red(dh, h))
    np.add.at(gR, ri_, red(dr, r))
    np.add.at(gE, ti_, red(dt, t))
    out = {"entity_embedding": gE, "relation_embedding": gR}
    if dmod is not None:
        out["modulus"] = dmod.astype(f)
    return out


# --------------------------------------------------------------------------------------------------
# Adam (torch.optim.Adam defaults as built at run.py:266-269; torch/optim/adam.py _multi_tensor_adam)
# --------------------------------------------------------------------------------------------------
def adam_update(p, g, m, v, step, lr, beta1=0.9, beta2=0.999, eps=1e-8):
    """One in-place Adam update of fp32 arrays; ``step`` is the 1-based counter after increment."""
    f = np.float32
    m += (g - m) * f(1 - beta1)                                   # lerp_
    v *= f(beta2)
    v += f(1 - beta2) * g * g                                     # addcmul_
    bc1 = 1 - beta1 ** step
    bc2 = 1 - beta2 ** step
    step_size = -(lr / bc1)
    denom = np.sqrt(v) / f(bc2 ** 0.5) + f(eps)
    p += f(step_size) * (m / denom)                               # addcdiv_
    return p, m, v


class TrainState:
    """Tables + lazily created Adam moments (torch creates state on the first step)."""

    def __init__(self, model_name, state, gamma, hidden_dim):
        self.model_name, self.gamma, self.hidden_dim = model_name, gamma, hidden_dim
        self.state = {k: np.array(v, dtype=np.float32) for k, v in state.items()}
        self.adam = {}

    def reset_optimizer(self):
        """run.py:318-321 re-creates Adam (moments and step counters dropped) at LR decay."""
        self.adam = {}


def train_step(ts: TrainState, batch, *, lr, adversarial, alpha=1.0, uni_weight=False,
               regularization=0.0, compute_dtype=np.float32, return_grads=False):
    """KGEModel.train_step (model.py:251-312) on one (positive, negative, weight, mode) batch."""
    positive, negative, weight, mode = batch
    positive, negative = np.asarray(positive), np.asarray(negative)
    neg = forward(ts.model_name, ts.state, (positive, negative), mode, ts.gamma, ts.hidden_dim, compute_dtype)
    pos = forward(ts.model_name, ts.state, positive, 'single', ts.gamma, ts.hidden_dim, compute_dtype)
    pl, nl, loss, dneg, dpos = loss_and_dscore(neg, pos, np.asarray(weight), adversarial, alpha,
                                               uni_weight, compute_dtype)
    g1 = score_backward(ts.model_name, ts.state, (positive, negative), mode, dneg, ts.gamma, ts.hidden_dim)
    g2 = score_backward(ts.model_name, ts.state, positive, 'single', dpos, ts.gamma, ts.hidden_dim)
    grads = {k: g1[k] + g2[k] for k in g1}
    log = {}
    if regularization != 0.0:
        reg, rE, rR = l3_regularization(ts.state, regularization, np.float64)
        grads["entity_embedding"] += rE
        grads["relation_embedding"] += rR
        loss = compute_dtype(loss + reg)
        log["regularization"] = float(reg)
    grads = {k: v.astype(np.float32) for k, v in grads.items()}
    for name in ("entity_embedding", "relation_embedding", "modulus"):
        if name not in ts.state:
            continue
        st = ts.adam.setdefault(name, {"step": 0, "m": np.zeros_like(ts.state[name]),
                                       "v": np.zeros_like(ts.state[name])})
        st["step"] += 1
        adam_update(ts.state[name], grads[name], st["m"], st["v"], st["step"], lr)
    log.update({"positive_sample_loss": float(pl), "negative_sample_loss": float(nl), "loss": float(loss)})
    return (log, grads) if return_grads else log


# --------------------------------------------------------------------------------------------------
# filtered ranking (model.py:346-427 with the candidate/bias encoding of dataloader.py:134-154)
# --------------------------------------------------------------------------------------------------
def build_true_sets(all_true_triples):
    """Sets used by TestDataset's membership test (dataloader.py:126,138,142), regrouped per query."""
    true_heads, true_tails = {}, {}
    for h, r, t in all_true_triples:
        true_heads.setdefault((r, t), set()).add(h)
        true_tails.setdefault((h, r), set()).add(t)
    return true_heads, true_tails


def candidates_and_bias(triple, mode, nentity, true_heads, true_tails):
    """dataloader.py:137-151: filtered columns are replaced by the positive id with bias -1."""
    h, r, t = triple
    cand = np.arange(nentity, dtype=np.int64)
    bias = np.zeros(nentity, dtype=np.float32)
    if mode == 'head-batch':
        filt = np.fromiter(true_heads.get((r, t), ()), dtype=np.int64)
        pos = h
    elif mode == 'tail-batch':
        filt = np.fromiter(true_tails.get((h, r), ()), dtype=np.int64)
        pos = t
    else:
        raise ValueError('negative batch mode %s not supported' % mode)
    filt = filt[filt != pos]
    cand[filt] = pos
    bias[filt] = -1.0
    return cand, bias, pos


def rank_from_scores(score_row, positive_arg):
    """model.py:396-411 on one row: 1 + position of the positive's column in a descending argsort.
    Tie policy = stable descending sort (lowest column first among equal scores), which is always
    inside the set of answers the reference's (unstable on CPU) argsort can give; SURVEY 7 hard part 2."""
    order = np.argsort(-score_row.astype(np.float64), kind='stable')
    hit = np.nonzero(order == positive_arg)[0]
    assert hit.size == 1                                          # model.py:408
    return 1 + int(hit[0])


def filtered_ranks(model_name, state, test_triples, all_true_triples, nentity, gamma, hidden_dim,
                   dtype=np.float32, return_scores=False):
    """Per-query ranks in the reference's order: all head-batch queries, then all tail-batch."""
    true_heads, true_tails = build_true_sets(all_true_triples)
    ranks, rows = [], []
    for mode in ('head-batch', 'tail-batch'):                    # model.py:375
        for triple in test_triples:
            cand, bias, pos = candidates_and_bias(triple, mode, nentity, true_heads, true_tails)
            positive = np.asarray([triple], dtype=np.int64)
            score = forward(model_name, state, (positive, cand[None, :]), mode, gamma, hidden_dim, dtype)[0]
            score = score + bias.astype(score.dtype)              # model.py:393
            ranks.append(rank_from_scores(score, pos))
            if return_scores:
                rows.append(score)
    ranks = np.asarray(ranks, dtype=np.int64)
    return (ranks, np.stack(rows)) if return_scores else ranks


def metrics_from_ranks(ranks):
    """model.py:412-427: python-float means in query order."""
    logs = [{'MRR': 1.0 / r, 'MR': float(r), 'HITS@1': 1.0 if r <= 1 else 0.0,
             'HITS@3': 1.0 if r <= 3 else 0.0, 'HITS@10': 1.0 if r <= 10 else 0.0}
            for r in (int(x) for x in ranks)]
    return {k: sum(l[k] for l in logs) / len(logs) for k in logs[0]}


def countries_samples(test_triples, regions):
    """model.py:325-330: (h, r, region) for every region, y_true = region == tail."""
    sample, y_true = [], []
    for h, r, t in test_triples:
        for region in regions:
            y_true.append(1 if region == t else 0)
            sample.append((h, r, region))
    return np.asarray(sample, dtype=np.int64), np.asarray(y_true)


def average_precision(y_true, y_score):
    """sklearn.metrics.average_precision_score for binary labels (model.py:342): step-wise sum of
    precision * recall increments over distinct thresholds, descending."""
    y_true = np.asarray(y_true)
    y_score = np.asarray(y_score, dtype=np.float64)
    order = np.argsort(-y_score, kind='mergesort')
    ys, yt = y_score[order], y_true[order]
    distinct = np.nonzero(np.diff(ys))[0]
    idx = np.r_[distinct, yt.size - 1]
    tps = np.cumsum(yt)[idx]
    fps = 1 + idx - tps
    precision = tps / (tps + fps)
    recall = tps / tps[-1]
    return float(np.sum(np.diff(np.r_[0.0, recall]) * precision))


# --------------------------------------------------------------------------------------------------
# negative sampling (dataloader.py:13-119) -- distribution restated, plus the bit-exact Philox stream of
# kge_sample_negatives so that device negatives can be compared with ==
# --------------------------------------------------------------------------------------------------
def count_frequency(triples, start=4):
    """dataloader.py:77-93."""
    count = {}
    for h, r, t in triples:
        count[(h, r)] = count.get((h, r), start - 1) + 1
        count[(t, -r - 1)] = count.get((t, -r - 1), start - 1) + 1
    return count


def subsampling_weight(triples):
    """dataloader.py:33-34: sqrt(1 / (count(h,r) + count(t,-r-1))) in fp32."""
    count = count_frequency(triples)
    c = np.array([count[(h, r)] + count[(t, -r - 1)] for h, r, t in triples], dtype=np.float32)
    return np.sqrt(np.float32(1) / c)


def true_head_and_tail(triples):
    """dataloader.py:95-119."""
    true_head, true_tail = {}, {}
    for h, r, t in triples:
        true_tail.setdefault((h, r), set()).add(t)
        true_head.setdefault((r, t), set()).add(h)
    return true_head, true_tail


def _philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Philox4x32-10 (Salmon et al. 2011) on uint32 numpy arrays."""
    M0, M1, W0, W1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57), 0x9E3779B9, 0xBB67AE85
    c0, c1, c2, c3 = (np.asarray(x, dtype=np.uint64) for x in (c0, c1, c2, c3))
    k0, k1 = int(k0) & 0xFFFFFFFF, int(k1) & 0xFFFFFFFF
    mask = np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0, p1 = M0 * c0, M1 * c2
        n0 = ((p1 >> np.uint64(32)) ^ c1 ^ np.uint64(k0)) & mask
        n1 = p1 & mask
        n2 = ((p0 >> np.uint64(32)) ^ c3 ^ np.uint64(k1)) & mask
        n3 = p0 & mask
        c0, c1, c2, c3 = n0, n1, n2, n3
        k0, k1 = (k0 + W0) & 0xFFFFFFFF, (k1 + W1) & 0xFFFFFFFF
    return c0, c1, c2, c3


def sample_negatives(triples, index, mode, nentity, N, seed, step):
    """Bit-exact restatement of kge_sample_negatives: per (row, negative) pair p, candidates come from
    Philox(counter=(p, 0, attempt, step), key=seed') four at a time, id = (u32 * nentity) >> 32, the first candidate
    outside the row's true set is kept (dataloader.py:38-61 keeps the first survivors of a uniform stream)."""
    true_head, true_tail = true_head_and_tail(triples)
    seed = (int(seed) << 1) | (1 if mode == 'head-batch' else 0)
    k0, k1 = seed & 0xFFFFFFFF, ((seed >> 32) ^ (int(step) >> 32)) & 0xFFFFFFFF
    out = np.zeros((len(index), N), dtype=np.int64)
    for b, ti in enumerate(index):
        h, r, t = triples[int(ti)]
        true = true_head[(r, t)] if mode == 'head-batch' else true_tail[(h, r)]
        for n in range(N):
            p = b * N + n
            attempt = 0
            while True:
                r4 = _philox4x32_10([p & 0xFFFFFFFF], [p >> 32], [attempt], [int(step) & 0xFFFFFFFF], k0, k1)
                cands = [int((int(x[0]) * nentity) >> 32) for x in r4]
                ok = [c for c in cands if c not in true]
                if ok:
                    out[b, n] = ok[0]
                    break
                attempt += 1
    return out
