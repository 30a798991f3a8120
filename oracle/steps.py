"""CPU oracle: one minibatch update step of BPRMF / CML / GBPRMF / WRMF.  TEST INFRASTRUCTURE.

Restates (numpy, fp32 forward math) what one ``sess.run(train_op)`` does in the reference:

* BPRMF  -- /root/reference/src/models/pl/models/bprmf.py:52-75 (loss), :83-88 (Adagrad)
* CML    -- /root/reference/src/models/pl/models/cml.py:55-109 (loss), :119-129 (Adagrad + whole-table clip)
* GBPRMF -- /root/reference/src/models/pl/models/gbprmf.py:58-93 (loss), :101-106 (Adagrad on U, V, b)
* WRMF   -- /root/reference/src/models/basic/models/wrmf.py:52-75 (loss), :83-88 (Adagrad)
* MF     -- /root/reference/src/models/basic/models/mf.py:54-78 = WRMF with weight 1 (oracle/rating.py)
* SVD    -- /root/reference/src/models/basic/models/svd.py:52-80 (loss with the d x d kernel matrix, dense Adagrad on it)
* PRIGP  -- /root/reference/src/models/pl/models/prigp.py:92-137 (5-tuples; Adagrad on U, V only)
* CPLR   -- /root/reference/src/models/pl/models/cplr_u.py:99-144 (coefficient-weighted 4-tuples; Adagrad on U, V, b)

TF1 semantics relied on (third-party): gradients are evaluated at the
pre-update parameters; IndexedSlices of every gather of a variable are concatenated and
duplicate rows SUMMED, then ``SparseApplyAdagrad`` runs once per unique row:
``acc += g*g; var -= lr * g / sqrt(acc)`` with ``acc0 = 0.1`` and no epsilon; ``lr`` is the
constant captured at graph build (bprmf.py:134 -- the ``*= .98`` at :159 is cosmetic).

Pinning: TensorFlow itself is not installable here and the reference has no tests for this path, so
PARITY IS UNPINNED AGAINST THE TENSORFLOW BINARY.  What pins it instead: (1) the reference's own model
files, imported unmodified and run through their train() on the TF-1.x stand-in of oracle/tf1_shim
(oracle/gen_refgraph_golden.py -> tests/golden/*_refgraph_golden.npz: state after every step and the
metric values train() returned); (2) a torch-autograd restatement of the graphs (oracle/gen_golden.py).

Duplicate-row gradient sums are accumulated in float64 and rounded once to float32, so the
oracle sits in the middle of any fp32 summation order (TF's segment-sum or the GPU's atomics).
"""
import numpy as np

ADAGRAD, SGD = 0, 1
F = np.float32


def _segment_sum(n_rows, rows, grads):
    """Sum ``grads[k]`` over equal ``rows[k]`` -> (unique_rows, summed fp32 grads)."""
    rows = np.asarray(rows).reshape(-1).astype(np.int64)
    grads = np.asarray(grads, dtype=np.float64).reshape(len(rows), -1)
    uniq, inv = np.unique(rows, return_inverse=True)
    out = np.zeros((len(uniq), grads.shape[1]), dtype=np.float64)
    np.add.at(out, inv, grads)
    return uniq, out.astype(F)


def apply_rows(table, acc, rows, grads, lr, optimizer=ADAGRAD):
    """TF1 ``_apply_sparse_duplicate_indices`` + ``SparseApplyAdagrad`` (or plain SGD)."""
    uniq, g = _segment_sum(table.shape[0], rows, grads)
    g = g.reshape((len(uniq),) + table.shape[1:])
    if optimizer == ADAGRAD:
        a = acc[uniq] + g * g
        acc[uniq] = a
        table[uniq] = table[uniq] - F(lr) * g / np.sqrt(a)
    else:
        table[uniq] = table[uniq] - F(lr) * g
    return uniq


def clip_rows(table, clip_norm):
    """``tf.clip_by_norm(t, c, axes=[1])`` = t * c / max(||t||_2, c) per row (cml.py:121-122)."""
    c = F(clip_norm)
    nrm = np.sqrt(np.sum(table * table, axis=1, dtype=F), dtype=F)
    table *= (c / np.maximum(nrm, c))[:, None]
    return table


def _softplus_neg(x):
    # -log(sigmoid(x)) (bprmf.py:70) in its overflow-safe form
    return np.logaddexp(F(0), -x).astype(F)


def _sigm1(x):
    # sigmoid(x) - 1 = d/dx[-log sigmoid(x)]
    return (-1.0 / (1.0 + np.exp(x.astype(np.float64)))).astype(F)


def bpr_step(U, V, accU, accV, pairs, negs, lr=0.1, reg=0.02, optimizer=ADAGRAD):
    """bprmf.py:52-88.  In-place update of U, V (and accumulators); returns the batch loss."""
    u, i = pairs[:, 0].astype(np.int64), pairs[:, 1].astype(np.int64)
    j = negs.astype(np.int64)                                   # [B, W]
    Uu, Vi, Vj = U[u], V[i], V[j]                               # [B,d] [B,d] [B,W,d]
    ui = np.sum(Uu * Vi, axis=1, dtype=F)
    uj = np.sum(Uu[:, None, :] * Vj, axis=2, dtype=F)
    x = ui[:, None] - uj                                        # [B, W]
    reg = F(reg)
    loss = np.sum(_softplus_neg(x), dtype=np.float64) + reg * 0.5 * (
        np.sum(Uu * Uu, dtype=np.float64) + np.sum(Vi * Vi, dtype=np.float64) + np.sum(Vj * Vj, dtype=np.float64))
    s = _sigm1(x)                                               # [B, W]
    S = np.sum(s, axis=1, dtype=F)
    gU = np.einsum('bw,bwd->bd', s, Vi[:, None, :] - Vj).astype(F) + reg * Uu
    gVi = S[:, None] * Uu + reg * Vi
    gVj = -s[:, :, None] * Uu[:, None, :] + reg * Vj
    apply_rows(U, accU, u, gU, lr, optimizer)
    apply_rows(V, accV, np.concatenate([i, j.reshape(-1)]),
               np.concatenate([gVi, gVj.reshape(-1, V.shape[1])]), lr, optimizer)
    return float(loss)


def cml_forward(U, V, pairs, negs, margin, use_rank_weight, n_items):
    """cml.py:55-88 forward quantities; also returns the smallest |hinge/impostor argument| so a test
    can assert its inputs are not within rounding distance of a relu/indicator kink."""
    u, i = pairs[:, 0].astype(np.int64), pairs[:, 1].astype(np.int64)
    j = negs.astype(np.int64)
    Uu, Vi, Vj = U[u], V[i], V[j]
    dp = np.sum((Uu - Vi) ** 2, axis=1, dtype=F)
    dn = np.sum((Uu[:, None, :] - Vj) ** 2, axis=2, dtype=F)    # [B, W]
    dmin = dn.min(axis=1)
    m = F(margin)
    h = dp - dmin + m
    imp_arg = dp[:, None] - dn + m
    if use_rank_weight:
        rank = np.mean((imp_arg > 0).astype(F), axis=1, dtype=F) * F(n_items)
        omega = np.log(rank + F(1.0)).astype(F)
    else:
        omega = np.ones_like(dp)
    kpp = np.min(np.abs(imp_arg), axis=1) if imp_arg.size else np.zeros(0, F)
    if dn.shape[1] > 1 and dn.shape[0]:
        # a near-tie of the two closest negatives is a kink of the min too (exact ties = the same item drawn twice are not)
        ds = np.sort(dn, axis=1)
        gap = ds[:, 1] - ds[:, 0]
        kpp = np.minimum(kpp, np.where(gap == 0, np.inf, gap).astype(F))
    return dict(Uu=Uu, Vi=Vi, Vj=Vj, dp=dp, dn=dn, dmin=dmin, h=h, omega=omega,
                kink=float(np.min(np.abs(imp_arg))) if imp_arg.size else np.inf, kink_per_pair=kpp)


def cml_step(U, V, accU, accV, pairs, negs, lr=0.1, reg_cov=1.0, margin=1.5, use_rank_weight=True,
             clip_norm=1.0, optimizer=ADAGRAD, clip_whole_table=True):
    """cml.py:55-129.  ``clip_whole_table`` True = the reference (both tables, every step);
    False clips only rows touched by this batch (what the fused kernel does after the first step)."""
    n_items = V.shape[0]
    f = cml_forward(U, V, pairs, negs, margin, use_rank_weight, n_items)
    u, i = pairs[:, 0].astype(np.int64), pairs[:, 1].astype(np.int64)
    j = negs.astype(np.int64)
    Uu, Vi, Vj, dn, dmin, h, omega = f['Uu'], f['Vi'], f['Vj'], f['dn'], f['dmin'], f['h'], f['omega']
    loss = np.sum(np.maximum(h, 0) * omega, dtype=np.float64)
    active = (h > 0).astype(F)
    ties = (dn == dmin[:, None]).astype(F)                      # reduce_min grad: equal split among ties
    ties /= ties.sum(axis=1, keepdims=True)
    coef = F(2.0) * omega * active                              # [B]
    dUj = Uu[:, None, :] - Vj                                   # [B, W, d]
    gU = coef[:, None] * ((Uu - Vi) - np.einsum('bw,bwd->bd', ties, dUj).astype(F))
    gVi = -coef[:, None] * (Uu - Vi)
    gVj = (coef[:, None] * ties)[:, :, None] * dUj
    if reg_cov > 0:                                             # cml.py:101-104,109 (an L2 term, not a covariance)
        c = F(reg_cov)
        loss += c * 0.5 * (np.sum(Uu * Uu, dtype=np.float64) + np.sum(Vi * Vi, dtype=np.float64)
                           + np.sum(Vj * Vj, dtype=np.float64))
        gU = gU + c * Uu
        gVi = gVi + c * Vi
        gVj = gVj + c * Vj
    tu = apply_rows(U, accU, u, gU, lr, optimizer)
    rows_v = np.concatenate([i, j.reshape(-1)])
    tv = apply_rows(V, accV, rows_v, np.concatenate([gVi, gVj.reshape(-1, V.shape[1])]), lr, optimizer)
    if clip_whole_table:
        clip_rows(U, clip_norm)
        clip_rows(V, clip_norm)
    else:
        U[tu] = clip_rows(U[tu], clip_norm)
        V[tv] = clip_rows(V[tv], clip_norm)
    return float(loss)


def gbpr_step(U, V, b, accU, accV, accb, pairs, negs, group, lr=0.1, reg=0.02, rho=0.5, optimizer=ADAGRAD):
    """gbprmf.py:58-106.  L2 covers U_u, U_g, V_i and b_j only (gbprmf.py:60-64)."""
    u, i = pairs[:, 0].astype(np.int64), pairs[:, 1].astype(np.int64)
    j, g = negs.astype(np.int64), group.astype(np.int64)
    G = g.shape[1]
    Uu, Vi, Vj, Ug = U[u], V[i], V[j], U[g]
    bi, bj = b[i], b[j]
    rho = F(rho)
    reg = F(reg)
    ui_u = np.sum(Uu * Vi, axis=1, dtype=F)
    Ugs = np.sum(Ug, axis=1, dtype=F)                           # [B, d]
    ui_g = np.sum(Ug * Vi[:, None, :], axis=(1, 2), dtype=F) / F(G)
    ui = rho * ui_g + (F(1) - rho) * ui_u + bi
    uj = np.sum(Uu[:, None, :] * Vj, axis=2, dtype=F) + bj
    x = ui[:, None] - uj
    loss = np.sum(_softplus_neg(x), dtype=np.float64) + reg * 0.5 * (
        np.sum(Uu * Uu, dtype=np.float64) + np.sum(Ug * Ug, dtype=np.float64)
        + np.sum(Vi * Vi, dtype=np.float64) + np.sum(bj * bj, dtype=np.float64))
    s = _sigm1(x)
    S = np.sum(s, axis=1, dtype=F)
    gUu = (F(1) - rho) * S[:, None] * Vi - np.einsum('bw,bwd->bd', s, Vj).astype(F) + reg * Uu
    gUg = (rho / F(G)) * S[:, None, None] * Vi[:, None, :] + reg * Ug
    gVi = S[:, None] * ((rho / F(G)) * Ugs + (F(1) - rho) * Uu) + reg * Vi
    gVj = -s[:, :, None] * Uu[:, None, :]
    gbi = S
    gbj = -s + reg * bj
    apply_rows(U, accU, np.concatenate([u, g.reshape(-1)]),
               np.concatenate([gUu, gUg.reshape(-1, U.shape[1])]), lr, optimizer)
    rows_v = np.concatenate([i, j.reshape(-1)])
    apply_rows(V, accV, rows_v, np.concatenate([gVi, gVj.reshape(-1, V.shape[1])]), lr, optimizer)
    apply_rows(b, accb, rows_v, np.concatenate([gbi, gbj.reshape(-1)]), lr, optimizer)
    return float(loss)


def wrmf_step(U, V, accU, accV, uir, lr=0.1, reg=0.02, weight=1.0, optimizer=ADAGRAD):
    """wrmf.py:52-88 (the minibatch-Adagrad WRMF the reference actually implements).
    ``uir`` is the sampler_rating batch: columns (user, item, rating)."""
    u, i = uir[:, 0].astype(np.int64), uir[:, 1].astype(np.int64)
    r = uir[:, 2].astype(F)
    Uu, Vi = U[u], V[i]
    reg = F(reg)
    w = F(weight)
    e = np.sum(Uu * Vi, axis=1, dtype=F) - r
    loss = 0.5 * w * np.sum(e * e, dtype=np.float64) + reg * 0.5 * (
        np.sum(Uu * Uu, dtype=np.float64) + np.sum(Vi * Vi, dtype=np.float64))
    gU = (w * e)[:, None] * Vi + reg * Uu
    gV = (w * e)[:, None] * Uu + reg * Vi
    apply_rows(U, accU, u, gU, lr, optimizer)
    apply_rows(V, accV, i, gV, lr, optimizer)
    return float(loss)


def svd_step(U, V, K, accU, accV, accK, uir, lr=0.1, reg=0.02, optimizer=ADAGRAD):
    """svd.py:52-80: ``pred = sum((U_u @ K) * V_i)``, ``L = l2_loss(pred - r) + reg * (l2_loss(U_u) + l2_loss(V_i))``
    (no L2 on K); Adagrad on U, V (sparse, duplicates summed) and on the whole kernel matrix K (dense gradient)."""
    u, i = uir[:, 0].astype(np.int64), uir[:, 1].astype(np.int64)
    r = uir[:, 2].astype(F)
    Uu, Vi = U[u], V[i]
    reg = F(reg)
    t = (Vi @ K.T).astype(F)            # t[b, a] = sum_c K[a, c] V_i[c]
    s = (Uu @ K).astype(F)              # s[b, c] = sum_a U_u[a] K[a, c]
    e = np.sum(Uu * t, axis=1, dtype=F) - r
    loss = 0.5 * np.sum(e * e, dtype=np.float64) + reg * 0.5 * (
        np.sum(Uu * Uu, dtype=np.float64) + np.sum(Vi * Vi, dtype=np.float64))
    gU = e[:, None] * t + reg * Uu
    gV = e[:, None] * s + reg * Vi
    gK = ((e[:, None] * Uu).astype(np.float64).T @ Vi.astype(np.float64)).astype(F)
    apply_rows(U, accU, u, gU, lr, optimizer)
    apply_rows(V, accV, i, gV, lr, optimizer)
    if optimizer == ADAGRAD:
        accK += gK * gK
        K -= F(lr) * gK / np.sqrt(accK)
    else:
        K -= F(lr) * gK
    return float(loss)


def truncated_normal(rng, shape, mean=0.0, stddev=0.1):
    """tf.truncated_normal_initializer: resample beyond +-2 sigma (bprmf.py:30)."""
    x = rng.standard_normal(shape)
    bad = np.abs(x) > 2
    while bad.any():
        x[bad] = rng.standard_normal(int(bad.sum()))
        bad = np.abs(x) > 2
    return (mean + stddev * x).astype(F)


def prigp_step(U, V, b, accU, accV, uijtk, lr=0.1, reg=0.01, alpha=1.0, optimizer=ADAGRAD):
    """/root/reference/src/models/pl/models/prigp.py:92-137.  x_um = <U_u, V_m> + b_m;
    L = sum -log s(x_ui - x_uj) + alpha sum -log s(x_ut - x_uk) + reg (l2(U_u) + l2(V[i,j,t,k]) + l2(b[i,j,t,k]));
    Adagrad on user_embed and item_embed ONLY (var_list, prigp.py:134): the bias is read, never updated."""
    u = uijtk[:, 0].astype(np.int64)
    m = uijtk[:, 1:].astype(np.int64)                           # [B, 4] = (i, j, t, k)
    Uu, Vm, bm = U[u], V[m], b[m]
    reg, alpha = F(reg), F(alpha)
    x = np.sum(Uu[:, None, :] * Vm, axis=2, dtype=F) + bm       # [B, 4]
    x1, x2 = x[:, 0] - x[:, 1], x[:, 2] - x[:, 3]
    loss = np.sum(_softplus_neg(x1), dtype=np.float64) + alpha * np.sum(_softplus_neg(x2), dtype=np.float64) + reg * 0.5 * (
        np.sum(Uu * Uu, dtype=np.float64) + np.sum(Vm * Vm, dtype=np.float64) + np.sum(bm * bm, dtype=np.float64))
    s1, s2 = _sigm1(x1), alpha * _sigm1(x2)
    g = np.stack([s1, -s1, s2, -s2], axis=1).astype(F)          # dL/dx_um
    gU = np.einsum('bm,bmd->bd', g, Vm).astype(F) + reg * Uu
    gV = g[:, :, None] * Uu[:, None, :] + reg * Vm
    apply_rows(U, accU, u, gU, lr, optimizer)
    apply_rows(V, accV, m.reshape(-1), gV.reshape(-1, V.shape[1]), lr, optimizer)
    return float(loss)


def cplr_step(U, V, b, accU, accV, accb, uitj, coefs, lr=0.1, reg=0.01, alpha=1.0, beta=1.0, gamma=1.0, optimizer=ADAGRAD):
    """/root/reference/src/models/pl/models/cplr_u.py:99-144.  c_ij = coef_ui + 1, c_tj = coef_ut + 1, c_it = c_ij / c_tj;
    L = alpha sum -log s(c_it (x_ui - x_ut)) + beta sum -log s(c_tj (x_ut - x_uj)) + gamma sum -log s(c_ij (x_ui - x_uj))
        + reg (l2(U_u) + l2(V[i,t,j]) + l2(b[i,t,j]));  Adagrad on user_embed, item_embed and item_bias (:141)."""
    u = uitj[:, 0].astype(np.int64)
    m = uitj[:, 1:].astype(np.int64)                            # [B, 3] = (i, t, j)
    Uu, Vm, bm = U[u], V[m], b[m]
    reg, alpha, beta, gamma = F(reg), F(alpha), F(beta), F(gamma)
    coefs = np.asarray(coefs, dtype=F)
    cij, ctj = coefs[:, 0] + F(1), coefs[:, 1] + F(1)
    cit = (cij / ctj).astype(F)
    x = np.sum(Uu[:, None, :] * Vm, axis=2, dtype=F) + bm
    z1, z2, z3 = cit * (x[:, 0] - x[:, 1]), ctj * (x[:, 1] - x[:, 2]), cij * (x[:, 0] - x[:, 2])
    loss = (alpha * np.sum(_softplus_neg(z1), dtype=np.float64) + beta * np.sum(_softplus_neg(z2), dtype=np.float64)
            + gamma * np.sum(_softplus_neg(z3), dtype=np.float64) + reg * 0.5 * (
                np.sum(Uu * Uu, dtype=np.float64) + np.sum(Vm * Vm, dtype=np.float64) + np.sum(bm * bm, dtype=np.float64)))
    a1, a2, a3 = alpha * cit * _sigm1(z1), beta * ctj * _sigm1(z2), gamma * cij * _sigm1(z3)
    g = np.stack([a1 + a3, a2 - a1, -a2 - a3], axis=1).astype(F)
    gU = np.einsum('bm,bmd->bd', g, Vm).astype(F) + reg * Uu
    gV = g[:, :, None] * Uu[:, None, :] + reg * Vm
    gb = g + reg * bm
    apply_rows(U, accU, u, gU, lr, optimizer)
    apply_rows(V, accV, m.reshape(-1), gV.reshape(-1, V.shape[1]), lr, optimizer)
    apply_rows(b, accb, m.reshape(-1), gb.reshape(-1), lr, optimizer)
    return float(loss)
