"""CPU oracle, multi-threaded variant: the same minibatch step as oracle/steps.py (BPRMF / CML) written with torch CPU
ops so that the CPU baseline of bench.py can use every host core (torch intra-op threads).  TEST INFRASTRUCTURE; it is
checked against oracle/steps.py in tests/test_oracle_steps.py and is only ever run by bench.py's CPU legs.

Restates bprmf.py:52-88 and cml.py:55-129 of the reference: gradients at pre-update parameters,
duplicate rows summed (index_add_), one TF1-Adagrad apply per touched row (acc0 = 0.1, no epsilon), CML whole-table clip."""
import torch


def _apply(table, acc, rows, grads, lr):
    uniq, inv = torch.unique(rows, return_inverse=True)
    g = torch.zeros(uniq.numel(), table.shape[1], dtype=torch.float32)
    g.index_add_(0, inv, grads)
    a = acc[uniq] + g * g
    acc[uniq] = a
    table[uniq] = table[uniq] - lr * g / torch.sqrt(a)


def bpr_step(U, V, accU, accV, pairs, negs, lr=0.1, reg=0.02):
    u, i, j = pairs[:, 0].long(), pairs[:, 1].long(), negs.long()
    Uu, Vi, Vj = U[u], V[i], V[j]
    x = (Uu * Vi).sum(1, keepdim=True) - (Uu[:, None, :] * Vj).sum(2)
    s = torch.sigmoid(x) - 1.0
    S = s.sum(1, keepdim=True)
    gU = (s[:, :, None] * (Vi[:, None, :] - Vj)).sum(1) + reg * Uu
    gVi = S * Uu + reg * Vi
    gVj = -s[:, :, None] * Uu[:, None, :] + reg * Vj
    _apply(U, accU, u, gU, lr)
    _apply(V, accV, torch.cat([i, j.reshape(-1)]), torch.cat([gVi, gVj.reshape(-1, V.shape[1])]), lr)


def cml_step(U, V, accU, accV, pairs, negs, lr=0.1, reg_cov=1.0, margin=1.5, use_rank_weight=True, clip_norm=1.0):
    n_items = V.shape[0]
    u, i, j = pairs[:, 0].long(), pairs[:, 1].long(), negs.long()
    Uu, Vi, Vj = U[u], V[i], V[j]
    dp = ((Uu - Vi) ** 2).sum(1)
    dn = ((Uu[:, None, :] - Vj) ** 2).sum(2)
    dmin, wmin = dn.min(1)
    h = dp - dmin + margin
    if use_rank_weight:
        omega = torch.log(((dp[:, None] - dn + margin) > 0).float().mean(1) * n_items + 1.0)
    else:
        omega = torch.ones_like(dp)
    coef = 2.0 * omega * (h > 0).float()
    tie = torch.zeros_like(dn)
    tie.scatter_(1, wmin[:, None], 1.0)
    dUj = Uu[:, None, :] - Vj
    gU = coef[:, None] * ((Uu - Vi) - (tie[:, :, None] * dUj).sum(1))
    gVi = -coef[:, None] * (Uu - Vi)
    gVj = (coef[:, None] * tie)[:, :, None] * dUj
    if reg_cov > 0:
        gU, gVi, gVj = gU + reg_cov * Uu, gVi + reg_cov * Vi, gVj + reg_cov * Vj
    _apply(U, accU, u, gU, lr)
    _apply(V, accV, torch.cat([i, j.reshape(-1)]), torch.cat([gVi, gVj.reshape(-1, V.shape[1])]), lr)
    for t in (U, V):                                        # cml.py:119-129: both whole tables, every step
        n = t.norm(dim=1, keepdim=True)
        t.mul_(clip_norm / torch.clamp(n, min=clip_norm))
