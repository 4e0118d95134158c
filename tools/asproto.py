"""numpy prototype of the block active-set iteration (qp_solver.h: qr_active_set) to try update rules on the CPU.
python scratch/asproto.py [B] [rule]"""
import sys, numpy as np
_R = __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))); sys.path.insert(0, _R); sys.path.insert(0, _R + '/tests')
import _pkg
pkg = _pkg.load()
from quadruped_robot_b200 import capi
import emul_binding as EB

FEAS, MULT = 1e-9, 1e-11


def dof_basis(act, mu_, ub):
    a0, a1, a2, a3, cap = act & 1, (act >> 1) & 1, (act >> 2) & 1, (act >> 3) & 1, (act >> 4) & 1
    if a0 + a1 == 2 or a2 + a3 == 2 or a0 + a1 + a2 + a3 >= 3:
        return None, np.zeros(3)
    im = 1.0 / mu_
    kx = -im if a0 else (im if a1 else 0.0)
    ky = -im if a2 else (im if a3 else 0.0)
    cols = []
    if not (a0 | a1): cols.append([1.0, 0, 0])
    if not (a2 | a3): cols.append([0, 1.0, 0])
    p = np.zeros(3)
    if cap:
        p = np.array([kx * ub, ky * ub, ub])
    else:
        cols.append([kx, ky, 1.0])
    Z = np.array(cols).T.reshape(3, len(cols)) if cols else np.zeros((3, 0))
    return Z, p


def round_solve(H, g, ub, mu_, acts):
    nf = len(acts)
    Zs, ps = [], []
    nred = 0
    for f in range(nf):
        Z, p = dof_basis(acts[f], mu_, ub[f])
        Zs.append(Z); ps.append(p)
        nred += 0 if Z is None else Z.shape[1]
    Zb = np.zeros((3 * nf, nred)); pv = np.zeros(3 * nf)
    o = 0
    for f in range(nf):
        pv[3 * f:3 * f + 3] = ps[f]
        if Zs[f] is not None:
            d = Zs[f].shape[1]
            Zb[3 * f:3 * f + 3, o:o + d] = Zs[f]; o += d
    q = H @ pv + g
    if nred:
        y = np.linalg.solve(Zb.T @ H @ Zb, -Zb.T @ q)
        x = Zb @ y + pv
    else:
        x = pv
    r = H @ x + g
    return x, r, nred, [Z is None for Z in Zs]


def update_base(acts, x, r, ub, mu_, vert):
    im = 1.0 / mu_
    new = list(acts)
    for f in range(len(acts)):
        rf = r[3 * f:3 * f + 3]; act = acts[f]; nact = act
        if vert[f]:
            if rf[2] < (abs(rf[0]) + abs(rf[1])) * im - MULT:
                l0 = rf[0] * im if rf[0] > 0 else 0.0; l1 = -rf[0] * im if rf[0] < 0 else 0.0
                l2 = rf[1] * im if rf[1] > 0 else 0.0; l3 = -rf[1] * im if rf[1] < 0 else 0.0
                nact = (1 if l0 > 0 else 0) | (2 if l1 > 0 else 0) | (4 if l2 > 0 else 0) | (8 if l3 > 0 else 0)
                if (nact & 3) and (nact & 12):
                    sx = 1.0 if nact & 1 else -1.0; sy = 1.0 if nact & 4 else -1.0
                    b0 = sx * mu_ * rf[0] + rf[2]; b1 = sy * mu_ * rf[1] + rf[2]
                    dd = mu_ * mu_ + 1.0; det = dd * dd - 1.0
                    lx = (dd * b0 - b1) / det; ly = (dd * b1 - b0) / det
                    if lx < 0 or ly < 0:
                        if lx < ly: nact &= ~3
                        else: nact &= ~12
        else:
            fx, fy, fz = x[3 * f:3 * f + 3]
            c = [mu_ * fx + fz, -mu_ * fx + fz, mu_ * fy + fz, -mu_ * fy + fz, ub[f] - fz]
            viol = 0
            for k in range(5):
                if not ((act >> k) & 1) and c[k] < -FEAS: viol |= 1 << k
            if viol:
                nact = act | viol
            elif act:
                lx = rf[0] * im if act & 1 else (-rf[0] * im if act & 2 else 0.0)
                ly = rf[1] * im if act & 4 else (-rf[1] * im if act & 8 else 0.0)
                lc = (lx + ly - rf[2]) if act & 16 else 0.0
                worst = -MULT; drop = 0
                if (act & 3) and lx < worst: worst = lx; drop = act & 3
                if (act & 12) and ly < worst: worst = ly; drop = act & 12
                if (act & 16) and lc < worst: worst = lc; drop = 16
                if drop: nact = act & ~drop
        new[f] = nact
    return new


def tiles(nb): return sum((nb - 1 - k) * (nb - k) // 2 for k in range(nb))


def solve(H, g, ub, mu_, links, rule, max_rounds=30, trace=False):
    nf = len(ub)
    acts = [0] * nf
    cost = 0; sizes = []
    for rd in range(max_rounds):
        x, r, nred, vert = round_solve(H, g, ub, mu_, acts)
        sizes.append(nred); cost += tiles((nred + 2) // 3)
        new = update_base(acts, x, r, ub, mu_, vert)
        if trace: print('   ', ''.join('%x' % a if a < 16 else chr(ord('A') + a - 16) for a in acts), nred)
        if new == acts:
            return rd + 1, cost, sizes, x
        if rule is not None:
            new = rule(acts, new, x, r, ub, mu_, links, rd)
        acts = new
    return max_rounds, cost, sizes, x


def rule_extrap(acts, new, x, r, ub, mu_, links, rd):
    """temporal extrapolation: a foot-step whose set GREW this round hands its new set to the same leg's next foot-step
    when that one is still a subset of it"""
    out = list(new)
    for f in range(len(acts)):
        if new[f] != acts[f] and (new[f] & acts[f]) == acts[f]:
            n = links[f]
            if n >= 0 and (new[n] | new[f]) == new[f] and new[n] != new[f]:
                out[n] = new[f]
    return out


RULES = {'base': None, 'extrap': rule_extrap}


def load(robot='a1', h=10, B=100, seed=5, gait='trot'):
    em = EB.load()
    batch = pkg.synth.make_mpc_batch(robot, h, 0.03, B, seed=seed, gait=gait)
    P = capi.params_of(batch["robot"], h, 0.03)
    H, g, ub = em.condense(P, batch)
    probs = []
    for i in range(B):
        gt = batch["gait"][i].reshape(h, 4)
        st = [(k, l) for k in range(h) for l in range(4) if gt[k, l] * P.f_max > 0]
        idx = np.array([12 * k + 3 * l + a for (k, l) in st for a in range(3)])
        Hs = H[i].astype(np.float64); Hs = 0.5 * (Hs + Hs.T)
        Hs = Hs[np.ix_(idx, idx)]; gs = g[i].astype(np.float64)[idx]
        ubs = np.array([float(np.float32(gt[k, l]) * np.float32(P.f_max)) for (k, l) in st])
        links = [-1] * len(st)
        for a, (k, l) in enumerate(st):
            for b, (k2, l2) in enumerate(st):
                if l2 == l and k2 == k + 1: links[a] = b
        mu_ = float(np.float32(1.0) / np.float32(P.mu))
        probs.append((Hs, gs, ubs, mu_, links))
    return probs


if __name__ == "__main__" and sys.argv[1:2] != ["more"]:
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    names = sys.argv[2:] or ['base', 'extrap']
    probs = load(B=B)
    for name in names:
        R, Cst = [], []
        for (H, g, ub, mu_, links) in probs:
            rd, cost, sizes, x = solve(H, g, ub, mu_, links, RULES[name])
            R.append(rd); Cst.append(cost)
        print(name, 'rounds', np.mean(R), 'max', max(R), 'tiles', np.mean(Cst))


def show(i, name, B=8):
    probs = load(B=B)
    H, g, ub, mu_, links = probs[i]
    print(links)
    print(solve(H, g, ub, mu_, links, RULES[name], trace=True)[:3])


def rule_fill0(acts, new, x, r, ub, mu_, links, rd):
    out = list(new)
    if rd == 0:
        for f in range(len(acts)):
            if new[f]:
                n = links[f]
                while n >= 0 and new[n] == 0:
                    out[n] = new[f]; n = links[n]
    return out


def mk_extrap(maxrd, steps=1):
    def rule(acts, new, x, r, ub, mu_, links, rd):
        out = list(new)
        if rd < maxrd:
            for f in range(len(acts)):
                if new[f] != acts[f] and (new[f] & acts[f]) == acts[f]:
                    n = links[f]; s = 0
                    while n >= 0 and s < steps and (new[n] | new[f]) == new[f] and new[n] != new[f]:
                        out[n] = new[f]; n = links[n]; s += 1
        return out
    return rule


RULES.update(fill0=rule_fill0, ex2=mk_extrap(2), ex3=mk_extrap(3), ex4=mk_extrap(4), ex2s2=mk_extrap(2, 2), ex3s2=mk_extrap(3, 2))
if __name__ == '__main__' and len(sys.argv) > 3 and sys.argv[1] == 'more':
    B = int(sys.argv[2])
    probs = load(B=B)
    for name in sys.argv[3:]:
        R, Cst = [], []
        for (H, g, ub, mu_, links) in probs:
            rd, cost, sizes, x = solve(H, g, ub, mu_, links, RULES[name])
            R.append(rd); Cst.append(cost)
        print(name, 'rounds', np.mean(R), 'max', max(R), 'tiles', np.mean(Cst))
