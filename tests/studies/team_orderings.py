"""Pair orderings of the team Jacobi kernel on 512 x 512 config-4 matrices (tests/studies/harvest_wide_thetas.py): blocks
of 8 rows, one CTA sweeping all blocks (p ascending, q > p ascending) against the cluster ordering (2 CS groups of
contiguous blocks: pairs inside the groups first, then a round-robin tournament between groups) and variants of the
latter.  A (p, q) block visit = 8 ring steps of 8 disjoint pairs; internal pairs of a block = circle method.
Counts sweeps with the device's thresholds and stopping rule."""
import pickle
import sys
import time

import numpy as np

sys.path.insert(0, '.')
from oracle import device_model as dm  # noqa: E402

BRW = 8
EPS = dm.EPS


class Sweeper:
    def __init__(self, X, thresholds=dm.THRESHOLDS):
        self.X = np.array(X, dtype=complex)
        self.M, self.N = self.X.shape
        self.tol2 = (2 * np.sqrt(self.N) * EPS) ** 2
        self.dead = dm.DEAD_REL2 * np.sum(np.abs(self.X) ** 2)
        self.thresholds = thresholds

    def begin(self, sweep):
        self.nrm2 = np.sum(np.abs(self.X) ** 2, axis=1)
        self.thr2 = max(self.tol2, self.thresholds[sweep]) if sweep < len(self.thresholds) else self.tol2
        self.small2 = self.tol2 if self.thr2 > self.tol2 else dm.SMALL_REL2
        self.nrot = self.nbig = 0

    def pairs(self, I, J):
        X, nrm2 = self.X, self.nrm2
        ai, aj = nrm2[I], nrm2[J]
        g = np.sum(X[I] * X[J].conj(), axis=1)
        g2 = np.abs(g) ** 2
        alive = (ai > self.dead) & (aj > self.dead)
        self.nbig += int(np.sum(alive & (g2 > self.small2 * ai * aj)))
        act = alive & (g2 > self.thr2 * ai * aj)
        if not act.any():
            return
        I, J, g, ai, aj = I[act], J[act], g[act], ai[act], aj[act]
        ga = np.abs(g)
        dd = aj - ai
        t = np.copysign(2 * ga / (np.abs(dd) + np.sqrt(dd * dd + 4 * ga * ga)), dd)
        cs = 1 / np.sqrt(1 + t * t)
        se = cs * t * g / ga
        xi, xj = X[I], X[J]
        X[I] = cs[:, None] * xi - se[:, None] * xj
        X[J] = se.conj()[:, None] * xi + cs[:, None] * xj
        nrm2[I], nrm2[J] = ai - t * ga, aj + t * ga
        self.nrot += int(act.sum())

    def internal(self, p):
        r0 = p * BRW
        n = min(BRW, self.M - r0)
        for r in range(n - 1):
            I, J = dm.rr_pairs(n, r)
            self.pairs(I + r0, J + r0)

    def visit(self, p, q):
        i = np.arange(BRW)
        for s in range(BRW):
            self.pairs(p * BRW + i, q * BRW + (i + s) % BRW)


def intra(sw, blocks):
    for a, p in enumerate(blocks):
        sw.internal(p)
        for q in blocks[a + 1:]:
            sw.visit(p, q)


def cross(sw, A, B):
    for p in A:
        for q in B:
            sw.visit(p, q)


def rr_groups(ng, rd):
    return [tuple(int(v) for v in (I, J)) for I, J in zip(*dm.rr_pairs(ng, rd))]


def run(X, groups, cross_first=False, rounds=None):
    """groups: list of lists of block indices.  One group = the single-CTA order."""
    sw = Sweeper(X)
    ng = len(groups)
    hist = []
    for sweep in range(dm.MAX_SWEEPS):
        sw.begin(sweep)
        def do_intra():
            for g in groups:
                intra(sw, g)
        def do_cross():
            for rd in (rounds if rounds is not None else range(ng - 1)):
                for ga, gb in rr_groups(ng, rd):
                    cross(sw, groups[ga], groups[gb])
        if ng == 1:
            do_intra()
        elif cross_first:
            do_cross(); do_intra()
        else:
            do_intra(); do_cross()
        hist.append(sw.nrot)
        if sw.nbig == 0:
            break
    return len(hist), sum(hist) / (sw.M * (sw.M - 1) / 2)


if __name__ == '__main__':
    hv = pickle.load(open('tests/studies/_thetas_wide.pkl', 'rb'))
    nmat = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    for (theta, chiR) in hv[:nmat]:
        R = np.linalg.qr(theta[:, dm.interleave_perm(chiR)], mode='r')
        nblk = R.shape[0] // BRW
        allb = list(range(nblk))
        variants = {
            'one CTA: all blocks, p ascending': ([allb], {}),
            'cluster of 4: 8 contiguous groups, intra first (the kernel)': ([allb[g * 8:(g + 1) * 8] for g in range(8)], {}),
            'cluster of 4: 8 contiguous groups, cross first': ([allb[g * 8:(g + 1) * 8] for g in range(8)], {'cross_first': True}),
            'cluster of 4: 8 interleaved groups (block b in group b mod 8)': ([allb[g::8] for g in range(8)], {}),
            'cluster of 4: rounds in reverse order': ([allb[g * 8:(g + 1) * 8] for g in range(8)], {'rounds': list(range(6, -1, -1))}),
            'cluster of 2: 4 contiguous groups': ([allb[g * 16:(g + 1) * 16] for g in range(4)], {}),
        }
        print('matrix', R.shape)
        for name, (groups, kw) in variants.items():
            t0 = time.time()
            ns, rot = run(R, groups, **kw)
            print(f'  {name:66s} sweeps {ns:2d}  rotations per pair {rot:.2f}  ({time.time() - t0:.0f} s)', flush=True)
