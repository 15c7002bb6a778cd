"""Independent numpy restatement of the SOGP recursion, written from matlab/sogp.m:58-191
(the second witness the reference ships), NOT from oracle/gpc_oracle.cpp.  Plain numpy
matrix products (BLAS summation order) and libm exp, so agreement with the oracle is to
rounding level only; it checks the algorithm, not the canonical arithmetic."""
import numpy as np


class SogpNumpy:
    def __init__(self, capacity=100, s20=float(np.float32(1e-1)), sigmaf_sq=100.0, l_sq=1.0,
                 eps_tol=float(np.float32(1e-6))):
        self.capacity, self.s20, self.p0, self.p1, self.eps_tol = capacity, s20, sigmaf_sq, l_sq, eps_tol
        self.n = 0
        self.alpha = np.zeros(0)
        self.C = np.zeros((0, 0))
        self.Q = np.zeros((0, 0))
        self.BV = np.zeros((2, 0))
        self.idx = []
        self.events = []

    def kern(self, x, B):
        d = B - x.reshape(2, 1)
        return self.p0 * np.exp(-0.5 / self.p1 * (d * d).sum(axis=0))

    def add(self, x, y, orig):
        x = np.asarray(x, float)
        kstar = self.p0
        if self.n == 0:
            self.alpha = np.array([y / (kstar + self.s20)])
            self.C = np.array([[-1.0 / (kstar + self.s20)]])
            self.Q = np.array([[1.0 / kstar]])
            self.BV = x.reshape(2, 1).copy()
            self.idx = [orig]
            self.n = 1
            self.events.append("first")
            return
        k = self.kern(x, self.BV)
        m = self.alpha @ k
        s2 = kstar + k @ self.C @ k
        r = -1.0 / (self.s20 + s2)
        q = -r * (y - m)
        e_hat = self.Q @ k
        gamma = kstar - k @ e_hat
        if gamma < float(np.float32(1e-12)):
            gamma = 0.0
        if gamma < self.eps_tol and self.capacity != -1:
            eta = 1.0 / (1.0 + gamma * r)
            s_hat = self.C @ k + e_hat
            self.alpha = self.alpha + s_hat * (q * eta)
            self.C = self.C + r * eta * np.outer(s_hat, s_hat)
            self.events.append("sparse")
        else:
            s = np.concatenate([self.C @ k, [1.0]])
            self.alpha = np.concatenate([self.alpha, [0.0]]) + q * s
            n = self.n
            Cn = np.zeros((n + 1, n + 1))
            Cn[:n, :n] = self.C
            self.C = Cn + r * np.outer(s, s)
            self.BV = np.concatenate([self.BV, x.reshape(2, 1)], axis=1)
            self.idx.append(orig)
            self.n += 1
            Qn = np.zeros((n + 1, n + 1))
            Qn[:n, :n] = self.Q
            e2 = np.concatenate([e_hat, [-1.0]])
            with np.errstate(divide="ignore", invalid="ignore"):
                self.Q = Qn + 1.0 / gamma * np.outer(e2, e2)
            self.events.append("full")
        while self.n > self.capacity and self.capacity > 0:
            sc = self.alpha ** 2 / (np.diag(self.Q) + np.diag(self.C))
            self.delete_bv(int(np.argmin(sc)))
            self.events.append("delcap")
        minscore = 0.0
        geo = float(np.float32(1e-9))
        while minscore < geo and self.n > 1:
            sc = 1.0 / np.diag(self.Q)
            loc = int(np.argmin(sc))
            minscore = sc[loc]
            if minscore < geo:
                self.delete_bv(loc)
                self.events.append("delgeo")

    def delete_bv(self, loc):
        n = self.n
        last = n - 1
        keep = [i for i in range(n) if i != loc]
        # move the last BV into slot loc (sogp.m / sparse_gp.hpp:252-281), as a permutation
        perm = list(range(n))
        perm[loc], perm[last] = perm[last], perm[loc]
        a = self.alpha[perm]
        Cp = self.C[np.ix_(perm, perm)]
        Qp = self.Q[np.ix_(perm, perm)]
        astar = a[last]
        cstar = Cp[last, last]
        qstar = Qp[last, last]
        Cs = Cp[:last, last]
        Qs = Qp[:last, last]
        a = a[:last]
        Cp = Cp[:last, :last]
        Qp = Qp[:last, :last]
        a = a - astar / (qstar + cstar) * (Qs + Cs)
        Cp = Cp + np.outer(Qs, Qs) / qstar - np.outer(Qs + Cs, Qs + Cs) / (qstar + cstar)
        Qp = Qp - np.outer(Qs, Qs) / qstar
        self.alpha, self.C, self.Q = a, Cp, Qp
        self.BV = self.BV[:, perm][:, :last]
        self.idx = [self.idx[i] for i in perm][:last]
        self.n = last

    def predict(self, x):
        if self.n == 0:
            return 0.0
        return float(self.alpha @ self.kern(np.asarray(x, float), self.BV))
