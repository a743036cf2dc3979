"""TEST INFRASTRUCTURE ONLY: an independent, slow restatement of ONE interior-point iteration straight from
the mathematical definition of the clustered low-rank SDP (MPMP.jl:642-657), in mpmath.

It shares no code and no algorithmic shortcut with oracle/clrsdp_ref.cpp: the constraint matrices A_i are
formed densely, the Schur complement is S[p,q] = Tr(A_p X^-1 A_q Y), the search direction comes from one
dense LU solve of the bordered system [S -B; B^T 0], and the step length from mpmath's symmetric
eigenvalue solver. Used to pin the C++ oracle on tiny instances (tests/test_oracle_pin.py).
"""
import mpmath


def _mp(a, i, ctx):
    m, e = a.get_int(i)
    return ctx.ldexp(ctx.mpf(m), e)


class DenseSDP:
    def __init__(self, constraints, b, blockinfo, prec, omega_p=None, omega_d=None, C=None):
        self.ctx = ctx = mpmath.mp.clone()
        ctx.prec = prec
        self.bi = bi = blockinfo
        self.n_y = bi.n_y
        self.b = [_mp(b, i, ctx) for i in range(bi.n_y)]
        self.blocks = []          # (j, l, nb)
        self.A = []               # A[i] = dict (j,l) -> dense nb x nb matrix, i over all (j,r,s,k)
        self.c = []
        self.Brows = []
        for j, con in enumerate(constraints):
            m, K = bi.m[j], bi.n_samples[j]
            for l in range(bi.L[j]):
                self.blocks.append((j, l, bi.Y_blocksizes[j][l]))
            for r in range(m):
                for s in range(r + 1):
                    for k in range(K):
                        mats = {}
                        for l in range(bi.L[j]):
                            dl = bi.delta[j][l]
                            nb = m * dl
                            M = ctx.zeros(nb, nb)
                            rs = bi.rank_sums[j][l]
                            for a in range(rs[k], rs[k + 1]):
                                v = [_mp(con.V[l], a * dl + i, ctx) for i in range(dl)]
                                h = _mp(con.H[l], a, ctx)
                                for i in range(dl):
                                    for i2 in range(dl):
                                        w = h * v[i] * v[i2] / 2
                                        M[r * dl + i, s * dl + i2] += w     # (v v^T) (x) E_rs, E_rs = (e_r e_s^T + e_s e_r^T)/2
                                        M[s * dl + i, r * dl + i2] += w
                            mats[(j, l)] = M
                        self.A.append(mats)
                        idx = (s + r * (r + 1) // 2) * K + k
                        self.c.append(_mp(con.c, idx, ctx))
                        self.Brows.append([_mp(con.B, idx * bi.n_y + q, ctx) for q in range(bi.n_y)])
        self.nx = len(self.A)
        op = ctx.mpf(10) ** 10 if omega_p is None else ctx.mpf(omega_p)
        od = ctx.mpf(10) ** 10 if omega_d is None else ctx.mpf(omega_d)
        self.x = [ctx.mpf(0)] * self.nx
        self.y = [ctx.mpf(0)] * self.n_y
        self.X = {(j, l): ctx.eye(nb) * op for j, l, nb in self.blocks}
        self.Y = {(j, l): ctx.eye(nb) * od for j, l, nb in self.blocks}
        self.ntot = sum(nb for _, _, nb in self.blocks)
        # objective matrix C (MPMP.jl:599): list over j of lists over l of MpArray (nb, nb); None = 0
        self.C = None
        if C is not None:
            self.C = {}
            for j, l, nb in self.blocks:
                flat = C[j][l].reshape(nb * nb)
                self.C[(j, l)] = ctx.matrix([[_mp(flat, r * nb + c, ctx) for c in range(nb)] for r in range(nb)])

    # ---- helpers -------------------------------------------------------------------------------------
    def _tr(self, i, Z):
        ctx = self.ctx
        t = ctx.mpf(0)
        for key, M in self.A[i].items():
            ZM = Z[key]
            n = M.rows
            for a in range(n):
                for bq in range(n):
                    if M[a, bq] != 0:
                        t += M[a, bq] * ZM[bq, a]
        return t

    def _sumA(self, coef):
        out = {(j, l): self.ctx.zeros(nb, nb) for j, l, nb in self.blocks}
        for i, mats in enumerate(self.A):
            for key, M in mats.items():
                out[key] += M * coef[i]
        return out

    def dot(self, P, Q):
        t = self.ctx.mpf(0)
        for key in P:
            n = P[key].rows
            for a in range(n):
                for bq in range(n):
                    t += P[key][a, bq] * Q[key][a, bq]
        return t

    def iterate(self, beta_inf=None, beta_feas=None, gamma=None, pd_feas=False):
        """one predictor-corrector step (MPMP.jl:754-887); returns a dict of everything computed"""
        ctx = self.ctx
        beta_inf = ctx.mpf(3) / 10 if beta_inf is None else beta_inf
        beta_feas = ctx.mpf(1) / 10 if beta_feas is None else beta_feas
        gamma = ctx.mpf(7) / 10 if gamma is None else gamma
        keys = [(j, l) for j, l, _ in self.blocks]
        X, Y, x, y = self.X, self.Y, self.x, self.y
        mu = self.dot(X, Y) / self.ntot
        Xinv = {k: ctx.inverse(X[k]) for k in keys}
        # residuals
        P = self._sumA(x)
        for k in keys:
            P[k] -= X[k]
            if self.C is not None:
                P[k] -= self.C[k]          # P = sum x_i A_i - X - C (MPMP.jl:1108-1118)
        d = [self.c[i] - sum(self.Brows[i][q] * y[q] for q in range(self.n_y)) - self._tr(i, Y) for i in range(self.nx)]
        p = [self.b[q] - sum(self.Brows[i][q] * x[i] for i in range(self.nx)) for q in range(self.n_y)]
        # Schur complement S[p,q] = Tr(A_p X^-1 A_q Y)
        S = ctx.zeros(self.nx, self.nx)
        for a in range(self.nx):
            for bq in range(self.nx):
                t = ctx.mpf(0)
                for k in self.A[a]:
                    if k in self.A[bq]:
                        prod = self.A[a][k] * Xinv[k] * self.A[bq][k] * Y[k]
                        t += sum(prod[i, i] for i in range(prod.rows))
                S[a, bq] = t
        Bm = ctx.matrix(self.Brows)

        def direction(Rm):
            Z = {k: Xinv[k] * (P[k] * Y[k] - Rm[k]) for k in keys}
            Z = {k: (Z[k] + Z[k].T) / 2 for k in keys}
            rhs_x = [-d[i] - self._tr(i, Z) for i in range(self.nx)]
            n = self.nx + self.n_y
            T = ctx.zeros(n, n)
            for a in range(self.nx):
                for bq in range(self.nx):
                    T[a, bq] = S[a, bq]
                for q in range(self.n_y):
                    T[a, self.nx + q] = -Bm[a, q]
                    T[self.nx + q, a] = Bm[a, q]
            sol = ctx.lu_solve(T, ctx.matrix(rhs_x + p))
            dx = [sol[i] for i in range(self.nx)]
            dy = [sol[self.nx + q] for q in range(self.n_y)]
            dX = self._sumA(dx)
            for k in keys:
                dX[k] += P[k]
            dY = {k: Xinv[k] * (Rm[k] - dX[k] * Y[k]) for k in keys}
            dY = {k: (dY[k] + dY[k].T) / 2 for k in keys}
            return dx, dX, dy, dY, Z

        mu_p = ctx.mpf(0) if pd_feas else beta_inf * mu
        R = {k: ctx.eye(X[k].rows) * mu_p - X[k] * Y[k] for k in keys}
        dxp, dXp, dyp, dYp, _ = direction(R)
        XdX = {k: X[k] + dXp[k] for k in keys}
        YdY = {k: Y[k] + dYp[k] for k in keys}
        r = self.dot(XdX, YdY) / (mu * self.ntot)
        beta = r * r if r < 1 else r
        beta_c = min(max(beta_feas, beta), ctx.mpf(1)) if pd_feas else max(beta_inf, beta)
        mu_c = beta_c * mu
        R = {k: ctx.eye(X[k].rows) * mu_c - X[k] * Y[k] - dXp[k] * dYp[k] for k in keys}
        dx, dX, dy, dY, Z = direction(R)

        def alpha(M, dM):
            lam = None
            for k in keys:
                L = ctx.cholesky(M[k])
                Li = ctx.inverse(L)
                W = Li * dM[k] * Li.T
                W = (W + W.T) / 2
                ev = ctx.eigsy(W, eigvals_only=True)
                m0 = min(ev)
                lam = m0 if lam is None else min(lam, m0)
            return (ctx.mpf(1) if lam > -gamma else -gamma / lam), lam

        ap, lam_x = alpha(X, dX)
        ad, lam_y = alpha(Y, dY)
        if pd_feas:
            ap = ad = min(ap, ad)
        self.x = [x[i] + ap * dx[i] for i in range(self.nx)]
        self.y = [y[q] + ad * dy[q] for q in range(self.n_y)]
        self.X = {k: X[k] + ap * dX[k] for k in keys}
        self.Y = {k: Y[k] + ad * dY[k] for k in keys}
        return dict(mu=mu, S=S, P=P, p=p, d=d, dx_pred=dxp, dy_pred=dyp, dx=dx, dy=dy, dX=dX, dY=dY, Z=Z, alpha_p=ap,
                    alpha_d=ad, beta_c=beta_c, lam_x=lam_x, lam_y=lam_y, Xinv=Xinv)

    def objectives(self):
        """(<c,x>, <C,Y> + <b,y>) of the current point, without b0 (MPMP.jl:1027-1034)"""
        po = sum((ci * xi for ci, xi in zip(self.c, self.x)), self.ctx.mpf(0))
        do = sum((bi * yi for bi, yi in zip(self.b, self.y)), self.ctx.mpf(0))
        if self.C is not None:
            do += self.dot(self.C, self.Y)
        return po, do

    def x_index_of(self, j, r, s, k):
        """position of (j,r,s,k) in self.x / the solver's x vector (same ordering: j, then (r,s) pairs, k fastest)"""
        bi = self.bi
        return bi.x_indices[j] + (s + r * (r + 1) // 2) * bi.n_samples[j] + k
