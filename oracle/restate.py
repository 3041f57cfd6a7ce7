"""oracle/restate.py -- CPU restatement (numpy, fp64) of the reference's Monte-Carlo hot path.

TEST INFRASTRUCTURE ONLY.  Imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline
leg as the CHECKER.  The product (compfinance_b200/) never imports this module; it has no CPU path.

Each function cites the reference file:line (relative to asavine/CompFinance) whose algorithm it
restates.  Paths are vectorised over numpy arrays, time steps are a Python loop.

Pinning: tests/test_oracle.py checks this module against (a) the known answers recorded in
SURVEY.md section 8c (direction numbers, Sobol states, mrg32k3a uniforms, invNormalCdf values,
config-1 and config-3 prices and risks) and (b) the reference itself compiled by
oracle/build_ref.py (oracle/_ref/libcfref.so), which in turn reproduces the reference's shipped
spreadsheet goldens (AutocallPricer.xlsx / testDLM.xlsx).
"""
import math
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
JK_FILE = os.path.join(_HERE, "..", "compfinance_b200", "data", "joe_kuo_old_1111.txt")

EPS = 1.0e-08                       # gaussians.h:8, utility.h:9
HALF_DAY = 0.00136986301369863      # mcMdlDupire.h:26
ONE_HOUR = 0.000114469              # mcPrd.h:26
ONEOVER2POW32 = 2.3283064365387E-10  # sobol.h:26 (NOT 2**-32)

# ------------------------------------------------------------------------------------------------
# Sobol  (sobol.h:31-151, direction numbers sobol.cpp:16-3672 = Joe-Kuo "old 1111" set)
# ------------------------------------------------------------------------------------------------
_dir_cache = None


def sobol_direction_numbers():
    """[32][1101] uint32, rebuilt from the Joe-Kuo initialisers with the published recurrence."""
    global _dir_cache
    if _dir_cache is not None:
        return _dir_cache
    rows = []
    with open(JK_FILE) as fh:
        next(fh)
        for line in fh:
            t = line.split()
            rows.append((int(t[1]), int(t[2]), [int(x) for x in t[3:]]))
    out = np.zeros((32, len(rows)), dtype=np.uint32)
    for d, (s, a, m0) in enumerate(rows):
        m = [0] * 33
        if s == 0:
            for i in range(1, 33):
                m[i] = 1
        else:
            for i in range(1, s + 1):
                m[i] = m0[i - 1]
            for i in range(s + 1, 33):
                x = m[i - s] ^ (m[i - s] << s)
                for k in range(1, s):
                    if (a >> (s - 1 - k)) & 1:
                        x ^= m[i - k] << k
                m[i] = x
        for i in range(1, 33):
            out[i - 1, d] = (m[i] << (32 - i)) & 0xFFFFFFFF
    _dir_cache = out
    return out


def sobol_states(dim, first, n):
    """Integer states of paths first..first+n-1: [n][dim] uint32.

    Sobol::next (sobol.h:77-101) XORs jkDir[ctz(~index)] into the state; after k calls the state
    is the XOR of jkDir[b] over the set bits b of Gray(k) = k ^ (k >> 1).  skipTo (sobol.h:119-150)
    builds the same value directly.  Path p (0-based) is the state after p + 1 calls."""
    dirs = sobol_direction_numbers()[:, :dim]
    idx = np.arange(first + 1, first + n + 1, dtype=np.uint64)
    gray = idx ^ (idx >> np.uint64(1))
    out = np.zeros((n, dim), dtype=np.uint32)
    for b in range(32):
        sel = ((gray >> np.uint64(b)) & np.uint64(1)).astype(bool)
        if sel.any():
            out[sel] ^= dirs[b][None, :]
    return out


def sobol_sequential(dim, n):
    """Literal restatement of Sobol::next (sobol.h:77-101), for cross-checking sobol_states."""
    dirs = sobol_direction_numbers()[:, :dim]
    state = np.zeros(dim, dtype=np.uint32)
    out = np.zeros((n, dim), dtype=np.uint32)
    for index in range(n):
        k, j = index, 0
        while k & 1:
            k >>= 1
            j += 1
        state = state ^ dirs[j]
        out[index] = state
    return out


def sobol_uniforms(dim, first, n):
    return ONEOVER2POW32 * sobol_states(dim, first, n).astype(np.float64)   # sobol.h:99-100


# ------------------------------------------------------------------------------------------------
# mrg32k3a  (mrg32k3a.h:23-394)
# ------------------------------------------------------------------------------------------------
M1, M2 = 4294967087, 4294944443
A12, A13, A21, A23 = 1403580, 810728, 527612, 1370589
M1P1 = 4294967088.0


def _matmul(a, b, mod):
    return [[sum(a[i][l] * b[l][k] for l in range(3)) % mod for k in range(3)] for i in range(3)]


def _matpow(a, e, mod):
    r = [[1, 0, 0], [0, 1, 0], [0, 0, 1]]
    while e:
        if e & 1:
            r = _matmul(r, a, mod)
        a = _matmul(a, a, mod)
        e >>= 1
    return r


def mrg32k3a_numerators(seed1, seed2, dim, first, n):
    """[n][dim] integer numerators z with u = z / (m1 + 1); odd paths repeat their even partner
    (antithetic cache, mrg32k3a.h:107-186).  nextNumber: mrg32k3a.h:55-81; skip-ahead by matrix
    powers: mrg32k3a.h:306-394 (stream offset of pair q is q * dim numbers, mrg32k3a.h:198-212)."""
    q0 = first // 2
    q1 = (first + n - 1) // 2
    A = [[0, A12, M1 - A13], [1, 0, 0], [0, 1, 0]]
    B = [[A21, 0, M2 - A23], [1, 0, 0], [0, 1, 0]]
    skip = q0 * dim
    Ab, Bb = _matpow(A, skip, M1), _matpow(B, skip, M2)
    x = [sum(Ab[i][l] * seed1 for l in range(3)) % M1 for i in range(3)]
    y = [sum(Bb[i][l] * seed2 for l in range(3)) % M2 for i in range(3)]
    pairs = np.zeros((q1 - q0 + 1, dim), dtype=np.uint32)
    for q in range(q1 - q0 + 1):
        for d in range(dim):
            xn = (A12 * x[1] - A13 * x[2]) % M1
            x = [xn, x[0], x[1]]
            yn = (A21 * y[0] - A23 * y[2]) % M2
            y = [yn, y[0], y[1]]
            pairs[q, d] = xn - yn if xn > yn else xn - yn + M1
    idx = (np.arange(first, first + n) // 2) - q0
    return pairs[idx]


def mrg32k3a_uniforms(seed1, seed2, dim, first, n):
    z = mrg32k3a_numerators(seed1, seed2, dim, first, n).astype(np.float64)
    u = z / M1P1
    odd = (np.arange(first, first + n) & 1).astype(bool)
    u[odd] = 1.0 - u[odd]                                   # mrg32k3a.h:112-117
    return u


# ------------------------------------------------------------------------------------------------
# invNormalCdf  (gaussians.h:47-87, Beasley-Springer-Moro)
# ------------------------------------------------------------------------------------------------
def inv_normal_cdf(p):
    p = np.asarray(p, dtype=np.float64)
    sup = p > 0.5
    up = np.where(sup, 1.0 - p, p)
    x = up - 0.5
    r = x * x
    num = ((-25.44106049637 * r + 41.39119773534) * r + -18.61500062529) * r + 2.50662823884
    den = (((3.13082909833 * r + -21.06224101826) * r + 23.08336743743) * r + -8.47351093090) * r + 1.0
    central = x * num / den
    with np.errstate(divide="ignore", invalid="ignore"):
        t = np.log(-np.log(up))
    c = [0.3374754822726147, 0.9761690190917186, 0.1607979714918209, 0.0276438810333863, 0.0038405729373609,
         0.0003951896511919, 0.0000321767881768, 0.0000002888167364, 0.0000003960315187]
    tail = c[8]
    for k in range(7, -1, -1):
        tail = c[k] + t * tail
    is_central = np.abs(x) < 0.42
    return np.where(is_central, np.where(sup, -central, central), np.where(sup, tail, -tail))


def gaussians(rng, dim, first, n):
    """RNG::nextG for paths first..first+n-1.  rng = ("sobol",) or ("mrg32k3a", seed1, seed2)."""
    if rng[0] == "sobol":
        return inv_normal_cdf(sobol_uniforms(dim, first, n))          # sobol.h:103-109
    z = mrg32k3a_numerators(rng[1], rng[2], dim, first, n).astype(np.float64)
    g = inv_normal_cdf(z / M1P1)                                       # mrg32k3a.h:166-186
    odd = (np.arange(first, first + n) & 1).astype(bool)
    g[odd] = -g[odd]
    return g


# ------------------------------------------------------------------------------------------------
# Host utilities: fillData (utility.h:11-79), interp (interp.h:26-63)
# ------------------------------------------------------------------------------------------------
def fill_data(original, max_dx, min_dx=0.0, add=None):
    """utility.h:24-79.  set_union with the tolerance comparator x < y - minDx (utility.h:38-46)."""
    original = list(original)
    if add:
        seq, i, j = [], 0, 0
        a, b = original, list(add)
        less = lambda x, y: x < y - min_dx   # noqa: E731
        while i < len(a) and j < len(b):
            if less(b[j], a[i]):
                seq.append(b[j]); j += 1
            elif less(a[i], b[j]):
                seq.append(a[i]); i += 1
            else:
                seq.append(a[i]); i += 1; j += 1
        seq += a[i:] + b[j:]
    else:
        seq = original
    filled = [seq[0]]
    for nxt in seq[1:]:
        cur = filled[-1]
        if nxt - cur > max_dx:
            add_points = int((nxt - cur) / max_dx - EPS) + 1
            spacing = (nxt - cur) / add_points
            t = cur + spacing
            while t < nxt - min_dx:
                filled.append(t)
                t += spacing
        filled.append(nxt)
    return filled


def interp1(xs, ys, x0):
    """interp.h:26-63: upper_bound, flat extrapolation, linear inside."""
    import bisect
    it = bisect.bisect_right(xs, x0)
    if it == len(xs):
        return ys[-1]
    if it == 0:
        return ys[0]
    n = it - 1
    t = (x0 - xs[n]) / (xs[n + 1] - xs[n])
    return ys[n] + (ys[n + 1] - ys[n]) * t


def interp1_weights(xs, x0):
    """(index, weight) pairs of interp1 as a linear map of ys."""
    import bisect
    it = bisect.bisect_right(xs, x0)
    if it == len(xs):
        return [(len(xs) - 1, 1.0)]
    if it == 0:
        return [(0, 1.0)]
    n = it - 1
    t = (x0 - xs[n]) / (xs[n + 1] - xs[n])
    return [(n, 1.0 - t), (n + 1, t)]


# ------------------------------------------------------------------------------------------------
# Products (host side): timelines
# ------------------------------------------------------------------------------------------------
def uoc_timeline(maturity, monitor_freq, system_time=0.0):
    """UOC constructor, mcPrd.h:165-176: floating-point accumulation t += freq."""
    tl = [system_time]
    t = system_time + monitor_freq
    while maturity - t > ONE_HOUR:
        tl.append(t)
        t += monitor_freq
    tl.append(maturity)
    return tl


# ------------------------------------------------------------------------------------------------
# Dupire  (mcMdlDupire.h:28-281)
# ------------------------------------------------------------------------------------------------
class DupireTables:
    """allocate (mcMdlDupire.h:165-193) + init (mcMdlDupire.h:195-217)."""

    def __init__(self, spot, spots, times, vols, max_dt, product_timeline, system_time=0.0):
        self.spot = float(spot)
        self.spots = [float(s) for s in spots]
        self.times = [float(t) for t in times]
        self.vols = np.asarray(vols, dtype=np.float64)
        self.log_spots = np.array([math.log(s) for s in self.spots])
        self.timeline = fill_data(product_timeline, max_dt, HALF_DAY, [system_time])
        pt = list(product_timeline)
        import bisect
        self.common = []
        for t in self.timeline:
            i = bisect.bisect_left(pt, t)
            self.common.append(i < len(pt) and pt[i] == t)     # binary_search, mcMdlDupire.h:181-184
        n = len(self.timeline) - 1
        self.n_steps = n
        self.sqrt_dt = np.array([math.sqrt(self.timeline[i + 1] - self.timeline[i]) for i in range(n)])
        self.interp_vols = np.zeros((n, len(self.spots)))
        for i in range(n):
            for j in range(len(self.spots)):
                self.interp_vols[i, j] = self.sqrt_dt[i] * interp1(self.times, list(self.vols[j]), self.timeline[i])

    def time_map(self):
        """init() (mcMdlDupire.h:202-216) as a linear map per step: interp_vols[i][j] =
        w1[i] * vols[j][col1[i]] + w2[i] * vols[j][col2[i]].  Returns (n_times, col1, col2, w1, w2)."""
        c1, c2, w1, w2 = [], [], [], []
        for i in range(self.n_steps):
            ws = interp1_weights(self.times, self.timeline[i])
            c1.append(ws[0][0]); w1.append(self.sqrt_dt[i] * ws[0][1])
            if len(ws) > 1:
                c2.append(ws[1][0]); w2.append(self.sqrt_dt[i] * ws[1][1])
            else:
                c2.append(ws[0][0]); w2.append(0.0)
        return len(self.times), np.array(c1, dtype=np.int32), np.array(c2, dtype=np.int32), np.array(w1), np.array(w2)

    def param_risks(self, spot_adj, ybar, n_paths):
        """propagateMarkToStart for init() (mcMdlDupire.h:202-216): interp_vols adjoints -> vols
        adjoints; then risks = adjoint / nPath (mcBase.h:745)."""
        vbar = np.zeros_like(self.vols)
        for i in range(self.n_steps):
            for (k, w) in interp1_weights(self.times, self.timeline[i]):
                vbar[:, k] += self.sqrt_dt[i] * w * ybar[i, :]
        return spot_adj / n_paths, vbar / n_paths


def _interp_rows(xs, y_row, x0):
    """Vectorised interp of one table row at many x0: value, bucket n, weight t, slope."""
    m = len(xs)
    ub = np.searchsorted(xs, x0, side="right")          # upper_bound
    hi = ub == m
    lo = ub == 0
    n = np.clip(ub - 1, 0, max(m - 2, 0))
    if m > 1:
        x1, x2 = xs[n], xs[n + 1]
        y1, y2 = y_row[n], y_row[n + 1]
        t = (x0 - x1) / (x2 - x1)
        v = y1 + (y2 - y1) * t
        slope = (y2 - y1) / (x2 - x1)
    else:
        t = np.zeros_like(x0)
        v = np.full_like(x0, y_row[0])
        slope = np.zeros_like(x0)
    v = np.where(hi, y_row[m - 1], np.where(lo, y_row[0], v))
    t = np.where(hi, 1.0 if m > 1 else 0.0, np.where(lo, 0.0, t))
    slope = np.where(hi | lo, 0.0, slope)
    return v, n, t, slope


def dupire_generate_paths(tab, g):
    """Dupire::generatePath (mcMdlDupire.h:238-280).  g: [n][n_steps].  Returns log-spot history
    L [n][n_steps+1] and the sampled spots on event dates S [n][n_events]."""
    n = g.shape[0]
    L = np.empty((n, tab.n_steps + 1))
    L[:, 0] = math.log(tab.spot)
    S = []
    if tab.common[0]:
        S.append(np.exp(L[:, 0]))
    for i in range(tab.n_steps):
        v, _, _, _ = _interp_rows(tab.log_spots, tab.interp_vols[i], L[:, i])
        L[:, i + 1] = L[:, i] + v * (-0.5 * v + g[:, i])
        if tab.common[i + 1]:
            S.append(np.exp(L[:, i + 1]))
    return L, np.stack(S, axis=1)


# ------------------------------------------------------------------------------------------------
# Black-Scholes  (mcMdlBS.h:25-350), risk-neutral measure
# ------------------------------------------------------------------------------------------------
class BSTables:
    """allocate (mcMdlBS.h:149-195) + init (mcMdlBS.h:197-277), first forward/discount per event."""

    def __init__(self, spot, vol, rate, div, product_timeline, fwd_mats, disc_mats, numeraire_flags,
                 system_time=0.0):
        self.spot, self.vol, self.rate, self.div = float(spot), float(vol), float(rate), float(div)
        pt = list(product_timeline)
        self.product_timeline = pt
        self.timeline = [system_time] + [t for t in pt if t > system_time]
        self.today_on_timeline = pt[0] == system_time
        n = len(self.timeline) - 1
        self.n_steps = n
        mu = self.rate - self.div
        self.dt = np.array([self.timeline[i + 1] - self.timeline[i] for i in range(n)])
        self.stds = self.vol * np.sqrt(self.dt)
        self.drifts = (mu - 0.5 * self.vol * self.vol) * self.dt
        E = len(pt)
        self.fwd_mats, self.disc_mats, self.num_flags = fwd_mats, disc_mats, numeraire_flags
        self.numeraires = np.array([math.exp(self.rate * pt[e]) if numeraire_flags[e] else 1.0 for e in range(E)])
        self.discounts = np.array([math.exp(-self.rate * (disc_mats[e] - pt[e])) if disc_mats[e] is not None else 1.0
                                   for e in range(E)])
        self.fwd_factors = np.array([math.exp(mu * (fwd_mats[e] - pt[e])) for e in range(E)])
        self.is_event = [self.today_on_timeline] + [True] * n

    def param_risks(self, adj, n_paths):
        """Chain rule of init() (mcMdlBS.h:203-276), SURVEY Appendix A.2.  adj layout:
        [spot, drift[D], std[D], num[E], ff[E], disc[E]] -> risks (spot, vol, rate, div)."""
        D, E = self.n_steps, len(self.product_timeline)
        pt = self.product_timeline
        sb = adj[0]
        db, stb = adj[1:1 + D], adj[1 + D:1 + 2 * D]
        nb, fb, cb = adj[1 + 2 * D:1 + 2 * D + E], adj[1 + 2 * D + E:1 + 2 * D + 2 * E], adj[1 + 2 * D + 2 * E:1 + 2 * D + 3 * E]
        vol_bar = float(np.sum(stb * np.sqrt(self.dt)) - self.vol * np.sum(db * self.dt))
        mu_bar = float(np.sum(db * self.dt))
        rate_bar, div_bar = mu_bar, -mu_bar
        for e in range(E):
            if self.num_flags[e]:
                rate_bar += nb[e] * pt[e] * self.numeraires[e]
            if self.disc_mats[e] is not None:
                rate_bar += -cb[e] * (self.disc_mats[e] - pt[e]) * self.discounts[e]
            tau = self.fwd_mats[e] - pt[e]
            rate_bar += fb[e] * tau * self.fwd_factors[e]
            div_bar += -fb[e] * tau * self.fwd_factors[e]
        return np.array([sb, vol_bar, rate_bar, div_bar]) / n_paths


def bs_generate_paths(tab, g):
    """BlackScholes::generatePath (mcMdlBS.h:321-350): spots S [n][n_steps+1] on the simulation timeline."""
    n = g.shape[0]
    S = np.empty((n, tab.n_steps + 1))
    S[:, 0] = tab.spot
    for i in range(tab.n_steps):
        S[:, i + 1] = S[:, i] * np.exp(tab.drifts[i] + tab.stds[i] * g[:, i])
    return S


# ------------------------------------------------------------------------------------------------
# Payoffs and their adjoints
# ------------------------------------------------------------------------------------------------
def european_payoff(F, strike, disc, num):
    """European::payoffs (mcPrd.h:113-125).  F = forwards[0][0] at the event date."""
    return np.maximum(F - strike, 0.0) * disc / num


def uoc_payoffs(F, strike, barrier, smooth_abs, num_last, is_put=False):
    """UOC::payoffs (mcPrd.h:235-288).  F [n][n_events] = forwards[0][0] on every event date,
    smooth_abs = double(F[0] * smoothFactor) (mcPrd.h:247).  Returns (pay [n][2], alive, killed)."""
    n, E = F.shape
    two, bar_s, min_s = 2 * smooth_abs, barrier + smooth_abs, barrier - smooth_abs
    alive = np.ones(n)
    killed = np.zeros(n, dtype=bool)
    for e in range(E):
        s = F[:, e]
        breach = (~killed) & (s > bar_s)
        fuzzy = (~killed) & (~breach) & (s > min_s)
        killed |= breach
        alive = np.where(breach, 0.0, alive)
        alive = np.where(fuzzy, alive * ((bar_s - s) / two), alive)
    fin = F[:, -1]
    euro = (np.maximum(strike - fin, 0.0) if is_put else np.maximum(fin - strike, 0.0)) / num_last
    return np.stack([alive * euro, euro], axis=1), alive, killed


def uoc_reverse(F, strike, barrier, smooth_abs, num_last, w, is_put=False):
    """Adjoint of sum_k w[k] * payoff[k] w.r.t. F[:, e] and the last numeraire."""
    n, E = F.shape
    pay, alive, killed = uoc_payoffs(F, strike, barrier, smooth_abs, num_last, is_put)
    euro = pay[:, 1]
    two, bar_s, min_s = 2 * smooth_abs, barrier + smooth_abs, barrier - smooth_abs
    Fbar = np.zeros_like(F)
    eurobar = w[0] * alive + w[1]
    fin = F[:, -1]
    x = (strike - fin) if is_put else (fin - strike)
    Fbar[:, -1] = np.where(x > 0.0, (-eurobar if is_put else eurobar) / num_last, 0.0)
    numbar = -eurobar * euro / num_last
    abar = np.where(killed, 0.0, w[0] * euro)
    alive_cur = alive.copy()
    for e in range(E - 1, -1, -1):
        s = F[:, e]
        fuzzy = (~killed) & (s > min_s)
        f = (bar_s - s) / two
        with np.errstate(divide="ignore", invalid="ignore"):
            alive_prev = np.where(f != 0.0, alive_cur / f, 0.0)
        Fbar[:, e] += np.where(fuzzy, abar * alive_prev * (-1.0 / two), 0.0)
        abar = np.where(fuzzy, abar * f, abar)
        alive_cur = np.where(fuzzy, alive_prev, alive_cur)
    return pay, Fbar, numbar


# ------------------------------------------------------------------------------------------------
# Whole-path runs: value and hand adjoints (what the tape computes on this path)
# ------------------------------------------------------------------------------------------------
def dupire_uoc_run(tab, prd, rng, first, n, w=None):
    """mcSimul / mcSimulAAD restated for Dupire x UOC.
    prd = dict(strike, barrier, smooth (factor), is_put).  Returns dict with per-path payoffs, and when
    w is given: agg per path, spot adjoint (sum over paths) and ybar [n_steps][n_knots] (sum over paths)."""
    g = gaussians(rng, tab.n_steps, first, n)
    L, S = dupire_generate_paths(tab, g)
    smooth_abs = float(math.exp(math.log(tab.spot)) * prd["smooth"]) if tab.common[0] else None
    if smooth_abs is None:
        raise ValueError("UOC always has today on its timeline")
    out = {}
    if w is None:
        pay, _, _ = uoc_payoffs(S, prd["strike"], prd["barrier"], smooth_abs, 1.0, prd.get("is_put", False))
        out["payoffs"] = pay
        return out
    pay, Sbar, _ = uoc_reverse(S, prd["strike"], prd["barrier"], smooth_abs, 1.0, w, prd.get("is_put", False))
    out["payoffs"] = pay
    out["agg"] = pay @ np.asarray(w)
    m = len(tab.log_spots)
    ybar = np.zeros((tab.n_steps, m))
    Lbar = np.zeros(n)
    e = S.shape[1] - 1
    for i in range(tab.n_steps - 1, -1, -1):
        if tab.common[i + 1]:
            Lbar = Lbar + Sbar[:, e] * S[:, e]
            e -= 1
        v, nn, t, slope = _interp_rows(tab.log_spots, tab.interp_vols[i], L[:, i])
        vbar = Lbar * (g[:, i] - v)
        np.add.at(ybar[i], nn, vbar * (1.0 - t))
        if m > 1:
            np.add.at(ybar[i], nn + 1, vbar * t)
        Lbar = Lbar + vbar * slope
    if tab.common[0]:
        Lbar = Lbar + Sbar[:, 0] * S[:, 0]
    out["spot_adj"] = float(np.sum(Lbar / tab.spot))
    out["ybar"] = ybar
    return out


def bs_run(tab, prd_kind, prd, rng, first, n, w=None):
    """mcSimul / mcSimulAAD restated for Black-Scholes x {European, UOC}.
    Returns per-path payoffs and, when w is given, the table adjoints
    [spot, drift[D], std[D], num[E], ff[E], disc[E]] summed over paths."""
    g = gaussians(rng, tab.n_steps, first, n)
    S = bs_generate_paths(tab, g)
    D, E = tab.n_steps, len(tab.product_timeline)
    ev = [i for i in range(D + 1) if tab.is_event[i]]          # timeline point of each event
    F = np.stack([S[:, ev[e]] * tab.fwd_factors[e] for e in range(E)], axis=1)
    out = {}
    if prd_kind == "european":
        pay = european_payoff(F[:, 0], prd["strike"], tab.discounts[0], tab.numeraires[0])[:, None]
    else:
        smooth_abs = float(F[0, 0] * prd["smooth"])
        pay, _, _ = uoc_payoffs(F, prd["strike"], prd["barrier"], smooth_abs, tab.numeraires[-1], prd.get("is_put", False))
    out["payoffs"] = pay
    if w is None:
        return out
    Fbar = np.zeros_like(F)
    numbar, discbar = np.zeros((n, E)), np.zeros((n, E))
    if prd_kind == "european":
        intrinsic = np.maximum(F[:, 0] - prd["strike"], 0.0)
        Fbar[:, 0] = np.where(F[:, 0] - prd["strike"] > 0.0, w[0] * tab.discounts[0] / tab.numeraires[0], 0.0)
        discbar[:, 0] = w[0] * intrinsic / tab.numeraires[0]
        numbar[:, 0] = -w[0] * intrinsic * tab.discounts[0] / tab.numeraires[0] / tab.numeraires[0]
    else:
        _, Fbar, nb = uoc_reverse(F, prd["strike"], prd["barrier"], smooth_abs, tab.numeraires[-1], w, prd.get("is_put", False))
        numbar[:, -1] = nb
    out["agg"] = pay @ np.asarray(w[:pay.shape[1]])
    adj = np.zeros(1 + 2 * D + 3 * E)
    Sbar = np.zeros(n)
    e = E - 1
    for i in range(D - 1, -1, -1):
        S1 = S[:, i + 1]
        Sbar = Sbar + Fbar[:, e] * tab.fwd_factors[e]
        adj[1 + 2 * D + E + e] += np.sum(Fbar[:, e] * S1)
        adj[1 + 2 * D + e] += np.sum(numbar[:, e])
        adj[1 + 2 * D + 2 * E + e] += np.sum(discbar[:, e])
        e -= 1
        abar = Sbar * S1
        adj[1 + i] += np.sum(abar)
        adj[1 + D + i] += np.sum(abar * g[:, i])
        Sbar = Sbar * np.exp(tab.drifts[i] + tab.stds[i] * g[:, i])
    if tab.is_event[0]:
        Sbar = Sbar + Fbar[:, 0] * tab.fwd_factors[0]
        adj[1 + 2 * D + E] += np.sum(Fbar[:, 0] * S[:, 0])
        adj[1 + 2 * D] += np.sum(numbar[:, 0])
        adj[1 + 2 * D + 2 * E] += np.sum(discbar[:, 0])
    adj[0] = np.sum(Sbar)
    out["table_adj"] = adj
    return out
