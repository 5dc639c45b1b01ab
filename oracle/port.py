"""CPU restatement ("port") of the reference hot path -- TEST INFRASTRUCTURE ONLY.

Every function cites the reference file:line it follows (paths relative to
/root/reference).  Pinned against the imported reference and the Stock fixtures by
tests/test_oracle_*.py and the committed vectors in tests/golden/ (made by
oracle/gen_golden.py).  The RK4 arithmetic comes from the absent third-party
library torchdiffeq==0.0.1 and is restated from its published source (parity of
that part is unpinned, see oracle/torchdiffeq_shim.py).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.
"""
import math

import numpy as np
import torch


# =============================================================================
# a1-a4: natural cubic spline coefficients
# =============================================================================
def natural_cubic_spline_coeffs(times, X):
    """controldiffeq/interpolate.py:159-226 (+ :7-53 no-NaN, :56-153 NaN path, misc.py:12-66 Thomas).

    Vectorised over all (batch, channel) scalar paths; the loop runs over knots.
    A fully observed path goes through the same arithmetic as the reference's
    vectorised branch (the NaN branch with offset 0 is the identity, :143-148).
    times: 1-D float tensor (L,), X: (..., L, C) with NaN = missing.
    Returns (a, b, two_c, three_d), each (..., L-1, C) in X's dtype.
    """
    if not times.is_floating_point() or not X.is_floating_point():
        raise ValueError("t and X must both be floating point/")
    if times.dim() != 1:
        raise ValueError("t must be one dimensional.")
    if X.dim() < 2:
        raise ValueError("X must have at least two dimensions, corresponding to time and channels.")
    if X.size(-2) != times.size(0):
        raise ValueError("The time dimension of X must equal the length of t.")
    if times.size(0) < 2:
        raise ValueError("Must have a time dimension of size at least 2.")
    L, C = X.shape[-2], X.shape[-1]
    xdt = X.detach().cpu().numpy().dtype
    tdt = times.detach().cpu().numpy().dtype
    t = times.detach().cpu().numpy()
    x = np.moveaxis(X.detach().cpu().numpy().reshape(-1, L, C), 1, 2).reshape(-1, L).copy()  # (P, L)
    P = x.shape[0]
    obs = ~np.isnan(x)
    nobs = obs.sum(1)
    any_obs = nobs > 0
    # imputation of the two end points, interpolate.py:104-114
    first_idx = np.where(any_obs, obs.argmax(1), 0)
    last_idx = np.where(any_obs, L - 1 - obs[:, ::-1].argmax(1), 0)
    rows = np.arange(P)
    x[:, 0] = np.where(obs[:, 0], x[:, 0], x[rows, first_idx])
    x[:, L - 1] = np.where(obs[:, L - 1], x[:, L - 1], x[rows, last_idx])
    obs = ~np.isnan(x)
    obs[~any_obs] = False

    # ---- forward sweep of the Thomas algorithm over the observed knots (misc.py:57-61)
    # rows of the system (interpolate.py:31-43): D_j = 2 (r_j + r_{j-1}), rhs_j = pds_j + pds_{j-1}
    # with r = 1/dt (in the dtype of `times`, :24-26), pds = 3 dx r^2.
    Dp = np.zeros((P, L), xdt)      # modified diagonal, stored at the knot position
    bp = np.zeros((P, L), xdt)      # modified rhs
    rr = np.zeros((P, L), xdt)      # r of the piece that STARTS at this observed knot
    nxt = np.full((P, L), -1, np.int64)   # index of the next observed knot
    prev_i = np.full(P, -1, np.int64)     # previous observed knot
    pprev_i = np.full(P, -1, np.int64)
    r_prev = np.zeros(P, xdt)
    pds_prev = np.zeros(P, xdt)
    with np.errstate(all='ignore'):
        for i in range(L):
            o = obs[:, i]
            have_prev = o & (prev_i >= 0)
            pi = np.where(have_prev, prev_i, 0)
            dt = (t[i] - t[pi]).astype(tdt)
            r = (tdt.type(1) / dt).astype(tdt)
            r2 = (r * r).astype(tdt)
            r = r.astype(xdt)
            pds = (xdt.type(3) * (x[:, i] - x[rows, pi])) * r2.astype(xdt)
            # finalise the row of the PREVIOUS observed knot now that its right piece is known
            D = (r + r_prev) * xdt.type(2)
            rhs = pds + pds_prev
            has_pp = have_prev & (pprev_i >= 0)
            ppi = np.where(has_pp, pprev_i, 0)
            w = np.where(has_pp, r_prev / Dp[rows, ppi], 0)
            Dn = np.where(has_pp, D - w * r_prev, D)
            bn = np.where(has_pp, rhs - w * bp[rows, ppi], rhs)
            Dp[rows, pi] = np.where(have_prev, Dn, Dp[rows, pi])
            bp[rows, pi] = np.where(have_prev, bn, bp[rows, pi])
            rr[rows, pi] = np.where(have_prev, r, rr[rows, pi])
            nxt[rows, pi] = np.where(have_prev, i, nxt[rows, pi])
            pprev_i = np.where(have_prev, prev_i, pprev_i)
            r_prev = np.where(have_prev, r, r_prev)
            pds_prev = np.where(have_prev, pds, pds_prev)
            prev_i = np.where(o, i, prev_i)
        # last row: D = 2 r_{m-2}, rhs = pds_{m-2}
        li = np.where(any_obs, prev_i, 0)
        D = (xdt.type(0) + r_prev) * xdt.type(2)
        rhs = pds_prev
        has_pp = any_obs & (pprev_i >= 0)
        ppi = np.where(has_pp, pprev_i, 0)
        w = np.where(has_pp, r_prev / Dp[rows, ppi], 0)
        Dp[rows, li] = np.where(has_pp, D - w * r_prev, D)
        bp[rows, li] = np.where(has_pp, rhs - w * bp[rows, ppi], rhs)
        # ---- back substitution (misc.py:63-65): k_j = (b'_j - U_j k_{j+1}) / D'_j
        kd = np.zeros((P, L), xdt)
        kd[rows, li] = bp[rows, li] / Dp[rows, li]
        for i in range(L - 2, -1, -1):
            o = obs[:, i] & (nxt[:, i] >= 0)
            ni = np.where(o, nxt[:, i], 0)
            val = (bp[:, i] - rr[:, i] * kd[rows, ni]) / Dp[:, i]
            kd[:, i] = np.where(o, val, kd[:, i])
        # ---- coefficients of each observed piece (interpolate.py:46-53) expanded to
        # every interval with the offset rule (:136-148)
        a = np.zeros((P, L - 1), xdt)
        b = np.zeros((P, L - 1), xdt)
        c2 = np.zeros((P, L - 1), xdt)
        d3 = np.zeros((P, L - 1), xdt)
        cur = np.zeros(P, np.int64)
        two_knots = obs.sum(1) == 2
        for i in range(L - 1):
            cur = np.where(obs[:, i], i, cur)
            ni = np.where(any_obs, nxt[rows, cur], 0)
            r = rr[rows, cur]
            r2 = ((r.astype(tdt) * r.astype(tdt)).astype(tdt)).astype(xdt)
            x0, x1 = x[rows, cur], x[rows, ni]
            k0, k1 = kd[rows, cur], kd[rows, ni]
            three_pd = xdt.type(3) * (x1 - x0)
            six_pd = xdt.type(2) * three_pd
            # a path with exactly two (imputed) knots is one linear piece (interpolate.py:17-21)
            pa = x0
            pb = np.where(two_knots, (x1 - x0) / (t[ni] - t[cur]).astype(tdt).astype(xdt), k0)
            pc = np.where(two_knots, 0, (six_pd * r - xdt.type(4) * k0 - xdt.type(2) * k1) * r)
            pd = np.where(two_knots, 0, (-six_pd * r + xdt.type(3) * (k0 + k1)) * r2)
            off = (t[cur] - t[i]).astype(tdt).astype(xdt)
            a_inner = (xdt.type(0.5) * pc - pd * off / xdt.type(3)) * off
            a[:, i] = pa + (a_inner - pb) * off
            b[:, i] = pb + (pd * off - pc) * off
            c2[:, i] = pc - xdt.type(2) * pd * off
            d3[:, i] = pd
    for arr in (a, b, c2, d3):
        arr[~any_obs] = 0   # interpolate.py:93-101: all-NaN path -> zero coefficients

    def back(arr):
        arr = arr.reshape(-1, C, L - 1)
        arr = np.moveaxis(arr, 1, 2).reshape(*X.shape[:-2], L - 1, C)
        return torch.from_numpy(np.ascontiguousarray(arr))
    return back(a), back(b), back(c2), back(d3)


# =============================================================================
# a5: spline evaluation
# =============================================================================
def _interpret_t(times, t, n_pieces):
    """controldiffeq/interpolate.py:261-267 (left piece at knots)."""
    index = int((t > times).sum()) - 1
    index = min(max(index, 0), n_pieces - 1)
    frac = t - times[index]
    return frac, index


def spline_evaluate(times, coeffs, t):
    """controldiffeq/interpolate.py:269-275."""
    a, b, two_c, three_d = coeffs
    frac, idx = _interpret_t(times, t, b.size(-2))
    inner = 0.5 * two_c[..., idx, :] + three_d[..., idx, :] * frac / 3
    inner = b[..., idx, :] + inner * frac
    return a[..., idx, :] + inner * frac


def spline_derivative(times, coeffs, t):
    """controldiffeq/interpolate.py:298-303."""
    a, b, two_c, three_d = coeffs
    frac, idx = _interpret_t(times, t, b.size(-2))
    inner = two_c[..., idx, :] + three_d[..., idx, :] * frac
    return b[..., idx, :] + inner * frac


# =============================================================================
# a7 / a11: the two vector-field MLPs, functional over a state_dict
# =============================================================================
def _lin(sd, name, x):
    return torch.nn.functional.linear(x, sd[name + '.weight'], sd[name + '.bias'])


def field_f(sd, z, C, prefix='func_f'):
    """experiments/models/vector_fields.py:205-218 (FinalTanh_ff6)."""
    z = _lin(sd, prefix + '.linear_in', z).relu()
    z = _lin(sd, prefix + '.linear_test', z).relu()
    i = 0
    while (prefix + '.linears.%d.weight' % i) in sd:
        z = _lin(sd, prefix + '.linears.%d' % i, z).relu()
        i += 1
    z = _lin(sd, prefix + '.linear_out', z)
    return z.view(*z.shape[:-1], C, C).tanh()


def field_g(sd, z, H, C, prefix='func_g'):
    """experiments/models/vector_fields.py:237-250 (FinalTanh)."""
    z = _lin(sd, prefix + '.linear_in', z).relu()
    i = 0
    while (prefix + '.linears.%d.weight' % i) in sd:
        z = _lin(sd, prefix + '.linears.%d' % i, z).relu()
        i += 1
    z = _lin(sd, prefix + '.linear_out', z)
    return z.view(*z.shape[:-1], H, C).tanh()


# =============================================================================
# a14: fixed-grid RK4 (3/8 rule) of torchdiffeq==0.0.1, restated
# =============================================================================
def fixed_grid(t, step_size):
    """torchdiffeq _grid_constructor_from_step_size: t0 + k*step, last point clipped."""
    niters = int(math.ceil(float((t[-1] - t[0]) / step_size + 1)))
    grid = torch.arange(0, niters).to(t) * step_size + t[0]
    if grid[-1] > t[-1]:
        grid[-1] = t[-1]
    return grid


def rk4_38(func, y0, t, step_size, grid=None):
    """torchdiffeq FixedGridODESolver.integrate + rk4_alt_step_func.

    func(t, y, n) -> dy, where n is the running call counter (0-based).  The grid is t[0] + k * step_size; with
    step_size None it is `grid` (options['grid_constructor']) or, with neither, `t` itself (FixedGridODESolver.__init__:
    grid_constructor = lambda f, y0, t: t)."""
    if step_size is not None:
        grid = fixed_grid(t, step_size)
    elif grid is None:
        grid = t
    assert grid[0] == t[0] and grid[-1] == t[-1]
    sol = [y0]
    j = 1
    y = y0
    n = 0
    for t0, t1 in zip(grid[:-1], grid[1:]):
        dt = t1 - t0
        k1 = func(t0, y, n)
        k2 = func(t0 + dt / 3, y + dt * k1 / 3, n + 1)
        k3 = func(t0 + dt * 2 / 3, y + dt * (k2 - k1 / 3), n + 2)
        k4 = func(t0 + dt, y + dt * (k1 - k2 + k3), n + 3)
        n += 4
        y1 = y + (k1 + 3 * k2 + 3 * k3 + k4) * (dt / 8)
        while j < len(t) and t1 >= t[j]:
            if t[j] == t0:
                sol.append(y)
            elif t[j] == t1:
                sol.append(y1)
            else:
                sol.append(y + (y1 - y) / (t1 - t0) * (t[j] - t0))
            j += 1
        y = y1
    return torch.stack(sol)


# =============================================================================
# plain NCDE: controldiffeq.cdeint with FinalTanh (a6, VectorField)
# =============================================================================
def cdeint_port(sd, times, coeffs, z0, t, step_size, H, C, prefix='func', grid=None):
    """controldiffeq/cdeint_module.py:25-32,141-151 with func = FinalTanh."""
    cf = tuple(c.to(z0.dtype) for c in coeffs)

    def field(tt, z, n):
        dX = spline_derivative(times, cf, tt)
        return (field_g(sd, z, H, C, prefix) @ dX.unsqueeze(-1)).squeeze(-1)

    return rk4_38(field, z0, t, step_size, grid)


# =============================================================================
# the dual NCDE (a6-a13)
# =============================================================================
def attention_head(sd, raw, cfg, slope):
    """experiments/models/metamodel.py:327-343 (ANCDE) / :650-666 (ANCDE_forecasting)."""
    att = raw
    if cfg['timewise']:
        att = _lin(sd, 'time_attention', att)
    if cfg['soft']:
        return torch.sigmoid(att)
    if cfg['slope_check']:
        hs = (torch.nn.functional.hardtanh(slope * att) + 1.0) / 2.0      # Hardsigmoid, :139-146
    else:
        hs = torch.sigmoid(att)
    return hs + (torch.round(hs) - hs).detach()                          # RoundFunctionST, :148-161


def dual_ncde_forward(sd, cfg, times, coeffs, final_index, slope, stream=False, z0=None,
                      adjoint=False, step_size=None):
    """Restates ANCDE.forward (metamodel.py:238-374) / ANCDE_forecasting.forward (:565-698)
    on top of ancde_bottom (cdeint_module.py:168-243) and ancde (:247-254), rk4 only.

    sd: state_dict-like mapping with the reference's parameter names.
    cfg: dict(kind, C, H, timewise, soft, slope_check, output_time).
    adjoint=True reproduces the reference's gradient FLOW in adjoint mode (SURVEY Q10:
    attention and h' are constants of the top solve) on the discrete solver.
    Returns dict(pred, bottom (L,B,C), h_prime (4S,B,C), attention, top (len(t),B,H)).
    """
    C, H = cfg['C'], cfg['H']
    cf = coeffs    # evaluated in their own dtype (float64 for Stock), cast per call like the reference (Q11)
    L = times.numel()
    if step_size is None:
        step_size = float((times[1:] - times[:-1]).min())                 # metamodel.py:311-313
    x0 = spline_evaluate(times, cf, times[0]).float()                       # metamodel.py:600,668
    if z0 is None:
        z0 = _lin(sd, 'initial_network', x0)                               # :276-277 / :600-602
    # ---- which output times the top solve is asked for (:287-302)
    if stream:
        t = times
        fi = None
    else:
        sorted_fi, inverse = final_index.unique(sorted=True, return_inverse=True)
        if 0 in sorted_fi:
            sorted_fi = sorted_fi[1:]
            fi = inverse
        else:
            fi = inverse + 1
        if len(times) - 1 in sorted_fi:
            sorted_fi = sorted_fi[:-1]
        t = torch.cat([times[0].unsqueeze(0), times[sorted_fi], times[-1].unsqueeze(0)])

    # ---- bottom NCDE over ALL knots (Q9), collecting h' in call order (Q5)
    hp = []

    def bottom_field(tt, z, n):
        dX = spline_derivative(times, cf, tt).float()                      # cdeint_module.py:58-60
        out = (field_f(sd, z, C) @ dX.unsqueeze(-1)).squeeze(-1)           # :61
        hp.append(out.detach())
        return out

    bottom = rk4_38(bottom_field, z0, times, 1.0 if step_size is None else step_size)
    h_prime = torch.stack(hp)                                              # (4S, B, C), constant
    attention = attention_head(sd, bottom, cfg, slope)
    y0 = _lin(sd, 'feature_extractor', x0 * attention[0])                  # :345-348
    att_solve = attention.detach() if adjoint else attention
    n_hp = h_prime.shape[0]

    def top_field(tt, z, n):
        G = field_g(sd, z, H, C)
        Xt = spline_derivative(times, cf, tt)                              # :112, NOT cast (Q4, Q11)
        dX = Xt.float()                                                    # :106
        q = int(math.floor(float(tt))) - 1                                 # cdeint_module.py:108-110 (Q3)
        a_t = att_solve[q]
        m = n % (n_hp - 1)                                                 # :134-137 (Q5), counter + reset rule
        dY = (dX * a_t + (a_t * (1 - a_t) * Xt) * h_prime[m]).float()      # :113-129
        return (G @ dY.unsqueeze(-1)).squeeze(-1)                          # :133

    top = rk4_38(top_field, y0, t, step_size)
    if stream:
        z_t = top.transpose(0, 1)                                          # (B, L, H)
        if cfg['kind'] == 'ancde_forecasting':
            pred = _lin(sd, 'linear', z_t[:, z_t.shape[1] - cfg['output_time']:, :])   # :694-696
        else:
            pred = _lin(sd, 'linear', z_t)
    else:
        idx = fi.unsqueeze(-1).expand(top.shape[1:]).unsqueeze(0)
        z_t = top.gather(dim=0, index=idx).squeeze(0)                      # :364-368
        pred = _lin(sd, 'linear', z_t)
    return dict(pred=pred, bottom=bottom, h_prime=h_prime, attention=attention, top=top, z_t=z_t, y0=y0, z0=z0)


def ivn_z0(sd, static):
    """experiments/sepsis_attentive.py:24-30 InitialValueNetwork (z0 = MLP(static))."""
    return _lin(sd, 'linear2', _lin(sd, 'linear1', static).relu())
