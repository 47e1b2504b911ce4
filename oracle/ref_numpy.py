"""ORACLE -- TEST INFRASTRUCTURE ONLY.  Not shipped, not measured, not a fallback.

CPU restatement (numpy + scipy, FP64) of the ray-integration hot path of
jdhare/turbulence_tracing.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import this module; the
product package ``turbulence_tracing_b200`` never does (a test enforces it).

Every function names the reference lines it restates (paths relative to the reference
checkout).  The arithmetic of the reference lives in third-party code that is NOT pinned by
the reference (no requirements file): ``scipy.integrate.solve_ivp`` (RK45),
``scipy.interpolate.RegularGridInterpolator``, ``numpy.gradient``, ``numpy.histogram2d``,
``numpy.fft.ifftn``.  The oracle calls the same library routines, as installed in this image
(numpy 2.3.5 / scipy 1.18.1).

Parity pin: ``tests/golden/*.npz`` were produced by importing the *live* reference in the
build container (``tests/golden/make_golden.py``); ``tests/test_oracle_golden.py`` checks this
restatement against every one of them.  Parts of the north-star path that have no
implementation in the reference checkout (phase / Faraday / inverse bremsstrahlung, the
imaging refractometer) are not restated here -- parity unpinned for those.
"""
from __future__ import annotations

import numpy as np
import scipy.constants as _sc
from scipy.integrate import solve_ivp
from scipy.interpolate import RegularGridInterpolator

C_LIGHT = _sc.c                    # particle_tracker.py:119
NC_OVER_OMEGA2 = 3.14207787e-4     # particle_tracker.py:228


# --------------------------------------------------------------------------------------
# density set-ups: particle_tracker.py:147-210
# --------------------------------------------------------------------------------------
def density(kind, x, y, z, **kw):
    """Analytic ``ne`` cubes on ``meshgrid(x, y, z, indexing='ij')``.

    kind: null (:147-152), slab (:154-165), linear_cos (:167-177), exponential_cos
    (:179-188), lens (:190-199), liner (:201-210); keyword defaults as in the reference.
    """
    X, Y, Z = np.meshgrid(x, y, z, indexing="ij")
    ex = x.max()
    if kind == "null":
        return np.zeros_like(X)
    if kind == "slab":
        s, n0 = kw.get("s", 1), kw.get("n_e0", 2e23)
        return n0 * (1.0 + s * X / ex)
    if kind == "linear_cos":
        s1, s2 = kw.get("s1", 0.1), kw.get("s2", 0.1)
        n0, Ly = kw.get("n_e0", 2e23), kw.get("Ly", 1)
        return n0 * (1.0 + s1 * X / ex) * (1 + s2 * np.cos(2 * np.pi * Y / Ly))
    if kind == "exponential_cos":
        n0, Ly, s = kw.get("n_e0", 1e24), kw.get("Ly", 1e-3), kw.get("s", 2e-3)
        return n0 * 10 ** (X / s) * (1 + np.cos(2 * np.pi * Y / Ly))
    if kind == "lens":
        n0, LR = kw.get("n_e0", 1e24), kw.get("LR", 1e-3)
        return n0 * np.exp(-(np.sqrt(X**2 + Y**2)) ** 2 / LR**2)
    if kind == "liner":
        n0, LR = kw.get("n_e0", 1e24), kw.get("LR", 1e-3)
        return n0 * np.exp(-(np.sqrt(X**2 + Z**2)) ** 2 / LR**2)
    raise ValueError(kind)


# --------------------------------------------------------------------------------------
# gradient grid: particle_tracker.py:220-241
# --------------------------------------------------------------------------------------
def critical_density(lwl=1053e-9):
    """omega and nc, particle_tracker.py:227-228."""
    omega = 2 * np.pi * (C_LIGHT / lwl)
    return omega, NC_OVER_OMEGA2 * omega**2


def calc_dndr(ne, x, y, z, lwl=1053e-9, ne_max=1):
    """ne -> (omega, ne_nc, dndx, dndy, dndz); particle_tracker.py:227-237.

    ``np.gradient`` with coordinate arrays: 2nd-order central inside, 1st-order one-sided at
    the faces, non-uniform formula whenever the axis spacing is not bit-constant.
    """
    omega, nc = critical_density(lwl)
    ne_nc = np.array(ne, dtype=np.float64) / nc
    ne_nc[ne_nc > ne_max] = ne_max
    k = -0.5 * C_LIGHT**2
    return dict(
        omega=omega,
        nc=nc,
        ne_nc=ne_nc,
        dndx=k * np.gradient(ne_nc, x, axis=0),
        dndy=k * np.gradient(ne_nc, y, axis=1),
        dndz=k * np.gradient(ne_nc, z, axis=2),
    )


class GradientField:
    """Three trilinear interpolators, zero outside the cube (faces inclusive).

    particle_tracker.py:239-241 (construction) and :243-256 (``dndr``).
    """

    def __init__(self, x, y, z, dndx, dndy, dndz):
        mk = lambda a: RegularGridInterpolator((x, y, z), a, bounds_error=False, fill_value=0.0)
        self.ix, self.iy, self.iz = mk(dndx), mk(dndy), mk(dndz)

    def dndr(self, pos):
        g = np.zeros_like(pos)
        p = pos.T
        g[0] = self.ix(p)
        g[1] = self.iy(p)
        g[2] = self.iz(p)
        return g


def make_field(ne, x, y, z, lwl=1053e-9, ne_max=1):
    d = calc_dndr(ne, x, y, z, lwl, ne_max)
    return GradientField(x, y, z, d["dndx"], d["dndy"], d["dndz"])


# --------------------------------------------------------------------------------------
# beam: particle_tracker.py:258-310
# --------------------------------------------------------------------------------------
def init_beam(Np, beam_size, divergence, extent, probing_direction="z", rand=None, randn=None):
    """Launch rays ``s0`` (6, Np).  Draw order t, u1, u2, phi, chi from the global numpy
    RNG (``np.random.rand`` / ``randn``) exactly like particle_tracker.py:273-278."""
    rand = rand or np.random.rand
    randn = randn or np.random.randn
    t = 2 * np.pi * rand(Np)
    u = rand(Np) + rand(Np)
    fold = u > 1
    u[fold] = 2 - u[fold]
    phi = np.pi * rand(Np)
    chi = divergence * randn(Np)
    a = beam_size * u * np.cos(t)
    b = beam_size * u * np.sin(t)
    vpar = C_LIGHT * np.cos(chi)
    v1 = C_LIGHT * np.sin(chi) * np.cos(phi)
    v2 = C_LIGHT * np.sin(chi) * np.sin(phi)
    s0 = np.zeros((6, Np))
    if probing_direction == "x":      # :280-289  (launch at +extent: quirk kept)
        s0[0], s0[1], s0[2] = extent, a, b
        s0[3], s0[4], s0[5] = vpar, v1, v2
    elif probing_direction == "y":    # :290-299
        s0[0], s0[1], s0[2] = a, -extent, b
        s0[3], s0[4], s0[5] = v1, vpar, v2
    elif probing_direction == "z":    # :300-309
        s0[0], s0[1], s0[2] = a, b, -extent
        s0[3], s0[4], s0[5] = v1, v2, vpar
    else:
        raise ValueError(probing_direction)
    return s0


# --------------------------------------------------------------------------------------
# ODE: particle_tracker.py:398-419 (rhs), :312-331 (solve), :333-380 (exit plane)
# --------------------------------------------------------------------------------------
def dsdt(t, s, field):
    n = s.size // 6
    s = s.reshape(6, n)
    out = np.zeros_like(s)
    out[3:6] = field.dndr(s[:3])
    out[:3] = s[3:]
    return out.flatten()


def ray_at_exit(sf, extent, probing_direction="z"):
    """Linear back-projection onto the plane ``axis = +extent``; rows
    (p1, atan(v1/vpar), p2, atan(v2/vpar)); particle_tracker.py:345-380."""
    par, a1, a2 = {"x": (0, 1, 2), "y": (1, 0, 2), "z": (2, 0, 1)}[probing_direction]
    p, v = sf[:3], sf[3:]
    tb = (p[par] - extent) / v[par]
    rf = np.zeros((4, sf.shape[1]))
    rf[0] = p[a1] - v[a1] * tb
    rf[2] = p[a2] - v[a2] * tb
    rf[1] = np.arctan(v[a1] / v[par])
    rf[3] = np.arctan(v[a2] / v[par])
    return rf


def solve(field, s0, extent, probing_direction="z", rtol=1e-3, atol=1e-6, batch=None,
          method="RK45"):
    """Integrate rays over [0, sqrt(8)*extent/c] and return (rf, sf, nfev_total).

    With ``rtol=1e-3, atol=1e-6, batch=None`` this is the reference's ``ElectronCube.solve``
    (particle_tracker.py:317-330: one ``solve_ivp`` over the flattened 6N system, default
    tolerances).  The parity oracle is the same call with tight tolerances and small
    ``batch`` (the step size is shared by all rays of a call).
    """
    T = np.sqrt(8.0) * extent / C_LIGHT
    n = s0.shape[1]
    batch = batch or n
    sf = np.empty_like(s0, dtype=np.float64)
    nfev = 0
    for lo in range(0, n, batch):
        hi = min(n, lo + batch)
        y0 = np.ascontiguousarray(s0[:, lo:hi]).flatten()
        sol = solve_ivp(lambda t, y: dsdt(t, y, field), [0, T], y0, t_eval=[0.0, T],
                        method=method, rtol=rtol, atol=atol)
        sf[:, lo:hi] = sol.y[:, -1].reshape(6, hi - lo)
        nfev += sol.nfev * (hi - lo)
    return ray_at_exit(sf, extent, probing_direction), sf, nfev


# --------------------------------------------------------------------------------------
# ray-transfer-matrix optics: ray_transfer_matrix.py:37-154
# --------------------------------------------------------------------------------------
def m_to_mm(r):                                   # :37-40
    out = np.array(r, dtype=np.float64, copy=True)
    out[0::2] *= 1e3
    return out


def _blockdiag(a, b):
    M = np.zeros((4, 4))
    M[:2, :2] = a
    M[2:, 2:] = b
    return M


def lens(r, f1, f2):                              # :42-54
    return _blockdiag([[1, 0], [-1 / f1, 1]], [[1, 0], [-1 / f2, 1]]) @ r


def sym_lens(r, f):                               # :56-60
    return lens(r, f, f)


def distance(r, d):                               # :62-71
    m = [[1, d], [0, 1]]
    return _blockdiag(m, m) @ r


def _reject(r, mask):
    r[:, mask] = np.nan                           # the reference writes None -> NaN
    return r


def circular_aperture(r, R):                      # :73-79   keep r2 <= R2
    return _reject(r, r[0] ** 2 + r[2] ** 2 > R**2)


def circular_stop(r, R):                          # :81-87   keep r2 >= R2
    return _reject(r, r[0] ** 2 + r[2] ** 2 < R**2)


def annular_stop(r, R1, R2):                      # :89-97   returns the mask only
    rr = r[0] ** 2 + r[2] ** 2
    return (rr > R1**2) & (rr < R2**2)


def angular_filter(r, Rs):                        # :99-111
    m = np.zeros(r.shape[1], dtype=bool)
    for i in range(len(Rs) // 2):
        m |= annular_stop(r, Rs[2 * i], Rs[2 * i + 1])
    return _reject(r, m)


def rect_aperture(r, Lx, Ly):                     # :128-136  (outside in BOTH axes)
    return _reject(r, (r[0] ** 2 > Lx**2) & (r[2] ** 2 > Ly**2))


def knife_edge(r, offset, axis, direction):       # :138-154
    a = {"x": 0, "y": 2}[axis]
    if direction > 0:
        return _reject(r, r[a] > offset)
    if direction < 0:
        return _reject(r, r[a] < offset)
    raise ValueError("direction must be <0 or >0")


def detector(kind, r0_m, focal_plane=0, L=400, R=25, **kw):
    """Element programs of ray_transfer_matrix.py:214-227 (Shadowgraphy), :236-250
    (Schlieren_DF), :259-273 (Schlieren_LF), :285-299 (AFR).  ``r0_m`` in metres/radians
    (``Rays.__init__`` converts, :171-172).  Returns rf (4, N) in mm/rad."""
    r = m_to_mm(r0_m)
    if kind == "shadowgraphy":
        r = distance(r, L - focal_plane)
        r = circular_aperture(r, R)
        r = sym_lens(r, L)
        r = distance(r, 2 * L)
        r = circular_aperture(r, R)
        r = sym_lens(r, L)
        return distance(r, L)
    if kind in ("schlieren_df", "schlieren_lf"):
        Rs = kw.get("R_stop", 1)
        r = distance(r, L - focal_plane)
        r = circular_aperture(r, R)
        r = sym_lens(r, L)
        r = distance(r, L)
        r = circular_stop(r, Rs) if kind == "schlieren_df" else circular_aperture(r, Rs)
        r = distance(r, L)
        r = circular_aperture(r, R)
        r = sym_lens(r, L)
        return distance(r, L)
    if kind == "afr":
        r = distance(r, L / 2 - focal_plane)
        r = circular_aperture(r, R)
        r = sym_lens(r, L / 2)
        r = distance(r, L / 4)
        r = angular_filter(r, kw["Rs"])
        r = distance(r, L / 4)
        r = circular_aperture(r, R)
        r = sym_lens(r, L / 2)
        return distance(r, L / 2)
    raise ValueError(kind)


def histogram(rf, Lx=18, Ly=13.5, bin_scale=10, pix_x=3448, pix_y=2574):
    """ray_transfer_matrix.py:173-191: NaN drop, histogram2d, transpose."""
    x, y = rf[0], rf[2]
    x = x[~np.isnan(x)]
    y = y[~np.isnan(y)]
    H, xe, ye = np.histogram2d(x, y, bins=[pix_x // bin_scale, pix_y // bin_scale],
                               range=[[-Lx / 2, Lx / 2], [-Ly / 2, Ly / 2]])
    return H.T, xe, ye


# --------------------------------------------------------------------------------------
# Gaussian random field: gaussian_fields/turboGen.py:388-434, 436-486, 488-538
# --------------------------------------------------------------------------------------
def gaussian_fft(N, k_func, ndim=3, Wr=None, Wi=None):
    """Timmer-Koenig synthesis on an odd grid M = 2N+1.  ``Wr``/``Wi`` default to
    ``np.random.randn(M,..)`` drawn in that order (turboGen.py:522-523)."""
    M = 2 * N + 1
    k = np.fft.fftfreq(M)
    grids = np.meshgrid(*([k] * ndim)) if ndim > 1 else [k]
    K = np.fft.fftshift(np.sqrt(sum(g**2 for g in grids)))
    shape = (M,) * ndim
    if Wr is None:
        Wr = np.random.randn(*shape)
    if Wi is None:
        Wi = np.random.randn(*shape)
    W = (Wr + np.flip(Wr)) + 1j * (Wi - np.flip(Wi))
    with np.errstate(divide="ignore", invalid="ignore"):
        F = W * np.sqrt(k_func(K))
    F = np.fft.ifftshift(F)
    F[(0,) * ndim] = 0
    return np.fft.ifftn(F).real


# --------------------------------------------------------------------------------------
# spectrum diagnostic: gaussian_fields/calculate_spectrum_3d.py:3-59
# --------------------------------------------------------------------------------------
def spectrum_3d_scalar(data, dx, k_bin_num=100):
    """Shell-averaged |fftn(data)|^2: shells of width K.max()/k_bin_num, mean power in shells
    0..k_bin_num-2 (the reference's loop leaves the last one at 0), volume-weighted centres."""
    P = np.abs(np.fft.fftn(data)) ** 2
    k = [np.fft.fftfreq(m, dx) for m in data.shape]
    K = np.sqrt(k[0][:, None, None] ** 2 + k[1][None, :, None] ** 2 + k[2][None, None, :] ** 2)
    w = K.max() / k_bin_num
    edges = w * np.arange(0, k_bin_num + 1)
    centres = (0.5 * (edges[:-1] ** 3 + edges[1:] ** 3)) ** (1 / 3)
    spec = np.zeros_like(centres)
    Kf, Pf = K.ravel(), P.ravel()
    with np.errstate(invalid="ignore"), __import__("warnings").catch_warnings():
        __import__("warnings").simplefilter("ignore")
        for i in range(1, k_bin_num):
            spec[i - 1] = Pf[(Kf < i * w) & (Kf >= (i - 1) * w)].mean()
    return centres, spec


# --------------------------------------------------------------------------------------
# Magnetised / absorbing extension -- PARITY UNPINNED.
# The reference checkout holds only call sites for these quantities
# (particle_tracking/example_kitchensink.py:72-101); there is nothing to restate.  This is an
# independent FP64 evaluation of the same textbook forms the CUDA path documents
# (include/tt_b200.h, tt_trace_aux), integrated by scipy along with the ray:
#   d ln(a)/dt = -c kappa / 2,  dphi/dt = omega (sqrt(1 - ne/nc) - 1),  dalpha/dt = V ne (B . v)
# --------------------------------------------------------------------------------------
VERDET = _sc.e**3 / (8 * np.pi**2 * _sc.epsilon_0 * _sc.m_e**2 * _sc.c**3)


def solve_aux(ne, B, kappa, x, y, z, s0, extent, probing_direction="z", lwl=1053e-9, ne_max=1,
              rtol=1e-10, atol=1e-13, batch=16):
    """Returns (rf, amplitude, phase, rotation) for rays s0 (6, N).  B: (nx, ny, nz, 3) or None,
    kappa: (nx, ny, nz) in 1/m or None."""
    d = calc_dndr(ne, x, y, z, lwl, ne_max)
    field = GradientField(x, y, z, d["dndx"], d["dndy"], d["dndz"])
    mk = lambda a: RegularGridInterpolator((x, y, z), a, bounds_error=False, fill_value=0.0)
    n_i = mk(d["ne_nc"])
    B_i = [mk(B[..., k]) for k in range(3)] if B is not None else None
    k_i = mk(kappa) if kappa is not None else None
    omega, nc = d["omega"], d["nc"]
    V = VERDET * lwl**2
    T = np.sqrt(8.0) * extent / C_LIGHT
    n = s0.shape[1]
    out = np.zeros((9, n))

    def rhs(t, y):
        m = y.size // 9
        s = y.reshape(9, m)
        o = np.zeros_like(s)
        p = s[:3].T
        o[:3] = s[3:6]
        o[3:6] = field.dndr(s[:3])
        nn = n_i(p)
        o[7] = omega * (np.sqrt(np.clip(1 - nn, 0, None)) - 1)
        if B_i is not None:
            o[8] = V * nc * nn * sum(B_i[k](p) * s[3 + k] for k in range(3))
        if k_i is not None:
            o[6] = -0.5 * C_LIGHT * k_i(p)
        return o.flatten()

    for lo in range(0, n, batch):
        hi = min(n, lo + batch)
        y0 = np.zeros((9, hi - lo))
        y0[:6] = s0[:, lo:hi]
        sol = solve_ivp(rhs, [0, T], y0.flatten(), t_eval=[0.0, T], method="RK45", rtol=rtol,
                        atol=np.tile(np.repeat([atol, atol * 1e12, 1e-10], [3, 3, 3]), (hi - lo, 1)).T.flatten())
        out[:, lo:hi] = sol.y[:, -1].reshape(9, hi - lo)
    return ray_at_exit(out[:6], extent, probing_direction), np.exp(out[6]), out[7], out[8]
