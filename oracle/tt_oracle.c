/* ORACLE -- TEST INFRASTRUCTURE ONLY.  Not shipped, not measured as the product, not a fallback.
 *
 * Plain-C restatement (FP64, no dependencies) of the ray-integration hot path of
 * jdhare/turbulence_tracing INCLUDING the third-party arithmetic it delegates to.  The reference pins no
 * versions (no requirements file); the algorithms restated here are those of numpy 2.3.5 / scipy 1.18.1,
 * the versions the golden fixtures under tests/golden/ were generated with:
 *
 *   tto_calc_dndr    particle_tracker.py:227-237            nc, clip, -c^2/2 * numpy.gradient(ne/nc, axis)
 *                    numpy/lib/_function_base_impl.py:995   (second-order interior -- uniform or non-uniform
 *                                                            spacing formula -- first-order one-sided faces)
 *   tto_dndr         particle_tracker.py:239-256            three RegularGridInterpolator(linear,
 *                    scipy/interpolate/_rgi.py:520-550       bounds_error=False, fill_value=0): bisection cell
 *                    + _rgi_cython.find_indices              search, weights multiplied and corners summed in
 *                                                            scipy's order, 0 outside, NaN for NaN input
 *   tto_solve        particle_tracker.py:312-331, 398-419   scipy.integrate.solve_ivp(method='RK45') over the
 *                    scipy/integrate/_ivp/rk.py,             flattened (6, Np) system: Dormand-Prince 5(4),
 *                    _ivp/common.py:select_initial_step      Hairer's initial step, RMS error norm over ALL
 *                    _ivp/ivp.py (t_eval handling)           6 Np components (one step sequence per bundle),
 *                                                            SAFETY 0.9, factors in [0.2, 10], final state
 *                                                            through the dense-output polynomial at t = T
 *   tto_ray_at_exit  particle_tracker.py:333-380            back-projection to the exit plane
 *
 * Pinned (tests/test_oracle_c.py) against the golden vectors produced by the live reference: bit-equal for the
 * gradient grid and the interpolation, rounding-level (BLAS summation order) for the integrator at scipy's
 * default and at tight tolerances.  Compile with -ffp-contract=off so that no FMA is formed where numpy rounds
 * twice.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may load this library.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <unistd.h>

#define TTO_C 299792458.0             /* scipy.constants.c, particle_tracker.py:119 */
#define TTO_NC_OVER_OMEGA2 3.14207787e-4 /* particle_tracker.py:228 */
#define TTO_PI 3.141592653589793

typedef struct {
    int nx, ny, nz;
    const double *x, *y, *z;          /* ascending node coordinates */
    const double *gx, *gy, *gz;       /* dndx, dndy, dndz, C order [ix][iy][iz] */
} tto_field;

/* ------------------------------------------------------------------------------------------------------
 * numpy.gradient(f, coord, axis=axis) * k        (edge_order = 1)
 * ------------------------------------------------------------------------------------------------------ */
static int gradient_axis(const double* f, const int dims[3], int axis, const double* coord, double k, double* out) {
    const long n = dims[axis];
    if (n < 2) return 1;
    const long stride = axis == 0 ? (long)dims[1] * dims[2] : axis == 1 ? dims[2] : 1;
    const long total = (long)dims[0] * dims[1] * dims[2];
    double* dx = (double*)malloc(sizeof(double) * (size_t)(n - 1));
    double *ca = NULL, *cb = NULL, *cc = NULL;
    if (!dx) return 2;
    int uniform = 1;
    for (long i = 0; i < n - 1; ++i) {
        dx[i] = coord[i + 1] - coord[i];
        if (dx[i] != dx[0]) uniform = 0;          /* (diffx == diffx[0]).all() */
    }
    if (!uniform && n > 2) {
        ca = (double*)malloc(sizeof(double) * (size_t)n);
        cb = (double*)malloc(sizeof(double) * (size_t)n);
        cc = (double*)malloc(sizeof(double) * (size_t)n);
        if (!ca || !cb || !cc) { free(dx); free(ca); free(cb); free(cc); return 2; }
        for (long i = 1; i < n - 1; ++i) {
            const double dx1 = dx[i - 1], dx2 = dx[i];
            ca[i] = -(dx2) / (dx1 * (dx1 + dx2));
            cb[i] = (dx2 - dx1) / (dx1 * dx2);
            cc[i] = dx1 / (dx2 * (dx1 + dx2));
        }
    }
    const double two_dx = 2. * dx[0];
    for (long e = 0; e < total; ++e) {
        const long i = (e / stride) % n;
        double g;
        if (i == 0) g = (f[e + stride] - f[e]) / dx[0];
        else if (i == n - 1) g = (f[e] - f[e - stride]) / dx[n - 2];
        else if (uniform) g = (f[e + stride] - f[e - stride]) / two_dx;
        else g = (ca[i] * f[e - stride] + cb[i] * f[e]) + cc[i] * f[e + stride];
        out[e] = k * g;
    }
    free(dx); free(ca); free(cb); free(cc);
    return 0;
}

/* particle_tracker.py:227-237.  ne_nc, gx, gy, gz: caller-allocated nx*ny*nz doubles. */
int tto_calc_dndr(const double* ne, int nx, int ny, int nz, const double* x, const double* y, const double* z,
                  double lwl, double ne_max, double* ne_nc, double* gx, double* gy, double* gz, double* omega_out,
                  double* nc_out) {
    const double omega = 2 * TTO_PI * (TTO_C / lwl);
    const double nc = TTO_NC_OVER_OMEGA2 * (omega * omega);
    const long total = (long)nx * ny * nz;
    for (long e = 0; e < total; ++e) {
        double v = ne[e] / nc;
        if (v > ne_max) v = ne_max;
        ne_nc[e] = v;
    }
    const int dims[3] = {nx, ny, nz};
    const double k = -0.5 * (TTO_C * TTO_C);
    int rc = gradient_axis(ne_nc, dims, 0, x, k, gx);
    if (!rc) rc = gradient_axis(ne_nc, dims, 1, y, k, gy);
    if (!rc) rc = gradient_axis(ne_nc, dims, 2, z, k, gz);
    if (omega_out) *omega_out = omega;
    if (nc_out) *nc_out = nc;
    return rc;
}

/* ------------------------------------------------------------------------------------------------------
 * RegularGridInterpolator, method='linear', bounds_error=False, fill_value=0.0
 * ------------------------------------------------------------------------------------------------------ */
/* _rgi_cython.find_indices -> find_interval_ascending(extrapolate=1): g[i] <= v < g[i+1], clamped to [0, n-2] */
static inline int find_interval(const double* g, int n, double v) {
    if (!(v >= g[0])) return 0;
    if (v >= g[n - 1]) return n - 2;
    int lo = 0, hi = n - 1;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (v >= g[mid]) lo = mid; else hi = mid;
    }
    return lo;
}

static inline void dndr_point(const tto_field* F, double px, double py, double pz, double* ox, double* oy, double* oz) {
    if (px != px || py != py || pz != pz) { *ox = *oy = *oz = NAN; return; }          /* f(nan) = nan */
    if (px < F->x[0] || px > F->x[F->nx - 1] || py < F->y[0] || py > F->y[F->ny - 1] || pz < F->z[0] ||
        pz > F->z[F->nz - 1]) { *ox = *oy = *oz = 0.0; return; }                     /* fill_value */
    const int i0 = find_interval(F->x, F->nx, px), i1 = find_interval(F->y, F->ny, py), i2 = find_interval(F->z, F->nz, pz);
    const double y0 = (px - F->x[i0]) / (F->x[i0 + 1] - F->x[i0]);
    const double y1 = (py - F->y[i1]) / (F->y[i1 + 1] - F->y[i1]);
    const double y2 = (pz - F->z[i2]) / (F->z[i2 + 1] - F->z[i2]);
    const double w0[2] = {1 - y0, y0}, w1[2] = {1 - y1, y1}, w2[2] = {1 - y2, y2};
    const long sy = F->nz, sx = (long)F->ny * F->nz;
    double vx = 0., vy = 0., vz = 0.;
    /* itertools.product over ((i, 1-y), (i+1, y)) per dimension, last dimension fastest;
       weight = ((1. * w_a) * w_b) * w_c;  value = value + values[corner] * weight */
    for (int a = 0; a < 2; ++a)
        for (int b = 0; b < 2; ++b)
            for (int c = 0; c < 2; ++c) {
                const double w = (w0[a] * w1[b]) * w2[c];
                const long e = (long)(i0 + a) * sx + (long)(i1 + b) * sy + (i2 + c);
                vx = vx + F->gx[e] * w;
                vy = vy + F->gy[e] * w;
                vz = vz + F->gz[e] * w;
            }
    *ox = vx; *oy = vy; *oz = vz;
}

/* ElectronCube.dndr, particle_tracker.py:243-256.  pos, out: (3, n) row-major. */
void tto_dndr(const tto_field* F, const double* pos, long n, double* out) {
    for (long r = 0; r < n; ++r)
        dndr_point(F, pos[r], pos[n + r], pos[2 * n + r], &out[r], &out[n + r], &out[2 * n + r]);
}

/* dsdt, particle_tracker.py:398-419: y = (6, n) row-major flattened */
static void rhs(const tto_field* F, const double* y, long n, double* f) {
    memcpy(f, y + 3 * n, sizeof(double) * (size_t)(3 * n));
    for (long r = 0; r < n; ++r)
        dndr_point(F, y[r], y[n + r], y[2 * n + r], &f[3 * n + r], &f[4 * n + r], &f[5 * n + r]);
}

/* ------------------------------------------------------------------------------------------------------
 * scipy.integrate.solve_ivp(fun, [0, T], y0, t_eval=[0, T], method='RK45', rtol, atol)
 * ------------------------------------------------------------------------------------------------------ */
static const double RK_C[6] = {0, 1. / 5, 3. / 10, 4. / 5, 8. / 9, 1};
static const double RK_A[6][5] = {
    {0, 0, 0, 0, 0},
    {1. / 5, 0, 0, 0, 0},
    {3. / 40, 9. / 40, 0, 0, 0},
    {44. / 45, -56. / 15, 32. / 9, 0, 0},
    {19372. / 6561, -25360. / 2187, 64448. / 6561, -212. / 729, 0},
    {9017. / 3168, -355. / 33, 46732. / 5247, 49. / 176, -5103. / 18656}};
static const double RK_B[6] = {35. / 384, 0, 500. / 1113, 125. / 192, -2187. / 6784, 11. / 84};
static const double RK_E[7] = {-71. / 57600, 0, 71. / 16695, -71. / 1920, 17253. / 339200, -22. / 525, 1. / 40};
static const double RK_P[7][4] = {
    {1, -8048581381. / 2820520608, 8663915743. / 2820520608, -12715105075. / 11282082432},
    {0, 0, 0, 0},
    {0, 131558114200. / 32700410799, -68118460800. / 10900136933, 87487479700. / 32700410799},
    {0, -1754552775. / 470086768, 14199869525. / 1410260304, -10690763975. / 1880347072},
    {0, 127303824393. / 49829197408, -318862633887. / 49829197408, 701980252875. / 199316789632},
    {0, -282668133. / 205662961, 2019193451. / 616988883, -1453857185. / 822651844},
    {0, 40617522. / 29380423, -110615467. / 29380423, 69997945. / 29380423}};
/* Guard of this restatement (scipy has none): a ray that starts exactly ON a face flying outwards with v = c can
 * make solve_ivp crawl in 1e-27 s steps for ever at tight tolerances (x + v h rounds back onto the face, where the
 * field jumps); such a bundle is reported as failed instead of hanging the test suite. */
#define TTO_MAX_ATTEMPTS 4000000L
#define RK_SAFETY 0.9
#define RK_MIN_FACTOR 0.2
#define RK_MAX_FACTOR 10.0

/* common.norm: np.linalg.norm(x) / x.size ** 0.5 of x_i = num_i / scale_i */
static double rms_ratio(const double* num, const double* scale, long N, double mul) {
    double s = 0.;
    for (long i = 0; i < N; ++i) {
        const double v = num[i] * mul / scale[i];
        s += v * v;
    }
    return sqrt(s) / pow((double)N, 0.5);
}

/* One bundle: n rays integrated with ONE adaptive step sequence (error norm over all 6n components).
 * y0, y_out: (6, n) row-major.  Returns nfev (>0), or -1 on allocation failure, -2 if the step size
 * underflowed (solve_ivp status -1).  n_steps / n_rejected / t_hist (the accepted step end times, at most
 * t_cap of them) are optional diagnostics. */
/* max_step: solve_ivp's option of that name (scipy/integrate/_ivp/rk.py: RungeKutta.__init__ -> validate_max_step,
 * _step_impl: "if self.h_abs > max_step: h_abs = max_step", select_initial_step bounded by it as well).  The reference
 * leaves it at its default, infinity (particle_tracker.py:322).  The sharp reference of the parity tests sets it to one
 * cell's transit time: behind an exactly flat stretch of the field (ne clipped to 0 over a few cells, free space) the
 * error estimate is 0, the step grows tenfold per step, and a millimetre-long step whose seven stage points all happen
 * to land in flat spots again is accepted at ANY rtol -- it leaps over the structure in between (observed on the
 * 513^3 benchmark cube: 8 of 8192 rays off by 0.1 .. 0.4 micrometres at rtol = 1e-13 while rtol = 1e-12 is right). */
long tto_solve_ivp_rk45_ms(const tto_field* F, const double* y0, long n, double T, double rtol, double atol, double max_step,
                           double* y_out, long* n_steps, long* n_rejected, double* t_hist, long t_cap) {
    const long N = 6 * n;
    if (n <= 0) return 0;
    double* buf = (double*)malloc(sizeof(double) * (size_t)N * 12);
    if (!buf) return -1;
    double* K[7];
    for (int s = 0; s < 7; ++s) K[s] = buf + (size_t)s * N;
    double *y = buf + 7 * (size_t)N, *ynew = y + N, *ytmp = ynew + N, *f = ytmp + N, *scale = f + N;
    long nfev = 0, steps = 0, rejected = 0;
    const double eps100 = 100 * 2.220446049250313e-16;
    if (rtol < eps100) rtol = eps100;                      /* validate_tol */
    memcpy(y, y0, sizeof(double) * (size_t)N);
    double t = 0.0;
    rhs(F, y, n, f); ++nfev;
    /* ---- select_initial_step (order = error_estimator_order = 4) ---- */
    double h_abs;
    {
        for (long i = 0; i < N; ++i) scale[i] = atol + fabs(y[i]) * rtol;
        const double d0 = rms_ratio(y, scale, N, 1.0), d1 = rms_ratio(f, scale, N, 1.0);
        double h0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * d0 / d1;
        if (h0 > T) h0 = T;
        for (long i = 0; i < N; ++i) ytmp[i] = y[i] + h0 * 1.0 * f[i];
        rhs(F, ytmp, n, K[0]); ++nfev;
        for (long i = 0; i < N; ++i) K[1][i] = K[0][i] - f[i];
        const double d2 = rms_ratio(K[1], scale, N, 1.0) / h0;
        double h1;
        if (d1 <= 1e-15 && d2 <= 1e-15) h1 = fmax(1e-6, h0 * 1e-3);
        else h1 = pow(0.01 / fmax(d1, d2), 1.0 / 5.0);
        h_abs = fmin(fmin(fmin(100 * h0, h1), T), max_step);
    }
    double t_old = 0.0;
    int failed = 0;
    while (t < T) {                                        /* solver.status == 'running' */
        const double min_step = 10 * fabs(nextafter(t, INFINITY) - t);
        if (h_abs > max_step) h_abs = max_step;
        else if (h_abs < min_step) h_abs = min_step;
        int step_rejected = 0;
        double h, t_new;
        for (;;) {
            if (h_abs < min_step || steps + rejected > TTO_MAX_ATTEMPTS) { failed = 1; break; }
            h = h_abs;
            t_new = t + h;
            if (t_new - T > 0) t_new = T;
            h = t_new - t;
            h_abs = fabs(h);
            /* ---- rk_step ---- */
            memcpy(K[0], f, sizeof(double) * (size_t)N);
            for (int s = 1; s < 6; ++s) {
                for (long i = 0; i < N; ++i) {
                    double dy = 0.;
                    for (int j = 0; j < s; ++j) dy += K[j][i] * RK_A[s][j];
                    ytmp[i] = y[i] + dy * h;
                }
                rhs(F, ytmp, n, K[s]); ++nfev;
            }
            for (long i = 0; i < N; ++i) {
                double acc = 0.;
                for (int j = 0; j < 6; ++j) acc += K[j][i] * RK_B[j];
                ynew[i] = y[i] + h * acc;
            }
            rhs(F, ynew, n, K[6]); ++nfev;
            /* ---- error estimate ---- */
            double ss = 0.;
            for (long i = 0; i < N; ++i) {
                double e = 0.;
                for (int j = 0; j < 7; ++j) e += K[j][i] * RK_E[j];
                const double sc = atol + fmax(fabs(y[i]), fabs(ynew[i])) * rtol;
                const double v = e * h / sc;
                ss += v * v;
            }
            const double error_norm = sqrt(ss) / pow((double)N, 0.5);
            if (error_norm < 1) {
                double factor = error_norm == 0 ? RK_MAX_FACTOR : fmin(RK_MAX_FACTOR, RK_SAFETY * pow(error_norm, -0.2));
                if (step_rejected) factor = fmin(1.0, factor);
                h_abs *= factor;
                break;
            }
            h_abs *= fmax(RK_MIN_FACTOR, RK_SAFETY * pow(error_norm, -0.2));
            step_rejected = 1;
            ++rejected;
        }
        if (failed) break;
        if (t_hist && steps < t_cap) t_hist[steps] = t_new;
        ++steps;
        t_old = t;
        t = t_new;
        if (t >= T) {
            /* solve_ivp evaluates t_eval points through the dense output of the step that contains them:
               y(T) = y_old + h * Q @ (1, 1, 1, 1),  Q = K.T @ P,  h = t - t_old (x = 1 exactly) */
            const double hd = t - t_old;
            for (long i = 0; i < N; ++i) {
                double q = 0.;
                for (int c = 0; c < 4; ++c) {
                    double qc = 0.;
                    for (int j = 0; j < 7; ++j) qc += K[j][i] * RK_P[j][c];
                    q += qc;
                }
                y_out[i] = hd * q + y[i];
            }
        }
        double* sw = y; y = ynew; ynew = sw;               /* y <- y_new */
        memcpy(f, K[6], sizeof(double) * (size_t)N);       /* f <- f_new */
    }
    free(buf);
    if (n_steps) *n_steps = steps;
    if (n_rejected) *n_rejected = rejected;
    return failed ? -2 : nfev;
}

long tto_solve_ivp_rk45(const tto_field* F, const double* y0, long n, double T, double rtol, double atol, double* y_out,
                        long* n_steps, long* n_rejected, double* t_hist, long t_cap) {
    return tto_solve_ivp_rk45_ms(F, y0, n, T, rtol, atol, INFINITY, y_out, n_steps, n_rejected, t_hist, t_cap);
}

/* Many bundles of `batch` rays each (the last one may be shorter), one solve_ivp call per bundle like the
 * reference run under multiprocessing.Pool (example_multiprocess.py:41-51); bundles are handed out to
 * `threads` POSIX threads (this image's gcc has no libgomp).  s0, sf: (6, n) row-major.  Returns the sum
 * over bundles of nfev * rays (ray-RHS evaluations), negative on allocation failure.  A bundle whose step
 * size underflows (solve_ivp status -1, e.g. a ray that slides along a face of the cube, where the field
 * jumps to zero) gets NaN outputs and is counted in *n_failed. */
typedef struct {
    const tto_field* F;
    const double* s0;
    double* sf;
    long n, batch, nb;
    double T, rtol, atol, max_step;
    long next;                 /* next bundle to hand out (guarded by mu) */
    long long total;
    int bad;                   /* 1: allocation failure */
    long n_failed;             /* bundles whose step size underflowed (solve_ivp status -1): outputs set to NaN */
    pthread_mutex_t mu;
} solve_job;

static void* solve_worker(void* arg) {
    solve_job* J = (solve_job*)arg;
    for (;;) {
        pthread_mutex_lock(&J->mu);
        const long b = J->next++;
        pthread_mutex_unlock(&J->mu);
        if (b >= J->nb) break;
        const long n = J->n, lo = b * J->batch, m = (lo + J->batch <= n ? J->batch : n - lo);
        double* in = (double*)malloc(sizeof(double) * (size_t)(12 * m));
        long nfev = -1;
        if (in) {
            double* out = in + 6 * m;
            for (int k = 0; k < 6; ++k) memcpy(in + k * m, J->s0 + (size_t)k * n + lo, sizeof(double) * (size_t)m);
            nfev = tto_solve_ivp_rk45_ms(J->F, in, m, J->T, J->rtol, J->atol, J->max_step, out, NULL, NULL, NULL, 0);
            if (nfev >= 0)
                for (int k = 0; k < 6; ++k) memcpy(J->sf + (size_t)k * n + lo, out + k * m, sizeof(double) * (size_t)m);
            else if (nfev == -2)
                for (int k = 0; k < 6; ++k)
                    for (long r = 0; r < m; ++r) J->sf[(size_t)k * n + lo + r] = NAN;
            free(in);
        }
        pthread_mutex_lock(&J->mu);
        if (nfev == -1) J->bad |= 1;
        else if (nfev == -2) J->n_failed += 1;
        else J->total += (long long)nfev * m;
        pthread_mutex_unlock(&J->mu);
    }
    return NULL;
}

int tto_max_threads(void) {
    const long c = sysconf(_SC_NPROCESSORS_ONLN);
    return c > 0 ? (int)c : 1;
}

long long tto_solve_ms(const tto_field* F, const double* s0, long n, long batch, double T, double rtol, double atol,
                       double max_step, double* sf, int threads, long* n_failed) {
    if (n_failed) *n_failed = 0;
    if (n <= 0) return 0;
    if (batch <= 0 || batch > n) batch = n;
    solve_job J;
    J.F = F; J.s0 = s0; J.sf = sf; J.n = n; J.batch = batch; J.nb = (n + batch - 1) / batch;
    J.T = T; J.rtol = rtol; J.atol = atol; J.max_step = max_step > 0 ? max_step : INFINITY; J.next = 0; J.total = 0; J.bad = 0; J.n_failed = 0;
    pthread_mutex_init(&J.mu, NULL);
    if (threads <= 0) threads = tto_max_threads();
    if (threads > J.nb) threads = (int)J.nb;
    if (threads <= 1) solve_worker(&J);
    else {
        pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)threads);
        int started = 0;
        if (th)
            for (; started < threads; ++started)
                if (pthread_create(&th[started], NULL, solve_worker, &J)) break;
        if (!started) solve_worker(&J);
        for (int i = 0; i < started; ++i) pthread_join(th[i], NULL);
        free(th);
    }
    pthread_mutex_destroy(&J.mu);
    if (n_failed) *n_failed = J.n_failed;
    return J.bad ? -(long long)J.bad : J.total;
}

long long tto_solve(const tto_field* F, const double* s0, long n, long batch, double T, double rtol, double atol,
                    double* sf, int threads, long* n_failed) {
    return tto_solve_ms(F, s0, n, batch, T, rtol, atol, INFINITY, sf, threads, n_failed);
}

/* ElectronCube.ray_at_exit, particle_tracker.py:345-380.  dir: 0 = 'x', 1 = 'y', 2 = 'z'. */
void tto_ray_at_exit(const double* sf, long n, double extent, int dir, double* rf) {
    const int par = dir, a1 = dir == 0 ? 1 : 0, a2 = dir == 2 ? 1 : 2;
    for (long r = 0; r < n; ++r) {
        const double pp = sf[(size_t)par * n + r], vp = sf[(size_t)(3 + par) * n + r];
        const double p1 = sf[(size_t)a1 * n + r], v1 = sf[(size_t)(3 + a1) * n + r];
        const double p2 = sf[(size_t)a2 * n + r], v2 = sf[(size_t)(3 + a2) * n + r];
        const double tb = (pp - extent) / vp;
        rf[r] = p1 - v1 * tb;
        rf[n + r] = atan(v1 / vp);
        rf[2 * n + r] = p2 - v2 * tb;
        rf[3 * n + r] = atan(v2 / vp);
    }
}
