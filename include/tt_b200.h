/*
 * tt_b200.h -- C ABI of libtt_b200.so: the B200 (sm_100a) implementation of the
 * ray-integration hot path of jdhare/turbulence_tracing.
 *
 * The reference has no FFI of its own (it is pure Python); each entry point
 * below names the reference function it replaces (paths relative to the reference checkout).
 * INTEGRATION.md shows the ctypes binding a maintainer of the reference would add.
 *
 * Conventions
 *   - plain pointers and sizes only; every *_dev pointer is DEVICE memory owned by the caller,
 *     every other pointer is HOST memory.  The library allocates nothing persistent: scratch
 *     space is passed in after a *_workspace() size query.  One exception: tt_trace / tt_trace_faces /
 *     tt_trace_aux / tt_trace_axes take stream-ordered scratch (cudaMallocFromPoolAsync / cudaFreeAsync on
 *     `stream`) from a memory pool OWNED by the library, which keeps what it was given across
 *     synchronisations (a fresh allocation costs milliseconds per launch): 8 bytes of flag words ("did the
 *     first pass defer a ray?", how many) and, for tt_trace / tt_trace_faces, a list of the deferred ray ids
 *     (4 bytes per ray of the launch, at most 64 MB).
 *   - work is enqueued on `stream` (a cudaStream_t passed as void*); entry points that return
 *     device results do not synchronise.  The *_host convenience entry points (host buffers
 *     in and out) synchronise before returning.
 *   - return value: TT_OK or an error code; tt_last_error() gives the thread-local message.
 *     No exceptions cross the ABI.  There is no CPU fallback: without a CUDA device every
 *     compute entry point returns TT_ERR_CUDA.
 *   - frame order: a cube with probing axis `par` (0=x, 1=y, 2=z) is stored in the ray frame
 *     (u, v, w) = (t1, t2, par) with (t1, t2) the transverse axes in the reference's `rf`
 *     order: par=z -> (x, y), par=y -> (x, z), par=x -> (y, z)   (particle_tracker.py:353-378).
 *     Gradient grid layout: grid[iw][iv][iu] of 4-vectors (g_u, g_v, g_w, ne/nc) with
 *     g = -1/2 * d(ne/nc)/d(coordinate) in 1/m, i.e. the reference's dndx/c^2
 *     (particle_tracker.py:235-237); 16 B per voxel in TT_F32, 32 B in TT_F64.
 */
#ifndef TT_B200_H
#define TT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TT_B200_ABI_VERSION 1

enum { TT_F32 = 0, TT_F64 = 1 };
enum { TT_OK = 0, TT_ERR_INVALID = 1, TT_ERR_CUDA = 2, TT_ERR_UNSUPPORTED = 3 };

/* per-ray status written by tt_trace (bit flags) */
enum {
    TT_RAY_EXIT_FACE = 1,  /* left through the far face of the probing axis               */
    TT_RAY_EXIT_SIDE = 2,  /* left through another face                                   */
    TT_RAY_TIME_CAP = 4,   /* still inside at c*T = s_max (particle_tracker.py:317)       */
    TT_RAY_MISSED = 8,     /* never entered the cube                                      */
    TT_RAY_GENERAL = 16    /* left the plane-marching fast path (steep or backward ray)   */
};

typedef void* tt_stream_t;

int tt_abi_version(void);
const char* tt_last_error(void);
/* number of CUDA devices visible, or -1 with tt_last_error() set */
int tt_device_count(void);
/* kernels of THIS library launched by this process so far (every successful <<<>>> of its own kernels; CUB's
 * radix-sort passes inside tt_sort_rays are library code and not counted).  Measurement aid: bench.py reports the
 * difference over its timed region as "gpu_launches".  No reference counterpart.                                 */
unsigned long long tt_launch_count(void);

/* Host -> device copy of PAGEABLE host memory (the plain numpy arrays the reference keeps its cube and rays in:
 * particle_tracker.py:212-218 external_ne, :258-310 init_beam -> self.s0) through a ring of pinned buffers filled by
 * worker threads (TT_H2D_THREADS, default: half the CPUs of the caller's affinity mask, at most 8), one DMA per 4 MB
 * piece queued on `stream`; sources under 8 MB: one plain cudaMemcpyAsync.  Same contract as cudaMemcpyAsync from
 * pageable memory: on return the source has been read completely, the copy is ordered on `stream`.  Measured on the
 * B200 box: 48 GB/s with 8 threads against 11 GB/s for the driver's own staging.  Pinned sources do not need it.      */
int tt_h2d_pageable(void* dst_dev, const void* src_host, size_t bytes, tt_stream_t stream);

/* ---- K1: ElectronCube.calc_dndr (particle_tracker.py:220-241) ------------------------------
 * ne_dev: the reference's C-ordered cube ne[ix][iy][iz], float (TT_F32) or double (TT_F64).
 * n_xyz / spacing_xyz: points and (uniform) node spacing per axis.  nc = critical density,
 * ne_max = clip of ne/nc (:231).  Differences follow numpy.gradient: central inside,
 * first-order one-sided on the faces (:235-237).  Output: frame-ordered grid (see top).      */
int tt_calc_dndr(const void* ne_dev, int ne_dtype, const int n_xyz[3],
                 const double spacing_xyz[3], int par, double nc, double ne_max,
                 void* grid4_dev, int grid_dtype, tt_stream_t stream);

/* ---- ElectronCube.dndr (particle_tracker.py:243-256) ----------------------------------------
 * Trilinear gradient at npts positions pos_dev[3][npts] (x, y, z rows, metres); zero outside
 * the cube, faces inclusive.  out_dev[3][npts] in the reference's units (m/s^2, = c^2 * g).   */
int tt_dndr(const void* grid4_dev, int grid_dtype, const int n_xyz[3], const double origin_xyz[3],
            const double spacing_xyz[3], int par, const double* pos_dev, long npts,
            double* out_dev, tt_stream_t stream);

/* ---- ElectronCube.init_beam (particle_tracker.py:258-310), device RNG variant ---------------
 * Same distribution and draw semantics (uniform disc radius beam_size via folded sum of two
 * uniforms, azimuth in [0, pi), Gaussian divergence), Philox4x32-10 counter RNG keyed by
 * (seed, first_ray + i) so that any shard of a beam can be generated on any rank.
 * s0_dev[6][np] rows x, y, z, vx, vy, vz (m, m/s).                                            */
int tt_init_beam(long np, long first_ray, uint64_t seed, double beam_size, double divergence,
                 double extent, int par, double* s0_dev, tt_stream_t stream);

/* ---- Morton order of launch positions (locality for the gather; no reference equivalent) ---
 * perm_dev[i] = index of the i-th ray along a Z-order curve over the transverse launch
 * position.  Workspace from tt_sort_rays_workspace().                                         */
int tt_sort_rays_workspace(long np, size_t* bytes);
int tt_sort_rays(const double* s0_dev, long np, int par, const double origin_xyz[3],
                 const double spacing_xyz[3], const int n_xyz[3], uint32_t* perm_dev,
                 void* workspace_dev, size_t workspace_bytes, tt_stream_t stream);

/* ---- K3+K4: ElectronCube.solve + dsdt + ray_at_exit (particle_tracker.py:312-331, 398-419,
 *      333-380) --------------------------------------------------------------------------------
 * Fixed-step RK4, marching plane to plane along the probing axis (steps_per_cell sub-planes
 * per cell), trilinear gradient with zero outside the cube, rays frozen when they leave, path
 * time capped at s_max = c*T = sqrt(8)*extent.  rf_dev[4][np] = (p1, atan(v1/vpar), p2,
 * atan(v2/vpar)) on the plane par-axis = +extent, SAME ray order as s0 (perm only changes
 * which thread integrates which ray).  sf_dev[6][np] (nullable) = state at time T as the
 * reference's cube.sf.  ray_steps_dev (nullable) is incremented by the number of RK4 steps
 * taken inside the cube.  status_dev[np] (nullable) receives TT_RAY_* flags.                  */
typedef struct tt_trace_params {
    int n_xyz[3];
    double origin_xyz[3];   /* coordinate of node 0 per axis (m)            */
    double spacing_xyz[3];  /* uniform node spacing per axis (m)            */
    int par;                /* probing axis 0/1/2                            */
    double extent;          /* exit plane = +extent, launch plane = -extent */
    double s_max;           /* c*T                                           */
    int steps_per_cell;     /* >= 1                                          */
    int dtype;              /* TT_F32 or TT_F64: grid element type and state arithmetic */
    int variant;            /* 0 = auto (event marching if status_dev is given, else cell cache);
                               1 = 8-corner gather per stage; 2 = cell cache; 3 = event marching
                               (packed FP32x2 arithmetic in TT_F32); 4 = event marching, scalar     */
} tt_trace_params;

int tt_trace(const tt_trace_params* p, const void* grid4_dev, const double* s0_dev, long np,
             const uint32_t* perm_dev, double* rf_dev, double* sf_dev,
             unsigned long long* ray_steps_dev, uint8_t* status_dev, tt_stream_t stream);

/* ---- K2f: the production path of ElectronCube.solve in TT_F32 at 1 step per cell ------------------------------
 * Same reference function as tt_trace (particle_tracker.py:312-331 solve, :398-419 dsdt, :333-380 ray_at_exit), same
 * arguments, outputs, status flags and ray order; the trilinear look-up of dsdt (:243-256) reads a second, derived
 * grid: per cell FACE (iw, cell row, cell column) the bilinear coefficients (A, B, C, D) of the three gradient
 * components, g(tu, tv) = A + tu B + tv (C + tu D) in cell coordinates tu, tv in [-1/2, 1/2], with the step-size
 * factors folded in -- three 16-byte words
 *   (A_u, A_v, B_u, B_v) (C_u, C_v, D_u, D_v) (A_w, C_w, B_w, D_w),   scaled by h_w^2/h_u, h_w^2/h_v, h_w
 * = 48 (nu-1)(nv-1)(nw+1) bytes (tt_face_grid_bytes; the last plane is a spare copy), formed ONCE per calc_dndr by tt_build_face_grid from the float4
 * grid of tt_calc_dndr instead of once per ray and plane inside the trace kernel.  tt_trace_faces marches every ray
 * over it (one RK4 step per cell, steps ending on cell faces as in tt_trace variant 3) and hands unusual rays (outside
 * / steep / side exit / time cap / non-finite) to the same general kernel over grid4_dev.  status_dev is required.
 * p->dtype must be TT_F32 and p->steps_per_cell 1; p->variant is ignored.                                         */
size_t tt_face_grid_bytes(const int n_xyz[3], int par);
int tt_build_face_grid(const void* grid4_dev, const int n_xyz[3], const double spacing_xyz[3], int par,
                       void* faces_dev, tt_stream_t stream);
int tt_trace_faces(const tt_trace_params* p, const void* grid4_dev, const void* faces_dev, const double* s0_dev,
                   long np, const uint32_t* perm_dev, double* rf_dev, double* sf_dev,
                   unsigned long long* ray_steps_dev, uint8_t* status_dev, tt_stream_t stream);

/* ---- magnetised / absorbing extension (BASELINE config 4) --------------------------------------
 * Evidence in the reference: call sites only (example_kitchensink.py:72-101: B_on, inv_brems,
 * phaseshift, external_B/Te/Z, Jf) -- there is no implementation in the checkout, so parity is
 * UNPINNED; textbook forms are used.  Along each ray, with s = c t, d = v/c:
 *   phase     dphi/ds   = (omega/c)(sqrt(1 - ne/nc) - 1)
 *   Faraday   dalpha/ds = verdet * ne * (B . d)        verdet = e^3 lambda^2/(8 pi^2 eps0 me^2 c^3)
 *   inverse bremsstrahlung  dln(a)/ds = -kappa/2       kappa = energy absorption coefficient (1/m)
 * aux4_dev: second grid (B_u, B_v, B_w, kappa) in the gradient grid's layout and dtype (nullable:
 * phase only).  aux_out_dev[3][np] = (amplitude factor, phase [rad], rotation [rad]).  Same
 * trajectory integration, outputs and edge-case handling as tt_trace (gather kernel).            */
typedef struct tt_aux_params {
    double omega;   /* laser angular frequency (rad/s)        */
    double nc;      /* critical density (m^-3)                */
    double verdet;  /* rad / (T m^2): rotation = verdet * int ne B_par ds */
} tt_aux_params;

int tt_trace_aux(const tt_trace_params* p, const tt_aux_params* a, const void* grid4_dev,
                 const void* aux4_dev, const double* s0_dev, long np, const uint32_t* perm_dev,
                 double* rf_dev, double* sf_dev, double* aux_out_dev,
                 unsigned long long* ray_steps_dev, uint8_t* status_dev, tt_stream_t stream);

/* tt_trace_aux in TT_F32 at 1 step per cell over FACE-coefficient grids (the production form of this extension, as
 * tt_trace_faces is for tt_trace): next to the gradient faces of tt_build_face_grid a second grid holds, per cell face,
 * the bilinear coefficients of ne/nc, B_u, B_v, B_w and kappa (five 16-byte words = 80 B per face cell,
 * tt_face_aux_grid_bytes), formed once by tt_build_face_aux_grid from grid4_dev and aux4_dev (nullable: phase only).
 * Same arguments, outputs, flags and ray order as tt_trace_aux; rays the event kernel cannot march are redone by the
 * general kernel over grid4_dev / aux4_dev.  Same evidence upstream (call sites only): parity unpinned.              */
size_t tt_face_aux_grid_bytes(const int n_xyz[3], int par);
int tt_build_face_aux_grid(const void* grid4_dev, const void* aux4_dev, const int n_xyz[3], const double spacing_xyz[3],
                           int par, void* faces_aux_dev, tt_stream_t stream);
int tt_trace_faces_aux(const tt_trace_params* p, const tt_aux_params* a, const void* grid4_dev, const void* aux4_dev,
                       const void* faces_dev, const void* faces_aux_dev, const double* s0_dev, long np,
                       const uint32_t* perm_dev, double* rf_dev, double* sf_dev, double* aux_out_dev,
                       unsigned long long* ray_steps_dev, uint8_t* status_dev, tt_stream_t stream);

/* The aux4_dev grid of tt_trace_aux from the user's cubes in one pass: (B_u, B_v, B_w, kappa) per node in the gradient
 * grid's layout and dtype (grid_dtype).  Reference: only the call sites exist upstream (example_kitchensink.py:72-101:
 * external_B / external_Te / external_Z, B_on / inv_brems); the formula is the NRL-formulary inverse-bremsstrahlung
 * coefficient documented at ElectronCube.kappa(): kappa[1/m] = 100 * 3.1e-7 Z ne^2 lnL Te^-3/2 / (omega^2 sqrt(1 - ne/nc)),
 * ne in cm^-3, Te in eV, ne/nc clipped at ne_max, lnL = coulomb_log or, if that is NaN, max(2, 24 - ln(sqrt(ne)/Te)).
 * ne_dev[ix][iy][iz] (ne_dtype); Te_dev / Z_dev: cubes of aux_dtype or NULL (then Te_scalar / Z_scalar);
 * B_dev[ix][iy][iz][3] (aux_dtype, xyz components) or NULL (B = 0); want_kappa = 0 leaves kappa = 0.               */
int tt_build_aux_grid(const void* ne_dev, int ne_dtype, const void* Te_dev, double Te_scalar, const void* Z_dev,
                      double Z_scalar, const void* B_dev, int aux_dtype, const int n_xyz[3], int par, double nc,
                      double ne_max, double omega, double coulomb_log, int want_kappa, void* aux4_dev, int grid_dtype,
                      tt_stream_t stream);

/* ---- rectilinear grids: axes whose nodes are NOT equally spaced ------------------------------------
 * The reference takes any ascending coordinate arrays: numpy.gradient(ne_nc, x, axis=0) uses the
 * non-uniform second-order stencil and RegularGridInterpolator((x, y, z), ...) locates cells by
 * bisection (particle_tracker.py:235-241).  These entry points are the same three operations with the
 * node coordinates x_dev[nx], y_dev[ny], z_dev[nz] (device, FP64, strictly ascending) instead of
 * origin + spacing; grid layout, outputs, status flags and ray order are those of tt_calc_dndr /
 * tt_trace / tt_dndr.  tt_trace_axes ignores p->origin_xyz and p->spacing_xyz.  p->variant: 0 = auto /
 * 3 = event marching in index space with the sizes of the current cell (needs status_dev; every RK4 step
 * inside one cell, corners loaded once per cell; state and arithmetic in the grid's element type, launch
 * and exit in FP64), followed by the gather kernel on the rays it defers (outside / steep / side exit /
 * time cap); 1, 2 = the gather kernel alone (FP64 arithmetic whatever the grid's type, cell search by
 * walking at every stage). */
int tt_calc_dndr_axes(const void* ne_dev, int ne_dtype, const int n_xyz[3], const double* x_dev,
                      const double* y_dev, const double* z_dev, int par, double nc, double ne_max,
                      void* grid4_dev, int grid_dtype, tt_stream_t stream);
int tt_trace_axes(const tt_trace_params* p, const double* x_dev, const double* y_dev, const double* z_dev,
                  const void* grid4_dev, const double* s0_dev, long np, const uint32_t* perm_dev,
                  double* rf_dev, double* sf_dev, unsigned long long* ray_steps_dev, uint8_t* status_dev,
                  tt_stream_t stream);
int tt_dndr_axes(const void* grid4_dev, int grid_dtype, const int n_xyz[3], const double* x_dev,
                 const double* y_dev, const double* z_dev, int par, const double* pos_dev, long npts,
                 double* out_dev, tt_stream_t stream);

/* ---- K5+K6: ray_transfer_matrix.py optics (:37-154), detector programs (:208-299) and
 *      Rays.histogram (:173-195) ---------------------------------------------------------------
 * One pass over the rays: scale positions (pos_scale = 1e3 is m_to_mm, :37-40), run the element
 * program, bin the survivors like numpy.histogram2d (right-open bins, last edge inclusive,
 * NaN dropped), H_dev[nby][nbx] (already transposed as :191) accumulated as uint64 counts.
 * rf_out_dev[4][np] nullable.  Rejected rays are NaN in all four rows (:78).                   */
enum {
    TT_OP_DISTANCE = 0,      /* a = d                          distance()          :62-71   */
    TT_OP_LENS = 1,          /* a = f1, b = f2                 lens()/sym_lens()   :42-60   */
    TT_OP_CIRC_APERTURE = 2, /* a = R   reject r^2 >  R^2      circular_aperture() :73-79   */
    TT_OP_CIRC_STOP = 3,     /* a = R   reject r^2 <  R^2      circular_stop()     :81-87   */
    TT_OP_ANNULAR_STOP = 4,  /* a = R1, b = R2 reject between  annular_stop() as used by
                                                               angular_filter()    :89-111  */
    TT_OP_RECT_APERTURE = 5, /* a = Lx, b = Ly (both outside)  rect_aperture()     :128-136 */
    TT_OP_KNIFE_EDGE = 6     /* a = offset, b = +-1 (x) or +-2 (y): sign = direction :138-154 */
};
typedef struct tt_optic {
    int op;
    int pad_;
    double a, b;
} tt_optic;

#define TT_MAX_OPTICS 64

int tt_optics_hist(const double* rf_in_dev, long np, double pos_scale, const tt_optic* program,
                   int n_ops, const double* xedges_dev, int nbx, const double* yedges_dev, int nby,
                   unsigned long long* H_dev, double* rf_out_dev, tt_stream_t stream);
/* same, visiting the rays in the order perm_dev[0..np) (the Morton order of tt_sort_rays), so
 * that a CTA's rays fall into a compact detector patch; rf_out keeps the original ray order. */
int tt_optics_hist_perm(const double* rf_in_dev, long np, const uint32_t* perm_dev, double pos_scale,
                        const tt_optic* program, int n_ops, const double* xedges_dev, int nbx,
                        const double* yedges_dev, int nby, unsigned long long* H_dev,
                        double* rf_out_dev, tt_stream_t stream);

/* same, plus a weighted image Hw_dev[nby][nbx] (double) += weights_dev[ray], i.e.
 * numpy.histogram2d(..., weights=w) as example_kitchensink.py:108-129 uses for amplitude- and
 * polarisation-weighted images.  H_dev may be NULL when only the weighted image is wanted.        */
int tt_optics_hist_weighted(const double* rf_in_dev, long np, const uint32_t* perm_dev, double pos_scale,
                            const tt_optic* program, int n_ops, const double* xedges_dev, int nbx,
                            const double* yedges_dev, int nby, unsigned long long* H_dev,
                            const double* weights_dev, double* Hw_dev, double* rf_out_dev,
                            tt_stream_t stream);

/* ---- K7: turboGen.gaussian3D_FFT (gaussian_fields/turboGen.py:488-538) -----------------------
 * Hermitian spectrum shaping + inverse real FFT (cuFFT) on the odd grid M = 2N+1.
 * sqrtP_lut_dev[3N^2+1]: sqrt(k_func(sqrt(q)/M)) for q = i^2+j^2+l^2 (the host evaluates the
 * user's Python k_func once per distinct |k|).  Wr_dev/Wi_dev: the two N(0,1) cubes of
 * :522-523 (M^3 doubles each) for bit-comparable runs, or both NULL to draw them from
 * Philox(seed).  out_dev: M^3 reals (dtype), = ifftn(F).real of :536-538.  Workspace from
 * tt_grf_workspace().                                                                          */
int tt_grf_workspace(int N, int dtype, size_t* bytes);
int tt_grf3d(int N, int dtype, const double* sqrtP_lut_dev, const double* Wr_dev,
             const double* Wi_dev, uint64_t seed, void* out_dev, void* workspace_dev,
             size_t workspace_bytes, tt_stream_t stream);

/* 1-D / 2-D / 3-D variants (gaussian1D_FFT :388-434, gaussian2D_FFT :436-486): ndim = 1, 2, 3; arrays
 * have shape (2N+1,)*ndim; the table needs ndim*N^2+1 entries.                                    */
int tt_grf_nd_workspace(int ndim, int N, int dtype, size_t* bytes);
int tt_grf_nd(int ndim, int N, int dtype, const double* sqrtP_lut_dev, const double* Wr_dev,
              const double* Wi_dev, uint64_t seed, void* out_dev, void* workspace_dev,
              size_t workspace_bytes, tt_stream_t stream);

/* ---- spectrum diagnostic: calculate_spectrum_3d.spectrum_3D_scalar
 *      (gaussian_fields/calculate_spectrum_3d.py:3-59) ----------------------------------------------
 * Shell-averaged power |FFT(data)|^2 of a real cube data_dev[nx][ny][nz] (TT_F32 / TT_F64): one R2C
 * cuFFT + one reduction over the half spectrum.  Shell i holds (i)*w <= |k| < (i+1)*w with
 * w = k_max / k_bin_num and |k| built from numpy.fft.fftfreq(n, dx); shells 0 .. k_bin_num-2 are
 * filled (the reference's loop stops there).  sum_dev / count_dev [k_bin_num]: sum of |F|^2 and number
 * of modes per shell (the caller divides).  The input is not modified.                              */
int tt_spectrum3d_workspace(const int n_xyz[3], int dtype, size_t* bytes);
int tt_spectrum3d(const void* data_dev, int dtype, const int n_xyz[3], double dx, double k_max,
                  int k_bin_num, double* sum_dev, unsigned long long* count_dev, void* workspace_dev,
                  size_t workspace_bytes, tt_stream_t stream);

/* ---- host-buffer convenience entry point (what a ctypes binding inside the reference calls) --
 * Whole path for one bundle of rays with HOST arrays: ne (C order, double) -> gradient grid ->
 * Morton sort -> trace -> rf (host, 4 x np doubles).  Allocates and frees its device scratch
 * internally, synchronises.  Same numerics as the device entry points above.                   */
int tt_solve_host(const double* ne_host, const int n_xyz[3], const double origin_xyz[3],
                  const double spacing_xyz[3], int par, double nc, double ne_max, double extent,
                  int steps_per_cell, int dtype, const double* s0_host, long np, double* rf_host,
                  double* sf_host, unsigned long long* ray_steps_host);

#ifdef __cplusplus
}
#endif
#endif /* TT_B200_H */
