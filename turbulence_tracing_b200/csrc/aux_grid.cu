// tt_build_aux_grid: the second node grid (B_u, B_v, B_w, kappa) of the magnetised / absorbing extension (SURVEY 8f1,
// BASELINE configs[3]) in ONE streaming pass over the user's cubes.
//
// The reference holds only the call sites (example_kitchensink.py:72-101: external_B / external_Te / external_Z,
// B_on / inv_brems / phaseshift); there the absorption coefficient and the field are interpolated like the gradient.
// Round 1 built this grid with eager tensor operations (kappa() formula: ~12 whole-cube FP64 temporaries, then stack /
// permute / cast / contiguous: four more): ~35 GB of temporaries at 1025^3.  Here every node is formed in registers:
// reads ne (+ Te, Z, B where they are cubes) once, writes 16 / 32 B once.
//
// Thread = one INPUT node, z fastest: the reads are coalesced; the write is one aligned 16- or 32-byte word at the
// node's frame position (u is never z, so writes of neighbouring threads are nu * 16 B apart -- a set-up kernel: every
// sector written carries 16 useful bytes of 32 in FP32, all of them in FP64).
#include "common.cuh"
#include "calc_dndr_tile.cuh"        // Vec4

namespace tt {

struct AuxGridArgs {
    long long total;       // nx ny nz
    int n[3];
    int fa[3];             // frame: (u, v, w) -> xyz axis
    double nc, ne_max, omega2;
    double lnL;            // Coulomb logarithm, or NaN: max(2, 24 - ln(sqrt(ne[cm^-3]) / Te[eV])) per node
    double Te_s, Z_s;      // used where the cube pointer is null
    int want_B, want_kappa;
};

// NRL formulary, inverse bremsstrahlung: kappa[1/m] = 100 * 3.1e-7 Z ne^2 lnL Te^-3/2 omega^-2 (1 - ne/nc)^-1/2, ne in cm^-3,
// Te in eV -- evaluated in the order ElectronCube.kappa() documents (FP64 throughout)
TT_HD double kappa_node(double ne, double Te, double Z, const AuxGridArgs& a) {
    const double ne_cc = ne * 1e-6;
    double lnL = a.lnL;
    if (!(lnL == lnL)) {
        const double c = ne_cc > 1e-30 ? ne_cc : 1e-30;
        lnL = 24.0 - log(sqrt(c) / Te);
        lnL = lnL > 2.0 ? lnL : 2.0;
    }
    double ne_nc = ne / a.nc;
    ne_nc = ne_nc < a.ne_max ? ne_nc : a.ne_max;
    double d = 1.0 - ne_nc;
    d = d > 1e-6 ? d : 1e-6;
    const double disp = sqrt(d);
    return (100.0 * 3.1e-7) * Z * (ne_cc * ne_cc) * lnL * pow(Te, -1.5) / (a.omega2 * disp);
}

template <typename TNe, typename TA, typename TOut>
__global__ void __launch_bounds__(256)
aux_grid_kernel(const TNe* __restrict__ ne, const TA* __restrict__ Te, const TA* __restrict__ Z, const TA* __restrict__ B,
                typename Vec4<TOut>::type* __restrict__ out, AuxGridArgs a) {
    typedef typename Vec4<TOut>::type V4;
    const long long plane_yz = (long long)a.n[1] * a.n[2];
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.total; i += (long long)gridDim.x * blockDim.x) {
        int idx[3];
        idx[0] = (int)(i / plane_yz);
        const long long r = i - (long long)idx[0] * plane_yz;
        idx[1] = (int)(r / a.n[2]);
        idx[2] = (int)(r - (long long)idx[1] * a.n[2]);
        double b[3] = {0.0, 0.0, 0.0}, kap = 0.0;
        if (a.want_B) {
            b[0] = (double)B[3 * i + a.fa[0]]; b[1] = (double)B[3 * i + a.fa[1]]; b[2] = (double)B[3 * i + a.fa[2]];
        }
        if (a.want_kappa)
            kap = kappa_node((double)ne[i], Te ? (double)Te[i] : a.Te_s, Z ? (double)Z[i] : a.Z_s, a);
        const size_t o = ((size_t)idx[a.fa[2]] * a.n[a.fa[1]] + idx[a.fa[1]]) * a.n[a.fa[0]] + idx[a.fa[0]];
        V4 v;
        v.x = (TOut)b[0]; v.y = (TOut)b[1]; v.z = (TOut)b[2]; v.w = (TOut)kap;
        out[o] = v;
    }
}

template <typename TNe, typename TA, typename TOut>
static int launch_aux_grid(const void* ne, const void* Te, const void* Z, const void* B, void* out, const AuxGridArgs& a,
                           cudaStream_t s) {
    long long blocks = (a.total + 255) / 256;
    if (blocks > 148LL * 32) blocks = 148LL * 32;
    aux_grid_kernel<TNe, TA, TOut><<<(unsigned)blocks, 256, 0, s>>>((const TNe*)ne, (const TA*)Te, (const TA*)Z, (const TA*)B,
                                                                  (typename Vec4<TOut>::type*)out, a);
    return launch_check("aux_grid_kernel");
}

}  // namespace tt

extern "C" int tt_build_aux_grid(const void* ne_dev, int ne_dtype, const void* Te_dev, double Te_scalar, const void* Z_dev,
                                 double Z_scalar, const void* B_dev, int aux_dtype, const int n_xyz[3], int par, double nc,
                                 double ne_max, double omega, double coulomb_log, int want_kappa, void* aux4_dev,
                                 int grid_dtype, tt_stream_t stream) {
    using namespace tt;
    TT_REQUIRE(aux4_dev && n_xyz, "tt_build_aux_grid: null pointer");
    TT_REQUIRE(par >= 0 && par <= 2, "tt_build_aux_grid: par must be 0, 1 or 2 (got %d)", par);
    TT_REQUIRE((ne_dtype == TT_F32 || ne_dtype == TT_F64) && (aux_dtype == TT_F32 || aux_dtype == TT_F64) &&
               (grid_dtype == TT_F32 || grid_dtype == TT_F64), "tt_build_aux_grid: dtype must be TT_F32 or TT_F64");
    TT_REQUIRE(B_dev || want_kappa, "tt_build_aux_grid: nothing to build (no B cube, no absorption)");
    AuxGridArgs a;
    a.total = 1;
    for (int i = 0; i < 3; ++i) {
        TT_REQUIRE(n_xyz[i] >= 2, "tt_build_aux_grid: every axis needs >= 2 points");
        a.n[i] = n_xyz[i];
        a.total *= n_xyz[i];
    }
    const Frame f = frame_of(par);
    for (int i = 0; i < 3; ++i) a.fa[i] = f.a[i];
    a.want_B = B_dev != nullptr;
    a.want_kappa = want_kappa != 0;
    if (a.want_kappa) {
        TT_REQUIRE(ne_dev, "tt_build_aux_grid: absorption needs the ne cube");
        TT_REQUIRE(nc > 0 && omega > 0, "tt_build_aux_grid: critical density and omega must be > 0");
        TT_REQUIRE(Te_dev || Te_scalar > 0, "tt_build_aux_grid: absorption needs Te > 0 (cube or scalar)");
    }
    a.nc = nc; a.ne_max = ne_max; a.omega2 = omega * omega; a.lnL = coulomb_log; a.Te_s = Te_scalar; a.Z_s = Z_scalar;
    cudaStream_t s = (cudaStream_t)stream;
#define TT_AUX_CASE(TNe, TA, TOut) return launch_aux_grid<TNe, TA, TOut>(ne_dev, Te_dev, Z_dev, B_dev, aux4_dev, a, s)
    const bool ne32 = ne_dtype == TT_F32, a32 = aux_dtype == TT_F32, o32 = grid_dtype == TT_F32;
    if (ne32 && a32 && o32) TT_AUX_CASE(float, float, float);
    if (ne32 && a32 && !o32) TT_AUX_CASE(float, float, double);
    if (ne32 && !a32 && o32) TT_AUX_CASE(float, double, float);
    if (ne32 && !a32 && !o32) TT_AUX_CASE(float, double, double);
    if (!ne32 && a32 && o32) TT_AUX_CASE(double, float, float);
    if (!ne32 && a32 && !o32) TT_AUX_CASE(double, float, double);
    if (!ne32 && !a32 && o32) TT_AUX_CASE(double, double, float);
    TT_AUX_CASE(double, double, double);
#undef TT_AUX_CASE
}
