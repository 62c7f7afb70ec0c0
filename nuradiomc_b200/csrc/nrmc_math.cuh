// nrmc_math.cuh -- scalar FP64 ray maths of the B200 analytic ray tracer (device code; also compiles for the
// host so that tests/cpu_harness can exercise exactly the same arithmetic without a GPU -- test-only, the product
// never runs it on the CPU).
//
// Reference behaviour being reproduced (NuRadioMC, /root/reference/NuRadioMC/SignalProp/analyticraytracing.py):
//   get_delta_y :204-272, find_solutions :1400-1547 (2-D) and :2118-2130 (mode loop), determine_solution_type
//   :1365-1398, get_path_segments :1091-1159, get_angle :1161-1199, get_reflection_angle :1201-1237,
//   get_path_length_analytic :602-690, get_travel_time_analytic :692-783, get_C_1 :487-491.
//
// This is NOT a translation.  The reference root-finds the miss distance delta_y(logC0) with hybr + Brent.  Here the
// problem is re-posed on Snell's invariant beta = n(z) sin(theta) = 1/C0:
//
//   y(z) - y(z') = beta/sqrt(c) * (U(z) - U(z')),   U(z) = z - z0 ln k1(z),   c = n_ice^2 - beta^2,
//   k1(z) = sqrt(c) s(z) + c - n_ice gamma(z),      s(z) = sqrt(n(z)^2 - beta^2),  gamma(z) = dn exp(z/z0),
//   U at the turning point: U_T = -z0 ln K_T,  K_T = k1(0) (surface reflection, beta <= n_s) or dn*beta (apex).
//
// so the horizontal range of every path family (n bottom bounces, launch up/down, arrival before/after the turning
// point) is one logarithm:  R(beta) = beta/sqrt(c) * [a1 z1 + a2 z2 + ar zr - z0 ln(k1(z1)^a1 k1(z2)^a2 k1(zr)^ar K_T^aT)]
// with small integer exponents (mode_coeffs).  R is traced along a closed curve of four smooth pieces, all of them
// parametrised by the rational parameter t of a circle beta^2 + sigma^2 = nX^2 (beta = nX 2t/(1+t^2), sigma = nX (1-t^2)/(1+t^2)):
//   P0: arrival up-going,   beta in (0, n_s]   nX = n_s, sigma = s(0),  t: 0 -> 1
//   P1: arrival up-going,   beta in [n_s, n2]  nX = n2,  sigma = s(z2), t: tmin -> 1
//   P2: arrival down-going, beta in [n_s, n2]  (refracted: apex inside the ice), t: 1 -> tmin
//   P3: arrival down-going, beta in (0, n_s]   (reflected off the surface),      t: 1 -> 0
// R starts and ends at 0 and is unimodal along the curve, so R = rho has 0 or 2 roots per mode: the three junction
// values bracket them; when all junctions are below rho the maximum is searched in the two pieces next to the
// largest junction.  All radicands are an offset assembled from differences of gamma's plus sigma^2 (no cancellation
// against n_ice), so ONE code path evaluates every piece, and dR/dt comes out of the same pass for ~25 % more
// arithmetic: roots are closed with safeguarded Newton steps (3-4 evaluations from straight-line starting points).
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define NRMC_HD __host__ __device__ __forceinline__
#define NRMC_HDN __host__ __device__ __noinline__
#else
#define NRMC_HD inline
#define NRMC_HDN inline
#endif

#define NRMC_MAX_REFLECTIONS 4
#define NRMC_SPEED_OF_LIGHT 0.299792458 /* m/ns */

namespace nrmc {

// Reciprocal, reciprocal square root and logarithm of the solver's inner loop.  On the device they are the fast paths only:
// CUDA's 1/x, rsqrt and log guard every call with an exponent-range test, a convergence barrier and a slow-path call, and
// log materialises its polynomial with two UMOVs per coefficient; the arguments here are positive normal numbers (sums of
// squares, the k factors of the range function), so the hardware seed (MUFU.RCP64H / RSQ64H, 2^-23) plus Newton steps is all
// that is needed, and log keeps its coefficients in constant memory (fdlibm's reduction and degree-14 odd polynomial,
// < 1 ulp).  Zero, inf and NaN (a horizontal ray in numerically homogeneous ice) give the IEEE results (inf, 0, NaN), so
// they propagate as before; denormals count as zero.  NRMC_STOCK_MATH: CUDA's functions everywhere (A/B builds).
#if defined(__CUDACC__)
__constant__ double c_logc[9] = {1.479819860511658591e-01, 1.531383769920937332e-01, 1.818357216161805012e-01, 2.222219843214978396e-01,
                                 2.857142874366239149e-01, 3.999999999940941908e-01, 6.666666666666735130e-01,
                                 6.93147180369123816490e-01, 1.90821492927058770002e-10};   // Lg7 .. Lg1, ln2_hi, ln2_lo
#endif
#if defined(__CUDA_ARCH__) && !defined(NRMC_STOCK_MATH)
__device__ __forceinline__ bool nrmc_is_normal(double x)            // finite, non-zero, not denormal (either sign)
{
    return (unsigned)((__double2hiint(x) & 0x7ff00000) - 0x00100000) < 0x7fe00000u;
}
__device__ __forceinline__ double nrmc_rcp(double x)
{
    double y0;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(x));          // 0 -> inf, inf -> 0, NaN -> NaN: returned as they are
    // branch-free: the refinement runs on every input and is discarded (two selects) where it would turn inf into NaN --
    // a divergent branch with its convergence barrier costs more issue slots than the four FMAs
    double e = fma(-x, y0, 1.0);
    double y = fma(y0, e, y0);
    e = fma(-x, y, 1.0);
    y = fma(y, e, y);
    return nrmc_is_normal(x) ? y : y0;
}
__device__ __forceinline__ double nrmc_rsqrt(double x)
{
    double y0;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(x));        // 0 -> inf, inf -> 0, negative / NaN -> NaN
    double e = fma(-(x * y0), y0, 1.0);                             // 1 - x y^2
    double y = fma(fma(0.375, e, 0.5) * e, y0, y0);                 // third order: y (1 + e/2 + 3 e^2/8)
    e = fma(-(x * y), y, 1.0);
    y = fma(0.5 * e, y, y);
    return (unsigned)(__double2hiint(x) - 0x00100000) < 0x7fe00000u ? y : y0;     // positive normal numbers only (branch-free)
}
__device__ __forceinline__ double nrmc_log(double x)
{
    int hi = __double2hiint(x);
    if ((unsigned)(hi - 0x00100000) >= 0x7fe00000u)                 // not a positive normal number
        return (x != x || x < 0.0) ? NAN : (x > 1.0 ? INFINITY : -INFINITY);   // (denormals count as 0)
    hi += 0x3ff00000 - 0x3fe6a09e;                                  // mantissa into [sqrt(1/2), sqrt(2))
    const double dk = (double)((hi >> 20) - 0x3ff);
    const double m = __hiloint2double((hi & 0x000fffff) + 0x3fe6a09e, __double2loint(x));
    const double f = m - 1.0;
    const double s = f * nrmc_rcp(2.0 + f);
    const double z = s * s, w = z * z;
    const double t1 = w * fma(w, fma(w, c_logc[1], c_logc[3]), c_logc[5]);
    const double t2 = z * fma(w, fma(w, fma(w, c_logc[0], c_logc[2]), c_logc[4]), c_logc[6]);
    const double hfsq = 0.5 * f * f;
    return fma(dk, c_logc[7], -((hfsq - fma(s, hfsq + (t2 + t1), dk * c_logc[8])) - f));
}
#define NRMC_RCP(x) nrmc_rcp(x)
#define NRMC_RSQRT(x) nrmc_rsqrt(x)
// the same for arguments KNOWN to be positive normal numbers (1 + t^2, n_ice^2 - beta^2, radicands clamped from below): no selects
__device__ __forceinline__ double nrmc_rcp_pn(double x)
{
    double y0;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(x));
    double e = fma(-x, y0, 1.0);
    double y = fma(y0, e, y0);
    e = fma(-x, y, 1.0);
    return fma(y, e, y);
}
__device__ __forceinline__ double nrmc_rsqrt_pn(double x)
{
    double y0;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(x));
    double e = fma(-(x * y0), y0, 1.0);
    double y = fma(fma(0.375, e, 0.5) * e, y0, y0);
    e = fma(-(x * y), y, 1.0);
    return fma(0.5 * e, y, y);
}
#define NRMC_RCP_PN(x) nrmc_rcp_pn(x)
#define NRMC_RSQRT_PN(x) nrmc_rsqrt_pn(x)
#define NRMC_LOG(x) nrmc_log(x)
// sqrt(x) for x >= 0 (0 -> 0) through the fast reciprocal square root
__device__ __forceinline__ double nrmc_sqrt_pos(double x) { const double r = x * nrmc_rsqrt(x); return x > 0.0 ? r : 0.0; }
#define NRMC_SQRT(x) nrmc_sqrt_pos(x)
#elif defined(__CUDA_ARCH__)
#define NRMC_RCP(x) (1.0 / (x))
#define NRMC_RSQRT(x) rsqrt(x)
#define NRMC_RCP_PN(x) (1.0 / (x))
#define NRMC_RSQRT_PN(x) rsqrt(x)
#define NRMC_LOG(x) log(x)
#define NRMC_SQRT(x) sqrt(x)
#else
#define NRMC_RCP(x) (1.0 / (x))
#define NRMC_RSQRT(x) (1.0 / sqrt(x))
#define NRMC_RCP_PN(x) (1.0 / (x))
#define NRMC_RSQRT_PN(x) (1.0 / sqrt(x))
#define NRMC_LOG(x) log(x)
#define NRMC_SQRT(x) sqrt(x)
#endif
// max / min of two numbers where neither is NaN or the NaN may propagate either way: one compare and a select.  (fmax / fmin
// canonicalise NaNs: DSETP + two moves + SEL + FSEL + LOP3 on sm_100, where FP64 min / max have no instruction.)
#define NRMC_MAX(a, b) ((a) > (b) ? (a) : (b))
#define NRMC_MIN(a, b) ((a) < (b) ? (a) : (b))

struct IceParams {            // medium_base.py:206-252 (+ add_reflective_bottom :47-66)
    double n_ice, dn, z0, inv_z0, inv_dn;
    double ns;                // n(0) = n_ice - dn
    double zr;                // reflective layer depth (<0) ; valid if n_refl > 0
    double gr, nr;            // gamma(zr), n(zr)
    int32_t n_refl;           // number of bottom reflections to consider
    int32_t att_model;        // 0 none, 1 SP1, 2 GL1, 3 MB1, 4 GL2, 5 GL3 (attenuation.py:14)
};

struct PairGeom {             // 2-D problem after set_start_and_end_point (py:2057-2090): z1 <= z2
    double z1, z2, rho;
    double g1, g2, n1, n2;    // gamma and n at z1, z2
    // radicand offsets (see ray_state)
    double A1, A2, Ar;        // (dn - gamma_i)(n_i + ns):  s_i^2 = A_i + sigma^2   ("sub" pieces,  sigma = s(0))
    double B1, Br;            // (g2 - gamma_i)(n_i + n2):  s_i^2 = B_i + sigma^2   ("band" pieces, sigma = s(z2))
    double c0_sub, c0_band;   // c = c0 + sigma^2
    double s2max;             // sqrt(n2^2 - ns^2)
    double tmin;              // band pieces: t in [tmin, 1]  <->  beta in [ns, n2]
};

struct ModeCoeffs { int aT, a1, a2, ar; };

// integer exponents of the range function for (reflection k, launch case, arrival after turning point)
NRMC_HD ModeCoeffs mode_coeffs(int k, int rcase, bool turned)
{
    ModeCoeffs m;
    if (k == 0) { m.aT = turned ? 2 : 0; m.a1 = -1; m.a2 = turned ? -1 : 1; m.ar = 0; }
    else if (rcase == 1) { m.aT = turned ? 2 * k + 2 : 2 * k; m.a1 = -1; m.a2 = turned ? -1 : 1; m.ar = -2 * k; }
    else { m.aT = turned ? 2 * k : 2 * k - 2; m.a1 = 1; m.a2 = turned ? -1 : 1; m.ar = -2 * k; }
    return m;
}

NRMC_HD double ipow(double x, int n) { double r = 1.0; for (int i = 0; i < n; ++i) r *= x; return r; }

// pair geometry from the depths and the two gammas (gamma = dn exp(z/z0) is the only transcendental in it)
NRMC_HD void make_pair_geom_g(const IceParams &ice, double z1, double z2, double rho, double g1, double g2, PairGeom &g)
{
    g.z1 = z1; g.z2 = z2; g.rho = rho;
    g.g1 = g1; g.g2 = g2;
    g.n1 = ice.n_ice - g.g1; g.n2 = ice.n_ice - g.g2;
    g.A1 = (ice.dn - g.g1) * (g.n1 + ice.ns);
    g.A2 = (ice.dn - g.g2) * (g.n2 + ice.ns);
    g.Ar = (ice.dn - ice.gr) * (ice.nr + ice.ns);
    g.B1 = (g.g2 - g.g1) * (g.n1 + g.n2);
    g.Br = (g.g2 - ice.gr) * (ice.nr + g.n2);
    g.c0_sub = ice.dn * (ice.n_ice + ice.ns);
    g.c0_band = g.g2 * (ice.n_ice + g.n2);
    g.s2max = NRMC_SQRT(NRMC_MAX(g.A2, 0.0));
    g.tmin = ice.ns * NRMC_RCP(g.n2 + g.s2max);
}

NRMC_HD void make_pair_geom(const IceParams &ice, double z1, double z2, double rho, PairGeom &g)
{
    make_pair_geom_g(ice, z1, z2, rho, ice.dn * exp(z1 * ice.inv_z0), ice.dn * exp(z2 * ice.inv_z0), g);
}

// Both piece classes use the rational parametrisation of the circle beta^2 + sigma^2 = nX^2:
//   beta = nX 2t/(1+t^2),  sigma = nX (1-t^2)/(1+t^2),  t in (0, 1],
// "sub" (P0, P3): nX = n_s, sigma = s(0);  "band" (P1, P2): nX = n2, sigma = s(z2), t in [tmin, 1].
// Every other radicand is an offset plus sigma^2, so one code path evaluates both classes (no divergence).
struct PieceConsts { double nX, c0, O1, O2, Or; bool band; };

NRMC_HD PieceConsts piece_consts(const IceParams &ice, const PairGeom &g, bool band)
{
    PieceConsts pc;
    pc.band = band;
    pc.nX = band ? g.n2 : ice.ns;
    pc.c0 = band ? g.c0_band : g.c0_sub;
    pc.O1 = band ? g.B1 : g.A1;
    pc.O2 = band ? 0.0 : g.A2;
    pc.Or = band ? g.Br : g.Ar;
    return pc;
}

// Quantities of one ray (one value of beta) that the property formulas need.
struct RayState {
    double beta, rc;          // Snell invariant, sqrt(n_ice^2 - beta^2)
    double s1, s2, sr, ss;    // n cos(theta) at z1, z2, zr and at the surface (ss = 0 for refracted rays)
    double k1_1, k1_2, k1_r, KT;
    bool reflected;           // turning point is the surface (beta <= n_s)
};

// Evaluate the ray state for the point t of a piece of class `band`.
NRMC_HD void ray_state(const IceParams &ice, const PairGeom &g, bool band, double t, RayState &r)
{
    const PieceConsts pc = piece_consts(ice, g, band);
    const double q = NRMC_RCP(1.0 + t * t);
    r.beta = pc.nX * (2.0 * t) * q;
    const double sig = pc.nX * ((1.0 - t) * (1.0 + t)) * q;
    const double sg2 = sig * sig;
    const double c = pc.c0 + sg2;
    r.ss = band ? 0.0 : sig;
    r.s1 = NRMC_SQRT(pc.O1 + sg2);
    r.s2 = band ? sig : NRMC_SQRT(pc.O2 + sg2);
    r.sr = ice.n_refl > 0 ? NRMC_SQRT(pc.Or + sg2) : 0.0;
    r.reflected = !band;
    r.rc = NRMC_SQRT(c);
    r.k1_1 = r.rc * r.s1 + (c - ice.n_ice * g.g1);
    r.k1_2 = r.rc * r.s2 + (c - ice.n_ice * g.g2);
    r.k1_r = ice.n_refl > 0 ? r.rc * r.sr + (c - ice.n_ice * ice.gr) : 1.0;
    r.KT = r.reflected ? r.rc * r.ss + (c - ice.n_ice * ice.dn) : ice.dn * r.beta;
}

// parameter t of the ray with invariant beta inside a class with radius nX (beta < nX)
NRMC_HD double t_of_beta(double nX, double beta) { return beta * NRMC_RCP(nX + NRMC_SQRT(NRMC_MAX((nX - beta) * (nX + beta), 0.0))); }

struct Curve {                // one (reflection, case) mode of one pair
    const IceParams *ice;
    const PairGeom *g;
    int k, rcase;
};

// g = R - rho on piece p in {0,1,2,3} at parameter t and (WITH_D) dg/dt, in ONE pass.  With h = sigma sigma',
//   d k1_i = h (s_i/rc + rc/s_i + 2),  d ln P = sum a_i d k_i / k_i  (one shared reciprocal for k = 0),
//   R = A B,  A = beta/rc,  B = lin - z0 ln P.
template <bool WITH_D>
NRMC_HD double curve_eval(const Curve &cv, int p, double t, double &dg)
{
#ifdef NRMC_COUNT_EVALS
    ++g_evals;
#endif
    const IceParams &ice = *cv.ice;
    const PairGeom &g = *cv.g;
    const bool band = (p == 1 || p == 2), turned = (p >= 2);
    const PieceConsts pc = piece_consts(ice, g, band);
    const double q = NRMC_RCP_PN(1.0 + t * t);
    const double beta = pc.nX * (2.0 * t) * q;
    const double sig = pc.nX * ((1.0 - t) * (1.0 + t)) * q;
    const double sg2 = sig * sig;
    const double c = pc.c0 + sg2;                       // n_ice^2 - beta^2 >= n_ice^2 - n(z2)^2 > 0
    const double irc = NRMC_RSQRT_PN(c), rc = c * irc;
    const double x1 = pc.O1 + sg2, x2 = pc.O2 + sg2;
    const double is1 = NRMC_RSQRT_PN(NRMC_MAX(x1, 1e-300)), is2 = NRMC_RSQRT_PN(NRMC_MAX(x2, 1e-300));
    const double s1 = x1 * is1, s2 = band ? sig : x2 * is2;
    const double k1 = rc * s1 + (c - ice.n_ice * g.g1), k2 = rc * s2 + (c - ice.n_ice * g.g2);
    const double KT = band ? ice.dn * beta : rc * sig + (c - ice.n_ice * ice.dn);
    const double A = beta * irc;
    // derivatives with respect to t
    double bp = 0, sp = 0, h = 0, drc = 0, dc = 0, dk1 = 0, dk2 = 0, dKT = 0;
    if (WITH_D) {
        bp = 2.0 * sig * q; sp = -2.0 * beta * q; h = sig * sp;
        drc = h * irc; dc = 2.0 * h;
        dk1 = drc * s1 + rc * (sp * (sig * is1)) + dc;
        dk2 = drc * s2 + rc * (band ? sp : sp * (sig * is2)) + dc;
        dKT = band ? ice.dn * bp : drc * sig + rc * sp + dc;
    }
    double P, lin, dlnP = 0.0;
    if (cv.k == 0) {
        const double Kx = turned ? KT : 1.0;
        const double iv = NRMC_RCP(k1 * k2 * Kx);
        P = (turned ? KT * KT * KT : k2 * k2) * iv;
        lin = turned ? -g.z1 - g.z2 : g.z2 - g.z1;
        if (WITH_D) dlnP = ((turned ? 2.0 * dKT * k1 * k2 - dk2 * k1 * KT : dk2 * k1) - dk1 * k2 * Kx) * iv;
    } else {
        const ModeCoeffs m = mode_coeffs(cv.k, cv.rcase, turned);
        const double xr = pc.Or + sg2, isr = NRMC_RSQRT_PN(NRMC_MAX(xr, 1e-300)), sr = xr * isr;
        const double kr = rc * sr + (c - ice.n_ice * ice.gr);
        double num = 1.0, den = 1.0;
        if (m.a1 > 0) num *= k1; else den *= k1;            // |a1| == 1
        if (m.a2 > 0) num *= k2; else den *= k2;            // |a2| == 1
        num *= ipow(KT, m.aT);
        den *= ipow(kr, -m.ar);
        P = num * NRMC_RCP(den);
        lin = m.a1 * g.z1 + m.a2 * g.z2 + m.ar * ice.zr;
        if (WITH_D) {
            // (fast-path reciprocals: every true division costs a slow-path branch in the kernel; 5 of them per evaluation here)
            const double dkr = drc * sr + rc * (sp * (sig * isr)) + dc;
            dlnP = m.a1 * dk1 * NRMC_RCP(k1) + m.a2 * dk2 * NRMC_RCP(k2) + m.ar * dkr * NRMC_RCP(kr) + (m.aT ? m.aT * dKT * NRMC_RCP(KT) : 0.0);
        }
    }
    const double Bk = lin - ice.z0 * NRMC_LOG(P);
    double R = A * Bk;
    if (WITH_D) dg = (bp - A * drc) * irc * Bk - A * ice.z0 * dlnP;
    if (!(R == R)) { R = 1e300; if (WITH_D) dg = 0.0; }   // horizontal ray in (numerically) homogeneous ice: infinite range
    return R - g.rho;
}

NRMC_HD double curve_g(const Curve &cv, int p, double t) { double d; return curve_eval<false>(cv, p, t, d); }
NRMC_HD double curve_gd(const Curve &cv, int p, double t, double &dg) { return curve_eval<true>(cv, p, t, dg); }

// Starting point for the root on piece p (NaN: none).  Direct rays (P0, P1): the straight line between the two points
// in ice of the path-averaged index n_bar = n_ice - z0 (gamma2 - gamma1)/(z2 - z1) has beta = n_bar sin(theta);
// reflected rays (P3): the same with the receiver mirrored at the surface.  Only used for k = 0.
NRMC_HD double guess_t(const IceParams &ice, const PairGeom &g, int p)
{
    if (p == 2) return NAN;
    const double dz = (p == 3) ? -g.z1 - g.z2 : g.z2 - g.z1;
    if (!(dz > 0.0)) return NAN;
    const double dgam = (p == 3) ? (ice.dn - g.g1) + (ice.dn - g.g2) : g.g2 - g.g1;
    const double nbar = ice.n_ice - ice.z0 * dgam * NRMC_RCP(dz);
    const double b0 = nbar * g.rho * NRMC_RSQRT(g.rho * g.rho + dz * dz);
    const double nX = (p == 1) ? g.n2 : ice.ns;
    return (b0 < nX) ? t_of_beta(nX, b0) : NAN;
}

// Bracketed root of g on piece p between a and b (ga, gb of opposite strict sign): Newton steps with the closed-form
// derivative, kept inside the bracket (bisection when a step leaves it).  x0: starting point (NaN: secant point).
// Stops on |g| <= 1e-10 m or when the Newton step is below 1e-6 relative (the step is then applied unevaluated:
// quadratic convergence puts the result at ~1e-11; scratch/tolstudy.cpp: against a 1e-9 threshold C0 moves by <= 1.1e-11,
// the path length by <= 5e-8 m on 4e5 pairs of the cfg5 geometry, for 6 % fewer evaluations).
// (Tried on the CPU harness, scratch/evalhist.cpp: the refracted rays' mean of 3.5 evaluations -- with 5 % of the solves at 6 - 7,
// which is what a warp waits for -- comes from the range curve bending over next to the junction t = 1; starting from the root of a
// parabola with its vertex there, plus a parabolic first step through the junction value, removes the tail (1 % above 4) but
// raises the mean to 3.85: the curve is close to linear over most of the piece.)
// (Tried: the zero of the inverse cubic Hermite interpolant through the last two points instead of the Newton step from the
// second evaluation on -- 3.60 -> 3.33 evaluations per root on the cfg5 geometry, the slowest lane of a warp 5.56 -> 5.09 --
// but the three extra live doubles and the longer loop body cost more on the B200 than the evaluations saved: 15.4 vs 15.1 ms.)
#ifndef NRMC_SOLVE_GTOL
#define NRMC_SOLVE_GTOL 1e-10
#endif
#ifndef NRMC_SOLVE_STEP_TOL
#define NRMC_SOLVE_STEP_TOL 1e-6
#endif
NRMC_HD double solve_piece(const Curve &cv, int p, double a, double ga, double b, double gb, double x0)
{
    const double gtol = NRMC_SOLVE_GTOL;
    // the bracket is kept ordered (lo < hi) from the start: no min / max per iteration
    double lo = a, glo = ga, hi = b, ghi = gb;
    if (b < a) { lo = b; glo = gb; hi = a; ghi = ga; }
    double x = x0;
    if (!(x > lo && x < hi)) x = (lo * ghi - hi * glo) * NRMC_RCP(ghi - glo);
    if (!(x > lo && x < hi)) x = 0.5 * (lo + hi);
    for (int it = 0; it < 100; ++it) {
        double dg;
        const double gx = curve_gd(cv, p, x, dg);
        if (fabs(gx) <= gtol) break;
        if ((gx > 0) == (ghi > 0)) { hi = x; ghi = gx; } else { lo = x; glo = gx; }
        double xn = x - gx * NRMC_RCP(dg);
        if (!(xn > lo && xn < hi)) xn = 0.5 * (lo + hi);
        else if (fabs(xn - x) <= NRMC_SOLVE_STEP_TOL * (fabs(x) + 1e-3)) { x = xn; break; }
        if (hi - lo <= 4e-16 * (fabs(lo) + fabs(hi))) { x = xn; break; }
        x = xn;
    }
    return x;
}

// Look for an interior maximum of g on piece p (end values <= 0): bracket the sign change of dg/dt and close it
// with Illinois steps on the derivative (the Illinois weights only steer the next abscissa; the bracket keeps the TRUE
// derivatives, which is what the termination test needs).  Returns true as soon as a point with g > 0 is found (xm, gm);
// false if the maximum is at an end point or stays <= 0.  `interior`: this piece holds the curve's maximum.
NRMC_HD bool maximise_piece(const Curve &cv, int p, double a, double b, double &xm, double &gm, bool &interior)
{
    double lo = fmin(a, b), hi = fmax(a, b);
    double dlo, dhi;
    double glo = curve_gd(cv, p, lo, dlo);
    double ghi = curve_gd(cv, p, hi, dhi);
    (void)ghi;
    interior = (dlo > 0.0) && (dhi < 0.0);
    xm = lo; gm = glo;
    if (!interior) return false;
    double wlo = 1.0, whi = 1.0;
    int side = 0;
    for (int it = 0; it < 100; ++it) {
        const double ea = wlo * dlo, eb = whi * dhi;
        double x = (lo * eb - hi * ea) / (eb - ea);
        if (!(x > lo && x < hi)) x = 0.5 * (lo + hi);
        double dx;
        const double gx = curve_gd(cv, p, x, dx);
        if (gx > gm || it == 0) { xm = x; gm = gx; }
        if (gx > 0.0) return true;
        if (dx > 0.0) { lo = x; dlo = dx; wlo = 1.0; if (side == 1) whi *= 0.5; side = 1; }
        else if (dx < 0.0) { hi = x; dhi = dx; whi = 1.0; if (side == -1) wlo *= 0.5; side = -1; }
        else break;
        // between lo and hi the derivative falls from dlo to dhi: g cannot rise above gm by more than slope x width
        if (fmax(dlo, -dhi) * (hi - lo) < 1e-11 || (hi - lo) <= 4e-16 * (fabs(lo) + fabs(hi))) break;
    }
    return gm > 0.0;
}

struct Root { double v; int piece; double beta; };
struct Bracket { double a, ga, b, gb; int piece; };

// piece end points in curve order: P0 t: 0 -> 1, P1 t: tmin -> 1, P2 t: 1 -> tmin, P3 t: 1 -> 0
NRMC_HD double piece_begin(const PairGeom &g, int p) { return p == 0 ? 0.0 : (p == 1 ? g.tmin : 1.0); }
NRMC_HD double piece_end(const PairGeom &g, int p) { return p <= 1 ? 1.0 : (p == 2 ? g.tmin : 0.0); }

// Junction values of the curve (J0 = J4 = -rho at the vertical ray) and the brackets they give.  Returns the number of
// brackets (0..2); need_hump: the whole curve was sampled below rho, the maximum has to be searched (hump_search).
NRMC_HD int classify_mode(const Curve &cv, double &J1, double &J2, double &J3, Bracket br[2], bool &need_hump)
{
    const PairGeom &g = *cv.g;
    const bool has_band = g.s2max > 0.0;
    const double J0 = -g.rho, J4 = -g.rho;
    J1 = curve_g(cv, 0, 1.0);
    J3 = curve_g(cv, 3, 1.0);
    J2 = has_band ? curve_g(cv, 1, 1.0) : J1;
    int nb = 0;
#define NRMC_PUSH_BRACKET(P, A, GA, B, GB)                                                         \
    do {                                                                                            \
        if (nb < 2) { br[nb].piece = (P); br[nb].a = (A); br[nb].ga = (GA); br[nb].b = (B); br[nb].gb = (GB); } \
        ++nb;                                                                                       \
    } while (0)
    // receiver exactly at the surface (no band): P0 and P3 coincide (py: one 'reflected' solution) -> only P3
    if (has_band && ((J0 > 0) != (J1 > 0))) NRMC_PUSH_BRACKET(0, 0.0, J0, 1.0, J1);
    if (has_band && ((J1 > 0) != (J2 > 0))) NRMC_PUSH_BRACKET(1, g.tmin, J1, 1.0, J2);
    if (has_band && ((J2 > 0) != (J3 > 0))) NRMC_PUSH_BRACKET(2, 1.0, J2, g.tmin, J3);
    if ((J3 > 0) != (J4 > 0)) NRMC_PUSH_BRACKET(3, 1.0, J3, 0.0, J4);
#undef NRMC_PUSH_BRACKET
    need_hump = (nb == 0 && J1 <= 0 && J2 <= 0 && J3 <= 0);
    return nb > 2 ? 2 : nb;
}

// The hump of a curve whose junctions all lie below rho: search the pieces adjacent to the largest junction.
// Returns 2 (two brackets around the maximum) or 0.
NRMC_HD int hump_search(const Curve &cv, double J1, double J2, double J3, Bracket br[2])
{
    const PairGeom &g = *cv.g;
    const bool has_band = g.s2max > 0.0;
    const double J0 = -g.rho, J4 = -g.rho;
    int jm = 1;
    double Jm = J1;
    if (J2 > Jm) { jm = 2; Jm = J2; }
    if (J3 > Jm) { jm = 3; Jm = J3; }
    bool interior = false;
    for (int side = 0; side < 2 && !interior; ++side) {
        const int p = jm - 1 + side;
        if (!has_band && p != 3) continue;
        const double pa = piece_begin(g, p), pb = piece_end(g, p);
        const double ja = (p == 0) ? J0 : (p == 1 ? J1 : (p == 2 ? J2 : J3));
        const double jb = (p == 0) ? J1 : (p == 1 ? J2 : (p == 2 ? J3 : J4));
        double xm, gm;
        if (maximise_piece(cv, p, pa, pb, xm, gm, interior)) {
            br[0].piece = p; br[0].a = pa; br[0].ga = ja; br[0].b = xm; br[0].gb = gm;
            br[1].piece = p; br[1].a = xm; br[1].ga = gm; br[1].b = pb; br[1].gb = jb;
            return 2;
        }
    }
    return 0;
}

// Largest horizontal range any ray of the k = 0 mode reaches between the depths of g (g.rho is ignored).  Every direct
// ray is shorter than the turned ray of the same beta (R_dir = I(z1) - I(z2) <= I(z1) + I(z2) = R_trn, I(z) = integral of
// beta/s from z to the turning point), so the maximum lies on the pieces P2 / P3 or at their junctions.  R_max grows with the
// depth of either end point (each I grows, and so does the admissible range beta <= n(z2)): a table of R_max on a depth
// grid bounds it from above at the deeper grid corner -- used by K_classify to discard shadow-zone pairs without a search.
NRMC_HD double range_max(const IceParams &ice, const PairGeom &g_in)
{
    PairGeom g = g_in;
    g.rho = 0.0;
    Curve cv;
    cv.ice = &ice; cv.g = &g; cv.k = 0; cv.rcase = 1;
    const bool has_band = g.s2max > 0.0;
    double best = curve_g(cv, 3, 1.0);                       // beta = n_s, reflected
    if (has_band) best = fmax(best, curve_g(cv, 2, 1.0));    // beta = n(z2): apex at the receiver
    for (int p = has_band ? 2 : 3; p <= 3; ++p) {
        double lo = fmin(piece_begin(g, p), piece_end(g, p)), hi = fmax(piece_begin(g, p), piece_end(g, p));
        if (p == 3) lo = 1e-9;                                // the vertical ray has range 0
        double dlo, dhi;
        const double glo = curve_gd(cv, p, lo, dlo), ghi = curve_gd(cv, p, hi, dhi);
        best = fmax(best, fmax(glo, ghi));
        if (!(dlo > 0.0 && dhi < 0.0)) continue;              // no interior maximum on this piece
        int side = 0;
        double wlo = 1.0, whi = 1.0;                          // Illinois weights; dlo / dhi stay the true derivatives
        for (int it = 0; it < 200; ++it) {
            const double ea = wlo * dlo, eb = whi * dhi;
            double x = (lo * eb - hi * ea) / (eb - ea);
            if (!(x > lo && x < hi)) x = 0.5 * (lo + hi);
            double dx;
            const double gx = curve_gd(cv, p, x, dx);
            best = fmax(best, gx);
            if (dx > 0.0) { lo = x; dlo = dx; wlo = 1.0; if (side == 1) whi *= 0.5; side = 1; }
            else if (dx < 0.0) { hi = x; dhi = dx; whi = 1.0; if (side == -1) wlo *= 0.5; side = -1; }
            else break;
            if (fmax(dlo, -dhi) * (hi - lo) < 1e-9 || (hi - lo) <= 4e-16 * (fabs(lo) + fabs(hi))) break;
        }
        best += fmax(fabs(dlo), fabs(dhi)) * (hi - lo);      // what the unconverged bracket could still add
    }
    return best;
}

NRMC_HD Root solve_bracket(const Curve &cv, const Bracket &b)
{
    Root r;
    r.piece = b.piece;
    // the straight-line starting points describe the whole piece; a bracket made by the hump search starts from its secant point
    const bool whole = (b.a == piece_begin(*cv.g, b.piece) && b.b == piece_end(*cv.g, b.piece));
    double x0 = (cv.k == 0 && whole) ? guess_t(*cv.ice, *cv.g, b.piece) : NAN;
#ifndef NRMC_NO_VERTEX_GUESS
    // Rays that turn (P2 refracted, P3 reflected beyond the straight-line estimate), whole piece, g > 0 at the junction t = 1 (apex
    // at the receiver / grazing the surface) and g << 0 at the other end: the range curve bends over next to t = 1, the secant
    // point lands beyond its maximum and the first Newton step is wasted -- 5 % of these solves need 6 - 7 evaluations, and a warp
    // waits for its slowest lane.  Start from the root of the parabola with its vertex at t = 1 instead: the mean number of
    // evaluations rises (3.46 -> 3.82, the curve is close to linear further down) but 98.6 % finish within 4 (scratch/evalhist.cpp).
    if (cv.k == 0 && whole && !(x0 == x0) && b.piece >= 2) {
        const bool b_is_top = b.b > b.a;
        const double top = b_is_top ? b.b : b.a, bot = b_is_top ? b.a : b.b, gtop = b_is_top ? b.gb : b.ga, gbot = b_is_top ? b.ga : b.gb;
        if (gtop > 0.0 && gbot < 0.0) {
            const double rg = gtop * NRMC_RCP(gtop - gbot), q = NRMC_SQRT(rg);
#ifdef NRMC_VERTEX_ONLY
            x0 = top - (top - bot) * q;
#else
            x0 = bot + (top - bot) * (1.0 - q) * (1.0 + rg);        // vertex parabola (1 - q) and chord (1 - q^2) blended with weight q
#endif
        }
    }
    // a bracket made by the hump search ends at a point next to the curve's maximum (g > 0 there, dg ~ 0): vertex parabola there
    if (!whole) {
        const bool b_pos = b.gb > 0.0;
        const double xp = b_pos ? b.b : b.a, xn = b_pos ? b.a : b.b, gp = b_pos ? b.gb : b.ga, gn = b_pos ? b.ga : b.gb;
        if (gp > 0.0 && gn < 0.0) x0 = xp - (xp - xn) * NRMC_SQRT(gp * NRMC_RCP(gp - gn));
    }
#endif
    r.v = solve_piece(cv, b.piece, b.a, b.ga, b.b, b.gb, x0);
    const double q = NRMC_RCP(1.0 + r.v * r.v);
    r.beta = ((b.piece == 1 || b.piece == 2) ? cv.g->n2 : cv.ice->ns) * (2.0 * r.v) * q;
    return r;
}

// All roots (0 or 2; 1 only on a tangency) of one mode, ordered by increasing C0 = 1/beta.
NRMC_HD int find_roots_mode(const IceParams &ice, const PairGeom &g, int k, int rcase, Root out[2])
{
    Curve cv;
    cv.ice = &ice; cv.g = &g; cv.k = k; cv.rcase = rcase;
    double J1, J2, J3;
    Bracket br[2];
    bool need_hump;
    int nb = classify_mode(cv, J1, J2, J3, br, need_hump);
    if (need_hump) nb = hump_search(cv, J1, J2, J3, br);
    Root r0, r1;
    r0.v = r1.v = 0; r0.piece = r1.piece = 0; r0.beta = r1.beta = 0;
    if (nb > 0) r0 = solve_bracket(cv, br[0]);
    if (nb > 1) r1 = solve_bracket(cv, br[1]);
    if (nb == 2 && r0.beta < r1.beta) { Root t = r0; r0 = r1; r1 = t; }  // ascending C0 (py:1547)
    out[0] = r0; out[1] = r1;
    return nb;
}

// ---------------------------------------------------------------------------------------------------------------
// per-solution properties
// ---------------------------------------------------------------------------------------------------------------
struct SolutionProps {
    double C0, C1;
    int type;                          // 1 direct, 2 refracted, 3 reflected (propagation.py:3-7), reference convention py:1365-1398
    double sin_l, cos_l;               // launch angle alpha_l (2-D frame): [sin, 0, cos]      py:2583
    double sin_r, cos_r;               // receive angle alpha_r                                  py:2617: [-sin, 0, cos]
    double path_length, travel_time;
    double refl_angle;                 // asin(beta/ns) if reflected else NaN; per-segment mask in refl_mask
    uint32_t refl_mask;                // bit i set: segment i reflects off the surface (py:1230)
    int n_segments;
};

NRMC_HD void solution_props(const IceParams &ice, const PairGeom &g, double x1y, int k, int rcase, const Root &root,
                             SolutionProps &o)
{
    const bool band = (root.piece == 1 || root.piece == 2);
    const bool turned = root.piece >= 2;
    RayState r;
    ray_state(ice, g, band, root.v, r);
    const ModeCoeffs m = mode_coeffs(k, rcase, turned);
    // (reciprocals through the fast path and shared: every true division costs a slow-path branch in the kernel)
    const double irc = NRMC_RCP(r.rc), in1 = NRMC_RCP(g.n1), in2 = NRMC_RCP(g.n2);
    const double A = r.beta * irc;
    o.C0 = NRMC_RCP(r.beta);
    // C_1 = y1 - y(z1; C_1 = 0),  y = z0 beta/sqrt(c) ln(gamma / (2 k1))       (py:487-491,118-125)
    o.C1 = x1y - ice.z0 * A * NRMC_LOG(g.g1 * NRMC_RCP(2.0 * r.k1_1));
    // solution type on the UNREFLECTED geometry (py:2146 -> :1386-1398): direct iff rho < y_turn - y1
    if (k == 0) o.type = turned ? (r.reflected ? 3 : 2) : 1;
    else {
        double T1 = A * (-g.z1 - ice.z0 * NRMC_LOG(r.KT * NRMC_RCP(r.k1_1)));
        o.type = (g.rho < T1) ? 1 : (r.reflected ? 3 : 2);
    }
    // launch: theta1 with sin = beta/n1, downward start (case 2, k>0) -> pi - theta1   (py:1161-1196)
    o.sin_l = r.beta * in1;
    o.cos_l = r.s1 * in1;
    if (k > 0 && rcase == 2) o.cos_l = -o.cos_l;
    // receive: pi - theta2 if the last segment arrives up-going, theta2 otherwise      (py:1198-1199)
    o.sin_r = r.beta * in2;
    o.cos_r = turned ? r.s2 * in2 : -r.s2 * in2;
    // path length / travel time (py:602-783): S(z) = n_ice/rc U(z) + z0 ln k2,  ct(z) = n_ice^2/rc U(z) + z0 (s + n_ice ln k2)
    // summed with the same integer coefficients as the range:  sum a U = rho rc / beta at the root.
    double k2_1 = r.s1 + g.n1, k2_2 = r.s2 + g.n2, k2_r = r.sr + ice.nr;
    double k2_T = r.reflected ? r.ss + ice.ns : r.beta;
    double sT = r.reflected ? r.ss : 0.0;
    double num = 1.0, den = 1.0;
    if (m.a1 > 0) num *= k2_1; else den *= k2_1;
    if (m.a2 > 0) num *= k2_2; else den *= k2_2;
    num *= ipow(k2_T, m.aT);
    double ssum = m.a1 * r.s1 + m.a2 * r.s2 + m.aT * sT;
    if (m.ar != 0) { den *= ipow(k2_r, -m.ar); ssum += m.ar * r.sr; }
    double lk2 = NRMC_LOG(num * NRMC_RCP(den));
    double nsumU = ice.n_ice * g.rho * o.C0;                  // n_ice / rc * sum a U,  sum a U = rho rc / beta
    o.path_length = nsumU + ice.z0 * lk2;
    o.travel_time = (ice.n_ice * nsumU + ice.z0 * (ssum + ice.n_ice * lk2)) * (1.0 / NRMC_SPEED_OF_LIGHT);
    // surface reflection angle per segment (py:1201-1237)
    o.n_segments = k + 1;
    o.refl_mask = 0;
    o.refl_angle = NAN;
    if (r.reflected) {
        o.refl_angle = atan2(r.beta, r.ss);
        for (int i = 0; i <= k; ++i) {
            bool has = true;
            if (i == 0 && k > 0 && rcase == 2) has = false;   // starts downward: no turning point on the first segment
            if (i == k && !turned) has = false;               // arrives before the turning point
            if (has) o.refl_mask |= (1u << i);
        }
    }
}


// ---------------------------------------------------------------------------------------------------------------
// Signal focusing (ray_tracing.get_focusing, py:2778-2888).  The reference re-traces the pair with the receiver moved by
// dz = -1 cm and forms  sqrt(D / sin(rec) |d launch angle / dz|) * sqrt(D sin(launch) / rho), limits it, and multiplies by
// sqrt(n_emitter / n_receiver).  Here the derivative is exact: on R(beta; z_e, z_r) = rho,
//   d beta / d z_r = -(dR/dz_r) / (dR/dbeta),  |dR/dz_r| = tan(theta_r) = beta / s_r,  |d launch / d beta| = 1 / s_e,
// so |d launch / d z_r| = beta / (s_r s_e |dR/dbeta|), with dR/dbeta = (dR/dt) / (dbeta/dt) from the same closed form the
// Newton solver uses.  s_2 / (dbeta/dt) is formed analytically: it stays finite when the apex approaches the shallower
// point (s_2 -> 0).  The branch (arrival before / after the turning point) is the one whose range reproduces rho.
// D: path length of the solution.  Symmetric in emitter / receiver up to the index factors, so `swap` only selects them.
// ---------------------------------------------------------------------------------------------------------------
NRMC_HD double focusing_factor(const IceParams &ice, const PairGeom &g, bool swap, int k, int rcase, double beta,
                                double path_length, double limit)
{
    Curve cv;
    cv.ice = &ice; cv.g = &g; cv.k = k; cv.rcase = rcase;
    const bool band = beta > ice.ns && g.s2max > 0.0;
    const double nX = band ? g.n2 : ice.ns;
    const double t = t_of_beta(nX, fmin(beta, nX));
    double d0, d1;
    const double r0 = curve_gd(cv, band ? 1 : 0, t, d0);      // arrival before the turning point
    const double r1 = curve_gd(cv, band ? 2 : 3, t, d1);      // after
    const double dg = fabs(r0) <= fabs(r1) ? d0 : d1;
    RayState r;
    ray_state(ice, g, band, t, r);
    const double q = 1.0 / (1.0 + t * t);
    const double s2_over_bp = band ? 0.5 / q : r.s2 / fmax(2.0 * r.ss * q, 1e-300);   // s_2 / (dbeta/dt)
    const double W = r.s1 * s2_over_bp * fabs(dg);            // s_e s_r |dR/dbeta|
    const double n_e = swap ? g.n2 : g.n1, n_r = swap ? g.n1 : g.n2;
    double f = sqrt(path_length * n_r / W) * sqrt(path_length * beta / (n_e * g.rho));
    if (!(f <= limit)) f = limit;                             // also the caustic (W -> 0) and 0/0
    return f * sqrt(n_e / n_r);
}

// ---------------------------------------------------------------------------------------------------------------
// one (vertex, antenna) pair: geometry -> modes -> roots -> properties -> SoA outputs          (py:2057-2146)
// ---------------------------------------------------------------------------------------------------------------
struct TraceOutputs {          // all [N,S] pair-major, S = 2 + 4 n_refl; any pointer may be null (skipped)
    int32_t *n_sol;            // [N]
    int32_t *status;           // [N] bit0: a point is above the surface (no solution, py:1445-1448); bit1: a point is below
                               //     the reflective layer (base:156-161 raises AttributeError); bit2: non-finite input
    int8_t *type, *reflection, *reflection_case;
    double *C0, *C1, *path_length, *travel_time;
    double *launch, *receive;  // [N,S,3]
    double *reflection_angle;  // [N,S,n_refl+1], NaN = None
    double *viewing_angle;     // [N,S] angle between the shower's propagation direction and the launch vector (simulation.py:191)
    const int64_t *row_offset; // compact layout: row of slot s of pair i = row_offset[i] + s (per-slot arrays are [rows,...]);
                               // null: padded layout, row = i * S + s
    int64_t row_limit;         // compact layout: rows the per-slot arrays can hold (solutions beyond it are dropped)
};

NRMC_HD int64_t row_of(const TraceOutputs &o, int64_t pair, int slot, int S)
{
    return o.row_offset ? o.row_offset[pair] + slot : pair * S + slot;
}

// Viewing-angle cut of the simulation loop (NuRadioMC/simulation/simulation.py:175-208): solutions whose launch direction
// is further than delta_C_cut from the Cherenkov cone of the shower are dropped before any further work.
struct ShowerCut { double sx, sy, sz, cut; bool on; };

// angle between the shower propagation direction and the launch vector (lx ex, lx ey, lz); radiotools.helper.get_angle
NRMC_HD double viewing_angle_of(const ShowerCut &sc, double ex, double ey, double lx, double lz)
{
    const double nrm = sqrt(sc.sx * sc.sx + sc.sy * sc.sy + sc.sz * sc.sz);
    double c = (sc.sx * lx * ex + sc.sy * lx * ey + sc.sz * lz) / nrm;
    c = fmin(fmax(c, -1.0), 1.0);
    return acos(c);
}

// |viewing angle - Cherenkov angle| <= cut, Cherenkov angle = arccos(1 / n(vertex))   (simulation.py:176-177, :195-208)
NRMC_HD bool passes_cut(const ShowerCut &sc, double viewing, double n_vertex)
{
    return fabs(viewing - acos(1.0 / n_vertex)) <= sc.cut;
}

struct SolRec {                // everything the attenuation kernels need about one solution: one aligned 64-byte load
    int64_t pair;
    double beta;               // Snell invariant of the ray
    double delta;              // n_ice - beta, formed without cancellation: (n_ice^2 - beta^2)/(n_ice + beta)
    double zv;                 // z0 ln(delta/dn): depth of the (possibly virtual, > 0) apex
    double z1, z2;             // depths of the deeper / shallower end point
    int32_t slot;
    uint8_t piece, k, rcase, pad;
    int64_t row;               // row of the solution in the per-slot output arrays (padded: pair * S + slot)
};

NRMC_HD void make_solrec(const IceParams &ice, const PairGeom &g, int64_t pair, int slot, int64_t row, int k, int rcase, const Root &root,
                         SolRec &r)
{
    r.pair = pair; r.slot = slot; r.piece = (uint8_t)root.piece; r.k = (uint8_t)k; r.rcase = (uint8_t)rcase; r.pad = 0; r.row = row;
    r.beta = root.beta;
    // c = n_ice^2 - beta^2 = c0 + sigma^2 with sigma from the curve parameter (ray_state)
    const bool band = (root.piece == 1 || root.piece == 2);
    const double q = NRMC_RCP(1.0 + root.v * root.v);
    const double sig = (band ? g.n2 : ice.ns) * ((1.0 - root.v) * (1.0 + root.v)) * q;
    const double c = (band ? g.c0_band : g.c0_sub) + sig * sig;
    r.delta = c * NRMC_RCP(ice.n_ice + r.beta);
    r.zv = ice.z0 * NRMC_LOG(r.delta * ice.inv_dn);
    r.z1 = g.z1; r.z2 = g.z2;
}

#define NRMC_STATUS_AIR 1
#define NRMC_STATUS_BELOW_REFLECTOR 2
#define NRMC_STATUS_NONFINITE 4

struct Frame2D { double z1, z2, rho, x1y, ex, ey; bool swap; };

// py:2057-2090: deeper point first, rotate about z so the shallower point lies at +rho
NRMC_HD void make_frame(double ax, double ay, double az, double bx, double by, double bz, Frame2D &f)
{
    f.swap = bz < az;
    if (f.swap) { double t; t = ax; ax = bx; bx = t; t = ay; ay = by; by = t; t = az; az = bz; bz = t; }
    double dx = bx - ax, dy = by - ay;
    const double d2 = dx * dx + dy * dy, rs = NRMC_RSQRT(d2);
    // (horizontal distances below 1e-150 m count as zero; NaN and inf coordinates propagate to rho for pair_status)
    if (d2 > 1e-300) { f.rho = d2 * rs; f.ex = dx * rs; f.ey = dy * rs; } else { f.rho = d2 == d2 ? 0.0 : d2; f.ex = 1.0; f.ey = 0.0; }  // atan2(0,0) = 0
    f.z1 = az; f.z2 = bz; f.x1y = ax;
}

// empty slots: 0 / NaN (HDF5 writer convention, output_writer_hdf5.py:272-275)
NRMC_HD void fill_empty_slot(const TraceOutputs &o, int64_t q, int K1)
{
    if (o.type) o.type[q] = 0;
    if (o.reflection) o.reflection[q] = 0;
    if (o.reflection_case) o.reflection_case[q] = 0;
    if (o.C0) o.C0[q] = NAN;
    if (o.C1) o.C1[q] = NAN;
    if (o.path_length) o.path_length[q] = NAN;
    if (o.travel_time) o.travel_time[q] = NAN;
    if (o.launch) { o.launch[3 * q] = NAN; o.launch[3 * q + 1] = NAN; o.launch[3 * q + 2] = NAN; }
    if (o.receive) { o.receive[3 * q] = NAN; o.receive[3 * q + 1] = NAN; o.receive[3 * q + 2] = NAN; }
    if (o.reflection_angle) for (int t = 0; t < K1; ++t) o.reflection_angle[q * K1 + t] = NAN;
    if (o.viewing_angle) o.viewing_angle[q] = NAN;
}

// launch direction in the 2-D frame (x along rho, z up): roles of launch and receive exchanged when the points were swapped
NRMC_HD void launch_2d(const Frame2D &f, const SolutionProps &p, double &lx, double &lz)
{
    lx = f.swap ? -p.sin_r : p.sin_l;
    lz = f.swap ? p.cos_r : p.cos_l;
}

NRMC_HD void write_solution(const TraceOutputs &o, int64_t q, int K1, const Frame2D &f, int k, int rcase, const SolutionProps &p)
{
    if (o.type) o.type[q] = (int8_t)p.type;
    if (o.reflection) o.reflection[q] = (int8_t)k;
    if (o.reflection_case) o.reflection_case[q] = (int8_t)rcase;
    if (o.C0) o.C0[q] = p.C0;
    if (o.C1) o.C1[q] = p.C1;
    if (o.path_length) o.path_length[q] = p.path_length;
    if (o.travel_time) o.travel_time[q] = p.travel_time;
    // 2-D vectors -> 3-D: R^T [vx,0,vz] = [vx ex, vx ey, vz]; roles exchanged when swapped (py:2583-2590,2617-2623)
    double lx = p.sin_l, lz = p.cos_l, rx = -p.sin_r, rz = p.cos_r;
    if (f.swap) { double tx = lx, tz = lz; lx = rx; lz = rz; rx = tx; rz = tz; }
    if (o.launch) { o.launch[3 * q] = lx * f.ex; o.launch[3 * q + 1] = lx * f.ey; o.launch[3 * q + 2] = lz; }
    if (o.receive) { o.receive[3 * q] = rx * f.ex; o.receive[3 * q + 1] = rx * f.ey; o.receive[3 * q + 2] = rz; }
    if (o.reflection_angle)
        for (int s = 0; s < K1; ++s) o.reflection_angle[q * K1 + s] = ((p.refl_mask >> s) & 1u) ? p.refl_angle : NAN;
}

// status of a pair before any ray is traced (0: traceable)
NRMC_HD int pair_status(const IceParams &ice, const Frame2D &f)
{
    if (!(f.rho == f.rho) || !(f.z1 == f.z1) || !(f.z2 == f.z2) || isinf(f.rho) || isinf(f.z1)) return NRMC_STATUS_NONFINITE;
    if (f.z2 > 0.0) return NRMC_STATUS_AIR;
    if (ice.n_refl > 0 && f.z1 < ice.zr) return NRMC_STATUS_BELOW_REFLECTOR;
    return 0;
}

// Returns the number of solutions; fills the SoA slots of pair i and (if recs != null) one SolRec per solution.
// Thread-per-pair form: used by the generic kernel (bottom reflections) and by the CPU test harness.
NRMC_HD int trace_pair(const IceParams &ice, double ax, double ay, double az, double bx, double by, double bz,
                        int64_t i, const TraceOutputs &o, SolRec *recs, const ShowerCut *sc = nullptr, int *n_recs = nullptr,
                        uint32_t *cut_mask = nullptr)
{
    const int S = 2 + 4 * ice.n_refl, K1 = ice.n_refl + 1;
    int n = 0, nrec = 0;
    uint32_t cut = 0;
    Frame2D f;
    make_frame(ax, ay, az, bx, by, bz, f);
    const int status = pair_status(ice, f);
    if (status == 0) {
        PairGeom g;
        make_pair_geom(ice, f.z1, f.z2, fmax(f.rho, 1e-12), g);
        for (int md = 0; md < 1 + 2 * ice.n_refl; ++md) {
            const int k = md == 0 ? 0 : (md - 1) / 2 + 1;
            const int rcase = md == 0 ? 1 : (md - 1) % 2 + 1;
            Root roots[2];
            const int nr = find_roots_mode(ice, g, k, rcase, roots);
            for (int j = 0; j < nr && n < S; ++j) {
                SolutionProps p;
                solution_props(ice, g, f.x1y, k, rcase, roots[j], p);
                write_solution(o, i * S + n, K1, f, k, rcase, p);
                bool keep = true;
                if (sc && sc->on) {
                    double lx, lz;
                    launch_2d(f, p, lx, lz);
                    const double va = viewing_angle_of(*sc, f.ex, f.ey, lx, lz);
                    if (o.viewing_angle) o.viewing_angle[i * S + n] = va;
                    keep = passes_cut(*sc, va, f.swap ? g.n2 : g.n1);
                    if (!keep) cut |= (1u << n);
                } else if (o.viewing_angle) o.viewing_angle[i * S + n] = NAN;
                if (recs && keep) make_solrec(ice, g, i, n, i * S + n, k, rcase, roots[j], recs[nrec++]);
                ++n;
            }
        }
    }
    if (o.n_sol) o.n_sol[i] = n;
    if (o.status) o.status[i] = status;
    for (int s = n; s < S; ++s) fill_empty_slot(o, i * S + s, K1);
    if (n_recs) *n_recs = nrec;
    if (cut_mask) *cut_mask = cut;
    return n;
}

}  // namespace nrmc
