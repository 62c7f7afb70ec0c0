// nrmc_att.cuh -- attenuation along the ray path: quadrature plan, node evaluation, attenuation length models.
//
// Reference behaviour: ray_tracing_2D.get_attenuation_along_path (analyticraytracing.py:933-1089) integrates
//   I(f) = int ds(z) / L(z, f)   per path segment with scipy.quad(epsrel=1e-2), takes exp(-I), interpolates to the
// output frequencies and multiplies the segments; L(z,f) is attenuation.get_attenuation_length (attenuation.py:145-262).
//
// B200 design (not a translation): with u = sqrt(z_v - z), z_v = z0 ln((n_ice - beta)/dn) the (possibly virtual)
// apex height, n(z) - beta = (n_ice - beta)(1 - exp(-u^2/z0)) and the line element becomes
//   ds = 2 u n / sqrt((n_ice - beta)(1 - exp(-u^2/z0))(n + beta)) du,
// analytic on the whole path -- the 1/sqrt turning-point singularity of ds/dz is gone, so a fixed 16-point
// Gauss-Legendre rule per panel converges spectrally (measured <= 1.3e-5 relative on the attenuation factor with one
// panel per leg, ~1e-8 with 24 points).  The path of any mode is a set of at most three u-panels
// [u_T,u_2], [u_2,u_1], [u_1,u_r] traversed an integer number of times per segment (att_plan).
#pragma once
#include "nrmc_math.cuh"

namespace nrmc {

// ---------------------------------------------------------------------------------------------------------------
// exp / expm1 for the attenuation kernels.  CUDA's exp() materialises every polynomial coefficient with two UMOVs
// (FP64 immediates do not fit the DFMA encoding), which costs more issue slots than the arithmetic; here the
// coefficients sit in constant memory (one LDCU.128 per two coefficients), the polynomial is the degree-11 Taylor
// series on |r| <= ln2/2 (truncation 6e-15 relative) and there is no slow path: the argument is clamped to [-700, 700],
// far outside anything an attenuation exponent needs (exp(-700) = 1e-304 stands in for 0).
// ---------------------------------------------------------------------------------------------------------------
#if defined(__CUDACC__)
__constant__ double c_expc[10] = {2.505210838544172e-08, 2.755731922398589e-07, 2.7557319223985893e-06, 2.48015873015873e-05, 0.0001984126984126984, 0.001388888888888889, 0.008333333333333333, 0.041666666666666664, 0.16666666666666666, 0.5};
__device__ __forceinline__ double exp_reduce(double x, int &k)
{
    const double t = fma(x, 1.4426950408889634, 6755399441055744.0);
    k = __double2loint(t);
    const double kd = t - 6755399441055744.0;
    double r = fma(kd, -6.93147180369123816490e-01, x);
    return fma(kd, -1.90821492927058770002e-10, r);
}
__device__ __forceinline__ double expm1_poly(double r)     // e^r - 1 on |r| <= ln2/2
{
    double p = c_expc[0];
#pragma unroll
    for (int i = 1; i < 10; ++i) p = fma(p, r, c_expc[i]);
    return fma(p * r, r, r);
}
__device__ __forceinline__ double exp_c(double x)
{
    int k;
    const double r = exp_reduce(fmin(fmax(x, -700.0), 700.0), k);
    const double p = expm1_poly(r) + 1.0;
    return __hiloint2double(__double2hiint(p) + (k << 20), __double2loint(p));
}
// arguments known to lie in (-700, 700): no clamp
__device__ __forceinline__ double exp_c_bounded(double x)
{
    int k;
    const double r = exp_reduce(x, k);
    const double p = expm1_poly(r) + 1.0;
    return __hiloint2double(__double2hiint(p) + (k << 20), __double2loint(p));
}
// x <= 0 of any size (attenuation exponents): clamp from below only
__device__ __forceinline__ double exp_c_neg(double x)
{
    int k;
    const double r = exp_reduce(fmax(x, -700.0), k);
    const double p = expm1_poly(r) + 1.0;
    return __hiloint2double(__double2hiint(p) + (k << 20), __double2loint(p));
}
// x <= 0 of any size, degree-8 polynomial (truncation 2e-10 relative): the factors the thread-per-solution kernels emit
__device__ __forceinline__ double exp_c_neg8(double x)
{
    int k;
    const double r = exp_reduce(x > -700.0 ? x : -700.0, k);
    double p = c_expc[3];
#pragma unroll
    for (int i = 4; i < 10; ++i) p = fma(p, r, c_expc[i]);
    const double q = fma(p * r, r, r) + 1.0;
    return __hiloint2double(__double2hiint(q) + (k << 20), __double2loint(q));
}
__device__ __forceinline__ double expm1_c_neg(double x)    // -700 < x <= 0
{
    int k;
    const double r = exp_reduce(fmax(x, -700.0), k);
    const double q = expm1_poly(r);
    const double s = __hiloint2double((1023 + k) << 20, 0);
    return fma(s, q, s - 1.0);
}
// -700 < x <= 0 guaranteed by the caller (no clamp: FP64 min / max are emulated with six instructions), degree-8 polynomial:
// truncation r^9 / 9! relative to r, 6e-10 -- the quadrature weights of the attenuation kernels need 1e-7
__device__ __forceinline__ double expm1_c_small(double x)
{
    int k;
    const double r = exp_reduce(x, k);
    double p = c_expc[3];
#pragma unroll
    for (int i = 4; i < 10; ++i) p = fma(p, r, c_expc[i]);
    const double q = fma(p * r, r, r);
    const double s = __hiloint2double((1023 + k) << 20, 0);
    return fma(s, q, s - 1.0);
}
__device__ __forceinline__ double expm1_c(double x)        // accurate for x -> 0: k = 0 returns the polynomial itself
{
    int k;
    const double r = exp_reduce(fmin(fmax(x, -700.0), 700.0), k);
    const double q = expm1_poly(r);
    const double s = __hiloint2double((1023 + k) << 20, 0);
    return fma(s, q, s - 1.0);
}
#endif
#if defined(__CUDA_ARCH__)
#define NRMC_EXP(x) exp_c(x)
#define NRMC_EXPM1(x) expm1_c(x)
#else
#define NRMC_EXP(x) exp(x)
#define NRMC_EXPM1(x) expm1(x)
#endif

#define NRMC_NQ 16            // Gauss-Legendre points per half-warp slot
#ifndef NRMC_GL1_SPP
#define NRMC_GL1_SPP 32       // sub-panels per panel for the rational GL1 model where a frequency comes within 60 m of the pole of
                              // 1/(A(z) - s_f) or crosses the 1 m floor along the path; worst dense-bin deviation on wide random
                              // geometry (scratch/stress_att.py): 8 -> 1.2e-4, 16 -> 1.0e-4, 32 -> 5e-5; cost is linear in it
#define NRMC_GL1_SPP_EASY 2   // everywhere else the integrand is smooth (two sub-panels: 4e-6)
#endif
#define NRMC_MAX_SEG (NRMC_MAX_REFLECTIONS + 1)

struct AttPlan {                          // all scalars: nothing here is indexed dynamically (no local memory)
    double beta, delta, zv;               // ray invariant, n_ice - beta, apex height (may be > 0: virtual)
    double uT, u2, u1, ur;                // panel boundaries in u = sqrt(zv - z): 0: [uT,u2]  1: [u2,u1]  2: [u1,ur]
    int act0, act1, act2, na;             // active panels (multiplicity > 0 and non-empty)
    int spp, n_slots;                     // sub-panels per active panel, 16-node slots in total
    int nseg, k, rcase;
    bool turned;
};

// times path segment s runs through panel p -- equivalent to get_path_segments (py:1091-1159) plus the
// first-segment mirroring for downward starts (py:943-950)
NRMC_HD int plan_mult(const AttPlan &p, int s, int panel)
{
    if (p.k == 0) return panel == 0 ? (p.turned ? 2 : 0) : (panel == 1 ? 1 : 0);
    if (s == 0) return p.rcase == 1 ? (panel == 2 ? 1 : 2) : (panel == 2 ? 1 : 0);
    if (s < p.k) return 2;
    return panel == 0 ? (p.turned ? 2 : 0) : 1;
}

NRMC_HD int plan_total_mult(const AttPlan &p, int panel)
{
    int t = 0;
    for (int s = 0; s < p.nseg; ++s) t += plan_mult(p, s, panel);
    return t;
}

NRMC_HD void plan_slot(const AttPlan &p, int slot, double &lo, double &hi, int &panel)
{
    const int ai = slot / p.spp, sub = slot - ai * p.spp;
    panel = ai == 0 ? p.act0 : (ai == 1 ? p.act1 : p.act2);
    const double plo = panel == 0 ? p.uT : (panel == 1 ? p.u2 : p.u1);
    const double phi = panel == 0 ? p.u2 : (panel == 1 ? p.u1 : p.ur);
    const double w = (phi - plo) / p.spp;
    lo = plo + sub * w;
    hi = (sub == p.spp - 1) ? phi : plo + (sub + 1) * w;
}

NRMC_HD void att_plan_core(const IceParams &ice, double z1, double z2, int piece, int k, int rcase, double beta, double delta, double zv,
                           AttPlan &p)
{
    const bool reflected = (piece == 0 || piece == 3);    // sub pieces: the turning point is the surface
    p.turned = piece >= 2;
    p.k = k; p.rcase = rcase; p.nseg = k + 1;
    p.beta = beta;
    p.delta = delta;
    p.zv = zv;
    p.uT = reflected ? sqrt(fmax(p.zv, 0.0)) : 0.0;
    p.u2 = sqrt(fmax(p.zv - z2, 0.0));
    p.u1 = sqrt(fmax(p.zv - z1, 0.0));
    p.ur = (k > 0) ? sqrt(fmax(p.zv - ice.zr, 0.0)) : p.u1;
    p.na = 0; p.act0 = p.act1 = p.act2 = 0;
    for (int q = 0; q < 3; ++q) {
        const double plo = q == 0 ? p.uT : (q == 1 ? p.u2 : p.u1), phi = q == 0 ? p.u2 : (q == 1 ? p.u1 : p.ur);
        if (plan_total_mult(p, q) > 0 && phi > plo) {
            if (p.na == 0) p.act0 = q; else if (p.na == 1) p.act1 = q; else p.act2 = q;
            ++p.na;
        }
    }
    // every active panel is cut into `spp` equal sub-panels of 16 nodes: 1 for the entire-function models (SP1, GL2,
    // MB1: <= 2e-7 measured), 32 for the rational GL1: 1/max(A(z) - s_f, 1) has a pole next to the path and a kink at
    // the 1 m floor once f approaches 75 MHz + A(z)/0.55 m (0.8-2 GHz, beyond the model's range).  Bins above 1e-3 are
    // good to 4e-6 with two sub-panels already; what needs resolution are the strongly attenuated bins (factor 1e-6) next to
    // them, because np.interp mixes them into dense bins that are still above 1e-3.  A path with a single panel gets twice
    // as many so that no half-warp idles.
    p.spp = (ice.att_model == 2) ? NRMC_GL1_SPP_EASY : 1;      // GL1: raised to NRMC_GL1_SPP by the kernel where a frequency needs it
    if (p.na == 1) p.spp *= 2;
    p.n_slots = p.na * p.spp;
}

NRMC_HD void att_plan(const IceParams &ice, const PairGeom &g, int piece, int k, int rcase, const RayState &r, AttPlan &p)
{
    const double delta = (r.rc * r.rc) / (ice.n_ice + r.beta);      // n_ice - beta without cancellation
    att_plan_core(ice, g.z1, g.z2, piece, k, rcase, r.beta, delta, ice.z0 * log(delta / ice.dn), p);
}

NRMC_HD void att_plan_rec(const IceParams &ice, const SolRec &rec, AttPlan &p)
{
    att_plan_core(ice, rec.z1, rec.z2, rec.piece, rec.k, rec.rcase, rec.beta, rec.delta, rec.zv, p);
}

// ---------------------------------------------------------------------------------------------------------------
// attenuation length models (attenuation.py:145-262), split into a depth part (per quadrature node) and a
// frequency part (per integration frequency, precomputed on the host, staged in shared memory)
// ---------------------------------------------------------------------------------------------------------------
struct AttNode { double p0, p1, p2; };

struct Gl3Table { const double *rows; int n; };   // (depth, slope, offset) x n

NRMC_HD double gl3_lookup(const Gl3Table &t, double depth, int col)   // attenuation.py:16-33 (interp1d, clamped ends)
{
    if (depth <= t.rows[0]) return t.rows[col];
    if (depth >= t.rows[3 * (t.n - 1)]) return t.rows[3 * (t.n - 1) + col];
    int lo = 0, hi = t.n - 1;
    while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (t.rows[3 * mid] <= depth) lo = mid; else hi = mid; }
    const double x0 = t.rows[3 * lo], x1 = t.rows[3 * hi], y0 = t.rows[3 * lo + col], y1 = t.rows[3 * hi + col];
    return y0 + (y1 - y0) * (depth - x0) / (x1 - x0);
}

NRMC_HD void att_node(int model, double z, const Gl3Table &gl3, AttNode &nd)
{
    nd.p0 = nd.p1 = nd.p2 = 0.0;
    switch (model) {
    case 1: {   // SP1: temperature profile attenuation.py:141-142, b-coefficients :176-178
        const double a = fabs(z);
        const double t = ((1.83415e-09 * a - 1.59061e-08) * a + 0.00267687) * a - 51.0696;
        const double b0 = -6.74890 + t * (0.026709 - t * 0.000884);
        const double b1 = -6.22121 - t * (0.070927 + t * 0.001773);
        const double b2 = -4.09468 - t * (0.002213 + t * 0.000332);
        nd.p0 = b1;                                   // ln(1/L) at 1 GHz
        nd.p1 = (b1 - b0) * (1.0 / 9.210340371976182);  // slope in ln f below 1 GHz: (b1-b0)/(0 - ln 1e-4)
        nd.p2 = (b2 - b1) * (1.0 / 1.1505720275988207); // above 1 GHz: (b2-b1)/ln 3.16
        break; }
    case 2: {   // GL1 attenuation.py:99-128: 75 MHz length, floored at 100 m
        double L = (((( -3.63912864e-14 * z - 2.21040482e-10) * z - 3.50628312e-07) * z - 9.82378264e-05) * z + 6.87257150e-02) * z + 1.16052586e+03;
        nd.p0 = L > 100.0 ? L : 100.0;
        break; }
    case 4: {   // GL2 attenuation.py:198-204
        nd.p0 = ((((-4.58987344e-17 * z - 2.89124473e-13) * z - 5.16435542e-10) * z - 2.58901767e-07) * z + 1.58815679e-05) * z + 1.20547286e+00;
        break; }
    case 3: {   // MB1 depth factor attenuation.py:239-244
        const double d = -z * (420.0 / 576.0);
        nd.p0 = (1250.0 * 0.08886 * exp(-0.048827 * (225.6746 - 86.517596 * log10(848.870 - d)))) / 231.21;
        break; }
    case 5: {   // GL3 attenuation.py:206-222
        nd.p0 = gl3_lookup(gl3, -z, 1);
        nd.p1 = gl3_lookup(gl3, -z, 2);
        break; }
    default: break;
    }
}

// host side: the per-frequency constants (fa, fb) the device needs
NRMC_HD void att_freq_consts(int model, double f, double &fa, double &fb)
{
    fa = 0.0; fb = 0.0;
    switch (model) {
    case 1: fa = log(f); fb = (f < 1.0) ? 0.0 : 1.0; break;                    // attenuation.py:175,180-185
    case 2: fa = 0.55 * (f / 1e-3 - 75.0); break;                               // :196
    case 4: fa = 852.0 + (-0.54 / 1e-3) * f; break;                             // :200-203
    case 3: { double L = 460.0 - 180.0 * f; fa = L * (1.0 / (1.0 + L / (2.0 * 576.0) * log(0.82))); break; }  // :231-232
    case 5: fa = f; break;
    default: break;
    }
}

// 1 / L(z, f) with the 1 m floor of attenuation.py:252-255
NRMC_HD double att_inv_length(int model, const AttNode &nd, double fa, double fb)
{
    double L;
    switch (model) {
    case 1: { const double e = exp(nd.p0 + (fb != 0.0 ? nd.p2 : nd.p1) * fa); return fmin(e, 1.0); }
    case 2: L = nd.p0 - fa; break;
    case 4: L = fa * nd.p0; break;
    case 3: L = fa * nd.p0; break;
    case 5: L = nd.p0 * fa + nd.p1; break;
    default: return 0.0;
    }
    return 1.0 / fmax(L, 1.0);
}

// one quadrature node of the u-interval [lo, hi]: abscissa x in [-1,1] with weight w -> depth z and ds-weight (without 1/L)
NRMC_HD void att_node_geometry(const IceParams &ice, const AttPlan &p, double lo, double hi, double x, double w, double &z, double &wds)
{
    const double half = 0.5 * (hi - lo);
    const double u = 0.5 * (hi + lo) + half * x;
    const double uu = u * u;
    z = fmin(p.zv - uu, 0.0);
    const double em = -NRMC_EXPM1(-uu * ice.inv_z0);
    const double n = p.beta + p.delta * em;
    wds = w * half * 2.0 * u * n / sqrt(p.delta * em * (n + p.beta));
}

// 16-point Gauss-Legendre rule on [-1,1] (positive half; symmetric)
#define NRMC_GL16_X {0.0950125098376374401853193354249581, 0.2816035507792589132304605014604961, 0.4580167776572273863424194429835775, \
                     0.6178762444026437484466717640487910, 0.7554044083550030338951011948474422, 0.8656312023878317438804678977123931, \
                     0.9445750230732325760779884155346083, 0.9894009349916499325961541734503326}
#define NRMC_GL16_W {0.1894506104550684962853967232082831, 0.1826034150449235888667636679692199, 0.1691565193950025381893120790303599, \
                     0.1495959888165767320815017305474785, 0.1246289712555338720524762821920164, 0.0951585116824927848099251076022462, \
                     0.0622535239386478928628438369943776, 0.0271524594117540948517805724560181}

}  // namespace nrmc
