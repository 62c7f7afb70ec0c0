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
// with small integer exponents (mode_coeffs).  R is traced along a closed curve of four smooth pieces
//   P0: arrival up-going,   beta in (0, n_s]   (t-parametrised:  beta = n_s 2t/(1+t^2), s(0) = n_s (1-t^2)/(1+t^2))
//   P1: arrival up-going,   beta in [n_s, n2]  (parametrised by s2 = s(z2) in [0, s2max])
//   P2: arrival down-going, beta in [n_s, n2]  (refracted: apex inside the ice)
//   P3: arrival down-going, beta in (0, n_s]   (reflected off the surface)
// R starts and ends at 0 and is unimodal along the curve, so R = rho has 0 or 2 roots per mode: the three junction
// values bracket them; when all junctions are below rho the maximum is searched in the two pieces next to the
// largest junction.  All radicands are assembled from differences of gamma's (no cancellation against n_ice).
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

struct IceParams {            // medium_base.py:206-252 (+ add_reflective_bottom :47-66)
    double n_ice, dn, z0, inv_z0;
    double ns;                // n(0) = n_ice - dn
    double zr;                // reflective layer depth (<0) ; valid if n_refl > 0
    double gr, nr;            // gamma(zr), n(zr)
    int32_t n_refl;           // number of bottom reflections to consider
    int32_t att_model;        // 0 none, 1 SP1, 2 GL1, 3 MB1, 4 GL2, 5 GL3 (attenuation.py:14)
};

struct PairGeom {             // 2-D problem after set_start_and_end_point (py:2057-2090): z1 <= z2
    double z1, z2, rho;
    double g1, g2, n1, n2;    // gamma and n at z1, z2
    // radicand offsets (see eval_range)
    double A1, A2, Ar;        // (dn - gamma_i)(n_i + ns):  s_i^2 = A_i + sigma_s^2   ("sub" pieces)
    double B1, Br;            // (g2 - gamma_i)(n_i + n2):  s_i^2 = B_i + s2^2        ("band" pieces)
    double c0_sub, c0_band;   // c = c0 + sigma^2
    double s2max;             // sqrt(n2^2 - ns^2)
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

NRMC_HD void make_pair_geom(const IceParams &ice, double z1, double z2, double rho, PairGeom &g)
{
    g.z1 = z1; g.z2 = z2; g.rho = rho;
    g.g1 = ice.dn * exp(z1 * ice.inv_z0);
    g.g2 = ice.dn * exp(z2 * ice.inv_z0);
    g.n1 = ice.n_ice - g.g1; g.n2 = ice.n_ice - g.g2;
    g.A1 = (ice.dn - g.g1) * (g.n1 + ice.ns);
    g.A2 = (ice.dn - g.g2) * (g.n2 + ice.ns);
    g.Ar = (ice.dn - ice.gr) * (ice.nr + ice.ns);
    g.B1 = (g.g2 - g.g1) * (g.n1 + g.n2);
    g.Br = (g.g2 - ice.gr) * (ice.nr + g.n2);
    g.c0_sub = ice.dn * (ice.n_ice + ice.ns);
    g.c0_band = g.g2 * (ice.n_ice + g.n2);
    g.s2max = sqrt(fmax(g.A2, 0.0));
}

// Quantities of one ray (one value of beta) that the property formulas need.
struct RayState {
    double beta, rc;          // Snell invariant, sqrt(n_ice^2 - beta^2)
    double s1, s2, sr, ss;    // n cos(theta) at z1, z2, zr and at the surface (ss = 0 for refracted rays)
    double k1_1, k1_2, k1_r, KT;
    bool reflected;           // turning point is the surface (beta <= n_s)
};

// Evaluate the ray state for a point of the curve.  sub: v = t in [0,1]; band: v = s2 in [0, s2max].
NRMC_HD void ray_state(const IceParams &ice, const PairGeom &g, bool band, double v, RayState &r)
{
    double c;
    if (!band) {
        double q = 1.0 / (1.0 + v * v);
        r.beta = ice.ns * (2.0 * v) * q;
        r.ss = ice.ns * ((1.0 - v) * (1.0 + v)) * q;
        double sg2 = r.ss * r.ss;
        c = g.c0_sub + sg2;
        r.s1 = sqrt(g.A1 + sg2);
        r.s2 = sqrt(g.A2 + sg2);
        r.sr = ice.n_refl > 0 ? sqrt(g.Ar + sg2) : 0.0;
        r.reflected = true;
    } else {
        double sg2 = v * v;
        r.beta = sqrt(fmax(g.n2 * g.n2 - sg2, 0.0));
        r.ss = 0.0;
        c = g.c0_band + sg2;
        r.s1 = sqrt(g.B1 + sg2);
        r.s2 = v;
        r.sr = ice.n_refl > 0 ? sqrt(g.Br + sg2) : 0.0;
        r.reflected = false;
    }
    r.rc = sqrt(c);
    r.k1_1 = r.rc * r.s1 + (c - ice.n_ice * g.g1);
    r.k1_2 = r.rc * r.s2 + (c - ice.n_ice * g.g2);
    r.k1_r = ice.n_refl > 0 ? r.rc * r.sr + (c - ice.n_ice * ice.gr) : 1.0;
    r.KT = r.reflected ? r.rc * r.ss + (c - ice.n_ice * ice.dn) : ice.dn * r.beta;
}

// horizontal range of the path family m for the ray r
NRMC_HD double range_of(const IceParams &ice, const PairGeom &g, const ModeCoeffs &m, const RayState &r)
{
    double num = 1.0, den = 1.0;
    if (m.a1 > 0) num *= r.k1_1; else den *= r.k1_1;            // |a1| == 1
    if (m.a2 > 0) num *= r.k1_2; else den *= r.k1_2;            // |a2| == 1
    num *= ipow(r.KT, m.aT);
    double lin = m.a1 * g.z1 + m.a2 * g.z2;
    if (m.ar != 0) { den *= ipow(r.k1_r, -m.ar); lin += m.ar * ice.zr; }
    double R = (r.beta / r.rc) * (lin - ice.z0 * log(num / den));
    return R;
}

struct Curve {                // one (reflection, case) mode of one pair
    const IceParams *ice;
    const PairGeom *g;
    ModeCoeffs m_dir, m_trn;
};

// piece p in {0,1,2,3}; value of R - rho at parameter v of that piece
NRMC_HD double curve_g(const Curve &cv, int p, double v)
{
#ifdef NRMC_COUNT_EVALS
    ++g_evals;
#endif
    RayState r;
    ray_state(*cv.ice, *cv.g, (p == 1 || p == 2), v, r);
    double R = range_of(*cv.ice, *cv.g, (p >= 2) ? cv.m_trn : cv.m_dir, r);
    if (!(R == R)) R = 1e300;  // horizontal ray in (numerically) homogeneous ice: infinite range
    return R - cv.g->rho;
}

// g = R - rho and its derivative with respect to the piece parameter v (closed form; used to locate the maximum of R
// when the whole curve was sampled below rho).
NRMC_HD double curve_gd(const Curve &cv, int p, double v, double &dg)
{
#ifdef NRMC_COUNT_EVALS
    ++g_evals;
#endif
    const IceParams &ice = *cv.ice;
    const PairGeom &g = *cv.g;
    const bool band = (p == 1 || p == 2);
    const ModeCoeffs &m = (p >= 2) ? cv.m_trn : cv.m_dir;
    RayState r;
    ray_state(ice, g, band, v, r);
    double db, dc, ds1, ds2, dsr, dKT;     // derivatives of beta, c, s1, s2, sr, KT with respect to v
    if (!band) {
        const double q = 1.0 / (1.0 + v * v);
        db = 2.0 * r.ss * q;
        const double dss = -2.0 * r.beta * q;
        const double h = r.ss * dss;        // d(ss^2)/dv / 2
        dc = 2.0 * h;
        ds1 = r.s1 > 0 ? h / r.s1 : 0.0; ds2 = r.s2 > 0 ? h / r.s2 : 0.0; dsr = r.sr > 0 ? h / r.sr : 0.0;
        const double drc = h / r.rc;
        dKT = drc * r.ss + r.rc * dss + dc;
    } else {
        db = -v / r.beta;
        dc = 2.0 * v;
        ds1 = r.s1 > 0 ? v / r.s1 : 0.0; ds2 = 1.0; dsr = r.sr > 0 ? v / r.sr : 0.0;
        dKT = ice.dn * db;
    }
    const double drc = 0.5 * dc / r.rc;
    const double dk1 = drc * r.s1 + r.rc * ds1 + dc, dk2 = drc * r.s2 + r.rc * ds2 + dc;
    double dlnP = m.a1 * dk1 / r.k1_1 + m.a2 * dk2 / r.k1_2 + (m.aT ? m.aT * dKT / r.KT : 0.0);
    if (m.ar != 0) dlnP += m.ar * (drc * r.sr + r.rc * dsr + dc) / r.k1_r;
    const double A = r.beta / r.rc;
    const double dA = (db - A * drc) / r.rc;
    double num = 1.0, den = 1.0;
    if (m.a1 > 0) num *= r.k1_1; else den *= r.k1_1;
    if (m.a2 > 0) num *= r.k1_2; else den *= r.k1_2;
    num *= ipow(r.KT, m.aT);
    double lin = m.a1 * g.z1 + m.a2 * g.z2;
    if (m.ar != 0) { den *= ipow(r.k1_r, -m.ar); lin += m.ar * ice.zr; }
    const double Bk = lin - ice.z0 * log(num / den);
    double R = A * Bk;
    dg = dA * Bk - A * ice.z0 * dlnP;
    if (!(R == R)) { R = 1e300; dg = 0.0; }
    return R - g.rho;
}

// Bracketed root of g on piece p between a and b (ga, gb of opposite strict sign): regula falsi with the Illinois
// modification, bisection safeguard; converges superlinearly on the smooth pieces.
NRMC_HD double solve_piece(const Curve &cv, int p, double a, double ga, double b, double gb)
{
    const double gtol = 1e-10;
    int side = 0;
    double x = a;
    for (int it = 0; it < 100; ++it) {
        double denom = gb - ga;
        x = (a * gb - b * ga) / denom;
        double lo = fmin(a, b), hi = fmax(a, b);
        if (!(x > lo && x < hi)) x = 0.5 * (a + b);
        double gx = curve_g(cv, p, x);
        if (fabs(gx) <= gtol) break;
        if ((gx > 0) == (gb > 0)) { b = x; gb = gx; if (side == 1) ga *= 0.5; side = 1; }
        else { a = x; ga = gx; if (side == -1) gb *= 0.5; side = -1; }
        if (fabs(b - a) <= 4e-16 * (fabs(a) + fabs(b))) { x = (fabs(ga) < fabs(gb)) ? a : b; break; }
    }
    return x;
}

// Look for an interior maximum of g on piece p (end values <= 0): bracket the sign change of dg/dv and close it
// with Illinois steps on the derivative.  Returns true as soon as a point with g > 0 is found (xm, gm); false if the
// maximum is at an end point or stays <= 0.  `interior` tells the caller whether this piece holds the curve's maximum.
NRMC_HD bool maximise_piece(const Curve &cv, int p, double a, double b, double &xm, double &gm, bool &interior)
{
    double lo = fmin(a, b), hi = fmax(a, b);
    double dlo, dhi;
    double glo = curve_gd(cv, p, lo, dlo);
    double ghi = curve_gd(cv, p, hi, dhi);
    interior = (dlo > 0.0) && (dhi < 0.0);
    xm = lo; gm = glo;
    if (!interior) return false;
    int side = 0;
    for (int it = 0; it < 60; ++it) {
        double x = (lo * dhi - hi * dlo) / (dhi - dlo);
        if (!(x > lo && x < hi)) x = 0.5 * (lo + hi);
        double dx;
        const double gx = curve_gd(cv, p, x, dx);
        if (gx > gm || it == 0) { xm = x; gm = gx; }
        if (gx > 0.0) return true;
        if (dx > 0.0) { lo = x; dlo = dx; if (side == 1) dhi *= 0.5; side = 1; }
        else if (dx < 0.0) { hi = x; dhi = dx; if (side == -1) dlo *= 0.5; side = -1; }
        else break;
        // the remaining gain is bounded by |slope| x bracket once the bracket lies in the concave cap around the maximum
        const double gain = fmax(fabs(dlo), fabs(dhi)) * (hi - lo);
        if (gain < 1e-11 || (hi - lo) <= 4e-16 * (fabs(lo) + fabs(hi))) break;
        if (it >= 2 && gm + 4.0 * gain < 0.0) break;     // deep in the shadow zone: cannot reach rho any more
    }
    return gm > 0.0;
}

struct Root { double v; int piece; double beta; };

// All roots (0 or 2; 1 only on a tangency) of one mode, ordered by increasing C0 = 1/beta.
// Written so that the solver and the maximum search each have ONE call site: everything inlines into the kernel and
// the pair geometry stays in registers.
NRMC_HD int find_roots_mode(const IceParams &ice, const PairGeom &g, int k, int rcase, Root out[2])
{
    Curve cv;
    cv.ice = &ice; cv.g = &g;
    cv.m_dir = mode_coeffs(k, rcase, false);
    cv.m_trn = mode_coeffs(k, rcase, true);
    const bool has_band = g.s2max > 0.0;
    // piece end points (parameter values) in curve order: P0 t:0->1, P1 s2:s2max->0, P2 s2:0->s2max, P3 t:1->0
    const double J0 = -g.rho, J4 = -g.rho;
    const double J1 = curve_g(cv, 0, 1.0);
    const double J3 = curve_g(cv, 3, 1.0);
    const double J2 = has_band ? curve_g(cv, 1, 0.0) : J1;
    // brackets: (piece, a, g(a), b, g(b)), at most two
    int bp0 = 0, bp1 = 0, nb = 0;
    double ba0 = 0, bga0 = 0, bb0 = 0, bgb0 = 0, ba1 = 0, bga1 = 0, bb1 = 0, bgb1 = 0;
#define NRMC_PUSH_BRACKET(P, A, GA, B, GB)                                               \
    do {                                                                                  \
        if (nb == 0) { bp0 = (P); ba0 = (A); bga0 = (GA); bb0 = (B); bgb0 = (GB); }       \
        else if (nb == 1) { bp1 = (P); ba1 = (A); bga1 = (GA); bb1 = (B); bgb1 = (GB); }  \
        ++nb;                                                                             \
    } while (0)
    // receiver exactly at the surface (no band): P0 and P3 coincide (py: one 'reflected' solution) -> only P3
    if (has_band && ((J0 > 0) != (J1 > 0))) NRMC_PUSH_BRACKET(0, 0.0, J0, 1.0, J1);
    if (has_band && ((J1 > 0) != (J2 > 0))) NRMC_PUSH_BRACKET(1, g.s2max, J1, 0.0, J2);
    if (has_band && ((J2 > 0) != (J3 > 0))) NRMC_PUSH_BRACKET(2, 0.0, J2, g.s2max, J3);
    if ((J3 > 0) != (J4 > 0)) NRMC_PUSH_BRACKET(3, 1.0, J3, 0.0, J4);
    if (nb == 0 && J1 <= 0 && J2 <= 0 && J3 <= 0) {
        // whole curve sampled below rho: look for a hump inside the pieces adjacent to the largest junction
        int jm = 1;
        double Jm = J1;
        if (J2 > Jm) { jm = 2; Jm = J2; }
        if (J3 > Jm) { jm = 3; Jm = J3; }
        bool interior = false;
        for (int side = 0; side < 2 && nb == 0 && !interior; ++side) {
            const int p = jm - 1 + side;
            if (!has_band && p != 3) continue;
            const double pa = (p == 1) ? g.s2max : ((p == 3) ? 1.0 : 0.0);
            const double pb = (p == 0) ? 1.0 : ((p == 2) ? g.s2max : 0.0);
            const double ja = (p == 0) ? J0 : (p == 1 ? J1 : (p == 2 ? J2 : J3));
            const double jb = (p == 0) ? J1 : (p == 1 ? J2 : (p == 2 ? J3 : J4));
            double xm, gm;
            if (maximise_piece(cv, p, pa, pb, xm, gm, interior)) {
                NRMC_PUSH_BRACKET(p, pa, ja, xm, gm);
                NRMC_PUSH_BRACKET(p, xm, gm, pb, jb);
            }
        }
    }
#undef NRMC_PUSH_BRACKET
    if (nb > 2) nb = 2;
    Root r0, r1;
    r0.v = r1.v = 0; r0.piece = r1.piece = 0; r0.beta = r1.beta = 0;
    for (int i = 0; i < nb; ++i) {
        const int p = i == 0 ? bp0 : bp1;
        const double v = solve_piece(cv, p, i == 0 ? ba0 : ba1, i == 0 ? bga0 : bga1, i == 0 ? bb0 : bb1, i == 0 ? bgb0 : bgb1);
        RayState rs;
        ray_state(ice, g, (p == 1 || p == 2), v, rs);
        if (i == 0) { r0.v = v; r0.piece = p; r0.beta = rs.beta; } else { r1.v = v; r1.piece = p; r1.beta = rs.beta; }
    }
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
    const double A = r.beta / r.rc;
    o.C0 = 1.0 / r.beta;
    // C_1 = y1 - y(z1; C_1 = 0),  y = z0 beta/sqrt(c) ln(gamma / (2 k1))       (py:487-491,118-125)
    o.C1 = x1y - ice.z0 * A * log(g.g1 / (2.0 * r.k1_1));
    // solution type on the UNREFLECTED geometry (py:2146 -> :1386-1398): direct iff rho < y_turn - y1
    if (k == 0) o.type = turned ? (r.reflected ? 3 : 2) : 1;
    else {
        double T1 = A * (-g.z1 - ice.z0 * log(r.KT / r.k1_1));
        o.type = (g.rho < T1) ? 1 : (r.reflected ? 3 : 2);
    }
    // launch: theta1 with sin = beta/n1, downward start (case 2, k>0) -> pi - theta1   (py:1161-1196)
    o.sin_l = r.beta / g.n1;
    o.cos_l = r.s1 / g.n1;
    if (k > 0 && rcase == 2) o.cos_l = -o.cos_l;
    // receive: pi - theta2 if the last segment arrives up-going, theta2 otherwise      (py:1198-1199)
    o.sin_r = r.beta / g.n2;
    o.cos_r = turned ? r.s2 / g.n2 : -r.s2 / g.n2;
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
    double lk2 = log(num / den);
    double sumU = g.rho / A;
    o.path_length = ice.n_ice / r.rc * sumU + ice.z0 * lk2;
    o.travel_time = (ice.n_ice * ice.n_ice / r.rc * sumU + ice.z0 * (ssum + ice.n_ice * lk2)) / NRMC_SPEED_OF_LIGHT;
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
};

struct SolRec {                // what the attenuation kernel needs to rebuild the ray
    double v;                  // curve parameter of the root
    int64_t pair;
    int32_t slot;
    uint8_t piece, k, rcase, pad;
};

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
    f.rho = sqrt(dx * dx + dy * dy);
    if (f.rho > 0) { f.ex = dx / f.rho; f.ey = dy / f.rho; } else { f.ex = 1.0; f.ey = 0.0; }  // atan2(0,0) = 0
    f.z1 = az; f.z2 = bz; f.x1y = ax;
}

// Returns the number of solutions; fills the SoA slots of pair i and (if recs != null) one SolRec per solution.
NRMC_HD int trace_pair(const IceParams &ice, double ax, double ay, double az, double bx, double by, double bz,
                        int64_t i, const TraceOutputs &o, SolRec *recs)
{
    const int S = 2 + 4 * ice.n_refl, K1 = ice.n_refl + 1;
    int status = 0, n = 0;
    Frame2D f;
    make_frame(ax, ay, az, bx, by, bz, f);
    if (!(f.rho == f.rho) || !(f.z1 == f.z1) || !(f.z2 == f.z2) || isinf(f.rho) || isinf(f.z1)) status |= NRMC_STATUS_NONFINITE;
    else if (f.z2 > 0.0) status |= NRMC_STATUS_AIR;
    else if (ice.n_refl > 0 && f.z1 < ice.zr) status |= NRMC_STATUS_BELOW_REFLECTOR;
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
                const int64_t q = i * S + n;
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
                    for (int s = 0; s < K1; ++s)
                        o.reflection_angle[q * K1 + s] = ((p.refl_mask >> s) & 1u) ? p.refl_angle : NAN;
                if (recs) { recs[n].v = roots[j].v; recs[n].pair = i; recs[n].slot = n; recs[n].piece = (uint8_t)roots[j].piece;
                            recs[n].k = (uint8_t)k; recs[n].rcase = (uint8_t)rcase; recs[n].pad = 0; }
                ++n;
            }
        }
    }
    if (o.n_sol) o.n_sol[i] = n;
    if (o.status) o.status[i] = status;
    for (int s = n; s < S; ++s) {   // empty slots: 0 / NaN (HDF5 writer convention, output_writer_hdf5.py:272-275)
        const int64_t q = i * S + s;
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
    }
    return n;
}

}  // namespace nrmc
