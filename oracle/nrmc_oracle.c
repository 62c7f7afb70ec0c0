/*
 * nrmc_oracle.c -- TEST INFRASTRUCTURE.  CPU restatement (plain C, scalar FP64) of the NuRadioMC analytic ray
 * tracer, used ONLY as the parity checker by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs.  The product (nuradiomc_b200/, libnrmc_rt.so) never links, imports or calls it.
 *
 * What it restates (all citations relative to /root/reference/, "py" = NuRadioMC/SignalProp/analyticraytracing.py,
 * "att" = NuRadioMC/utilities/attenuation.py, "base" = NuRadioMC/SignalProp/propagation_base_class.py):
 *   - the closed-form ray functions py:99-370, C_1 / mirroring py:487-511
 *   - the objective get_delta_y py:204-272 (with the F4 fix: x1 is copied, never mutated)
 *   - solution typing py:1365-1398, path segments py:1091-1159, angles py:1161-1237
 *   - analytic path length / travel time py:602-783
 *   - sparse attenuation frequencies py:885-931, attenuation integral py:933-1089, L(z,f) att:99-262
 *   - the 3-D wrapper: geometry py:2057-2090, mode loop py:2118-2130, vectors py:2560-2624
 *
 * Third-party numerics.  The reference delegates root finding and quadrature to scipy (pyproject.toml:23 pins
 * scipy="*"; 1.18.1 in the build image): optimize.root(hybr) + optimize.brentq (py:1479,1504,1526) and
 * integrate.quad (py:1071).  They are restated here from their published algorithms:
 *   - roots: the objective has the shape "negative at both ends, one positive hump" (SURVEY.md App. A); the
 *     reference locates one root with hybr on f^2 and then runs Brent on (root+1e-4, 100) and (-100, root-1e-4).
 *     Which root hybr lands on is chaotic near the shadow boundary, so the oracle implements the
 *     F6 *arbiter* instead: a scan of the reference's own objective in logC0, refinement of every negative
 *     local maximum, and Brent (Brent 1973 "zeroin", as in scipy.optimize.brentq) on every sign change.
 *   - quadrature: adaptive bisection with the 21-point Gauss-Kronrod rule and QUADPACK's error heuristic
 *     (Piessens et al. 1983, routine QAG/QK21), break point at the turning point as with quad(points=[z_turn]).
 *     mode "reference": the reference's variable and tolerance (epsrel=1e-2, py:1071-1072);
 *     mode "tight": same integrand ds/L, legs that end on a turning point are substituted t = z_turn -+ u^2
 *     (removes the integrable 1/sqrt singularity), epsrel=1e-11 -- the "tightened oracle" of SURVEY.md F5.
 *
 * Parity pinning: tests/test_oracle_golden.py checks this file against the reference's own goldens
 * reference_C0.pkl / reference_C0_MooresBay.pkl (T05, T06) and against fixtures produced by running the
 * reference's Python path in the build container (tests/golden/make_golden.py).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define ORA_MAX_ROOTS 4      /* per mode; the reference assumes <= 2, more are reported so tests can assert */
#define ORA_MAX_SEG   8
#define ORA_SPEED_OF_LIGHT 0.299792458 /* m/ns, scipy.constants.c * units.m/units.s, py:56 */

typedef struct {
    double n_ice, delta_n, z_0;
    double reflection;      /* z of the reflective bottom layer [m]; NAN if the medium has none */
    int32_t att_model;      /* att:14 : SP1=1 GL1=2 MB1=3 GL2=4 GL3=5 ; 0 = no attenuation */
    int32_t n_reflections;
    int32_t n_freq;         /* n_frequencies_integration (base:118-126 default 100) */
    int32_t quad_mode;      /* 0 = "reference" (epsrel 1e-2 in z), 1 = "tight" (substitution, epsrel 1e-11) */
    int32_t scan_n;         /* number of scan points in logC0 (0 -> default 4401) */
    double scan_lo, scan_hi;/* scan window in logC0 (0,0 -> default [-22, 22]) */
    const double *gl3;      /* 300x3 row-major table (depth, slope, offset) or NULL */
    int32_t gl3_rows;
    int32_t pad_;
} ora_cfg;

/* ------------------------------------------------------------------------------------------------ */
/* closed-form ray functions (py:99-370)                                                             */
/* ------------------------------------------------------------------------------------------------ */

static double get_C0_from_log(double logC0, double n_ice) { return exp(logC0) + 1. / n_ice; }           /* py:99 */
static double get_gamma(double z, const ora_cfg *m) { return m->delta_n * exp(z / m->z_0); }             /* py:127 */
static double n_of_z(double z, const ora_cfg *m) { return m->n_ice - m->delta_n * exp(z / m->z_0); }     /* py:358 */

static double get_y(double gamma, double C_0, double C_1, const ora_cfg *m)                              /* py:105 */
{
    double b = 2 * m->n_ice;
    double c = m->n_ice * m->n_ice - pow(C_0, -2);
    double root = fabs(gamma * gamma - gamma * b + c);
    double logargument = gamma / (2 * sqrt(c) * sqrt(root) - b * gamma + 2 * c);
    return m->z_0 * pow(m->n_ice * m->n_ice * C_0 * C_0 - 1, -0.5) * log(logargument) + C_1;
}

static void get_turning_point(double c, const ora_cfg *m, double *gamma2, double *z2)                    /* py:133 */
{
    double b = 2 * m->n_ice;
    double g = b * 0.5 - sqrt(0.25 * b * b - c);
    double z = log(g / m->delta_n) * m->z_0;
    if (z > 0) { z = 0; g = get_gamma(0, m); }
    *gamma2 = g; *z2 = z;
}

static double get_y_with_z_mirror(double z, double C_0, double C_1, const ora_cfg *m)                    /* py:160 */
{
    double c = m->n_ice * m->n_ice - pow(C_0, -2);
    double gamma_turn, z_turn;
    get_turning_point(c, m, &gamma_turn, &z_turn);
    double y_turn = get_y(gamma_turn, C_0, C_1, m);
    if (z < z_turn) return get_y(get_gamma(z, m), C_0, C_1, m);
    return 2 * y_turn - get_y(get_gamma(2 * z_turn - z, m), C_0, C_1, m);
}

static double get_C_1(const double x1[2], double C_0, const ora_cfg *m)                                  /* py:487 */
{
    return x1[0] - get_y_with_z_mirror(x1[1], C_0, 0.0, m);
}

static double get_y_turn(double C_0, const double x1[2], const ora_cfg *m)                               /* py:186 */
{
    double c = m->n_ice * m->n_ice - pow(C_0, -2);
    double gamma_turn, z_turn;
    get_turning_point(c, m, &gamma_turn, &z_turn);
    double C_1 = get_C_1(x1, C_0, m);
    return get_y(gamma_turn, C_0, C_1, m);
}

static void get_reflection_point(double C_0, double C_1, const ora_cfg *m, double out[2])                /* py:281 */
{
    double c = m->n_ice * m->n_ice - pow(C_0, -2);
    double gamma_turn, z_turn;
    get_turning_point(c, m, &gamma_turn, &z_turn);
    out[1] = m->reflection;
    out[0] = get_y_with_z_mirror(-m->reflection + 2 * z_turn, C_0, C_1, m);
}

/* py:204-272.  x1 is taken by value (F4: the reference mutates its argument for case 2; the harness patch and the
 * reference's C++ path work on a copy, and so does this restatement). */
static double get_delta_y(double C_0, const double x1_in[2], const double x2[2], const ora_cfg *m,
                          int reflection, int reflection_case)
{
    double x1[2] = {x1_in[0], x1_in[1]};
    if (C_0 < 1. / m->n_ice) return -INFINITY;
    double c = m->n_ice * m->n_ice - pow(C_0, -2);
    if (reflection > 0 && reflection_case == 2) {
        double y_turn = get_y_turn(C_0, x1, m);
        double dy = y_turn - x1[0];
        x1[0] = x1[0] - 2.0 * dy;
    }
    for (int i = 0; i < reflection; ++i) {
        double C_1 = get_C_1(x1, C_0, m);
        double p[2];
        get_reflection_point(C_0, C_1, m, p);
        x1[0] = p[0]; x1[1] = p[1];
    }
    double C_1 = get_C_1(x1, C_0, m);
    double gamma_turn, z_turn;
    get_turning_point(c, m, &gamma_turn, &z_turn);
    double y_turn = get_y(gamma_turn, C_0, C_1, m);
    if (z_turn < x2[1]) {
        double dz = z_turn - x2[1], dy = y_turn - x2[0];
        return -(sqrt(dz * dz + dy * dy) + 10 * fabs(dz));
    }
    if (y_turn > x2[0]) {
        double y2_fit = get_y(get_gamma(x2[1], m), C_0, C_1, m);
        return x2[0] - y2_fit;
    } else {
        double y2_raw = get_y(get_gamma(x2[1], m), C_0, C_1, m);
        double y2_fit = 2 * y_turn - y2_raw;
        return -(x2[0] - y2_fit);
    }
}

static double get_z_unmirrored(double z, double C_0, const ora_cfg *m)                                   /* py:293 */
{
    double c = m->n_ice * m->n_ice - pow(C_0, -2);
    double gamma_turn, z_turn;
    get_turning_point(c, m, &gamma_turn, &z_turn);
    return (z > z_turn) ? 2 * z_turn - z : z;
}

static double get_y_diff(double z_raw, double C_0, const ora_cfg *m)                                     /* py:306 (in_air=False) */
{
    double z = get_z_unmirrored(z_raw, C_0, m);
    double n_z = n_of_z(z, m);
    double res;
    if (C_0 * C_0 * n_z * n_z > 1) res = 1 / sqrt(C_0 * C_0 * n_z * n_z - 1);
    else res = INFINITY;
    if (z != z_raw) res *= -1;
    return res;
}

static void get_z_mirrored(const double x1[2], const double x2[2], double C_0, const ora_cfg *m, double out[2]) /* py:496 */
{
    double c = m->n_ice * m->n_ice - pow(C_0, -2);
    double C_1 = get_C_1(x1, C_0, m);
    double gamma_turn, z_turn;
    get_turning_point(c, m, &gamma_turn, &z_turn);
    double y_turn = get_y(gamma_turn, C_0, C_1, m);
    double zstart = x1[1], zstop = x2[1];
    if (y_turn < x2[0]) zstop = zstart + fabs(z_turn - x1[1]) + fabs(z_turn - x2[1]);
    out[0] = x2[0]; out[1] = zstop;
}

static double ds_of_t(double t, double C_0, const ora_cfg *m)                                            /* py:513 */
{
    double d = get_y_diff(t, C_0, m);
    return sqrt(d * d + 1);
}

static int determine_solution_type(const double x1[2], const double x2[2], double C_0, const ora_cfg *m) /* py:1365 */
{
    double c = m->n_ice * m->n_ice - pow(C_0, -2);
    double C_1 = get_C_1(x1, C_0, m);
    double gamma_turn, z_turn;
    get_turning_point(c, m, &gamma_turn, &z_turn);
    double y_turn = get_y(gamma_turn, C_0, C_1, m);
    if (x2[0] < y_turn) return 1;
    if (z_turn == 0) return 3;
    return 2;
}

typedef struct { double x1_orig[2], x1[2], x2_orig[2], x2[2], C_0, C_1; } ora_segment;

static int get_path_segments(const double x1_in[2], const double x2_in[2], double C_0, const ora_cfg *m,
                             int reflection, int reflection_case, ora_segment *seg)                      /* py:1091 */
{
    double x1[2] = {x1_in[0], x1_in[1]}, x2[2] = {x2_in[0], x2_in[1]};
    if (reflection == 0) {
        ora_segment s = {{x1_in[0], x1_in[1]}, {x1[0], x1[1]}, {x2_in[0], x2_in[1]}, {x2[0], x2[1]}, C_0, get_C_1(x1, C_0, m)};
        seg[0] = s;
        return 1;
    }
    if (reflection_case == 2) {
        double y_turn = get_y_turn(C_0, x1, m);
        double dy = y_turn - x1[0];
        x1[0] = x1[0] - 2 * dy;
    }
    int n = 0;
    for (int i = 0; i < reflection + 1 && n < ORA_MAX_SEG; ++i) {
        double C_1 = get_C_1(x1, C_0, m);
        get_reflection_point(C_0, C_1, m, x2);
        int stop_loop = 0;
        if (x2[0] > x2_in[0]) { stop_loop = 1; x2[0] = x2_in[0]; x2[1] = x2_in[1]; }
        ora_segment s = {{x1_in[0], x1_in[1]}, {x1[0], x1[1]}, {x2_in[0], x2_in[1]}, {x2[0], x2[1]}, C_0, C_1};
        seg[n++] = s;
        if (stop_loop) break;
        x1[0] = x2[0]; x1[1] = x2[1];
    }
    return n;
}

static double get_angle(const double x[2], const double x_start_in[2], double C_0, const ora_cfg *m,
                        int reflection, int reflection_case)                                             /* py:1161 */
{
    ora_segment seg[ORA_MAX_SEG];
    int n = get_path_segments(x_start_in, x, C_0, m, reflection, reflection_case, seg);
    const double *x_start = seg[n - 1].x1;
    double zm[2];
    get_z_mirrored(x_start, x, C_0, m, zm);
    double dy = get_y_diff(zm[1], C_0, m);
    double angle = atan(dy);
    if (angle < 0) angle = M_PI + angle;
    return angle;
}

static double get_launch_angle(const double x1[2], double C_0, const ora_cfg *m, int reflection, int reflection_case)
{ return get_angle(x1, x1, C_0, m, reflection, reflection_case); }                                       /* py:1195 */

static double get_receive_angle(const double x1[2], const double x2[2], double C_0, const ora_cfg *m, int reflection, int reflection_case)
{ return M_PI - get_angle(x2, x1, C_0, m, reflection, reflection_case); }                                /* py:1198 */

/* py:1201-1237; out[i] = angle or NAN (None) per segment; returns number of segments */
static int get_reflection_angle(const double x1[2], const double x2[2], double C_0, const ora_cfg *m,
                                int reflection, int reflection_case, double *out)
{
    ora_segment seg[ORA_MAX_SEG];
    double c = m->n_ice * m->n_ice - pow(C_0, -2);
    int n = get_path_segments(x1, x2, C_0, m, reflection, reflection_case, seg);
    for (int i = 0; i < n; ++i) {
        double gamma_turn, z_turn;
        get_turning_point(c, m, &gamma_turn, &z_turn);
        double y_turn = get_y_turn(C_0, seg[i].x1, m);
        if (z_turn >= 0 && y_turn > seg[i].x1_orig[0] && y_turn < seg[i].x2_orig[0]) {
            double p[2] = {y_turn, 0};
            out[i] = get_angle(p, seg[i].x1, C_0, m, 0, 1);
        } else out[i] = NAN;
    }
    return n;
}

/* first-segment mirroring for downward-starting rays, shared by py:630-639, :720-729, :943-952 */
static void segment_endpoints(const ora_segment *s, int iS, int reflection_case, double x1[2], double x2[2])
{
    if (iS == 0 && reflection_case == 2) {
        x1[0] = s->x1_orig[0]; x1[1] = s->x2[1];
        x2[0] = s->x2[0];      x2[1] = s->x1_orig[1];
    } else {
        x1[0] = s->x1[0]; x1[1] = s->x1[1];
        x2[0] = s->x2[0]; x2[1] = s->x2[1];
    }
}

/* py:602-690 (path length, which=0) and py:692-783 (travel time, which=1) */
static double path_or_time_analytic(const double x1_in[2], const double x2_in[2], double C_0, const ora_cfg *m,
                                    int reflection, int reflection_case, int which)
{
    ora_segment seg[ORA_MAX_SEG];
    int nseg = get_path_segments(x1_in, x2_in, C_0, m, reflection, reflection_case, seg);
    double n_ice = m->n_ice, z_0 = m->z_0;
    double acc = 0;
    for (int iS = 0; iS < nseg; ++iS) {
        double x1[2], x2[2];
        segment_endpoints(&seg[iS], iS, reflection_case, x1, x2);
        double z1 = x1[1], z2 = x2[1];
        int solution_type = determine_solution_type(x1, x2, C_0, m);
        double launch_angle = get_launch_angle(x1, C_0, m, reflection, reflection_case);
        double beta = n_of_z(x1[1], m) * sin(launch_angle);
        double alpha = n_ice * n_ice - beta * beta;
#define ORA_GAMMA(z) fmax(0.0, n_of_z((z), m) * n_of_z((z), m) - beta * beta)
#define ORA_L1(z) (sqrt(alpha * ORA_GAMMA(z)) + n_ice * n_of_z((z), m) - beta * beta)
#define ORA_L2(z) (sqrt(ORA_GAMMA(z)) + n_of_z((z), m))
#define ORA_S(z) (n_ice / sqrt(alpha) * ((z) - z_0 * log(ORA_L1(z))) + z_0 * log(ORA_L2(z)))
#define ORA_CT(z) (z_0 * (sqrt(ORA_GAMMA(z)) - n_ice * n_ice / sqrt(alpha) * log(ORA_L1(z)) + n_ice * log(ORA_L2(z))) + n_ice * n_ice * (z) / sqrt(alpha))
#define ORA_F(z) (which ? ORA_CT(z) : ORA_S(z))
        if (solution_type == 1) acc += ORA_F(z2) - ORA_F(z1);
        else {
            double z_turn = 0;
            if (solution_type != 3) {
                double gamma_turn;
                get_turning_point(n_ice * n_ice - pow(C_0, -2), m, &gamma_turn, &z_turn);
            }
            acc += 2 * ORA_F(z_turn) - ORA_F(z1) - ORA_F(z2);
        }
    }
    return which ? acc / ORA_SPEED_OF_LIGHT : acc;
}

/* ------------------------------------------------------------------------------------------------ */
/* attenuation length L(z,f)  (att:99-262; scalar branch)                                            */
/* ------------------------------------------------------------------------------------------------ */

static double gl3_interp(const ora_cfg *m, double x, int col)                                            /* att:16-33 */
{
    const double *t = m->gl3; int n = m->gl3_rows;
    if (x <= t[0]) return t[col];
    if (x >= t[3 * (n - 1)]) return t[3 * (n - 1) + col];
    int lo = 0, hi = n - 1;
    while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (t[3 * mid] <= x) lo = mid; else hi = mid; }
    double x0 = t[3 * lo], x1 = t[3 * hi];
    double y0 = t[3 * lo + col], y1 = t[3 * hi + col];
    return y0 + (y1 - y0) * (x - x0) / (x1 - x0);
}

double ora_attenuation_length(const ora_cfg *m, double z, double frequency)
{
    double L;
    switch (m->att_model) {
    case 1: { /* SP1 att:168-192 */
        double z2 = fabs(z);
        double t = 1.83415e-09 * z2 * z2 * z2 + (-1.59061e-08 * z2 * z2) + 0.00267687 * z2 + (-51.0696); /* att:141-142 */
        double w0 = log(0.0001), w1 = 0.0, w2 = log(3.16);
        double w = log(frequency);
        double b0 = -6.74890 + t * (0.026709 - t * 0.000884);
        double b1 = -6.22121 - t * (0.070927 + t * 0.001773);
        double b2 = -4.09468 - t * (0.002213 + t * 0.000332);
        double a, bb;
        if (frequency < 1.) { a = (b1 * w0 - b0 * w1) / (w0 - w1); bb = (b1 - b0) / (w1 - w0); }
        else { a = (b2 * w1 - b1 * w2) / (w1 - w2); bb = (b2 - b1) / (w2 - w1); }
        L = 1. / exp(a + bb * w);
        break; }
    case 2: { /* GL1 att:99-128,194-196 */
        static const double fit[6] = {1.16052586e+03, 6.87257150e-02, -9.82378264e-05, -3.50628312e-07, -2.21040482e-10, -3.63912864e-14};
        double att75 = 0, zp = 1;
        for (int p = 0; p < 6; ++p) { att75 += fit[p] * zp; zp *= z; }
        if (att75 < 100.) att75 = 100.;
        L = att75 - 0.55 * (frequency / 1e-3 - 75);
        break; }
    case 4: { /* GL2 att:198-204 */
        static const double fit[6] = {1.20547286e+00, 1.58815679e-05, -2.58901767e-07, -5.16435542e-10, -2.89124473e-13, -4.58987344e-17};
        double bulk = 852.0 + (-0.54 / 1e-3) * frequency;
        double p = 0; for (int k = 5; k >= 0; --k) p = p * z + fit[k];
        L = bulk * p;
        break; }
    case 5: { /* GL3 att:206-222 */
        L = gl3_interp(m, -z, 1) * frequency + gl3_interp(m, -z, 2);
        break; }
    case 3: { /* MB1 att:224-244 */
        double R = 0.82, d_ice = 576.;
        L = 460. - 180. * frequency;
        L *= 1. / (1 + L / (2 * d_ice) * log(R));
        double d = -z * 420. / d_ice;
        double Lz = (1250. * 0.08886 * exp(-0.048827 * (225.6746 - 86.517596 * log10(848.870 - d))));
        L *= Lz / 231.21;
        break; }
    default: return INFINITY;
    }
    if (L < 1.) L = 1.;      /* att:252-255 */
    if (z > 0) L = INFINITY; /* att:256-257 */
    return L;
}

/* ------------------------------------------------------------------------------------------------ */
/* adaptive Gauss-Kronrod 21 (QUADPACK QK21 nodes/weights and error heuristic; QAG-style bisection)   */
/* ------------------------------------------------------------------------------------------------ */

static const double XGK[11] = {0.995657163025808080735527280689003, 0.973906528517171720077964012084452,
    0.930157491355708226001207180059508, 0.865063366688984510732096688423493, 0.780817726586416897063717578345042,
    0.679409568299024406234327365114874, 0.562757134668604683339000099272694, 0.433395394129247190799265943165784,
    0.294392862701460198131126603103866, 0.148874338981631210884826001129720, 0.0};
static const double WGK[11] = {0.011694638867371874278064396062192, 0.032558162307964727478818972459390,
    0.054755896574351996031381300244580, 0.075039674810919952767043140916190, 0.093125454583697605535065465083366,
    0.109387158802297641899210590325805, 0.123491976262065851077958109585166, 0.134709217311473325928054001771707,
    0.142775938577060080797094273138717, 0.147739104901338491374841515972068, 0.149445554002916905664936468389821};
static const double WG[5] = {0.066671344308688137593568809893332, 0.149451349150580593145776339657697,
    0.219086362515982043995534934228163, 0.269266719309996355091226921569469, 0.295524224714752870173815619188769};

typedef double (*ora_fn)(double, void *);

static void qk21(ora_fn f, void *ctx, double a, double b, double *result, double *abserr)
{
    double center = 0.5 * (a + b), half = 0.5 * (b - a), ahalf = fabs(half);
    double fc = f(center, ctx);
    double resg = 0, resk = WGK[10] * fc, resabs = fabs(resk);
    double fv1[10], fv2[10];
    for (int j = 0; j < 5; ++j) {
        int jtw = 2 * j + 1;
        double absc = half * XGK[jtw];
        double f1 = f(center - absc, ctx), f2 = f(center + absc, ctx);
        fv1[jtw] = f1; fv2[jtw] = f2;
        resg += WG[j] * (f1 + f2);
        resk += WGK[jtw] * (f1 + f2);
        resabs += WGK[jtw] * (fabs(f1) + fabs(f2));
    }
    for (int j = 0; j < 5; ++j) {
        int jtwm1 = 2 * j;
        double absc = half * XGK[jtwm1];
        double f1 = f(center - absc, ctx), f2 = f(center + absc, ctx);
        fv1[jtwm1] = f1; fv2[jtwm1] = f2;
        resk += WGK[jtwm1] * (f1 + f2);
        resabs += WGK[jtwm1] * (fabs(f1) + fabs(f2));
    }
    double reskh = resk * 0.5;
    double resasc = WGK[10] * fabs(fc - reskh);
    for (int j = 0; j < 10; ++j) resasc += WGK[j] * (fabs(fv1[j] - reskh) + fabs(fv2[j] - reskh));
    *result = resk * half;
    resabs *= ahalf; resasc *= ahalf;
    double err = fabs((resk - resg) * half);
    if (resasc != 0 && err != 0) { double s = pow(200 * err / resasc, 1.5); err = resasc * (s < 1 ? s : 1); }
    if (resabs > 2.2250738585072014e-308 / (50 * 2.220446049250313e-16)) { double e = 50 * 2.220446049250313e-16 * resabs; if (e > err) err = e; }
    *abserr = err;
}

#define ORA_QLIMIT 2000
static double quad_adaptive(ora_fn f, void *ctx, const double *breaks, int nbreaks, double epsabs, double epsrel, long *neval)
{
    static __thread double A[ORA_QLIMIT], B[ORA_QLIMIT], R[ORA_QLIMIT], E[ORA_QLIMIT];
    int n = 0;
    double total = 0, errsum = 0;
    for (int i = 0; i + 1 < nbreaks; ++i) {
        if (breaks[i] == breaks[i + 1]) continue;
        A[n] = breaks[i]; B[n] = breaks[i + 1];
        qk21(f, ctx, A[n], B[n], &R[n], &E[n]);
        if (neval) *neval += 21;
        total += R[n]; errsum += E[n]; ++n;
    }
    while (n < ORA_QLIMIT - 1) {
        double tol = fmax(epsabs, epsrel * fabs(total));
        if (errsum <= tol) break;
        int k = 0;
        for (int i = 1; i < n; ++i) if (E[i] > E[k]) k = i;
        double a = A[k], b = B[k], mid = 0.5 * (a + b);
        if (!(mid > a && mid < b)) break;
        double r1, e1, r2, e2;
        qk21(f, ctx, a, mid, &r1, &e1);
        qk21(f, ctx, mid, b, &r2, &e2);
        if (neval) *neval += 42;
        total += r1 + r2 - R[k]; errsum += e1 + e2 - E[k];
        A[k] = a; B[k] = mid; R[k] = r1; E[k] = e1;
        A[n] = mid; B[n] = b; R[n] = r2; E[n] = e2; ++n;
    }
    total = 0;
    for (int i = 0; i < n; ++i) total += R[i];
    return total;
}

/* ------------------------------------------------------------------------------------------------ */
/* attenuation along the path (py:885-1089)                                                          */
/* ------------------------------------------------------------------------------------------------ */

typedef struct { const ora_cfg *m; double C_0, f, z_turn; int sub; /* 0 none, -1: t=z_turn-u^2, +1: t=z_turn+u^2 */ } att_ctx;

static double att_integrand(double v, void *p)                                                           /* py:986-988 */
{
    att_ctx *c = (att_ctx *)p;
    double t = v, jac = 1;
    if (c->sub < 0) { t = c->z_turn - v * v; jac = 2 * v; }
    else if (c->sub > 0) { t = c->z_turn + v * v; jac = 2 * v; }
    double z = get_z_unmirrored(t, c->C_0, c->m);
    double ds = ds_of_t(t, c->C_0, c->m);
    if (!isfinite(ds)) return 0.0; /* only reachable exactly on the turning point (jac = 0 there) */
    return jac * ds / ora_attenuation_length(c->m, z, c->f);
}

/* ds alone in the substituted variable (window cell of the optimized discretisation) */
static double att_integrand_ds(double v, void *p)
{
    att_ctx *c = (att_ctx *)p;
    double t = v, jac = 1;
    if (c->sub < 0) { t = c->z_turn - v * v; jac = 2 * v; }
    else if (c->sub > 0) { t = c->z_turn + v * v; jac = 2 * v; }
    double ds = ds_of_t(t, c->C_0, c->m);
    if (!isfinite(ds)) return 0.0;
    return jac * ds;
}

/* numpy.linspace(start, stop, n) */
static void linspace(double a, double b, int n, double *out)
{
    if (n == 1) { out[0] = a; return; }
    double step = (b - a) / (n - 1);
    for (int i = 0; i < n; ++i) out[i] = a + i * step;
    out[n - 1] = b;
}

/* py:885-931; returns number of sparse frequencies written to freqs (capacity >= n_freq + n_freq/2) */
int ora_sparse_frequencies(int n_freq_int, const double *frequency, int nf, double max_detector_freq, double *freqs)
{
    int n_nonnull = 0; double flo = INFINITY, fhi = -INFINITY;
    for (int i = 0; i < nf; ++i) if (frequency[i] > 0) { ++n_nonnull; flo = fmin(flo, frequency[i]); fhi = fmax(fhi, frequency[i]); }
    if (n_nonnull == 0) return 0;
    int n = n_freq_int < n_nonnull ? n_freq_int : n_nonnull;
    linspace(flo, fhi, n, freqs);
    if (n < n_nonnull && !isnan(max_detector_freq)) {
        int n_tot = 0, n_above = 0; double tmin = INFINITY, tmax = -INFINITY, amin = INFINITY, amax = -INFINITY;
        for (int i = 0; i < nf; ++i) {
            int det = frequency[i] <= max_detector_freq;
            if (det && frequency[i] > 0) { ++n_tot; tmin = fmin(tmin, frequency[i]); tmax = fmax(tmax, frequency[i]); }
            if (!det) { ++n_above; amin = fmin(amin, frequency[i]); amax = fmax(amax, frequency[i]); }
        }
        n = n_freq_int < n_tot ? n_freq_int : n_tot;
        linspace(tmin, tmax, n, freqs);
        if (n_above > 1) { linspace(amin, amax, n / 2, freqs + n); n += n / 2; }
    }
    return n;
}

/* np.interp(x, xp, fp) for ascending xp */
static double interp1(double x, const double *xp, const double *fp, int n)
{
    if (x <= xp[0]) return fp[0];
    if (x >= xp[n - 1]) return fp[n - 1];
    int lo = 0, hi = n - 1;
    while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (xp[mid] <= x) lo = mid; else hi = mid; }
    double slope = (fp[hi] - fp[lo]) / (xp[hi] - xp[lo]);
    return slope * (x - xp[lo]) + fp[lo];
}

/* ---- the "optimized" discretisation the reference switches to for GL3 (py:62, :458, :998-1064) ---- */

/* py:65-67 */
static int get_n_steps(double x1, double x2, double dx)
{
    int n = (int)floor(fabs(x1 - x2) / dx);   /* int(abs(x1 - x2) // dx) */
    return n > 3 ? n : 3;
}

/* py:70-76: np.linspace(x1, x2, get_n_steps(...)), or the single point if x1 == x2; returns the count */
static int get_equidistant_steps(double x1, double x2, double dx, double *out)
{
    if (x1 == x2) { out[0] = x1; return 1; }
    int n = get_n_steps(x1, x2, dx);
    linspace(x1, x2, n, out);
    return n;
}

static double ds_ctx(double t, void *p) { att_ctx *c = (att_ctx *)p; return ds_of_t(t, c->C_0, c->m); }

/* exponent of the attenuation factor of one segment for the sparse frequencies: 10 m midpoint sum of ds/L in the
 * mirrored depth coordinate, the cell that holds the turning point replaced by (integral of ds) / L(z_turn)  (py:998-1064) */
#define ORA_MAX_STEPS 4096
static void optimized_exponent(const double x1[2], const double x2m[2], double z_turn, double C_0, const ora_cfg *m,
                               const double *freqs, int nsp, double *expo, long *neval)
{
    const double dx = 10.0, window = 20.0;
    static __thread double steps[ORA_MAX_STEPS];
    int n = 0;
    int fallback = (x1[1] - window / 2 < z_turn && z_turn < x2m[1] + window / 2);
    if (fallback) {
        double w0 = fmax(x1[1], z_turn - window / 2), w1 = fmin(z_turn + window / 2, x2m[1]);
        n = get_equidistant_steps(x1[1], w0, dx, steps);
        n += get_equidistant_steps(w1, x2m[1], dx, steps + n);
    } else {
        n = get_equidistant_steps(x1[1], x2m[1], dx, steps);
    }
    for (int k = 0; k < nsp; ++k) expo[k] = 0.0;
    int idx = -2;
    if (fallback) {
        /* np.digitize(z_turn, path_steps) - 1 for increasing bins, clamped (py:1044-1050) */
        int d = 0;
        while (d < n && steps[d] <= z_turn) ++d;
        idx = d - 1;
        if (idx == n - 1) idx -= 1;
        else if (idx == -1) idx = 0;
    }
    for (int i = 0; i + 1 < n; ++i) {
        double dxa = steps[i + 1] - steps[i];
        if (i == idx) {
            att_ctx c = {m, C_0, 0, z_turn, 0};
            double S;
            if (m->quad_mode == 0) {
                double br[3] = {steps[i], z_turn, steps[i + 1]};
                if (steps[i] < z_turn && z_turn < steps[i + 1]) S = quad_adaptive(ds_ctx, &c, br, 3, 1.49e-8, 1e-2, neval);
                else { double b2[2] = {steps[i], steps[i + 1]}; S = quad_adaptive(ds_ctx, &c, b2, 2, 1.49e-8, 1e-2, neval); }
            } else {
                /* tight: ds = 1/sqrt(1 - (beta/n)^2) has a 1/sqrt singularity at a refracted apex: substitute t = z_turn -+ u^2 */
                S = 0;
                const double a = steps[i], b = steps[i + 1];
                if (a < z_turn) {
                    double lo_t = b < z_turn ? b : z_turn;
                    double b2[2] = {sqrt(z_turn - lo_t), sqrt(z_turn - a)};
                    c.sub = -1; S += quad_adaptive(att_integrand_ds, &c, b2, 2, 0, 1e-12, neval);
                }
                if (b > z_turn) {
                    double lo_t = a > z_turn ? a : z_turn;
                    double b2[2] = {sqrt(lo_t - z_turn), sqrt(b - z_turn)};
                    c.sub = +1; S += quad_adaptive(att_integrand_ds, &c, b2, 2, 0, 1e-12, neval);
                }
            }
            for (int k = 0; k < nsp; ++k) expo[k] += S / ora_attenuation_length(m, z_turn, freqs[k]);
        } else {
            double mid = steps[i] + dxa / 2;
            double z = get_z_unmirrored(mid, C_0, m);
            double ds = ds_of_t(mid, C_0, m);
            for (int k = 0; k < nsp; ++k) expo[k] += ds / ora_attenuation_length(m, z, freqs[k]) * dxa;
        }
    }
}

/* py:933-1089 (python branch, not the GL3 "optimized" discretisation).  out[nf] dense factors on `frequency`;
 * if sparse_out != NULL it additionally receives the product over segments of the factors at the sparse
 * frequencies (what a consumer that interpolates later needs). */
static void attenuation_along_path(const double x1_in[2], const double x2_in[2], double C_0, const ora_cfg *m,
                                   const double *frequency, int nf, double max_detector_freq,
                                   int reflection, int reflection_case, double *out, double *sparse_out, long *neval)
{
    double freqs[4096], fac[4096];
    int nsp = ora_sparse_frequencies(m->n_freq, frequency, nf, max_detector_freq, freqs);
    for (int i = 0; i < nf; ++i) out[i] = 1.0;
    if (sparse_out) for (int i = 0; i < nsp; ++i) sparse_out[i] = 1.0;
    ora_segment seg[ORA_MAX_SEG];
    int nseg = get_path_segments(x1_in, x2_in, C_0, m, reflection, reflection_case, seg);
    for (int iS = 0; iS < nseg; ++iS) {
        double x1[2], x2[2];
        segment_endpoints(&seg[iS], iS, reflection_case, x1, x2);
        double x2m[2];
        get_z_mirrored(x1, x2, C_0, m, x2m);
        double gamma_turn, z_turn;
        get_turning_point(m->n_ice * m->n_ice - pow(C_0, -2), m, &gamma_turn, &z_turn);
        int interior = (x1[1] < z_turn && z_turn < x2m[1]);
        int refracted_apex = (z_turn < 0); /* unclamped turning point -> 1/sqrt singularity of ds */
        if (m->att_model == 5) {   /* speedup_attenuation_models = ["GL3"] (py:62, :458) */
            double expo[4096];
            optimized_exponent(x1, x2m, z_turn, C_0, m, freqs, nsp, expo, neval);
            for (int k = 0; k < nsp; ++k) fac[k] = exp(-expo[k]);
        } else
        for (int k = 0; k < nsp; ++k) {
            att_ctx c = {m, C_0, freqs[k], z_turn, 0};
            double I = 0;
            if (m->quad_mode == 0) {
                double br[3] = {x1[1], z_turn, x2m[1]};
                if (interior) I = quad_adaptive(att_integrand, &c, br, 3, 1.49e-8, 1e-2, neval);
                else { double b2[2] = {x1[1], x2m[1]}; I = quad_adaptive(att_integrand, &c, b2, 2, 1.49e-8, 1e-2, neval); }
            } else {
                double epsrel = 1e-11;
                if (interior) {
                    if (refracted_apex) {
                        double bl[2] = {0, sqrt(z_turn - x1[1])}, brr[2] = {0, sqrt(x2m[1] - z_turn)};
                        c.sub = -1; I += quad_adaptive(att_integrand, &c, bl, 2, 0, epsrel, neval);
                        c.sub = +1; I += quad_adaptive(att_integrand, &c, brr, 2, 0, epsrel, neval);
                    } else {
                        double br[3] = {x1[1], z_turn, x2m[1]};
                        I = quad_adaptive(att_integrand, &c, br, 3, 0, epsrel, neval);
                    }
                } else if (refracted_apex && z_turn > x2m[1] && (z_turn - x2m[1]) < 0.5 * (x2m[1] - x1[1])) {
                    /* direct ray that ends just below its apex: same substitution tames the near-singularity */
                    double bl[2] = {sqrt(z_turn - x2m[1]), sqrt(z_turn - x1[1])};
                    c.sub = -1; I = quad_adaptive(att_integrand, &c, bl, 2, 0, epsrel, neval);
                } else {
                    double b2[2] = {x1[1], x2m[1]};
                    I = quad_adaptive(att_integrand, &c, b2, 2, 0, epsrel, neval);
                }
            }
            fac[k] = exp(-I);                                                                            /* py:1075 */
        }
        for (int i = 0; i < nf; ++i) if (frequency[i] > 0) out[i] *= interp1(frequency[i], freqs, fac, nsp); /* py:1077-1078,1086 */
        if (sparse_out) for (int k = 0; k < nsp; ++k) sparse_out[k] *= fac[k];
    }
}

/* ------------------------------------------------------------------------------------------------ */
/* root finding in logC0: scan + local-maximum refinement + Brent                                    */
/* ------------------------------------------------------------------------------------------------ */

typedef struct { const ora_cfg *m; const double *x1, *x2; int reflection, reflection_case; long neval; } obj_ctx;

static double obj_delta_y(double logC0, void *p)                                                         /* py:1357-1363 */
{
    obj_ctx *c = (obj_ctx *)p;
    c->neval++;
    double C_0 = get_C0_from_log(logC0, c->m->n_ice);
    double v = get_delta_y(C_0, c->x1, c->x2, c->m, c->reflection, c->reflection_case);
    return v;
}

/* Brent (1973) zeroin, the algorithm behind scipy.optimize.brentq (xtol=2e-12, rtol=4*eps, maxiter=100). */
static double brentq(ora_fn f, void *ctx, double xa, double xb, double fa, double fb)
{
    const double xtol = 2e-12, rtol = 8.881784197001252e-16;
    double xpre = xa, xcur = xb, fpre = fa, fcur = fb, xblk = 0, fblk = 0, spre = 0, scur = 0;
    if (fpre == 0) return xpre;
    if (fcur == 0) return xcur;
    for (int i = 0; i < 100; ++i) {
        if (fpre != 0 && fcur != 0 && (signbit(fpre) != signbit(fcur))) { xblk = xpre; fblk = fpre; spre = scur = xcur - xpre; }
        if (fabs(fblk) < fabs(fcur)) { xpre = xcur; xcur = xblk; xblk = xpre; fpre = fcur; fcur = fblk; fblk = fpre; }
        double delta = (xtol + rtol * fabs(xcur)) / 2, sbis = (xblk - xcur) / 2;
        if (fcur == 0 || fabs(sbis) < delta) return xcur;
        if (fabs(spre) > delta && fabs(fcur) < fabs(fpre)) {
            double stry;
            if (xpre == xblk) stry = -fcur * (xcur - xpre) / (fcur - fpre);
            else { double dpre = (fpre - fcur) / (xpre - xcur), dblk = (fblk - fcur) / (xblk - xcur);
                   stry = -fcur * (fblk * dblk - fpre * dpre) / (dblk * dpre * (fblk - fpre)); }
            if (2 * fabs(stry) < fmin(fabs(spre), 3 * fabs(sbis) - delta)) { spre = scur; scur = stry; }
            else { spre = sbis; scur = sbis; }
        } else { spre = sbis; scur = sbis; }
        xpre = xcur; fpre = fcur;
        if (fabs(scur) > delta) xcur += scur; else xcur += (sbis > 0 ? delta : -delta);
        fcur = f(xcur, ctx);
    }
    return xcur;
}

/* golden-section maximisation of f on [a,c] around interior sample b (f(b) >= f(a), f(c)); stops early when f > 0 */
static double maximise(ora_fn f, void *ctx, double a, double b, double c, double fb, double *xmax)
{
    const double gr = 0.3819660112501051;
    double x = b, fx = fb;
    for (int it = 0; it < 80 && (c - a) > 1e-13 * (1 + fabs(x)); ++it) {
        double u = (x - a > c - x) ? x - gr * (x - a) : x + gr * (c - x);
        double fu = f(u, ctx);
        if (fu > fx) { if (u > x) a = x; else c = x; x = u; fx = fu; if (fx > 0) break; }
        else { if (u > x) c = u; else a = u; }
    }
    *xmax = x;
    return fx;
}

/* all roots of the reference objective for one (reflection, case) mode; returns count, C0 ascending */
int ora_find_roots_2d(const ora_cfg *m, const double x1[2], const double x2[2], int reflection, int reflection_case,
                      double *C0_out, long *neval)
{
    obj_ctx ctx = {m, x1, x2, reflection, reflection_case, 0};
    int n = m->scan_n > 0 ? m->scan_n : 4401;
    double lo = -22, hi = 22;
    if (m->scan_lo != 0 || m->scan_hi != 0) { lo = m->scan_lo; hi = m->scan_hi; }
    double h = (hi - lo) / (n - 1);
    double roots[16]; int nr = 0;
    double fm2 = NAN, fm1 = NAN, lm1 = 0;
    for (int i = 0; i < n && nr < 14; ++i) {
        double l = (i == n - 1) ? hi : lo + i * h;
        double f = obj_delta_y(l, &ctx);
        if (i > 0 && isfinite(f) && isfinite(fm1)) {
            if ((fm1 < 0 && f > 0) || (fm1 > 0 && f < 0)) roots[nr++] = brentq(obj_delta_y, &ctx, lm1, l, fm1, f);
            else if (f == 0) roots[nr++] = l;
            else if (i > 1 && isfinite(fm2) && fm1 < 0 && fm1 >= fm2 && fm1 >= f) {
                /* negative local maximum between l-2h and l: does the hump poke through zero? */
                double xm, fmx = maximise(obj_delta_y, &ctx, l - 2 * h, lm1, l, fm1, &xm);
                if (fmx > 0) {
                    roots[nr++] = brentq(obj_delta_y, &ctx, l - 2 * h, xm, fm2, fmx);
                    roots[nr++] = brentq(obj_delta_y, &ctx, xm, l, fmx, f);
                } else if (fmx >= -1e-9) {
                    /* tangency: the objective touches zero without changing sign.  Happens systematically for a
                     * receiver exactly at the surface (z2 = 0: both branches of py:255-272 give -|rho - y_turn|);
                     * the reference's hybr on f^2 (py:1479-1483) reports this single root. */
                    roots[nr++] = xm;
                }
            }
        }
        fm2 = fm1; fm1 = f; lm1 = l;
    }
    /* sort + de-duplicate (a refined hump can straddle a scan node) */
    for (int i = 1; i < nr; ++i) { double v = roots[i]; int j = i - 1; while (j >= 0 && roots[j] > v) { roots[j + 1] = roots[j]; --j; } roots[j + 1] = v; }
    int k = 0;
    for (int i = 0; i < nr; ++i) if (k == 0 || fabs(roots[i] - roots[k - 1]) > 1e-9) roots[k++] = roots[i];
    if (k > ORA_MAX_ROOTS) k = ORA_MAX_ROOTS;
    for (int i = 0; i < k; ++i) C0_out[i] = get_C0_from_log(roots[i], m->n_ice);
    if (neval) *neval += ctx.neval;
    return k;
}

/* ------------------------------------------------------------------------------------------------ */
/* 3-D wrapper (py:1932-2146, 2560-2776)                                                             */
/* ------------------------------------------------------------------------------------------------ */

typedef struct {
    int32_t *n_sol;          /* [N] */
    int8_t *type, *reflection, *reflection_case;   /* [N,S] */
    double *C0, *C1, *path_length, *travel_time;   /* [N,S] */
    double *launch, *receive;                      /* [N,S,3] */
    double *reflection_angle;                      /* [N,S,n_reflections+1], NAN = None */
    double *attenuation;                           /* [N,S,nf] or NULL */
    double *attenuation_sparse;                    /* [N,S,nsp] or NULL */
    int32_t *status;                               /* [N] bit0: point below reflective layer, bit1: >2 roots in a mode, bit2: overflow */
} ora_out;

static void trace_one(const ora_cfg *m, const double *X1in, const double *X2in, const double *frequency, int nf, int nsp,
                      double max_detector_freq, const ora_out *o, int64_t i, long *neval)
{
    const int S = 2 + 4 * m->n_reflections, K1 = m->n_reflections + 1;
    double X1[3] = {X1in[0], X1in[1], X1in[2]}, X2[3] = {X2in[0], X2in[1], X2in[2]};
    o->n_sol[i] = 0; o->status[i] = 0;
    if (m->n_reflections && (X1[2] < m->reflection || X2[2] < m->reflection)) { o->status[i] |= 1; return; }  /* base:156-161 */
    int swap = 0;
    if (X2[2] < X1[2]) { swap = 1; for (int k = 0; k < 3; ++k) { double t = X1[k]; X1[k] = X2[k]; X2[k] = t; } } /* py:2072-2077 */
    double dX[3] = {X2[0] - X1[0], X2[1] - X1[1], X2[2] - X1[2]};
    double dPhi = -atan2(dX[1], dX[0]);                                                                  /* py:2080 */
    double c = cos(dPhi), s = sin(dPhi);
    double x1[2] = {X1[0], X1[2]};
    double x2[2] = {c * dX[0] - s * dX[1] + X1[0], X2[2]};                                                /* py:2084-2089 */
    /* mode loop py:2118-2125 */
    int n = 0;
    double C0s[64]; int refl[64], rcase[64];
    for (int md = 0; md < 1 + 2 * m->n_reflections; ++md) {
        int reflection = md == 0 ? 0 : (md - 1) / 2 + 1;
        int reflection_case = md == 0 ? 1 : (md - 1) % 2 + 1;
        double r[ORA_MAX_ROOTS];
        int k = 0;
        if (!(x2[1] > 0))      /* receiver in air: get_delta_y is negative everywhere (py:247-253) -> py:1445-1448 returns [] */
            k = ora_find_roots_2d(m, x1, x2, reflection, reflection_case, r, neval);
        if (k > 2) o->status[i] |= 2;
        for (int j = 0; j < k && n < 64; ++j) { C0s[n] = r[j]; refl[n] = reflection; rcase[n] = reflection_case; ++n; }
    }
    if (n > S) { o->status[i] |= 4; n = 0; }                                                              /* py:2128-2130 */
    o->n_sol[i] = n;
    for (int iS = 0; iS < n; ++iS) {
        int64_t q = i * S + iS;
        double C_0 = C0s[iS];
        o->type[q] = (int8_t)determine_solution_type(x1, x2, C_0, m);                                     /* py:2146 */
        o->reflection[q] = (int8_t)refl[iS]; o->reflection_case[q] = (int8_t)rcase[iS];
        o->C0[q] = C_0; o->C1[q] = get_C_1(x1, C_0, m);
        o->path_length[q] = path_or_time_analytic(x1, x2, C_0, m, refl[iS], rcase[iS], 0);
        o->travel_time[q] = path_or_time_analytic(x1, x2, C_0, m, refl[iS], rcase[iS], 1);
        double al = get_launch_angle(x1, C_0, m, refl[iS], rcase[iS]);
        double ar = get_receive_angle(x1, x2, C_0, m, refl[iS], rcase[iS]);
        double l2[3] = {sin(al), 0, cos(al)}, r2[3] = {-sin(ar), 0, cos(ar)};                            /* py:2583,2617 */
        if (swap) { l2[0] = -sin(ar); l2[2] = cos(ar); r2[0] = sin(al); r2[2] = cos(al); }                /* py:2584-2588,2618-2622 */
        /* R^T v, R = ((c,-s,0),(s,c,0),(0,0,1)) py:2082,2590 */
        o->launch[3 * q + 0] = c * l2[0] + s * l2[1]; o->launch[3 * q + 1] = -s * l2[0] + c * l2[1]; o->launch[3 * q + 2] = l2[2];
        o->receive[3 * q + 0] = c * r2[0] + s * r2[1]; o->receive[3 * q + 1] = -s * r2[0] + c * r2[1]; o->receive[3 * q + 2] = r2[2];
        double ra[ORA_MAX_SEG];
        int nra = get_reflection_angle(x1, x2, C_0, m, refl[iS], rcase[iS], ra);
        for (int k = 0; k < K1; ++k) o->reflection_angle[q * K1 + k] = k < nra ? ra[k] : NAN;
        if (m->att_model > 0 && nf > 0 && (o->attenuation || o->attenuation_sparse)) {
            double dense[8192], sparse[4096];
            attenuation_along_path(x1, x2, C_0, m, frequency, nf, max_detector_freq, refl[iS], rcase[iS], dense, sparse, neval);
            if (o->attenuation) memcpy(o->attenuation + q * nf, dense, sizeof(double) * nf);
            if (o->attenuation_sparse) memcpy(o->attenuation_sparse + q * nsp, sparse, sizeof(double) * nsp);
        }
    }
}

/* N pairs, AoS inputs X1[N,3], X2[N,3] (row-major).  Outputs must be pre-filled by the caller (NaN / 0).
 * Returns the number of objective + integrand evaluations (work counter for the CPU baseline report). */
int64_t ora_trace(const ora_cfg *m, int64_t N, const double *X1, const double *X2, const double *frequency, int32_t nf,
                  double max_detector_freq, const ora_out *o, int32_t n_threads)
{
    int nsp = 0;
    if (nf > 0) { double tmp[4096]; nsp = ora_sparse_frequencies(m->n_freq, frequency, nf, max_detector_freq, tmp); }
    int64_t total = 0;
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#pragma omp parallel for schedule(dynamic, 16) reduction(+ : total)
#endif
    for (int64_t i = 0; i < N; ++i) {
        long ne = 0;
        trace_one(m, X1 + 3 * i, X2 + 3 * i, frequency, nf, nsp, max_detector_freq, o, i, &ne);
        total += ne;
    }
    return total;
}

/* thin exports for unit tests of the 2-D pieces */
double ora_delta_y(const ora_cfg *m, double logC0, const double *x1, const double *x2, int reflection, int reflection_case)
{ return get_delta_y(get_C0_from_log(logC0, m->n_ice), x1, x2, m, reflection, reflection_case); }
int ora_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
