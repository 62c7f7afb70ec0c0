// TEST-ONLY: host build of the kernels' scalar maths (nuradiomc_b200/csrc/nrmc_math.cuh) so that the algorithm can be
// compared with the oracle on machines without a GPU (`pytest -m "not gpu"`).  Never loaded by the product.
static long g_evals = 0;      // evaluations of the range curve (host build only: harness_solver_evaluations)
#define NRMC_COUNT_EVALS
#include "../../nuradiomc_b200/csrc/nrmc_math.cuh"
#include <vector>
using namespace nrmc;
extern "C" int harness_trace(double n_ice, double dn, double z0, double zr, int n_refl, int64_t N, const double *X1,
                             const double *X2, int32_t *n_sol, int32_t *status, int8_t *type, int8_t *reflection,
                             int8_t *reflection_case, double *C0, double *C1, double *path_length, double *travel_time,
                             double *launch, double *receive, double *reflection_angle)
{
    IceParams ice;
    ice.n_ice = n_ice; ice.dn = dn; ice.z0 = z0; ice.inv_z0 = 1.0 / z0; ice.inv_dn = 1.0 / dn; ice.ns = n_ice - dn;
    ice.n_refl = n_refl; ice.zr = n_refl > 0 ? zr : -1e30;
    ice.gr = n_refl > 0 ? dn * exp(zr / z0) : 0.0; ice.nr = n_ice - ice.gr; ice.att_model = 0;
    TraceOutputs o = {n_sol, status, type, reflection, reflection_case, C0, C1, path_length, travel_time, launch, receive, reflection_angle};
    for (int64_t i = 0; i < N; ++i)
        trace_pair(ice, X1[3 * i], X1[3 * i + 1], X1[3 * i + 2], X2[3 * i], X2[3 * i + 1], X2[3 * i + 2], i, o, nullptr);
    return 0;
}

// focusing factor of every solution of a padded [N,S] result (same device function as K_focusing)
extern "C" int harness_focusing(double n_ice, double dn, double z0, double zr, int n_refl, int64_t N, const double *X1,
                                const double *X2, const int32_t *n_sol, const double *C0, const int8_t *reflection,
                                const int8_t *reflection_case, const double *path_length, double limit, double *focusing)
{
    IceParams ice;
    ice.n_ice = n_ice; ice.dn = dn; ice.z0 = z0; ice.inv_z0 = 1.0 / z0; ice.inv_dn = 1.0 / dn; ice.ns = n_ice - dn;
    ice.n_refl = n_refl; ice.zr = n_refl > 0 ? zr : -1e30;
    ice.gr = n_refl > 0 ? dn * exp(zr / z0) : 0.0; ice.nr = n_ice - ice.gr; ice.att_model = 0;
    const int S = 2 + 4 * n_refl;
    for (int64_t i = 0; i < N; ++i) {
        Frame2D f;
        make_frame(X1[3 * i], X1[3 * i + 1], X1[3 * i + 2], X2[3 * i], X2[3 * i + 1], X2[3 * i + 2], f);
        PairGeom g;
        make_pair_geom(ice, f.z1, f.z2, fmax(f.rho, 1e-12), g);
        for (int s = 0; s < S; ++s) {
            const int64_t q = i * S + s;
            focusing[q] = s < n_sol[i] ? focusing_factor(ice, g, f.swap, reflection[q], reflection_case[q], 1.0 / C0[q], path_length[q], limit) : NAN;
        }
    }
    return 0;
}

// evaluations of the range curve the Newton solver spends on each bracket of a k = 0 pair (starting points, termination
// thresholds): evals[2 i + b], piece[2 i + b] (-1: no such bracket)
extern "C" int harness_solver_evaluations(double n_ice, double dn, double z0, int64_t N, const double *X1, const double *X2,
                                          int32_t *evals, int8_t *piece)
{
    IceParams ice;
    ice.n_ice = n_ice; ice.dn = dn; ice.z0 = z0; ice.inv_z0 = 1.0 / z0; ice.inv_dn = 1.0 / dn; ice.ns = n_ice - dn;
    ice.n_refl = 0; ice.zr = -1e30; ice.gr = 0.0; ice.nr = n_ice; ice.att_model = 0;
    for (int64_t i = 0; i < N; ++i) {
        evals[2 * i] = evals[2 * i + 1] = 0; piece[2 * i] = piece[2 * i + 1] = -1;
        Frame2D f;
        make_frame(X1[3 * i], X1[3 * i + 1], X1[3 * i + 2], X2[3 * i], X2[3 * i + 1], X2[3 * i + 2], f);
        if (pair_status(ice, f) != 0) continue;
        PairGeom g;
        make_pair_geom(ice, f.z1, f.z2, fmax(f.rho, 1e-12), g);
        Curve cv;
        cv.ice = &ice; cv.g = &g; cv.k = 0; cv.rcase = 1;
        double J1, J2, J3;
        Bracket br[2];
        bool need_hump;
        int nb = classify_mode(cv, J1, J2, J3, br, need_hump);
        if (need_hump) nb = hump_search(cv, J1, J2, J3, br);
        for (int b = 0; b < nb; ++b) {
            g_evals = 0;
            (void)solve_bracket(cv, br[b]);
            evals[2 * i + b] = (int32_t)g_evals; piece[2 * i + b] = (int8_t)br[b].piece;
        }
    }
    return 0;
}
