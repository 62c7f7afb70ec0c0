// CPU emulation of K_att's GL1 quadrature schemes (test/prototype only).
#include "/root/repo/nuradiomc_b200/csrc/nrmc_att.cuh"
#include <vector>
#include <cstring>
using namespace nrmc;
static const double GX[8] = NRMC_GL16_X, GW[8] = NRMC_GL16_W;
static inline void node_xw(int q, double &x, double &w) { if (q < 8) { x = -GX[7 - q]; w = GW[7 - q]; } else { x = GX[q - 8]; w = GW[q - 8]; } }

struct Slot { double lo, hi; int panel; };
static void slot_of(const AttPlan &p, int spp, int slot, Slot &s)
{
    AttPlan q = p; q.spp = spp; q.n_slots = q.na * spp;
    plan_slot(q, slot, s.lo, s.hi, s.panel);
}
// integrate all frequencies in `sel` over one 16-node slot, add sign * contribution to H[panel][j]; returns min A over nodes and ends
static double do_slot(const IceParams &ice, const AttPlan &p, const Slot &s, int Fs, const double *fa, const unsigned char *sel, double sign,
                      double *H, int Fs_pad, long &work)
{
    Gl3Table gl3 = {nullptr, 0};
    double amin = 1e300;
    for (int e = 0; e < 2; ++e) { const double u = e ? s.hi : s.lo; AttNode nd; att_node(2, fmin(p.zv - u * u, 0.0), gl3, nd); amin = fmin(amin, nd.p0); }
    for (int q = 0; q < 16; ++q) {
        double x, w, z, wds; node_xw(q, x, w);
        att_node_geometry(ice, p, s.lo, s.hi, x, w, z, wds);
        AttNode nd; att_node(2, z, gl3, nd);
        amin = fmin(amin, nd.p0);
        for (int j = 0; j < Fs; ++j) if (!sel || sel[j]) { H[s.panel * Fs_pad + j] += sign * wds * att_inv_length(2, nd, fa[j], 0.0); ++work; }
    }
    return amin;
}

extern "C" int gl1_emul(double n_ice, double dn, double z0, int64_t N, const double *X1, const double *X2, int Fs, const double *fa,
                        int scheme, double margin, double margin_ratio, double *out, int32_t *n_sol, long *work_out)
{
    IceParams ice; ice.n_ice = n_ice; ice.dn = dn; ice.z0 = z0; ice.inv_z0 = 1.0 / z0; ice.ns = n_ice - dn; ice.n_refl = 0; ice.zr = -1e30; ice.gr = 0; ice.nr = n_ice; ice.att_model = 2;
    const int Fs_pad = Fs;
    long work = 0, n_redo = 0, n_solutions = 0;
    std::vector<double> H(3 * Fs_pad);
    std::vector<unsigned char> sel(Fs);
    for (int64_t i = 0; i < N; ++i) {
        Frame2D f; make_frame(X1[3*i], X1[3*i+1], X1[3*i+2], X2[3*i], X2[3*i+1], X2[3*i+2], f);
        n_sol[i] = 0;
        for (int s = 0; s < 2; ++s) for (int j = 0; j < Fs; ++j) out[(i * 2 + s) * Fs + j] = NAN;
        if (pair_status(ice, f) != 0) continue;
        PairGeom g; make_pair_geom(ice, f.z1, f.z2, fmax(f.rho, 1e-12), g);
        Root roots[2];
        const int nr = find_roots_mode(ice, g, 0, 1, roots);
        n_sol[i] = nr;
        for (int r = 0; r < nr; ++r) {
            const bool band = roots[r].piece == 1 || roots[r].piece == 2;
            RayState rs; ray_state(ice, g, band, roots[r].v, rs);
            AttPlan p; att_plan(ice, g, roots[r].piece, 0, 1, rs, p);
            Gl3Table gl3 = {nullptr, 0};
            AttNode a_deep, a_top;
            att_node(2, g.z1, gl3, a_deep); att_node(2, roots[r].piece >= 2 ? fmin(p.zv, 0.0) : g.z2, gl3, a_top);
            const double a_min = fmin(a_deep.p0, a_top.p0);
            std::fill(H.begin(), H.end(), 0.0);
            const int spp0 = p.spp;                         // EASY (x2 for single-panel paths)
            const int n0 = p.na * spp0;
            const int m0 = plan_total_mult(p, 0), m1 = plan_total_mult(p, 1), m2 = plan_total_mult(p, 2);
            const int fine_per_coarse = NRMC_GL1_SPP / NRMC_GL1_SPP_EASY;
            bool redo_generic = false;
            if (scheme == 3) {
                int j_hard = Fs; for (int j = 0; j < Fs; ++j) if (fa[j] > a_min - 10.0) { j_hard = j; break; }
                for (int pa = 0; pa < 2; ++pa) {
                    if (plan_total_mult(p, pa) == 0) continue;
                    const double plo = pa == 0 ? p.uT : p.u2, phi = pa == 0 ? p.u2 : p.u1;
                    if (!(phi > plo)) continue;
                    if (pa == 0) { Slot sl; sl.lo = plo; sl.hi = phi; sl.panel = 0; do_slot(ice, p, sl, Fs, fa, nullptr, 1.0, H.data(), Fs_pad, work); }
                    else { const double w0 = (phi - plo) / 1.56; const double e[4] = {plo, plo + w0, plo + 1.4 * w0, phi};
                        for (int k = 0; k < 3; ++k) { Slot sl; sl.lo = e[k]; sl.hi = e[k + 1]; sl.panel = 1; do_slot(ice, p, sl, Fs, fa, nullptr, 1.0, H.data(), Fs_pad, work); } }
                }
                for (int j = j_hard; j < Fs; ++j) if (m0 * H[j] + m1 * H[Fs_pad + j] + m2 * H[2 * Fs_pad + j] < 30.0) redo_generic = true;
                ++n_solutions; if (redo_generic) { ++n_redo; std::fill(H.begin(), H.end(), 0.0); }
            }
            if (scheme == 0 || redo_generic) {
                int j_hard = Fs; for (int j = 0; j < Fs; ++j) if (fa[j] > a_min - 60.0) { j_hard = j; break; }
                for (int c = 0; c < n0; ++c) { Slot sl; slot_of(p, spp0, c, sl); do_slot(ice, p, sl, Fs, fa, nullptr, 1.0, H.data(), Fs_pad, work); }
                bool any = false;
                for (int j = 0; j < Fs; ++j) { sel[j] = j >= j_hard && (m0 * H[j] + m1 * H[Fs_pad + j] + m2 * H[2 * Fs_pad + j] < 30.0); if (sel[j]) { any = true; H[j] = H[Fs_pad + j] = H[2 * Fs_pad + j] = 0.0; } }
                if (scheme == 0) { ++n_solutions; if (any) ++n_redo; }
                if (any) for (int c = 0; c < n0 * fine_per_coarse; ++c) { Slot sl; slot_of(p, spp0 * fine_per_coarse, c, sl); do_slot(ice, p, sl, Fs, fa, sel.data(), 1.0, H.data(), Fs_pad, work); }
            } else if (scheme == 3) {
            } else if (scheme == 2) {
                // graded slots: panel 1 = [u2,u1] (deep end u1) in n1 slots shrinking by `ratio` toward u1; panel 0 uniform n0p slots
                const int n1 = (int)work_out[3], n0p = (int)work_out[4]; const double ratio = margin_ratio;
                int j_hard = Fs; for (int j = 0; j < Fs; ++j) if (fa[j] > a_min - margin) { j_hard = j; break; }
                for (int pa = 0; pa < 3; ++pa) {
                    if (plan_total_mult(p, pa) == 0) continue;
                    const double plo = pa == 0 ? p.uT : (pa == 1 ? p.u2 : p.u1), phi = pa == 0 ? p.u2 : (pa == 1 ? p.u1 : p.ur);
                    if (!(phi > plo)) continue;
                    const int ns = pa == 1 ? n1 : n0p;
                    double tot = 0, wdt = 1; for (int k = 0; k < ns; ++k) { tot += wdt; wdt *= (pa == 1 ? ratio : 1.0); }
                    double a = plo; wdt = (phi - plo) / tot;
                    for (int k = 0; k < ns; ++k) { Slot sl; sl.lo = a; sl.hi = (k == ns - 1) ? phi : a + wdt; sl.panel = pa; a = sl.hi; wdt *= (pa == 1 ? ratio : 1.0);
                        do_slot(ice, p, sl, Fs, fa, nullptr, 1.0, H.data(), Fs_pad, work); }
                }
                bool any = false;
                for (int j = 0; j < Fs; ++j) { sel[j] = j >= j_hard && (m0 * H[j] + m1 * H[Fs_pad + j] + m2 * H[2 * Fs_pad + j] < 30.0); if (sel[j]) { any = true; H[j] = H[Fs_pad + j] = H[2 * Fs_pad + j] = 0.0; } }
                ++n_solutions; if (any) ++n_redo;
                if (any) for (int c = 0; c < n0 * fine_per_coarse; ++c) { Slot sl; slot_of(p, spp0 * fine_per_coarse, c, sl); do_slot(ice, p, sl, Fs, fa, sel.data(), 1.0, H.data(), Fs_pad, work); }
            } else {
                // per coarse slot flags
                std::vector<double> amin_c(n0);
                for (int c = 0; c < n0; ++c) { Slot sl; slot_of(p, spp0, c, sl); amin_c[c] = do_slot(ice, p, sl, Fs, fa, nullptr, 1.0, H.data(), Fs_pad, work); }
                std::vector<unsigned char> redo(Fs);
                for (int j = 0; j < Fs; ++j) redo[j] = (m0 * H[j] + m1 * H[Fs_pad + j] + m2 * H[2 * Fs_pad + j] < 30.0);
                for (int c = 0; c < n0; ++c) {
                    bool any = false;
                    for (int j = 0; j < Fs; ++j) { sel[j] = redo[j] && fa[j] > amin_c[c] - margin; any = any || sel[j]; }
                    if (!any) continue;
                    Slot sl; slot_of(p, spp0, c, sl);
                    do_slot(ice, p, sl, Fs, fa, sel.data(), -1.0, H.data(), Fs_pad, work);         // take the coarse contribution out again
                    for (int k = 0; k < fine_per_coarse; ++k) { Slot fs; slot_of(p, spp0 * fine_per_coarse, c * fine_per_coarse + k, fs); do_slot(ice, p, fs, Fs, fa, sel.data(), 1.0, H.data(), Fs_pad, work); }
                }
            }
            // reference result order: ascending C0 = find_roots_mode order
            for (int j = 0; j < Fs; ++j) out[(i * 2 + r) * Fs + j] = exp(-(m0 * H[j] + m1 * H[Fs_pad + j] + m2 * H[2 * Fs_pad + j]));
        }
    }
    work_out[0] = work; work_out[1] = n_redo; work_out[2] = n_solutions;
    return 0;
}
