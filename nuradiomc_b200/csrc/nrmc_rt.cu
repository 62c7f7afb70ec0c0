// nrmc_rt.cu -- sm_100a kernels and the C ABI (include/nrmc_rt.h) of the batched analytic ray tracer.
//
//   K_classify / K_hump / K_roots (+ _m variants for media with a reflective bottom)
//                  binned solver: junction values of the range curve per (pair, mode) -> bracket / hump items in SoA queues in HBM
//                  (block- / warp-aggregated appends) -> maximum search -> safeguarded Newton per bracket, closed-form properties,
//                  SoA stores, solution records for the attenuation kernels (two-ended SoA work list).
//   K_att_sp1      SP1: thread per solution, 12-node panels, 12 frequency-independent Chebyshev moments in the ice temperature,
//                  coefficient table staged by TMA bulk copies (cp.async.bulk + mbarrier), table exponentials, staged coalesced
//                  row stores.
//   K_att_gl1 / K_gl1_item / K_gl1_fine
//                  GL1: thread per solution, closed-form Chebyshev series per frequency; hard (solution, frequency) items per
//                  thread / per warp.
//   K_att_sep      MB1 / GL2 (separable), sparse output: thread per solution, one depth integral per panel.
//   K_att          everything else (GL3: the reference's 10 m discretisation cell for cell; dense output of bottom-reflected
//                  paths; fall-back lists of the kernels above): warp per solution, half-warp per 16-node slot.
//   K_small        N <= 2048 pairs: one cooperative launch (thread-per-pair trace, grid barrier, generic attenuation).
//   Persistent kernels draw their work from ticket counters (warp_ticket), not from a static stride.
//   K_apply_effects, K_pack_*, K_att_expand, K_rmax_table, K_att_length, K_fp64_peak: see the comments at each kernel.
//
// No tensor cores: nothing here is a contraction.  There is no CPU fallback: every entry point
// fails with NRMC_ERR_NO_DEVICE / NRMC_ERR_CUDA if the device path is unavailable.
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <cmath>
#include <string>
#include <vector>

#include "../../include/nrmc_rt.h"
#include "nrmc_att.cuh"

using namespace nrmc;

// ---------------------------------------------------------------------------------------------------------------
// device side
// ---------------------------------------------------------------------------------------------------------------
struct KInput {
    const double *vx, *vy, *vz, *ax, *ay, *az;
    int64_t n_pairs;       // pairs in this launch
    int64_t n_antennas;    // outer-product divisor (1 pair = vertex p / n_antennas, antenna p % n_antennas) if outer
    int32_t outer;
    const double *sx, *sy, *sz;   // optional, per vertex: propagation direction of the shower (viewing-angle cut)
    double delta_C_cut;
};

__device__ __forceinline__ ShowerCut load_cut(const KInput &in, int64_t p)
{
    ShowerCut sc;
    sc.on = in.sx != nullptr;
    sc.sx = sc.sy = sc.sz = 0.0; sc.cut = in.delta_C_cut;
    if (sc.on) {
        const int64_t iv = in.outer ? p / in.n_antennas : p;
        sc.sx = __ldg(in.sx + iv); sc.sy = __ldg(in.sy + iv); sc.sz = __ldg(in.sz + iv);
    }
    return sc;
}

__device__ __forceinline__ void load_pair(const KInput &in, int64_t p, double &x1, double &y1, double &z1, double &x2,
                                          double &y2, double &z2)
{
    int64_t iv = p, ia = p;
    if (in.outer) { iv = p / in.n_antennas; ia = p - iv * in.n_antennas; }
    x1 = __ldg(in.vx + iv); y1 = __ldg(in.vy + iv); z1 = __ldg(in.vz + iv);
    x2 = __ldg(in.ax + ia); y2 = __ldg(in.ay + ia); z2 = __ldg(in.az + ia);
}

// The work list of the attenuation kernels is filled from both ends: solutions whose path has one quadrature panel
// (direct rays) from the front, two-panel paths (refracted / reflected) from the back, so that the warps of the
// thread-per-solution kernel are homogeneous.  counts[0] = front entries, counts[WL_BACK] = back entries.
// Per-lane counter block (16 x u64): [0] work front, [1] work back, [2] fallback front, [3] fallback back (unused, 0),
// [4] root queue (entries), [5] hump queue.
#define CNT_WORK 0
#define CNT_FALLBACK 2
#define CNT_ROOTS 4
#define CNT_HUMPS 5
#define CNT_ITEMS 6             // GL1: hard (solution, frequency) items, and those of them that need the fine quadrature
#define CNT_FINE 7
#define CNT_TICKET_ATT 8        // ticket counters of the persistent kernels (dynamic work distribution)
#define CNT_TICKET_ROOTS 9
#define CNT_TICKET_HUMP 10
#define CNT_TICKET_KATT 11
#define CNT_TICKET_ITEM 12
#define CNT_TICKET_FINE 13
#define CNT_WORK_PACKED 14       // K_roots: front | back << 32 (one atomic per warp), split into CNT_WORK / CNT_WORK + WL_BACK afterwards
#define CNT_STRIDE 16
static_assert(CNT_WORK_PACKED < CNT_STRIDE && CNT_TICKET_FINE < CNT_STRIDE, "ticket / packed counters must lie inside the lane's counter block (zeroed per chunk)");
#define N_LANES 3               // two pipeline lanes of the host-memory calls + the lane of device-resident calls
#define DEV_LANE 2
#define CNT_ROWBASE (N_LANES * CNT_STRIDE)   // running row base of a compact device-resident call (not reset per chunk)
#define WL_BACK 1
#define FULL_MASK 0xffffffffu

// Work list in HBM, structure of arrays: the thread-per-solution kernels read entry w with lane w, so every field is one
// contiguous 256-byte (4-byte fields: 128-byte) access per warp.  (Round 1 kept 64-byte records: every 8-byte field access
// of a warp touched 32 sectors, two thirds of the L2 traffic of K_roots was "excessive" in ncu's terms.)
struct WorkList {
    double *beta, *delta, *zv, *z1, *z2;
    int64_t *row;
    int32_t *meta;              // slot | piece << 8 | k << 16 | rcase << 24
    unsigned long long cap;
};
#define WORKLIST_BYTES_PER_ENTRY 52
static WorkList carve_worklist(void *p, unsigned long long cap)
{
    WorkList w;
    const unsigned long long c = (cap + 1ull) & ~1ull;          // keeps every array 16-byte aligned
    double *d = (double *)p;
    w.beta = d; w.delta = d + c; w.zv = d + 2 * c; w.z1 = d + 3 * c; w.z2 = d + 4 * c;
    w.row = (int64_t *)(d + 5 * c);
    w.meta = (int32_t *)(d + 6 * c);
    w.cap = cap;
    return w;
}
static size_t worklist_bytes(unsigned long long cap) { return (size_t)((cap + 1ull) & ~1ull) * WORKLIST_BYTES_PER_ENTRY + 64; }

__device__ __forceinline__ unsigned long long worklist_index(const WorkList &wl, unsigned long long n_front, unsigned long long w)
{
    return w < n_front ? w : wl.cap - 1ull - (w - n_front);
}
__device__ __forceinline__ SolRec worklist_load(const WorkList &wl, unsigned long long i)
{
    SolRec r;
    r.pair = 0;
    r.beta = wl.beta[i]; r.delta = wl.delta[i]; r.zv = wl.zv[i]; r.z1 = wl.z1[i]; r.z2 = wl.z2[i];
    r.row = wl.row[i];
    const int32_t m = wl.meta[i];
    r.slot = m & 0xff; r.piece = (uint8_t)((m >> 8) & 0xff); r.k = (uint8_t)((m >> 16) & 0xff); r.rcase = (uint8_t)((m >> 24) & 0xff); r.pad = 0;
    return r;
}
__device__ __forceinline__ void worklist_store(const WorkList &wl, unsigned long long i, const SolRec &r)
{
    wl.beta[i] = r.beta; wl.delta[i] = r.delta; wl.zv[i] = r.zv; wl.z1[i] = r.z1; wl.z2[i] = r.z2;
    wl.row[i] = r.row;
    wl.meta[i] = (r.slot & 0xff) | ((int32_t)r.piece << 8) | ((int32_t)r.k << 16) | ((int32_t)r.rcase << 24);
}
__device__ __forceinline__ SolRec worklist_get(const WorkList &wl, unsigned long long n_front, unsigned long long w)
{
    return worklist_load(wl, worklist_index(wl, n_front, w));
}
// pull the record a lane will need in its NEXT trip through a persistent loop into L2 (no registers, no shared memory): the loads
// at the top of the trip then cost an L2 hit instead of a DRAM round trip (20 % of K_att_sp1's stall samples sat on them)
__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ void worklist_prefetch(const WorkList &wl, unsigned long long i)
{
    prefetch_l2(wl.beta + i); prefetch_l2(wl.delta + i); prefetch_l2(wl.zv + i); prefetch_l2(wl.z1 + i); prefetch_l2(wl.z2 + i);
    prefetch_l2(wl.row + i); prefetch_l2(wl.meta + i);
}

// next work of a warp of a persistent kernel: one atomic per warp on a ticket counter (zeroed with the counter block of the chunk)
__device__ __forceinline__ unsigned long long warp_ticket(unsigned long long *ticket, unsigned long long n, unsigned lane)
{
    unsigned long long g = 0;
    if (lane == 0) g = atomicAdd(ticket, n);
    return __shfl_sync(FULL_MASK, g, 0);
}

#define N_OUT 15            // output arrays of nrmc_rt_output (out_ptr)
// ---------------------------------------------------------------------------------------------------------------
// Binned solver for media without bottom reflections (one mode per pair).  Pairs differ wildly in the work they need
// (shadow zone: a maximum search; lit zone: two root solves), so a thread-per-pair kernel runs with half-empty warps.
// Here the work is split at the points where it diverges and re-packed through queues in HBM (HBM is idle in this workload):
//   K_classify  thread per pair:  frame, gamma(z1), gamma(z2), the three junction values -> one root-queue ENTRY (two brackets),
//                                 or one hump item (curve entirely below rho), or "no solution"
//   K_hump      thread per hump item: maximum search with the closed-form derivative -> one root-queue entry or "no solution"
//   K_roots     thread per bracket (two lanes per entry): safeguarded Newton, closed-form properties, SoA stores, work-list
//               record for the attenuation kernels.  One shuffle orders the two roots of a pair by C0 (py:1547).
// Queue appends are warp-aggregated (ballot + one atomic per warp).  The queues are structures of arrays indexed by the
// entry, the two brackets of an entry packed as double2: every queue access of a warp is one contiguous run of sectors.
// ---------------------------------------------------------------------------------------------------------------
struct RootQ {
    int64_t *pair;              // [cap]
    double *g1, *g2;            // [cap] gamma at the two depths (the only transcendentals of the pair geometry)
    double2 *a, *ga, *b, *gb;   // [cap] bracket end points and values; .x: first bracket, .y: second
    int32_t *meta;              // [cap] bits 0-1 piece of bracket 0, 2-3 piece of bracket 1, 4: bracket 1 exists,
                                //       8-15 k bounces, 16-23 launch case, 24-31 mode index (mode_bits)
};
#define ROOTQ_BYTES_PER_ENTRY 92
static RootQ carve_rootq(void *p, size_t cap)
{
    RootQ q;
    const size_t c = (cap + 3) & ~(size_t)3;
    double *d = (double *)p;
    q.pair = (int64_t *)d; q.g1 = d + c; q.g2 = d + 2 * c;
    q.a = (double2 *)(d + 3 * c); q.ga = (double2 *)(d + 5 * c); q.b = (double2 *)(d + 7 * c); q.gb = (double2 *)(d + 9 * c);
    q.meta = (int32_t *)(d + 11 * c);
    return q;
}
static size_t rootq_bytes(size_t cap) { return ((cap + 3) & ~(size_t)3) * ROOTQ_BYTES_PER_ENTRY + 64; }

struct HumpQ { int64_t *pair; double *g1, *g2, *J1, *J2, *J3; int32_t *mode; };
#define HUMPQ_BYTES_PER_ENTRY 52
static HumpQ carve_humpq(void *p, size_t cap)
{
    HumpQ q;
    const size_t c = (cap + 3) & ~(size_t)3;
    double *d = (double *)p;
    q.pair = (int64_t *)d; q.g1 = d + c; q.g2 = d + 2 * c; q.J1 = d + 3 * c; q.J2 = d + 4 * c; q.J3 = d + 5 * c;
    q.mode = (int32_t *)(d + 6 * c);
    return q;
}
static size_t humpq_bytes(size_t cap) { return ((cap + 3) & ~(size_t)3) * HUMPQ_BYTES_PER_ENTRY + 64; }

// mode of a work item for media with bottom reflections: k bounces (bits 8-15), launch case (16-23), mode index (24-31)
__device__ __forceinline__ int mode_bits(int k, int rcase, int md) { return (k << 8) | (rcase << 16) | (md << 24); }
__device__ __forceinline__ void mode_of(int md, int &k, int &rcase) { k = md == 0 ? 0 : (md - 1) / 2 + 1; rcase = md == 0 ? 1 : (md - 1) % 2 + 1; }

// Upper bounds of the maximal range on a depth grid (range_max): entry [i1, i2] belongs to the depths (-i1 dz, -i2 dz).
struct RmaxTable { const double *t; int32_t n; double dz; };
#define RMAX_DZ 8.0
#define RMAX_N 401            // 0 ... 3200 m

__global__ void K_rmax_table(IceParams ice, int n, double dz, double *table)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * n) return;
    const int i1 = i / n, i2 = i - i1 * n;
    if (i1 < i2) return;                                     // filled by the symmetric entry
    PairGeom g;
    make_pair_geom(ice, -i1 * dz, -i2 * dz, 0.0, g);
    double r = range_max(ice, g);
    if (!(r < 1e6)) r = INFINITY;      // deep, numerically homogeneous ice: a horizontal ray reaches any distance
    table[i1 * n + i2] = r;
    table[i2 * n + i1] = r;
}

__global__ void K_set_u64(unsigned long long *p, unsigned long long v) { *p = v; }
__global__ void K_unpack_work(unsigned long long *cnt)
{
    const unsigned long long v = cnt[CNT_WORK_PACKED];
    cnt[CNT_WORK] = v & 0xffffffffull; cnt[CNT_WORK + WL_BACK] = v >> 32;
}

__device__ __forceinline__ void push_brackets(bool have, int64_t pair, const PairGeom &g, const Bracket *br, int nb, const RootQ &q,
                                              unsigned long long *root_count, unsigned lane, int mode = 0)
{
    const unsigned m = __ballot_sync(FULL_MASK, have);
    if (m == 0) return;
    const int leader = __ffs(m) - 1;
    unsigned long long base = 0;
    if ((int)lane == leader) base = atomicAdd(root_count, (unsigned long long)__popc(m));
    base = __shfl_sync(FULL_MASK, base, leader);
    if (have) {
        const unsigned long long e = base + __popc(m & ((1u << lane) - 1u));
        const Bracket &b0 = br[0], &b1 = br[nb > 1 ? 1 : 0];
        q.pair[e] = pair; q.g1[e] = g.g1; q.g2[e] = g.g2;
        q.a[e] = make_double2(b0.a, b1.a); q.ga[e] = make_double2(b0.ga, b1.ga);
        q.b[e] = make_double2(b0.b, b1.b); q.gb[e] = make_double2(b0.gb, b1.gb);
        q.meta[e] = (b0.piece & 3) | ((b1.piece & 3) << 2) | ((nb > 1 ? 1 : 0) << 4) | mode;
    }
}

__device__ __forceinline__ void push_hump(bool have, int64_t pair, const PairGeom &g, double J1, double J2, double J3, int mode, const HumpQ &q,
                                          unsigned long long *hump_count, unsigned lane)
{
    const unsigned mh = __ballot_sync(FULL_MASK, have);
    if (mh == 0) return;
    const int leader = __ffs(mh) - 1;
    unsigned long long base = 0;
    if ((int)lane == leader) base = atomicAdd(hump_count, (unsigned long long)__popc(mh));
    base = __shfl_sync(FULL_MASK, base, leader);
    if (have) {
        const unsigned long long e = base + __popc(mh & ((1u << lane) - 1u));
        q.pair[e] = pair; q.g1[e] = g.g1; q.g2[e] = g.g2; q.J1[e] = J1; q.J2[e] = J2; q.J3[e] = J3; q.mode[e] = mode;
    }
}

// NaN rows of the attenuation outputs for pairs without solution, written by the whole warp (coalesced) inside the
// compute-bound solver kernels, where the store traffic is free
struct AttFill { double *sparse, *dense; int32_t Fs, F; };

__device__ __forceinline__ void warp_fill_nan_rows(bool mine, int64_t pair, const AttFill &af, unsigned lane)
{
    if (!af.sparse && !af.dense) return;
    unsigned m = __ballot_sync(FULL_MASK, mine);
    while (m) {
        const int l = __ffs(m) - 1;
        m &= m - 1;
        const int64_t pr = __shfl_sync(FULL_MASK, pair, l);
        // both slots of the pair: 2 Fs doubles = Fs 16-byte stores (the pair's rows start on a 16-byte boundary), streaming
        const double2 nan2 = make_double2(NAN, NAN);
        if (af.sparse) { double2 *d = reinterpret_cast<double2 *>(af.sparse + pr * 2 * af.Fs); for (int j = lane; j < af.Fs; j += 32) __stcs(d + j, nan2); }
        if (af.dense) { double2 *d = reinterpret_cast<double2 *>(af.dense + pr * 2 * af.F); for (int j = lane; j < af.F; j += 32) __stcs(d + j, nan2); }
    }
}

__device__ __forceinline__ void warp_fill_nan_slot(bool mine, int64_t q, const AttFill &af, unsigned lane)
{
    if (!af.sparse && !af.dense) return;
    unsigned m = __ballot_sync(FULL_MASK, mine);
    while (m) {
        const int l = __ffs(m) - 1;
        m &= m - 1;
        const int64_t qq = __shfl_sync(FULL_MASK, q, l);
        if (af.sparse) { double *d = af.sparse + qq * af.Fs; for (int j = lane; j < af.Fs; j += 32) __stcs(d + j, NAN); }
        if (af.dense) { double *d = af.dense + qq * af.F; for (int j = lane; j < af.F; j += 32) __stcs(d + j, NAN); }
    }
}

__device__ __forceinline__ void write_no_solution(const TraceOutputs &out, int64_t p, int status)
{
    if (out.n_sol) out.n_sol[p] = 0;
    if (out.status) out.status[p] = status;
    if (out.row_offset) return;          // compact layout: empty slots have no rows
    fill_empty_slot(out, 2 * p, 1);
    fill_empty_slot(out, 2 * p + 1, 1);
}

#define CLASSIFY_THREADS 256
#ifndef CLASSIFY_MIN_BLOCKS
#define CLASSIFY_MIN_BLOCKS 6   // 40 registers + spills; A/B on the B200: 4 blocks 3.30, 5: 3.13, 6: 3.04, 8: 3.39 ms per 1e8 pairs
#endif
__global__ void __launch_bounds__(CLASSIFY_THREADS, CLASSIFY_MIN_BLOCKS)
K_classify(IceParams ice, KInput in, TraceOutputs out, AttFill af, RmaxTable rmax, RootQ rootq, unsigned long long *root_count,
           HumpQ humpq, unsigned long long *hump_count)
{
    const int64_t p = (int64_t)blockIdx.x * CLASSIFY_THREADS + threadIdx.x;
    const unsigned lane = threadIdx.x & 31u;
    int kind = 0;     // 0: nothing to queue, 1: brackets, 2: hump search
    PairGeom g;
    Bracket br[2];
    int nb = 0;
    double J1 = 0, J2 = 0, J3 = 0;
    if (p < in.n_pairs) {
        double x1, y1, z1, x2, y2, z2;
        load_pair(in, p, x1, y1, z1, x2, y2, z2);
        Frame2D f;
        make_frame(x1, y1, z1, x2, y2, z2, f);
        const int status = pair_status(ice, f);
        // shadow zone: no ray of the pair reaches further than R_max at the deeper grid corner (monotone in both depths, range_max).
        // Tested FIRST: a pair beyond the bound needs no geometry and no junction values (47 % of the cfg5 pairs have no
        // solution, and a fifth of the warps consist of such pairs only: vertices far from every station of the warp).
        bool beyond = false;
#ifndef CLASSIFY_LATE_RMAX
        if (status == 0 && rmax.t) {
            const int i1 = (int)ceil(-f.z1 / rmax.dz), i2 = (int)ceil(-f.z2 / rmax.dz);
            if (i1 < rmax.n && i2 < rmax.n && i1 >= 0 && i2 >= 0) {
                const double bound = __ldg(rmax.t + i1 * rmax.n + i2);
                beyond = f.rho > bound * (1.0 + 1e-9) + 1e-6;
            }
        }
#endif
        if (status == 0 && !beyond) {
            make_pair_geom(ice, f.z1, f.z2, fmax(f.rho, 1e-12), g);
            Curve cv;
            cv.ice = &ice; cv.g = &g; cv.k = 0; cv.rcase = 1;
            bool need_hump;
            nb = classify_mode(cv, J1, J2, J3, br, need_hump);
#ifdef CLASSIFY_LATE_RMAX
            if (need_hump && rmax.t) {
                const int i1 = (int)ceil(-f.z1 / rmax.dz), i2 = (int)ceil(-f.z2 / rmax.dz);
                if (i1 < rmax.n && i2 < rmax.n && i1 >= 0 && i2 >= 0) {
                    const double bound = __ldg(rmax.t + i1 * rmax.n + i2);
                    if (f.rho > bound * (1.0 + 1e-9) + 1e-6) need_hump = false;
                }
            }
#endif
            kind = nb > 0 ? 1 : (need_hump ? 2 : 0);
        }
        if (kind == 0) write_no_solution(out, p, status);
        if (kind == 1) { if (out.n_sol) out.n_sol[p] = nb; if (out.status) out.status[p] = 0; }
    }
    if (!out.row_offset) warp_fill_nan_rows(p < in.n_pairs && kind == 0, p, af, lane);
    // Queue appends aggregated over the BLOCK (one atomic per queue and block): the entries of 256 consecutive pairs stay
    // contiguous and in pair order, so the 16 entries a warp of K_roots works on are consecutive pairs with solutions --
    // consecutive output rows -- except where a warp straddles two blocks' runs (1 in 12).  With warp-level appends a quarter
    // of the warps of K_roots fell off the coalesced store path.
    __shared__ unsigned s_cnt[2][CLASSIFY_THREADS / 32];
    __shared__ unsigned long long s_base[2];
    const unsigned warp = threadIdx.x >> 5;
    const unsigned mr = __ballot_sync(FULL_MASK, kind == 1), mh = __ballot_sync(FULL_MASK, kind == 2);
    if (lane == 0) { s_cnt[0][warp] = __popc(mr); s_cnt[1][warp] = __popc(mh); }
    __syncthreads();
    if (threadIdx.x < 2) {
        unsigned total = 0;
        for (int w = 0; w < CLASSIFY_THREADS / 32; ++w) total += s_cnt[threadIdx.x][w];
        s_base[threadIdx.x] = total ? atomicAdd(threadIdx.x == 0 ? root_count : hump_count, (unsigned long long)total) : 0ull;
    }
    __syncthreads();
    if (kind == 1 || kind == 2) {
        const int qi = kind - 1;
        unsigned before = __popc((qi == 0 ? mr : mh) & ((1u << lane) - 1u));
        for (unsigned w = 0; w < warp; ++w) before += s_cnt[qi][w];
        const unsigned long long e = s_base[qi] + before;
        if (kind == 1) {
            const Bracket &b0 = br[0], &b1 = br[nb > 1 ? 1 : 0];
            rootq.pair[e] = p; rootq.g1[e] = g.g1; rootq.g2[e] = g.g2;
            rootq.a[e] = make_double2(b0.a, b1.a); rootq.ga[e] = make_double2(b0.ga, b1.ga);
            rootq.b[e] = make_double2(b0.b, b1.b); rootq.gb[e] = make_double2(b0.gb, b1.gb);
            rootq.meta[e] = (b0.piece & 3) | ((b1.piece & 3) << 2) | ((nb > 1 ? 1 : 0) << 4);
        } else {
            humpq.pair[e] = p; humpq.g1[e] = g.g1; humpq.g2[e] = g.g2; humpq.J1[e] = J1; humpq.J2[e] = J2; humpq.J3[e] = J3; humpq.mode[e] = 0;
        }
    }
}

#define HUMP_THREADS 128
#ifndef HUMP_MIN_BLOCKS
#define HUMP_MIN_BLOCKS 4
#endif
__global__ void __launch_bounds__(HUMP_THREADS, HUMP_MIN_BLOCKS)
K_hump(IceParams ice, KInput in, TraceOutputs out, AttFill af, HumpQ humpq, const unsigned long long *hump_count, RootQ rootq,
       unsigned long long *root_count)
{
    const unsigned lane = threadIdx.x & 31u;
    const unsigned long long n = *hump_count;
    unsigned long long *ticket = const_cast<unsigned long long *>(hump_count) + (CNT_TICKET_HUMP - CNT_HUMPS);
    unsigned long long w_next = warp_ticket(ticket, 32ull, lane);
    while (w_next < n) {
        const unsigned long long w0 = w_next;
        w_next = warp_ticket(ticket, 32ull, lane);
        const unsigned long long w = w0 + lane;
        const bool active = w < n;
        PairGeom g;
        Bracket br[2];
        int nb = 0;
        int64_t pair = 0;
        if (active) {
            pair = humpq.pair[w];
            double x1, y1, z1, x2, y2, z2;
            load_pair(in, pair, x1, y1, z1, x2, y2, z2);
            Frame2D f;
            make_frame(x1, y1, z1, x2, y2, z2, f);
            make_pair_geom_g(ice, f.z1, f.z2, fmax(f.rho, 1e-12), humpq.g1[w], humpq.g2[w], g);
            Curve cv;
            cv.ice = &ice; cv.g = &g; cv.k = 0; cv.rcase = 1;
            nb = hump_search(cv, humpq.J1[w], humpq.J2[w], humpq.J3[w], br);
            if (nb == 0) write_no_solution(out, pair, 0);
            else { if (out.n_sol) out.n_sol[pair] = nb; if (out.status) out.status[pair] = 0; }
        }
        if (!out.row_offset) warp_fill_nan_rows(active && nb == 0, pair, af, lane);
        push_brackets(active && nb > 0, pair, g, br, nb, rootq, root_count, lane);
    }
}

// bracket j of root-queue entry e (lane pairs read the two halves of the same double2: one contiguous run per warp)
__device__ __forceinline__ Bracket rootq_bracket(const RootQ &q, unsigned long long w, int meta)
{
    Bracket b;
    b.a = reinterpret_cast<const double *>(q.a)[w]; b.ga = reinterpret_cast<const double *>(q.ga)[w];
    b.b = reinterpret_cast<const double *>(q.b)[w]; b.gb = reinterpret_cast<const double *>(q.gb)[w];
    b.piece = (meta >> (2 * (int)(w & 1ull))) & 3;
    return b;
}

// launch / receive vectors of the solutions of a warp.  When the rows of the warp's solutions form one contiguous range (the
// rule: consecutive queue entries are consecutive pairs with solutions) the 3-vectors are transposed through shared memory
// and written as three fully coalesced runs; lanes storing their own 24-byte rows touch three times as many sectors.
__device__ __forceinline__ void warp_store_vec3(double *dst, bool valid, int64_t row, bool dense, int64_t row_min, int n_valid, double vx,
                                                double vy, double vz, double *stage, unsigned lane)
{
    if (!dst) return;
    if (dense) {
        if (valid) { double *s = stage + 3 * (int)(row - row_min); s[0] = vx; s[1] = vy; s[2] = vz; }
        __syncwarp();
        double *d = dst + 3 * row_min;
        for (int i = lane; i < 3 * n_valid; i += 32) d[i] = stage[i];
        __syncwarp();
    } else if (valid) {
        dst[3 * row] = vx; dst[3 * row + 1] = vy; dst[3 * row + 2] = vz;
    }
}

#define ROOTS_THREADS 128
#ifndef ROOTS_MIN_BLOCKS
#define ROOTS_MIN_BLOCKS 8      // 64 registers: 32 warps per SM (A/B on the B200 with the ticket loop: 7 blocks 10.05, 8 blocks 9.82, 9 blocks 9.81, 10 blocks 9.85 ms per 1e8 pairs)
#endif
template <bool CUT>
__global__ void __launch_bounds__(ROOTS_THREADS, ROOTS_MIN_BLOCKS)
K_roots(IceParams ice, KInput in, TraceOutputs out, AttFill af, RootQ rootq, const unsigned long long *root_count, WorkList worklist,
        unsigned long long *work_count)
{
    __shared__ double s_stage[ROOTS_THREADS / 32][96];
    const unsigned lane = threadIdx.x & 31u;
    double *stage = s_stage[threadIdx.x >> 5];
    const unsigned long long n = 2ull * *root_count;     // two lanes per entry
#ifdef ROOTS_STATIC
    const unsigned long long stride = (unsigned long long)gridDim.x * ROOTS_THREADS;
    for (unsigned long long w0 = (unsigned long long)blockIdx.x * ROOTS_THREADS + (threadIdx.x & ~31u); w0 < n; w0 += stride) {
#else
    // warps draw their next 32 brackets from a ticket counter (see K_att_sp1: a static stride leaves the SMs with fewer and
    // fewer warps towards the end of the launch, the Newton iteration counts differ from warp to warp on top)
    unsigned long long *ticket = const_cast<unsigned long long *>(root_count) + (CNT_TICKET_ROOTS - CNT_ROOTS);
    unsigned long long w_next = warp_ticket(ticket, 32ull, lane);
    while (w_next < n) {
        const unsigned long long w0 = w_next;
        w_next = warp_ticket(ticket, 32ull, lane);
#endif
        const unsigned long long w = w0 + lane;
        const bool active = w < n;
        bool valid = false;
        int64_t pair = 0;
        double viewing = NAN;
        bool keep = true;      // passes the viewing-angle cut (always, when no shower axes were given)
        // everything the stores need, computed once, before the two roots of the pair are ordered
        SolutionProps pr;
        SolRec rec;
        double ex = 1.0, ey = 0.0;
        bool swap = false;
        pr.C0 = pr.C1 = pr.path_length = pr.travel_time = pr.refl_angle = NAN;
        pr.sin_l = pr.cos_l = pr.sin_r = pr.cos_r = 0.0; pr.type = 0; pr.refl_mask = 0; pr.n_segments = 1;
        rec.beta = rec.delta = rec.zv = rec.z1 = rec.z2 = 0.0; rec.piece = 0; rec.k = 0; rec.rcase = 1; rec.pad = 0; rec.slot = 0; rec.row = 0; rec.pair = 0;
        if (active) {
            const unsigned long long e = w >> 1;
            const int meta = rootq.meta[e];
            pair = rootq.pair[e];
            valid = (w & 1ull) == 0 || ((meta >> 4) & 1) != 0;
            if (valid) {
                double x1, y1, z1, x2, y2, z2;
                load_pair(in, pair, x1, y1, z1, x2, y2, z2);
                Frame2D f;
                make_frame(x1, y1, z1, x2, y2, z2, f);
                PairGeom g;
                make_pair_geom_g(ice, f.z1, f.z2, fmax(f.rho, 1e-12), rootq.g1[e], rootq.g2[e], g);
                Curve cv;
                cv.ice = &ice; cv.g = &g; cv.k = 0; cv.rcase = 1;
                const Bracket b = rootq_bracket(rootq, w, meta);
                const Root root = solve_bracket(cv, b);
                solution_props(ice, g, f.x1y, 0, 1, root, pr);
                make_solrec(ice, g, pair, 0, 0, 0, 1, root, rec);
                ex = f.ex; ey = f.ey; swap = f.swap;
                if (CUT) {
                    // launch direction from Snell's invariant: sin = beta / n, cos = s / n at the emitter (py:1161-1199, :2583-2590)
                    const ShowerCut sc = load_cut(in, pair);
                    viewing = viewing_angle_of(sc, ex, ey, swap ? -pr.sin_r : pr.sin_l, swap ? pr.cos_r : pr.cos_l);
                    keep = passes_cut(sc, viewing, swap ? g.n2 : g.n1);
                }
            }
        }
        // order the two roots of the pair by ascending C0 = descending beta (py:1547); the partner sits in lane ^ 1
        const double beta_other = __shfl_xor_sync(FULL_MASK, valid ? rec.beta : -1.0, 1);
        int slot = lane & 1;
        if (valid) slot = (rec.beta > beta_other) ? 0 : ((rec.beta < beta_other) ? 1 : (int)(lane & 1u));
        int64_t row = valid ? row_of(out, pair, slot, 2) : 0;
        if (out.row_offset && row >= out.row_limit) valid = false;      // caller's compact arrays are full: drop (reported as NRMC_ERR_CAPACITY)
        // work-list slots of the warp's solutions (front: one quadrature panel, back: two): ONE atomic per warp on a packed counter
        // (front count in the low, back count in the high 32 bits; K_unpack_work splits it after the launch).  Its value is not
        // needed before the end of the trip: the round trip to L2 (7 % of the kernel's stall samples when the two bases were
        // fetched and used on the spot) hides behind the stores of the properties.
        const bool two_panel = rec.piece >= 2;
        unsigned mf = 0, mb = 0;
        unsigned long long wl_base = 0;
        if (worklist.beta) {
            mf = __ballot_sync(FULL_MASK, valid && keep && !two_panel); mb = __ballot_sync(FULL_MASK, valid && keep && two_panel);
            if ((mf | mb) && (int)lane == __ffs(mf | mb) - 1)
                wl_base = atomicAdd(work_count + (CNT_WORK_PACKED - CNT_WORK), (unsigned long long)__popc(mf) | ((unsigned long long)__popc(mb) << 32));
        }
        if (CUT) warp_fill_nan_slot(valid && !keep, row, af, lane);     // cut solutions: NaN attenuation rows
        // are the rows of this warp's solutions one contiguous range?
        const unsigned mv = __ballot_sync(FULL_MASK, valid);
        bool dense = false;
        int64_t row_min = 0;
        const int n_valid = __popc(mv);
        if (mv) {
            const int64_t row_ref = __shfl_sync(FULL_MASK, row, __ffs(mv) - 1);
            const int64_t dd = row - row_ref;
            const int d = valid ? (int)(dd < -64 ? -64 : (dd > 64 ? 64 : dd)) : 0;
            const int dmin = __reduce_min_sync(FULL_MASK, d), dmax = __reduce_max_sync(FULL_MASK, d);
            dense = (dmax - dmin + 1 == n_valid);
            row_min = row_ref + dmin;
        }
        if (valid) {
            if (out.type) out.type[row] = (int8_t)pr.type;
            if (out.reflection) out.reflection[row] = 0;
            if (out.reflection_case) out.reflection_case[row] = 1;
            if (out.C0) out.C0[row] = pr.C0;
            if (out.C1) out.C1[row] = pr.C1;
            if (out.path_length) out.path_length[row] = pr.path_length;
            if (out.travel_time) out.travel_time[row] = pr.travel_time;
            if (out.reflection_angle) out.reflection_angle[row] = (pr.refl_mask & 1u) ? pr.refl_angle : NAN;
            if (out.viewing_angle) out.viewing_angle[row] = viewing;
        }
        {
            // 2-D vectors -> 3-D: R^T [vx,0,vz] = [vx ex, vx ey, vz]; roles exchanged when swapped (py:2583-2590,2617-2623)
            double lx = pr.sin_l, lz = pr.cos_l, rx = -pr.sin_r, rz = pr.cos_r;
            if (swap) { const double tx = lx, tz = lz; lx = rx; lz = rz; rx = tx; rz = tz; }
            warp_store_vec3(out.launch, valid, row, dense, row_min, n_valid, lx * ex, lx * ey, lz, stage, lane);
            warp_store_vec3(out.receive, valid, row, dense, row_min, n_valid, rx * ex, rx * ey, rz, stage, lane);
        }
        if (mf | mb) {
            wl_base = __shfl_sync(FULL_MASK, wl_base, __ffs(mf | mb) - 1);
            const unsigned below = (1u << lane) - 1u;
            if (valid && keep) {
                rec.slot = slot; rec.row = row;
                worklist_store(worklist, two_panel ? worklist.cap - 1ull - ((wl_base >> 32) + __popc(mb & below))
                                                   : (wl_base & 0xffffffffull) + __popc(mf & below), rec);
            }
        }
        if (active && !valid && !out.row_offset && (w & 1ull)) {
            fill_empty_slot(out, 2 * pair + 1, 1);    // single root (tangency at the surface): second slot stays empty
            if (af.sparse) for (int j = 0; j < af.Fs; ++j) af.sparse[(2 * pair + 1) * af.Fs + j] = NAN;
            if (af.dense) for (int j = 0; j < af.F; ++j) af.dense[(2 * pair + 1) * af.F + j] = NAN;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// The same binned pipeline for media with a reflective bottom: the work item is a (pair, mode) with M = 1 + 2 n_reflections
// modes (k bounces, launch up / down).  K_classify_m / K_hump_m record the number of roots of every mode, K_slots_m turns
// the counts of a pair into slot offsets (the reference's result order: mode first, then C0 ascending, py:2122-2125) and
// fills the unused slots, K_roots_m writes every solution to its slot.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(CLASSIFY_THREADS, 4)      // (5 blocks the same, 6 and 8 slower: A/B on cfg4)
K_classify_m(IceParams ice, KInput in, TraceOutputs out, RmaxTable rmax, int M, int8_t *mode_count, RootQ rootq,
             unsigned long long *root_count, HumpQ humpq, unsigned long long *hump_count)
{
    const int64_t t = (int64_t)blockIdx.x * CLASSIFY_THREADS + threadIdx.x;
    const unsigned lane = threadIdx.x & 31u;
#ifdef CLASSIFY_M_PAIR_MAJOR
    const bool in_range = t < in.n_pairs * M;
    const int64_t p = in_range ? t / M : 0;
    const int md = in_range ? (int)(t - p * M) : 0;
#else
    // a WARP works on one mode of 32 consecutive pairs (tiles of 32 pairs x M modes): the k = 0 and k > 0 branches of the range
    // curve, the launch cases and the bounce counts do not diverge inside a warp, and the queue entries a warp appends -- hence
    // the entries a warp of K_roots_m reads -- are of one mode too.  (Pair-major items, t = pair M + mode, put all modes of a
    // pair into neighbouring lanes: every warp ran both branches.)
    const int64_t tile = t / (32 * M);
    const int r = (int)(t - tile * (32 * M));
    const int md_raw = r >> 5;
    const int64_t p_raw = tile * 32 + (r & 31);
    const bool in_range = p_raw < in.n_pairs;
    const int64_t p = in_range ? p_raw : 0;
    const int md = in_range ? md_raw : 0;
#endif
    int k, rcase;
    mode_of(md, k, rcase);
    int kind = 0, nb = 0;
    PairGeom g;
    Bracket br[2];
    double J1 = 0, J2 = 0, J3 = 0;
    if (in_range) {
        double x1, y1, z1, x2, y2, z2;
        load_pair(in, p, x1, y1, z1, x2, y2, z2);
        Frame2D f;
        make_frame(x1, y1, z1, x2, y2, z2, f);
        const int status = pair_status(ice, f);
        if (md == 0 && out.status) out.status[p] = status;
        if (status == 0) {
            make_pair_geom(ice, f.z1, f.z2, fmax(f.rho, 1e-12), g);
            Curve cv;
            cv.ice = &ice; cv.g = &g; cv.k = k; cv.rcase = rcase;
            bool need_hump;
            nb = classify_mode(cv, J1, J2, J3, br, need_hump);
            if (need_hump && md == 0 && rmax.t) {
                const int i1 = (int)ceil(-f.z1 / rmax.dz), i2 = (int)ceil(-f.z2 / rmax.dz);
                if (i1 < rmax.n && i2 < rmax.n && i1 >= 0 && i2 >= 0 && f.rho > __ldg(rmax.t + i1 * rmax.n + i2) * (1.0 + 1e-9) + 1e-6) need_hump = false;
            }
            kind = nb > 0 ? 1 : (need_hump ? 2 : 0);
        }
        if (kind != 2) mode_count[p * M + md] = (int8_t)(kind == 1 ? nb : 0);
    }
    push_brackets(kind == 1, p, g, br, nb, rootq, root_count, lane, mode_bits(k, rcase, md));
    push_hump(kind == 2, p, g, J1, J2, J3, mode_bits(k, rcase, md), humpq, hump_count, lane);
}

__global__ void __launch_bounds__(HUMP_THREADS, HUMP_MIN_BLOCKS)
K_hump_m(IceParams ice, KInput in, int M, int8_t *mode_count, HumpQ humpq, const unsigned long long *hump_count, RootQ rootq,
         unsigned long long *root_count)
{
    const unsigned lane = threadIdx.x & 31u;
    const unsigned long long n = *hump_count;
    unsigned long long *ticket = const_cast<unsigned long long *>(hump_count) + (CNT_TICKET_HUMP - CNT_HUMPS);
    unsigned long long w_next = warp_ticket(ticket, 32ull, lane);
    while (w_next < n) {
        const unsigned long long w0 = w_next;
        w_next = warp_ticket(ticket, 32ull, lane);
        const unsigned long long w = w0 + lane;
        const bool active = w < n;
        PairGeom g;
        Bracket br[2];
        int nb = 0, mode = 0;
        int64_t pair = 0;
        if (active) {
            pair = humpq.pair[w]; mode = humpq.mode[w];
            double x1, y1, z1, x2, y2, z2;
            load_pair(in, pair, x1, y1, z1, x2, y2, z2);
            Frame2D f;
            make_frame(x1, y1, z1, x2, y2, z2, f);
            make_pair_geom_g(ice, f.z1, f.z2, fmax(f.rho, 1e-12), humpq.g1[w], humpq.g2[w], g);
            Curve cv;
            cv.ice = &ice; cv.g = &g; cv.k = (mode >> 8) & 0xff; cv.rcase = (mode >> 16) & 0xff;
            nb = hump_search(cv, humpq.J1[w], humpq.J2[w], humpq.J3[w], br);
            mode_count[pair * M + ((mode >> 24) & 0xff)] = (int8_t)nb;
        }
        push_brackets(active && nb > 0, pair, g, br, nb, rootq, root_count, lane, mode & ~0xff);
    }
}

// counts of the modes of a pair -> slot offsets (in place), n_sol, and the fill of the unused slots (padded layout)
__global__ void __launch_bounds__(256)
K_slots_m(KInput in, TraceOutputs out, AttFill af, int M, int S, int K1, int8_t *mode_count)
{
    const int64_t p = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (p >= in.n_pairs) return;
    int n = 0;
    for (int md = 0; md < M; ++md) { const int c = mode_count[p * M + md]; mode_count[p * M + md] = (int8_t)n; n += c; }
    if (out.n_sol) out.n_sol[p] = n;
    if (out.row_offset) return;          // compact layout: unused slots have no rows
    for (int sl = n; sl < S; ++sl) {
        fill_empty_slot(out, p * S + sl, K1);
        if (af.sparse) for (int j = 0; j < af.Fs; ++j) af.sparse[(p * S + sl) * af.Fs + j] = NAN;
        if (af.dense) for (int j = 0; j < af.F; ++j) af.dense[(p * S + sl) * af.F + j] = NAN;
    }
}

#ifndef ROOTS_M_MIN_BLOCKS
#define ROOTS_M_MIN_BLOCKS 8    // 64 registers (151 unconstrained: 3 blocks); cfg4 on the B200: 3 blocks 7.39, 4: 6.42, 5: 5.90, 6: 5.81, 8: 5.59 ms
#endif
__global__ void __launch_bounds__(ROOTS_THREADS, ROOTS_M_MIN_BLOCKS)
K_roots_m(IceParams ice, KInput in, TraceOutputs out, AttFill af, int M, int S, int K1, const int8_t *slot_base, RootQ rootq,
          const unsigned long long *root_count, WorkList worklist, unsigned long long *work_count)
{
    const unsigned lane = threadIdx.x & 31u;
    const unsigned long long n = 2ull * *root_count;
    unsigned long long *ticket = const_cast<unsigned long long *>(root_count) + (CNT_TICKET_ROOTS - CNT_ROOTS);
    unsigned long long w_next = warp_ticket(ticket, 32ull, lane);
    while (w_next < n) {
        const unsigned long long w0 = w_next;
        w_next = warp_ticket(ticket, 32ull, lane);
        const unsigned long long w = w0 + lane;
        const bool active = w < n;
        bool valid = false, keep = true;
        int64_t pair = 0;
        int k = 0, rcase = 1, md = 0;
        double viewing = NAN;
        SolutionProps pr;
        SolRec rec;
        Frame2D f;
        f.ex = 1.0; f.ey = 0.0; f.swap = false; f.z1 = f.z2 = f.rho = f.x1y = 0.0;
        rec.beta = 0.0;
        if (active) {
            const unsigned long long e = w >> 1;
            const int meta = rootq.meta[e];
            pair = rootq.pair[e];
            valid = (w & 1ull) == 0 || ((meta >> 4) & 1) != 0;
            k = (meta >> 8) & 0xff; rcase = (meta >> 16) & 0xff; md = (meta >> 24) & 0xff;
            if (valid) {
                double x1, y1, z1, x2, y2, z2;
                load_pair(in, pair, x1, y1, z1, x2, y2, z2);
                make_frame(x1, y1, z1, x2, y2, z2, f);
                PairGeom g;
                make_pair_geom_g(ice, f.z1, f.z2, fmax(f.rho, 1e-12), rootq.g1[e], rootq.g2[e], g);
                Curve cv;
                cv.ice = &ice; cv.g = &g; cv.k = k; cv.rcase = rcase;
                const Bracket b = rootq_bracket(rootq, w, meta);
                const Root root = solve_bracket(cv, b);
                solution_props(ice, g, f.x1y, k, rcase, root, pr);
                make_solrec(ice, g, pair, 0, 0, k, rcase, root, rec);
                if (in.sx) {
                    const ShowerCut sc = load_cut(in, pair);
                    double lx, lz;
                    launch_2d(f, pr, lx, lz);
                    viewing = viewing_angle_of(sc, f.ex, f.ey, lx, lz);
                    keep = passes_cut(sc, viewing, f.swap ? g.n2 : g.n1);
                }
            }
        }
        // the two roots of a mode sit in adjacent lanes: C0 ascending = beta descending (py:1547)
        const double beta_other = __shfl_xor_sync(FULL_MASK, valid ? rec.beta : -1.0, 1);
        int rank = lane & 1;
        if (valid) rank = (rec.beta > beta_other) ? 0 : ((rec.beta < beta_other) ? 1 : (int)(lane & 1u));
        int64_t row = 0;
        if (valid) {
            const int slot = slot_base[pair * M + md] + rank;
            row = row_of(out, pair, slot, S);
            if (out.row_offset && row >= out.row_limit) valid = false;       // caller's compact arrays are full
            if (valid) {
                write_solution(out, row, K1, f, k, rcase, pr);
                if (out.viewing_angle) out.viewing_angle[row] = viewing;
                if (worklist.beta && keep) {
                    rec.slot = slot; rec.row = row;
                    worklist_store(worklist, atomicAdd(work_count, 1ull), rec);
                }
            }
        }
        warp_fill_nan_slot(valid && !keep, row, af, lane);
    }
}

// ---- TMA bulk copy helpers (1-D, global -> shared, completion on an mbarrier) --------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t done = 0;
    while (!done) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    }
}
__device__ __forceinline__ void tma_bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

__constant__ double c_glx[16] = {-0.9894009349916499325961541734503326, -0.9445750230732325760779884155346083,
    -0.8656312023878317438804678977123931, -0.7554044083550030338951011948474422, -0.6178762444026437484466717640487910,
    -0.4580167776572273863424194429835775, -0.2816035507792589132304605014604961, -0.0950125098376374401853193354249581,
    0.0950125098376374401853193354249581, 0.2816035507792589132304605014604961, 0.4580167776572273863424194429835775,
    0.6178762444026437484466717640487910, 0.7554044083550030338951011948474422, 0.8656312023878317438804678977123931,
    0.9445750230732325760779884155346083, 0.9894009349916499325961541734503326};
__constant__ double c_glw[16] = {0.0271524594117540948517805724560181, 0.0622535239386478928628438369943776,
    0.0951585116824927848099251076022462, 0.1246289712555338720524762821920164, 0.1495959888165767320815017305474785,
    0.1691565193950025381893120790303599, 0.1826034150449235888667636679692199, 0.1894506104550684962853967232082831,
    0.1894506104550684962853967232082831, 0.1826034150449235888667636679692199, 0.1691565193950025381893120790303599,
    0.1495959888165767320815017305474785, 0.1246289712555338720524762821920164, 0.0951585116824927848099251076022462,
    0.0622535239386478928628438369943776, 0.0271524594117540948517805724560181};

// 12-point Gauss-Legendre rule, positive half (the rule is symmetric)
__constant__ double c_glx12h[6] = {1.25233408511468913e-01, 3.67831498998180184e-01, 5.87317954286617483e-01, 7.69902674194304693e-01, 9.04117256370474798e-01, 9.81560634246719244e-01};
__constant__ double c_glw12h[6] = {2.49147045813402690e-01, 2.33492536538354611e-01, 2.03167426723065730e-01, 1.60078328543346415e-01, 1.06939325995319065e-01, 4.71753363865114114e-02};
// SP1 depth dependence (attenuation.py:141-142 temperature profile, :176-178 b-coefficients) in constant memory
__constant__ double c_sp1[16] = {1.83415e-09, -1.59061e-08, 0.00267687, -51.0696,
                                 -6.74890, 0.026709, -0.000884, -6.22121, -0.070927, -0.001773, -4.09468, -0.002213, -0.000332,
                                 1.0 / 9.210340371976182, 1.0 / 1.1505720275988207, 0.0};

struct AttTables {            // device pointers, every array padded to a multiple of 16 bytes
    const double *fa, *fb;    // [Fs_pad] per integration frequency constants (att_freq_consts)
    const double *it;         // [F_pad]  interpolation weight t of every output bin
    const int32_t *ii;        // [F_pad]  left sparse index of every output bin; -1: bin <= 0 Hz (factor 1)
    int32_t Fs, Fs_pad, F, F_pad;
    Gl3Table gl3;
};

#define ATT_WARPS 4
#define ATT_THREADS (ATT_WARPS * 32)

__device__ __forceinline__ void stage_tables(uint64_t *bar, void *dst0, const void *src0, uint32_t b0, void *dst1, const void *src1,
                                             uint32_t b1, void *dst2, const void *src2, uint32_t b2, void *dst3, const void *src3,
                                             uint32_t b3)
{
    // per-frequency tables -> shared memory with TMA bulk copies: one elected thread issues, everybody waits on the mbarrier
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_expect_tx(bar, b0 + b1 + b2 + b3);
        if (b0) tma_bulk_g2s(dst0, src0, b0, bar);
        if (b1) tma_bulk_g2s(dst1, src1, b1, bar);
        if (b2) tma_bulk_g2s(dst2, src2, b2, bar);
        if (b3) tma_bulk_g2s(dst3, src3, b3, bar);
    }
    mbar_wait(bar, 0);
}

// ---------------------------------------------------------------------------------------------------------------
// GL3: the reference does not integrate ds/L for this model (300 rows of measured, rough data) but DEFINES the result
// by a discretisation (analyticraytracing.py:62, :458, :998-1064): per path segment, in the mirrored depth coordinate
// t in [z_a, t_b], np.linspace cells of ~10 m, ds at the cell centre times the cell width over L(z_centre, f); the
// cell around the turning point (+-10 m) is replaced by its exact path length over L(z_turn, f).  Reproduced cell for
// cell: lanes = cells, warp reduction per frequency.  fac[s][j] receives exp(-I) of segment s.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int gl3_n_steps(double a, double b) { const int n = (int)floor(fabs(a - b) / 10.0); return n > 3 ? n : 3; }  // py:65-67

// path-length antiderivative F(z) of the ray (py:602-690), z <= apex, with n - beta formed without cancellation
__device__ __forceinline__ double gl3_F(const IceParams &ice, double beta, double delta, double zv, double c, double rc, double z)
{
    const double nmb = -delta * expm1(-(zv - z) * ice.inv_z0);      // n(z) - beta >= 0
    const double n = beta + nmb;
    const double sz = sqrt(fmax(nmb * (n + beta), 0.0));
    const double k1 = rc * sz + (c - ice.n_ice * (ice.n_ice - n));
    const double k2 = sz + n;
    return ice.n_ice / rc * (z - ice.z0 * log(k1)) + ice.z0 * log(k2);
}

__device__ __forceinline__ void gl3_path(const IceParams &ice, const SolRec &rec, const Gl3Table &gl3, const double *s_f, int Fs, int Fs_pad,
                                         double *fac, int lane)
{
    const double zT = fmin(rec.zv, 0.0);                            // get_turning_point clamps to the surface (py:152-156)
    const double beta = rec.beta, delta = rec.delta;
    const double c = delta * (ice.n_ice + beta), rc = sqrt(c);
    const int k = rec.k;
    const bool turned_last = rec.piece >= 2;
    for (int s = 0; s <= k; ++s) {
        // the segment as the reference integrates it (get_path_segments py:1091-1159, first-segment mirroring py:943-950)
        double za, zb;
        bool turned;
        if (k == 0) { za = rec.z1; zb = rec.z2; turned = turned_last; }
        else if (s == 0) {
            if (rec.rcase == 1) { za = rec.z1; zb = ice.zr; turned = true; }
            else { za = ice.zr; zb = rec.z1; turned = false; }
        } else if (s < k) { za = ice.zr; zb = ice.zr; turned = true; }
        else { za = ice.zr; zb = rec.z2; turned = turned_last; }
        const double ta = za, tb = turned ? 2.0 * zT - zb : zb;       // get_z_mirrored (py:496-511)
        for (int j = lane; j < Fs; j += 32) fac[s * Fs_pad + j] = 0.0;
        __syncwarp();
        const bool fallback = (ta - 10.0 < zT) && (zT < tb + 10.0);
        double w0 = ta, w1 = tb;
        int nA = 1, nB = 1, n_cells;
        if (fallback) {
            w0 = fmax(ta, zT - 10.0); w1 = fmin(zT + 10.0, tb);
            nA = (ta == w0) ? 1 : gl3_n_steps(ta, w0);
            nB = (w1 == tb) ? 1 : gl3_n_steps(w1, tb);
            n_cells = nA + nB - 1;
        } else {
            nA = (ta == tb) ? 1 : gl3_n_steps(ta, tb);
            n_cells = nA - 1;
        }
        for (int base = 0; base < n_cells; base += 32) {
            const int cell = base + lane;
            const bool live = cell < n_cells;
            double w = 0.0, slope = 0.0, off = 1.0;
            if (live) {
                double lo, hi;
                bool window = false;
                if (!fallback) {
                    const double step = (tb - ta) / (nA - 1);
                    lo = ta + cell * step; hi = (cell + 1 == nA - 1) ? tb : ta + (cell + 1) * step;
                } else if (cell < nA - 1) {
                    const double step = (w0 - ta) / (nA - 1);
                    lo = ta + cell * step; hi = (cell + 1 == nA - 1) ? w0 : ta + (cell + 1) * step;
                } else if (cell == nA - 1) {
                    lo = w0; hi = w1; window = true;
                } else {
                    const int i = cell - nA;
                    const double step = (tb - w1) / (nB - 1);
                    lo = w1 + i * step; hi = (i + 1 == nB - 1) ? tb : w1 + (i + 1) * step;
                }
                double depth;
                if (!window) {
                    const double mid = lo + (hi - lo) / 2.0;
                    const double z = mid > zT ? 2.0 * zT - mid : mid;            // get_z_unmirrored (py:293-304)
                    const double nmb = -delta * expm1(-(rec.zv - z) * ice.inv_z0);
                    const double n = beta + nmb;
                    w = n / sqrt(nmb * (n + beta)) * (hi - lo);                 // ds(mid) * dx   (py:513-517, :1030-1035)
                    depth = -z;
                } else {
                    const double zl = lo > zT ? 2.0 * zT - lo : lo, zh = hi > zT ? 2.0 * zT - hi : hi;
                    const double Fl = gl3_F(ice, beta, delta, rec.zv, c, rc, zl), Fh = gl3_F(ice, beta, delta, rec.zv, c, rc, zh);
                    if (hi <= zT) w = Fh - Fl;
                    else if (lo >= zT) w = Fl - Fh;
                    else w = 2.0 * gl3_F(ice, beta, delta, rec.zv, c, rc, zT) - Fl - Fh;   // integral of ds over the cell (py:1058)
                    depth = -zT;
                }
                slope = gl3_lookup(gl3, depth, 1);
                off = gl3_lookup(gl3, depth, 2);
            }
            for (int j = 0; j < Fs; ++j) {
                double term = w / fmax(slope * s_f[j] + off, 1.0);             // attenuation.py:206-222, 1 m floor :252-255
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) term += __shfl_xor_sync(0xffffffffu, term, d);
                if (lane == 0) fac[s * Fs_pad + j] += term;
            }
            __syncwarp();
        }
        for (int j = lane; j < Fs; j += 32) fac[s * Fs_pad + j] = exp(-fac[s * Fs_pad + j]);
        __syncwarp();
    }
}

// deepest point of the path of a work-list record: the emitter side end point, or the reflective layer for bottom-reflected paths
__device__ __forceinline__ double k_deepest(const IceParams &ice, const SolRec &rec) { return rec.k > 0 ? ice.zr : rec.z1; }

// Generic attenuation kernel (all models, any number of bottom reflections): one warp per solution.
// dynamic shared memory (doubles): fa[Fs_pad] fb[Fs_pad] it[F_pad] | ii[F_pad] (int32) | per warp: H[3][Fs_pad] fac[nseg][Fs_pad]
template <bool GL3>
__device__ __forceinline__ void att_generic_body(const IceParams &ice, const AttTables &tb, const WorkList &worklist, const unsigned long long *work_count,
                                                 unsigned long long *ticket, int nseg_max, double *att_sparse, double *att_dense,
                                                 unsigned char *smem_raw, uint64_t &bar)
{
    const bool dense = att_dense != nullptr;
    const int Fd_pad = dense ? tb.F_pad : 0;
    double *s_fa = reinterpret_cast<double *>(smem_raw);
    double *s_fb = s_fa + tb.Fs_pad;
    double *s_it = s_fb + tb.Fs_pad;
    int32_t *s_ii = reinterpret_cast<int32_t *>(s_it + Fd_pad);
    double *s_warp = reinterpret_cast<double *>(s_ii + Fd_pad);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int per_warp = (3 + nseg_max) * tb.Fs_pad;
    double *H = s_warp + warp * per_warp;            // [3][Fs_pad] panel integrals
    double *fac = H + 3 * tb.Fs_pad;                 // [nseg][Fs_pad] exp(-I_seg)
    stage_tables(&bar, s_fa, tb.fa, (uint32_t)tb.Fs_pad * 8u, s_fb, tb.fb, (uint32_t)tb.Fs_pad * 8u, s_it, tb.it,
                 (uint32_t)Fd_pad * 8u, s_ii, tb.ii, (uint32_t)Fd_pad * 4u);

    const int q = lane & 15, half = lane >> 4;
    const double xq = c_glx[q], wq = c_glw[q];
    const unsigned long long n_front = work_count[0], n_work = n_front + work_count[WL_BACK];
    // one solution per warp: drawn from the ticket counter (persistent grid of K_att), or by a static stride when there is no
    // counter (K_small: the grid covers the work list, a ticket would only add two atomic round trips to a 0.1 ms launch)
    const unsigned long long w_stride = (unsigned long long)gridDim.x * ATT_WARPS;
    unsigned long long w_next = ticket ? warp_ticket(ticket, 1ull, (unsigned)lane) : (unsigned long long)blockIdx.x * ATT_WARPS + warp;
    while (w_next < n_work) {
        const unsigned long long w = w_next;
        w_next = ticket ? warp_ticket(ticket, 1ull, (unsigned)lane) : w + w_stride;
        const SolRec rec = worklist_get(worklist, n_front, w);
        AttPlan plan;
        att_plan_rec(ice, rec, plan);
        // GL1: frequencies [j_hard, Fs) come close to the pole of 1/max(A(z) - s_f, 1) or cross its 1 m floor somewhere on the
        // path and are integrated on NRMC_GL1_SPP sub-panels; the others on NRMC_GL1_SPP_EASY (s_f ascends with frequency).
        // A(z) (the 75 MHz length) falls with depth below ~1 km and is flat above: its minimum sits at an end of the path.
        int j_hard = tb.Fs;
        if (ice.att_model == NRMC_ATT_GL1) {
            AttNode a_deep, a_top;
            const double z_lo = k_deepest(ice, rec), z_top = rec.piece >= 2 ? fmin(rec.zv, 0.0) : rec.z2;
            att_node(NRMC_ATT_GL1, z_lo, tb.gl3, a_deep);
            att_node(NRMC_ATT_GL1, z_top, tb.gl3, a_top);
            const double a_min = fmin(a_deep.p0, a_top.p0);
            for (int j = lane; j < tb.Fs; j += 32) if (s_fa[j] > a_min - 60.0) j_hard = min(j_hard, j);
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) j_hard = min(j_hard, __shfl_xor_sync(0xffffffffu, j_hard, d));
        }
        const int64_t slot_index = rec.row;
        if (GL3) {
            gl3_path(ice, rec, tb.gl3, s_fa, tb.Fs, tb.Fs_pad, fac, lane);
            for (int j = lane; j < tb.Fs; j += 32) {
                double prod = 1.0;
                for (int s = 0; s < plan.nseg; ++s) prod *= fac[s * tb.Fs_pad + j];
                if (att_sparse) att_sparse[slot_index * tb.Fs + j] = prod;
            }
        } else {
        for (int j = lane; j < 3 * tb.Fs_pad; j += 32) H[j] = 0.0;
        __syncwarp();
        // MB1 and GL2 are separable, L(z, f) = fa(f) p0(z) (attenuation.py:198-204, :229-244): away from the 1 m floor the
        // exponent of every frequency is ONE depth integral over fa(f), G = int ds / p0; where fa p0 <= 1 on the whole path
        // (GL2 above 1.58 GHz: negative length, floored) it is the path length S = int ds.  Only a frequency whose floor is
        // crossed ON the path needs the per-(node, frequency) loop below.
        bool need_loop = true;
        if (ice.att_model == NRMC_ATT_MB1 || ice.att_model == NRMC_ATT_GL2) {
            double Gp0 = 0, Gp1 = 0, Gp2 = 0, Sp0 = 0, Sp1 = 0, Sp2 = 0, pmin = INFINITY, pmax = -INFINITY;
            for (int pass = 0; pass * 2 < plan.n_slots; ++pass) {
                const int slot = pass * 2 + half;
                const bool live = slot < plan.n_slots;
                double lo = 0.0, hi = 0.0, z = 0.0, wds = 0.0, g = 0.0;
                int panel = -1;
                if (live) {
                    AttNode nd;
                    plan_slot(plan, slot, lo, hi, panel);
                    att_node_geometry(ice, plan, lo, hi, xq, wq, z, wds);
                    att_node(ice.att_model, z, tb.gl3, nd);
                    g = wds / nd.p0;
                    pmin = fmin(pmin, nd.p0); pmax = fmax(pmax, nd.p0);
                }
#pragma unroll
                for (int d = 8; d > 0; d >>= 1) { g += __shfl_xor_sync(0xffffffffu, g, d); wds += __shfl_xor_sync(0xffffffffu, wds, d); }
                const double g_o = __shfl_xor_sync(0xffffffffu, g, 16), w_o = __shfl_xor_sync(0xffffffffu, wds, 16);
                const int panel_o = __shfl_xor_sync(0xffffffffu, panel, 16);
                // every lane keeps the panel sums (its own half-warp's slot and the other one's)
                if (panel == 0) { Gp0 += g; Sp0 += wds; } else if (panel == 1) { Gp1 += g; Sp1 += wds; } else if (panel == 2) { Gp2 += g; Sp2 += wds; }
                if (panel_o == 0) { Gp0 += g_o; Sp0 += w_o; } else if (panel_o == 1) { Gp1 += g_o; Sp1 += w_o; } else if (panel_o == 2) { Gp2 += g_o; Sp2 += w_o; }
            }
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) { pmin = fmin(pmin, __shfl_xor_sync(0xffffffffu, pmin, d)); pmax = fmax(pmax, __shfl_xor_sync(0xffffffffu, pmax, d)); }
            bool mixed = false;
            for (int j = lane; j < tb.Fs; j += 32) {
                const double fa = s_fa[j];
                if (fa * pmin >= 1.0 && fa * pmax >= 1.0) { const double inv = 1.0 / fa; H[j] = Gp0 * inv; H[tb.Fs_pad + j] = Gp1 * inv; H[2 * tb.Fs_pad + j] = Gp2 * inv; }
                else if (fa * pmin <= 1.0 && fa * pmax <= 1.0) { H[j] = Sp0; H[tb.Fs_pad + j] = Sp1; H[2 * tb.Fs_pad + j] = Sp2; }
                else mixed = true;
            }
            need_loop = __any_sync(0xffffffffu, mixed);
            __syncwarp();
            if (need_loop) { for (int j = lane; j < 3 * tb.Fs_pad; j += 32) H[j] = 0.0; __syncwarp(); }
        }
        // quadrature: each half-warp integrates one 16-node slot per pass; slot sums are added to their panel.
        // Phase 0: all frequencies below j_hard on the plan's sub-panels; phase 1 (GL1 only): the hard ones on the fine ones.
        for (int phase = 0; phase < 2 && need_loop; ++phase) {
            // phase 0: every frequency on the plan's sub-panels.  Phase 1 (GL1): the frequencies from j_hard on are redone on
            // the fine sub-panels -- unless the coarse exponent already exceeds 30: such a factor (< 1e-13) stays invisible
            // even if the coarse quadrature is off by a third, while a visible factor (I < 16) never passes the test.
            const int j_begin = phase == 0 ? 0 : j_hard, j_end = tb.Fs;
            if (j_begin >= j_end) continue;
            if (phase == 1) {
                bool redo = false;
                const int m0 = plan_total_mult(plan, 0), m1 = plan_total_mult(plan, 1), m2 = plan_total_mult(plan, 2);
                for (int j = j_hard + lane; j < tb.Fs; j += 32) {
                    const bool r = m0 * H[j] + m1 * H[tb.Fs_pad + j] + m2 * H[2 * tb.Fs_pad + j] < 30.0;
                    fac[j] = r ? 1.0 : 0.0;                  // fac is free until the factors are formed
                    if (r) { H[j] = 0.0; H[tb.Fs_pad + j] = 0.0; H[2 * tb.Fs_pad + j] = 0.0; }
                    redo = redo || r;
                }
                __syncwarp();
                if (!__any_sync(0xffffffffu, redo)) continue;
            }
            if (phase == 1) {
                plan.spp = (plan.spp / NRMC_GL1_SPP_EASY) * NRMC_GL1_SPP;
                plan.n_slots = plan.na * plan.spp;
            }
        for (int pass = 0; pass * 2 < plan.n_slots; ++pass) {
            const int slot = pass * 2 + half;
            const bool live = slot < plan.n_slots;
            double lo = 0.0, hi = 0.0, z = 0.0, wds = 0.0;
            int panel = 0;
            AttNode nd;
            nd.p0 = nd.p1 = nd.p2 = 0.0;
            if (live) {
                plan_slot(plan, slot, lo, hi, panel);
                att_node_geometry(ice, plan, lo, hi, xq, wq, z, wds);
                att_node(ice.att_model, z, tb.gl3, nd);
            }
            const int panel_other = __shfl_xor_sync(0xffffffffu, live ? panel : -1, 16);
            const bool merge = (panel_other == panel);       // both half-warps work on the same panel
            for (int j = j_begin; j < j_end; ++j) {
                if (phase == 1 && fac[j] == 0.0) continue;       // warp-uniform
                double term = live ? wds * att_inv_length(ice.att_model, nd, s_fa[j], s_fb[j]) : 0.0;
                term += __shfl_xor_sync(0xffffffffu, term, 8);
                term += __shfl_xor_sync(0xffffffffu, term, 4);
                term += __shfl_xor_sync(0xffffffffu, term, 2);
                term += __shfl_xor_sync(0xffffffffu, term, 1);
                const double other = __shfl_xor_sync(0xffffffffu, term, 16);
                if (merge) { if (lane == 0) H[panel * tb.Fs_pad + j] += term + other; }
                else if (q == 0 && live) H[panel * tb.Fs_pad + j] += term;
            }
            __syncwarp();
        }
        }
        // per segment: I_seg = sum_panel mult[seg][panel] * H[panel];  factor = exp(-I_seg)   (py:1075)
        for (int j = lane; j < tb.Fs; j += 32) {
            const double h0 = H[j], h1 = H[tb.Fs_pad + j], h2 = H[2 * tb.Fs_pad + j];
            double prod = 1.0;
            for (int s = 0; s < plan.nseg; ++s) {
                const double I = plan_mult(plan, s, 0) * h0 + plan_mult(plan, s, 1) * h1 + plan_mult(plan, s, 2) * h2;
                const double e = exp(-I);
                fac[s * tb.Fs_pad + j] = e;
                prod *= e;
            }
            if (att_sparse) att_sparse[slot_index * tb.Fs + j] = prod;
        }
        }
        __syncwarp();
        if (dense) {
            // np.interp of every segment's factors onto the output grid, product over segments (py:1077-1078,1086)
            double *dst = att_dense + slot_index * tb.F;
            for (int b = lane; b < tb.F; b += 32) {
                const int i0 = s_ii[b];
                double val = 1.0;
                if (i0 >= 0) {
                    const double t = s_it[b];
                    for (int s = 0; s < plan.nseg; ++s) {
                        const double f0 = fac[s * tb.Fs_pad + i0];
                        val *= t != 0.0 ? (fac[s * tb.Fs_pad + i0 + 1] - f0) * t + f0 : f0;     // t == 0: no right neighbour needed (Fs == 1, np.interp ends)
                    }
                }
                dst[b] = val;
            }
        }
        __syncwarp();
    }
}

#ifndef ATT_MIN_BLOCKS
#define ATT_MIN_BLOCKS 6         // 80 registers (106 unconstrained); cfg4 + MB1 on the B200: 4 blocks 86.0, 5: 81.8, 6: 77.1, 8: 77.2 ms
#endif
template <bool GL3>
__global__ void __launch_bounds__(ATT_THREADS, ATT_MIN_BLOCKS)
K_att(IceParams ice, KInput in, AttTables tb, WorkList worklist, const unsigned long long *work_count, unsigned long long *ticket, int nseg_max,
      double *att_sparse, double *att_dense)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t bar;
    att_generic_body<GL3>(ice, tb, worklist, work_count, ticket, nseg_max, att_sparse, att_dense, smem_raw, bar);
}

// ---------------------------------------------------------------------------------------------------------------
// Small batches (the scalar API of the reference's production caller, simulation.py:173-210, is a batch of ONE pair): the binned
// pipeline costs 5-8 launches whose fixed cost dwarfs the arithmetic.  One cooperative launch instead: phase 1 traces a pair per
// thread (all modes, padded layout, work-list records), a grid-wide barrier, phase 2 integrates the attenuation with the generic
// warp-per-solution code.  Used for N <= NRMC_SMALL_PAIRS pairs in the padded layout.
// ---------------------------------------------------------------------------------------------------------------
#define NRMC_SMALL_PAIRS 2048
template <bool GL3>
__global__ void __launch_bounds__(ATT_THREADS)
K_small(IceParams ice, KInput in, TraceOutputs out, AttFill af, AttTables tb, WorkList worklist, unsigned long long *work_count, int nseg_max,
        double *att_sparse, double *att_dense)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t bar;
    const int S = 2 + 4 * ice.n_refl;
    for (int64_t p = (int64_t)blockIdx.x * ATT_THREADS + threadIdx.x; p < in.n_pairs; p += (int64_t)gridDim.x * ATT_THREADS) {
        double x1, y1, z1, x2, y2, z2;
        load_pair(in, p, x1, y1, z1, x2, y2, z2);
        const ShowerCut sc = load_cut(in, p);
        SolRec recs[2 + 4 * NRMC_MAX_REFLECTIONS];
        int n_recs = 0;
        uint32_t cut_mask = 0;
        const int n = trace_pair(ice, x1, y1, z1, x2, y2, z2, p, out, worklist.beta ? recs : nullptr, &sc, &n_recs, &cut_mask);
        for (int sl = 0; sl < S; ++sl) {              // attenuation rows of the empty slots and of the cut solutions: NaN
            if (sl < n && !((cut_mask >> sl) & 1u)) continue;
            if (af.sparse) for (int j = 0; j < af.Fs; ++j) af.sparse[(p * S + sl) * af.Fs + j] = NAN;
            if (af.dense) for (int j = 0; j < af.F; ++j) af.dense[(p * S + sl) * af.F + j] = NAN;
        }
        if (worklist.beta) for (int r = 0; r < n_recs; ++r) worklist_store(worklist, atomicAdd(work_count, 1ull), recs[r]);
    }
    if (!worklist.beta) return;
    __threadfence();
    cooperative_groups::this_grid().sync();
    att_generic_body<GL3>(ice, tb, worklist, work_count, nullptr, nseg_max, att_sparse, att_dense, smem_raw, bar);
}

// ---------------------------------------------------------------------------------------------------------------
// SP1 fast path (no bottom reflections).  1/L(z, f) = exp(b1(T) + p(T) ln f) with b1 and the slope p quadratics in the ice
// temperature T (attenuation.py:170-192) and T a monotone cubic in depth (:141-142).  For every integration frequency f_j the
// function g_j(tau) = 1/L(T(tau), f_j) of the normalised temperature tau in [-1, 1] (surface ... SP1_DEPTH_MAX) is entire and varies
// by a factor e^{+-1.3} at most (b1 is a parabola in T with its maximum at -20 C), so its Chebyshev series converges fast:
//     int ds / L(z, f_j) = sum_q w_q g_j(tau_q) = sum_k A_jk M_k,      M_k = sum_q w_q T_k(tau_q)        (w_q: ds-weights)
// with SP1_KT = 12 frequency-INDEPENDENT moments (sum of the dropped coefficients <= 2.2e-7 A_j0 for every f in 1 MHz ... 3 GHz, 7e-6
// with 10; the host measures it for the frequency set at hand and disables the fast path if it exceeds SP1_TRUNCATION).  The reference's per-(node, frequency)
// exponential disappears AND so does the per-node exp(b1): a node costs its geometry (one expm1, one rsqrt), a cubic for tau and
// a three-term recurrence.  (Round 1 expanded in the slope per band, 2 x 8 moments plus one exponential per node: 88 FP64
// instructions per node, now 60; tests/test_sp1_moment_form.py restates the mathematics on the CPU.)
// One THREAD per solution: the moments sit in registers (no reduction at all), then every integration frequency is a SP1_KT-term
// dot product with a table staged in shared memory by TMA.  Paths that reach below SP1_DEPTH_MAX go to the generic kernel
// (fall-back list).  The 1 m floor of attenuation.py:252-255 cannot be reached where the fast path is enabled: the host checks
// g_j < 1 on the whole temperature range.
// ---------------------------------------------------------------------------------------------------------------
#ifndef SP1_KT
#define SP1_KT 12
#endif
#define SP1_K SP1_KT          // row pitch of the coefficient table
#define SP1_DEPTH_MAX 2800.0
#define SP1_TRUNCATION 3e-7
#define SP1_NQ 12             // Gauss-Legendre nodes per panel: <= 6e-6 on the factor (16: 1e-8; 10: 9e-5) -- scratch/attconv.cpp
struct Sp1Tables {
    const double *wk;         // [Fs_pad][SP1_KT]  Chebyshev coefficients A_jk of g_j(tau)
    double tau[4];            // tau(a) = ((tau[0] a + tau[1]) a + tau[2]) a + tau[3], a = depth [m]
    double depth_max;
};
#define SP1_THREADS 128

// one quadrature node of the SP1 kernel: geometry (attenuation-independent) and the moment update
// 1 - exp(-y), 0 <= y < 700, with the power of two 2^(i/256) from the shared-memory table: exp(-y) = A (1 + q), A = 2^k 2^(i/256),
// q = expm1(r) = r + r^2/2 + r^3/6 on |r| <= ln2/512 (truncation r^3/24 <= 1e-10 relative to q: for y -> 0, where k = i = 0 and the
// result IS -q, that is its relative accuracy; elsewhere 1 - A >= 0.0027 is formed without cancellation).  8 FP64 instructions
// instead of the 14 of the reduction to |r| <= ln2/2 with a degree-8 polynomial.
__device__ __forceinline__ double one_minus_exp_neg_tab(double y, const double *s_exp2)
{
    const double t = fma(y, -369.3299304675746, 6755399441055744.0);          // -256 / ln 2
    const int n = __double2loint(t);                                            // <= 0
    const double r = fma(t - 6755399441055744.0, -0.0027076061740622863, -y);
    const double q = fma(r * r, fma(r, 0.16666666666666666, 0.5), r);
    const double T = s_exp2[n & 255];
    const double A = __hiloint2double(__double2hiint(T) + ((n >> 8) << 20), __double2loint(T));
    return fma(-A, q, 1.0 - A);
}

__device__ __forceinline__ void sp1_node(const IceParams &ice, const AttPlan &plan, const Sp1Tables &sp, const double *s_exp2, double u, double wscale,
                                         double (&M)[SP1_KT])
{
    const double uu = u * u;
    const double a = fabs(plan.zv - uu);                   // depth of the node (zv - uu <= 0 up to rounding at a virtual apex)
#ifndef SP1_NODE_POLY
    const double em = one_minus_exp_neg_tab(uu * ice.inv_z0, s_exp2);    // u^2 / z0 <= 3 km / z0
#else
    const double em = -expm1_c_small(-uu * ice.inv_z0);    // u^2 / z0 <= 3 km / z0: far inside the range of the reduction
#endif
    const double n = plan.beta + plan.delta * em;
    const double wds = wscale * u * n * rsqrt(plan.delta * em * (n + plan.beta));
    const double x = fma(fma(fma(sp.tau[0], a, sp.tau[1]), a, sp.tau[2]), a, sp.tau[3]), x2 = x + x;
    double t0 = wds, t1 = wds * x;
    M[0] += t0; M[1] += t1;
#pragma unroll
    for (int k = 2; k < SP1_KT; ++k) { const double t2 = fma(x2, t1, -t0); M[k] += t2; t0 = t1; t1 = t2; }
}

// Factors exp(-sum_k M_k A_jk) of the integration frequencies [j_begin, j_end) (at most SP1_SEG of them) for the 32
// solutions of a warp.  Each lane computes its own solution four frequencies at a time and parks the values in the warp's
// staging rows in shared memory; the warp then writes row after row with consecutive lanes on consecutive frequencies.
// (A lane storing its own row directly makes every 8-byte store a separate 32-byte sector write: measured 60 % of the
// kernel's time.)
#ifndef SP1_EW
#define SP1_EW 3             // frequencies per lane and emit step (independent chains)
#endif
#ifndef SP1_SEG
#define SP1_SEG 24
#endif
#ifndef SP1_EXP_SKIP
#define SP1_EXP_SKIP 3      // leading Taylor coefficients left out of the emit's exponential: degree 8 on |r| <= ln2/2,
                            // truncation r^9/9! <= 2e-10 relative on the factor (tolerance 1e-4); 0.8 ms per 1e8 pairs
#endif
#define SP1_ROW (SP1_SEG + 1)        // odd row pitch: conflict-free column writes
__device__ __forceinline__ void sp1_emit(const double (&M)[SP1_KT], const double *s_wk, const double *s_exp2, int j_begin, int j_end,
                                         double *stage, double *dst, unsigned lane)
{
    double *mine = stage + lane * SP1_ROW;
    for (int j0 = j_begin; j0 < j_end; j0 += SP1_EW) {
        double acc[SP1_EW];
        int jj[SP1_EW];
#pragma unroll
        for (int u = 0; u < SP1_EW; ++u) { jj[u] = min(j0 + u, j_end - 1); acc[u] = 0.0; }
#pragma unroll
        for (int k = 0; k + 1 < SP1_KT; k += 2) {
#pragma unroll
            for (int u = 0; u < SP1_EW; ++u) {
                const double2 w2 = *reinterpret_cast<const double2 *>(s_wk + jj[u] * SP1_K + k);
                acc[u] = fma(M[k], w2.x, acc[u]);
                acc[u] = fma(M[k + 1], w2.y, acc[u]);
            }
        }
        if (SP1_KT & 1) {
#pragma unroll
            for (int u = 0; u < SP1_EW; ++u) acc[u] = fma(M[SP1_KT - 1], s_wk[jj[u] * SP1_K + SP1_KT - 1], acc[u]);
        }
        // four exponentials as interleaved chains (values first, the shared-memory stores after: a store between them would
        // order the chains, the compiler cannot prove that the staging row does not alias the tables)
        double x[SP1_EW], r[SP1_EW], pv[SP1_EW];
        int kk[SP1_EW];
#ifndef SP1_EXP_POLY
        // exp(x) = 2^k 2^(i/256) e^r, |r| <= ln2/512: the power of two from a 256-entry table in shared memory, e^r = 1 + r + r^2/2
        // (truncation r^3/6 <= 4.1e-10 relative; the tolerance on the factor is 1e-4): 6 FP64 instructions instead of 13
#pragma unroll
        for (int u = 0; u < SP1_EW; ++u) {
            x[u] = acc[u] < 700.0 ? -acc[u] : -700.0;
            const double t = fma(x[u], 369.3299304675746, 6755399441055744.0);      // 256 / ln 2
            kk[u] = __double2loint(t);
            r[u] = fma(t - 6755399441055744.0, -0.0027076061740622863, x[u]);          // ln 2 / 256
        }
#pragma unroll
        for (int u = 0; u < SP1_EW; ++u) {
            const double p1 = s_exp2[kk[u] & 255] * fma(r[u], fma(r[u], 0.5, 1.0), 1.0);
            pv[u] = __hiloint2double(__double2hiint(p1) + ((kk[u] >> 8) << 20), __double2loint(p1));
        }
#else
#pragma unroll
        for (int u = 0; u < SP1_EW; ++u) { x[u] = fmax(-acc[u], -700.0); r[u] = exp_reduce(x[u], kk[u]); pv[u] = c_expc[SP1_EXP_SKIP]; }
#pragma unroll
        for (int i = 1 + SP1_EXP_SKIP; i < 10; ++i) {
#pragma unroll
            for (int u = 0; u < SP1_EW; ++u) pv[u] = fma(pv[u], r[u], c_expc[i]);
        }
#pragma unroll
        for (int u = 0; u < SP1_EW; ++u) {
            const double p1 = fma(pv[u] * r[u], r[u], r[u]) + 1.0;
            pv[u] = __hiloint2double(__double2hiint(p1) + (kk[u] << 20), __double2loint(p1));
        }
#endif
#pragma unroll
        for (int u = 0; u < SP1_EW; ++u) mine[jj[u] - j_begin] = pv[u];
    }
    __syncwarp();
    const int len = j_end - j_begin;
#pragma unroll 4
    for (int r = 0; r < 32; ++r) {
        double *row = reinterpret_cast<double *>(__shfl_sync(0xffffffffu, reinterpret_cast<unsigned long long>(dst), r));
        if (row != nullptr && (int)lane < len) __stcs(row + j_begin + lane, stage[r * SP1_ROW + lane]);
    }
    __syncwarp();
}

// Work distribution: the warps of the persistent grid draw their next 32 x SP1_CHUNK solutions from a ticket counter.  (A static
// stride gave every warp the same number of solutions, but the warp schedulers are not fair: ncu showed 20.5 of the 28 resident
// warps per SM active on average -- warps that ran ahead had finished while the others still had a quarter of their share to do.)
// Two-panel paths (the back of the work list, twice the nodes) are drawn first, so the tail of the launch consists of cheap ones.
#ifndef SP1_CHUNK
#define SP1_CHUNK 1
#endif

#ifdef SP1_MIN_BLOCKS
__global__ void __launch_bounds__(SP1_THREADS, SP1_MIN_BLOCKS)
#else
__global__ void __launch_bounds__(SP1_THREADS)
#endif
K_att_sp1(IceParams ice, KInput in, AttTables tb, Sp1Tables sp, WorkList worklist, const unsigned long long *work_count, int sparse_is_tmp,
          double *att_sparse, WorkList fallback, unsigned long long *fallback_count)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ double s_exp2[256];                          // 2^(i/256) for the emit's exponential
    double *s_wk = reinterpret_cast<double *>(smem_raw);
    double *stage = s_wk + tb.Fs_pad * SP1_K + (threadIdx.x >> 5) * (32 * SP1_ROW);
    for (int i = threadIdx.x; i < 256; i += SP1_THREADS) s_exp2[i] = exp2((double)i * (1.0 / 256.0));
    stage_tables(&bar, s_wk, sp.wk, (uint32_t)tb.Fs_pad * SP1_K * 8u, nullptr, nullptr, 0u, nullptr, nullptr, 0u, nullptr, nullptr, 0u);
    const unsigned lane = threadIdx.x & 31u;
    const unsigned long long n_front = work_count[0], n_work = n_front + work_count[WL_BACK];
    unsigned long long *ticket = const_cast<unsigned long long *>(work_count) + (CNT_TICKET_ATT - CNT_WORK);
    unsigned long long g = warp_ticket(ticket, 32ull * SP1_CHUNK, lane);
    while (g < n_work) {
        const unsigned long long g_next = warp_ticket(ticket, 32ull * SP1_CHUNK, lane);     // in flight during this chunk
#pragma unroll 1
        for (int sub = 0; sub < SP1_CHUNK; ++sub) {
            const unsigned long long idx = g + 32ull * sub + lane;
            const bool active = idx < n_work;
            const unsigned long long w = active ? n_work - 1ull - idx : 0ull;                // back of the list first
            double M[SP1_KT];
#pragma unroll
            for (int k = 0; k < SP1_KT; ++k) M[k] = 0.0;
            double *dst = nullptr;
            if (active) {
                SolRec rec = worklist_get(worklist, n_front, w);
                if (sparse_is_tmp) rec.row = (int64_t)worklist_index(worklist, n_front, w);   // scratch rows: work-list position
                AttPlan plan;
                att_plan_rec(ice, rec, plan);
                const bool ok = rec.z1 >= -sp.depth_max;        // the whole path lies inside the temperature range of the series
                // k = 0 paths: panel 0 = [u_T, u_2] twice (after the turning point), panel 1 = [u_2, u_1] once
#pragma unroll 1
                for (int panel = plan.turned ? 0 : 1; panel < 2 && ok; ++panel) {
                    const double lo = panel == 0 ? plan.uT : plan.u2, hi = panel == 0 ? plan.u2 : plan.u1;
                    if (!(hi > lo)) continue;
                    const double half = 0.5 * (hi - lo), mid = 0.5 * (hi + lo);
                    const double scale = (panel == 0 ? 4.0 : 2.0) * half;      // multiplicity x du/dx x the 2 of ds = 2 u n / ... du
#pragma unroll 1
                    for (int i = 0; i < SP1_NQ / 2; ++i) {                     // the symmetric node pair mid -+ half x_i: two independent chains
                        const double hx = half * c_glx12h[i], ws = scale * c_glw12h[i];
                        sp1_node(ice, plan, sp, s_exp2, mid - hx, ws, M);
                        sp1_node(ice, plan, sp, s_exp2, mid + hx, ws, M);
                    }
                }
                if (ok) dst = att_sparse + rec.row * (int64_t)tb.Fs;
                else worklist_store(fallback, atomicAdd(fallback_count, 1ull), rec);          // below the series' depth range: generic kernel
            }
#ifndef SP1_NO_PREFETCH
            {   // the record this lane reads in its next trip -> L2 (the ticket of the next chunk has arrived by now)
                const unsigned long long nidx = (sub + 1 < SP1_CHUNK ? g + 32ull * (sub + 1) : g_next) + lane;
                if (nidx < n_work) worklist_prefetch(worklist, worklist_index(worklist, n_front, n_work - 1ull - nidx));
            }
#endif
            for (int jb = 0; jb < tb.Fs; jb += SP1_SEG) sp1_emit(M, s_wk, s_exp2, jb, min(jb + SP1_SEG, tb.Fs), stage, dst, lane);
        }
        g = g_next;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// GL1 fast path (no bottom reflections).  1/L = 1 / max(A(z) - s_f, 1 m), A the 75 MHz length (a quintic in depth, attenuation.py:
// 99-128) and s_f = 0.55 m/MHz (f - 75 MHz) (:195-196): rational in the frequency, so there is no entire-function series as for
// SP1 -- but for a frequency whose pole A = s_f lies BELOW the values A takes along the path, 1/(A - s_f) is analytic on the
// path's A-range [A_lo, A_hi] and its Chebyshev series there is a closed form: with x = (A - A_mid) / A_half in [-1, 1] and the
// pole at -a, a = (A_mid - s_f) / A_half > 1,
//     1 / (A - s_f) = 1 / (A_half (a + x)) = (1 / A_half) (2 / sqrt(a^2 - 1)) sum'_k (-r)^k T_k(x),   r = a - sqrt(a^2 - 1) < 1,
// so   int ds / L = (2 / (A_half sqrt(a^2 - 1))) [M_0 / 2 + sum_{k >= 1} (-r)^k M_k],   M_k = sum_q w_q T_k(x_q):
// GL1_K frequency-independent moments per solution (one thread, 16-node panels, each leg in GL1_SUB sub-panels), then one
// Horner chain per frequency.  Per (solution, frequency):
//   easy       A_lo - s_f >= max(1 m, (GL1_XPMIN - 1) A_half): the series (r <= 0.47: r^16 = 6e-6 of the last term);
//   floor      s_f >= A_hi - 1 m: L = 1 m on the whole path, exponent = path length = M_0, exact;
//   invisible  the part of the path with A <= s_f + 1 m (L = 1 m) spans at least GL1_IVIS metres of depth (A_lo is reached inside
//              the path, |dA/dz| <= dadz_max), or path length / max(A_hi - s_f, 1) >= GL1_IVIS: factor < 2.1e-9, written as
//              exp(-lower bound) (the parity floor for factors below 1e-3 is 1e-7 absolute);
//   hard       the rest (1.3 % of the items on cfg3, 0.5 per solution): queued as (solution, frequency) ITEMS for K_gl1_item,
//              which integrates ONE frequency per thread on sub-panels graded towards the deep end of the path (where A falls and
//              the pole is approached); items whose pole comes within GL1_MARGIN of the path and whose factor is still visible
//              go on to K_gl1_fine (warp per item, 32 sub-panels per leg).
// Round 1 integrated every frequency directly (37 x 64 reciprocals per solution, 23.7 kFLOP) and handed every solution with a
// near-pole frequency to the generic warp-per-solution kernel, which redid ALL its hard frequencies: 5 % of the solutions cost as
// much as the other 95 %.  scratch/gl1_pole_emul.py: the scheme against the reference integrand at tight quadrature tolerance, on cfg3 and on wide random geometry
// (easy items: 3.7e-6 / 3.1e-5 worst relative deviation).
// ---------------------------------------------------------------------------------------------------------------
// quadrature node of the u-interval [lo, hi] for the thread-per-solution kernels: as att_node_geometry, with the bounded expm1
// and the reciprocal square root instead of sqrt + division (the IEEE versions cost two slow-path branches per node)
__device__ __forceinline__ void node_geometry_fast(const IceParams &ice, const AttPlan &p, double lo, double hi, double x, double w, double &z, double &wds)
{
    const double half = 0.5 * (hi - lo);
    const double u = fma(half, x, 0.5 * (hi + lo));
    const double uu = u * u;
    z = NRMC_MIN(p.zv - uu, 0.0);
    const double em = -expm1_c_small(-uu * ice.inv_z0);
    const double n = fma(p.delta, em, p.beta);
    wds = (w * half) * (2.0 * u) * n * rsqrt(p.delta * em * (n + p.beta));
}

struct Gl1Tables {
    double zx[4];            // depths in (z_min, 0) where dA/dz = 0: interior extrema of A along a path
    int32_t nzx, pad;
    double inv_dadz_max;     // 1 / max |dA/dz| on [z_min, 0]
    double z_min;            // below this depth A reaches its 100 m floor (attenuation.py:113,124-126): generic kernel
};
#define GL1_K 16
#define GL1_XPMIN 1.3
#define GL1_SUB 2
#define GL1_IVIS 20.0
#define GL1_SEG 13
#define GL1_ROW GL1_SEG             // odd row pitch: conflict-free column writes
#define GL1_MARGIN 10.0
#define GL1_FINE_SPP 32

__device__ __forceinline__ double gl1_A(double z)
{
    return (((( -3.63912864e-14 * z - 2.21040482e-10) * z - 3.50628312e-07) * z - 9.82378264e-05) * z + 6.87257150e-02) * z + 1.16052586e+03;
}

#ifndef GL1_MIN_BLOCKS
#define GL1_MIN_BLOCKS 4
#endif
__global__ void __launch_bounds__(SP1_THREADS, GL1_MIN_BLOCKS)
K_att_gl1(IceParams ice, AttTables tb, Gl1Tables gt, WorkList worklist, const unsigned long long *work_count, int sparse_is_tmp, double *att_sparse,
          WorkList fallback, unsigned long long *fallback_count, unsigned long long *items, unsigned long long *item_count,
          unsigned long long item_cap)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *s_fa = reinterpret_cast<double *>(smem_raw);
    double *stage = s_fa + tb.Fs_pad + (threadIdx.x >> 5) * (32 * GL1_ROW);
    for (int j = threadIdx.x; j < tb.Fs_pad; j += SP1_THREADS) s_fa[j] = tb.fa[j];
    __syncthreads();
    const unsigned lane = threadIdx.x & 31u;
    const unsigned long long n_front = work_count[0], n_work = n_front + work_count[WL_BACK];
    // work distribution as K_att_sp1: tickets, two-panel paths (the back of the list) first
    unsigned long long *ticket = const_cast<unsigned long long *>(work_count) + (CNT_TICKET_ATT - CNT_WORK);
    unsigned long long g_next = warp_ticket(ticket, 32ull, lane);
    while (g_next < n_work) {
        const unsigned long long idx = g_next + lane;
        g_next = warp_ticket(ticket, 32ull, lane);
        const unsigned long long w = idx < n_work ? n_work - 1ull - idx : n_work;
        double M[GL1_K];
#pragma unroll
        for (int k = 0; k < GL1_K; ++k) M[k] = 0.0;
        double *dst = nullptr;
        double A_lo = 1.0, A_hi = 2.0, A_mid = 1.5, A_half = 0.5, inv_half = 2.0, depth_span = 0.0;
        unsigned long long wi = 0;
        SolRec rec;
        bool active = false, to_generic = false;
        if (w < n_work) {
            wi = worklist_index(worklist, n_front, w);
            rec = worklist_load(worklist, wi);
            if (sparse_is_tmp) rec.row = (int64_t)wi;               // scratch rows: work-list position
            AttPlan plan;
            att_plan_rec(ice, rec, plan);
            active = rec.z1 >= gt.z_min;
            if (active) {
                // values A takes along the path: end points and the interior extrema of the quintic
                const double z_top = plan.turned ? NRMC_MIN(rec.zv, 0.0) : rec.z2;
                const double Aa = gl1_A(rec.z1), Ab = gl1_A(z_top);
                A_lo = NRMC_MIN(Aa, Ab); A_hi = NRMC_MAX(Aa, Ab);
#pragma unroll
                for (int e = 0; e < 4; ++e)
                    if (gt.zx[e] > rec.z1 && gt.zx[e] < z_top) { const double Ax = gl1_A(gt.zx[e]); A_lo = NRMC_MIN(A_lo, Ax); A_hi = NRMC_MAX(A_hi, Ax); }
                A_mid = 0.5 * (A_hi + A_lo);
                A_half = NRMC_MAX(0.5 * (A_hi - A_lo), 1e-3 * A_mid);
                inv_half = 1.0 / A_half;
                depth_span = z_top - rec.z1;
                // leg 0 = [u_T, u_2] (after the turning point, run through twice), leg 1 = [u_2, u_1]
#pragma unroll 1
                for (int leg = plan.turned ? 0 : 1; leg < 2; ++leg) {
                    const double lo = leg == 0 ? plan.uT : plan.u2, hi = leg == 0 ? plan.u2 : plan.u1;
                    if (!(hi > lo)) continue;
                    const double mult = leg == 0 ? 2.0 : 1.0, wsub = (hi - lo) * (1.0 / GL1_SUB);
#pragma unroll 1
                    for (int sp = 0; sp < GL1_SUB; ++sp) {
                        const double a0 = lo + sp * wsub, a1 = sp == GL1_SUB - 1 ? hi : lo + (sp + 1) * wsub;
#pragma unroll 1
                        for (int q = 0; q < NRMC_NQ / 2; ++q) {            // the symmetric node pair: two independent chains
                            double za, wa, zb, wb;
                            node_geometry_fast(ice, plan, a0, a1, c_glx[q], c_glw[q], za, wa);
                            node_geometry_fast(ice, plan, a0, a1, c_glx[NRMC_NQ - 1 - q], c_glw[NRMC_NQ - 1 - q], zb, wb);
                            const double xa = (gl1_A(za) - A_mid) * inv_half, xa2 = xa + xa, xb = (gl1_A(zb) - A_mid) * inv_half, xb2 = xb + xb;
                            double a_t0 = wa * mult, a_t1 = a_t0 * xa, b_t0 = wb * mult, b_t1 = b_t0 * xb;
                            M[0] += a_t0 + b_t0; M[1] += a_t1 + b_t1;
#pragma unroll
                            for (int k = 2; k < GL1_K; ++k) {
                                const double a_t2 = fma(xa2, a_t1, -a_t0), b_t2 = fma(xb2, b_t1, -b_t0);
                                M[k] += a_t2 + b_t2;
                                a_t0 = a_t1; a_t1 = a_t2; b_t0 = b_t1; b_t1 = b_t2;
                            }
                        }
                    }
                }
                dst = att_sparse + rec.row * (int64_t)tb.Fs;
            } else to_generic = true;
        }
        if (g_next + lane < n_work) worklist_prefetch(worklist, worklist_index(worklist, n_front, n_work - 1ull - (g_next + lane)));
        // factors, GL1_SEG frequencies at a time through the warp's staging rows (coalesced row stores, as K_att_sp1)
        for (int jb = 0; jb < tb.Fs; jb += GL1_SEG) {
            const int je = min(jb + GL1_SEG, tb.Fs);
            double *mine = stage + lane * GL1_ROW;
            for (int j = jb; j < je; ++j) {
                const double s = s_fa[j];
                double v = NAN;                                       // hard items: K_gl1_item / K_gl1_fine write the factor later
                if (active) {
                    const double d = A_lo - s;
                    if (d >= NRMC_MAX(1.0, (GL1_XPMIN - 1.0) * A_half)) {
                        const double a = (A_mid - s) * inv_half, a2m1 = fma(a, a, -1.0), isq = rsqrt(a2m1), mr = -NRMC_RCP(fma(a2m1, isq, a));
                        double acc = M[GL1_K - 1];
#pragma unroll
                        for (int k = GL1_K - 2; k >= 1; --k) acc = fma(acc, mr, M[k]);
                        acc = fma(acc, mr, 0.5 * M[0]);
                        v = exp_c_neg8(-acc * 2.0 * inv_half * isq);
                    } else if (s >= A_hi - 1.0) {
                        v = exp_c_neg8(-M[0]);
                    } else {
                        const double lb = NRMC_MAX(NRMC_MIN((s + 1.0 - A_lo) * gt.inv_dadz_max, depth_span), M[0] / NRMC_MAX(A_hi - s, 1.0));
                        if (lb >= GL1_IVIS) v = exp_c_neg8(-lb);
                        else if (!to_generic) {
                            const unsigned long long idx = atomicAdd(item_count, 1ull);
                            if (idx < item_cap) items[idx] = (wi << 16) | (unsigned long long)j;
                            else to_generic = true;                   // queue full: the generic kernel redoes the whole row
                        }
                    }
                }
                mine[j - jb] = v;
            }
            __syncwarp();
            const int len = je - jb;
#pragma unroll 4
            for (int r = 0; r < 32; ++r) {
                double *row = reinterpret_cast<double *>(__shfl_sync(FULL_MASK, reinterpret_cast<unsigned long long>(dst), r));
                if (row != nullptr && (int)lane < len) __stcs(row + jb + lane, stage[r * GL1_ROW + lane]);
            }
            __syncwarp();
        }
        if (w < n_work && to_generic) worklist_store(fallback, atomicAdd(fallback_count, 1ull), rec);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Separable models (MB1, GL2: L(z, f) = fa(f) p0(z), attenuation.py:198-204, :229-244), any number of bottom reflections, sparse
// output.  Away from the 1 m floor of :252-255 the exponent of every frequency and every path segment is ONE depth integral per
// panel, G_p = int ds / p0, over fa(f); where fa p0 <= 1 on the whole path (GL2 above 1.58 GHz: negative length, floored) it is the
// path length S_p = int ds.  The product of the segments' factors is exp(-sum_p m_p H_p) with m_p the number of times the path runs
// through panel p (plan_total_mult) -- the per-segment factors themselves are only needed for a DENSE output of a bottom-reflected
// path (np.interp per segment, py:1077-1086): that case stays with the generic kernel.  One THREAD per solution: the panels of the
// plan as 16-node slots exactly as the generic warp-per-solution kernel cuts them (same nodes, same floor decisions), no shuffle
// reductions, no shared-memory round trips per solution; a solution with a frequency whose floor is crossed ON the path goes to
// the generic kernel through the fall-back list.  cfg4 + MB1 (3.4e7 solutions): 77 ms in the generic kernel.
// ---------------------------------------------------------------------------------------------------------------
#ifndef SEP_MIN_BLOCKS
#define SEP_MIN_BLOCKS 8        // 64 registers; cfg4 + MB1 on the B200: 4 blocks 20.5, 5: 19.4, 6: 18.9, 8: 18.6 ms (generic kernel: 77.0)
#endif
// depth factor p0(z) of the separable models as att_node, MB1 in power form: 1250 * 0.08886 * exp(-0.048827 * (225.6746 - 86.517596 *
// log10(x))) / 231.21 = A x^kappa, x = 848.870 - d (attenuation.py:239-244), with the solver's log and the table-free exponential
// instead of the library's exp and log10 (two slow-path branches per node); differs from att_node by 1e-15 relative
__device__ __forceinline__ double sep_p0(int model, double z)
{
    if (model == NRMC_ATT_MB1) {
        const double x = NRMC_MAX(fma(z, 420.0 / 576.0, 848.870), 1e-3);   // 848.870 - d, d = -z 420 / 576 (> 428 above the 576 m shelf bottom)
        return 7.872503543690096e-06 * exp_c_bounded(1.8346312901726596 * NRMC_LOG(x));
    }
    return ((((-4.58987344e-17 * z - 2.89124473e-13) * z - 5.16435542e-10) * z - 2.58901767e-07) * z + 1.58815679e-05) * z + 1.20547286e+00;   // GL2
}
__global__ void __launch_bounds__(SP1_THREADS, SEP_MIN_BLOCKS)
K_att_sep(IceParams ice, AttTables tb, WorkList worklist, const unsigned long long *work_count, int sparse_is_tmp, double *att_sparse,
          WorkList fallback, unsigned long long *fallback_count)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *s_fa = reinterpret_cast<double *>(smem_raw);
    double *stage = s_fa + tb.Fs_pad + (threadIdx.x >> 5) * (32 * GL1_ROW);
    for (int j = threadIdx.x; j < tb.Fs_pad; j += SP1_THREADS) s_fa[j] = tb.fa[j];
    __syncthreads();
    const unsigned lane = threadIdx.x & 31u;
    const unsigned long long n_front = work_count[0], n_work = n_front + work_count[WL_BACK];
    unsigned long long *ticket = const_cast<unsigned long long *>(work_count) + (CNT_TICKET_ATT - CNT_WORK);
    unsigned long long g_next = warp_ticket(ticket, 32ull, lane);
    while (g_next < n_work) {
        const unsigned long long idx = g_next + lane;
        g_next = warp_ticket(ticket, 32ull, lane);
        const unsigned long long w = idx < n_work ? n_work - 1ull - idx : n_work;
        double GG = 0.0, SS = 0.0, pmin = INFINITY, pmax = -INFINITY;
        double *dst = nullptr;
        SolRec rec;
        bool active = false, to_generic = false;
        if (w < n_work) {
            const unsigned long long wi = worklist_index(worklist, n_front, w);
            rec = worklist_load(worklist, wi);
            if (sparse_is_tmp) rec.row = (int64_t)wi;               // scratch rows: work-list position
            AttPlan plan;
            att_plan_rec(ice, rec, plan);
            active = true;
            double G0 = 0, G1 = 0, G2 = 0, S0 = 0, S1 = 0, S2 = 0;
#pragma unroll 1
            for (int slot = 0; slot < plan.n_slots; ++slot) {
                double lo, hi;
                int panel;
                plan_slot(plan, slot, lo, hi, panel);
                double ga = 0.0, gb = 0.0, sa = 0.0, sb = 0.0;
#pragma unroll 1
                for (int q = 0; q < NRMC_NQ / 2; ++q) {          // two nodes per trip: independent chains
                    double za, wa, zb, wb;
                    node_geometry_fast(ice, plan, lo, hi, c_glx[q], c_glw[q], za, wa);
                    node_geometry_fast(ice, plan, lo, hi, c_glx[NRMC_NQ - 1 - q], c_glw[NRMC_NQ - 1 - q], zb, wb);
#ifdef SEP_STOCK_NODE
                    AttNode na, nb;
                    att_node(ice.att_model, za, tb.gl3, na);
                    att_node(ice.att_model, zb, tb.gl3, nb);
                    const double pa = na.p0, pb = nb.p0;
#else
                    const double pa = sep_p0(ice.att_model, za), pb = sep_p0(ice.att_model, zb);
#endif
                    ga = fma(wa, NRMC_RCP(pa), ga); gb = fma(wb, NRMC_RCP(pb), gb);
                    sa += wa; sb += wb;
                    pmin = NRMC_MIN(pmin, NRMC_MIN(pa, pb)); pmax = NRMC_MAX(pmax, NRMC_MAX(pa, pb));
                }
                ga += gb; sa += sb;
                if (panel == 0) { G0 += ga; S0 += sa; } else if (panel == 1) { G1 += ga; S1 += sa; } else { G2 += ga; S2 += sa; }
            }
            const double m0 = (double)plan_total_mult(plan, 0), m1 = (double)plan_total_mult(plan, 1), m2 = (double)plan_total_mult(plan, 2);
            GG = m0 * G0 + m1 * G1 + m2 * G2;
            SS = m0 * S0 + m1 * S1 + m2 * S2;
            dst = att_sparse + rec.row * (int64_t)tb.Fs;
        }
        if (g_next + lane < n_work) worklist_prefetch(worklist, worklist_index(worklist, n_front, n_work - 1ull - (g_next + lane)));
        // factors, GL1_SEG frequencies at a time through the warp's staging rows (coalesced row stores, as K_att_sp1)
        for (int jb = 0; jb < tb.Fs; jb += GL1_SEG) {
            const int je = min(jb + GL1_SEG, tb.Fs);
            double *mine = stage + lane * GL1_ROW;
            for (int j = jb; j < je; ++j) {
                const double fa = s_fa[j];
                double v = NAN;                                       // floor crossed on the path: the generic kernel redoes the row
                if (active) {
                    const double a = fa * pmin, b = fa * pmax;
                    if (a >= 1.0 && b >= 1.0) v = exp_c_neg8(-GG * NRMC_RCP(fa));
                    else if (a <= 1.0 && b <= 1.0) v = exp_c_neg8(-SS);
                    else to_generic = true;
                }
                mine[j - jb] = v;
            }
            __syncwarp();
            const int len = je - jb;
#pragma unroll 4
            for (int r = 0; r < 32; ++r) {
                double *row = reinterpret_cast<double *>(__shfl_sync(FULL_MASK, reinterpret_cast<unsigned long long>(dst), r));
                if (row != nullptr && (int)lane < len) __stcs(row + jb + lane, stage[r * GL1_ROW + lane]);
            }
            __syncwarp();
        }
        if (active && to_generic) worklist_store(fallback, atomicAdd(fallback_count, 1ull), rec);
    }
}

// one hard (solution, frequency) item per thread: direct quadrature of ds / max(A - s_f, 1) on sub-panels graded towards the deep
// end of the path (widths w, 0.4 w, 0.16 w on the up-going leg; the leg after the turning point is one panel), 16 nodes each
__global__ void __launch_bounds__(128)
K_gl1_item(IceParams ice, AttTables tb, WorkList worklist, int sparse_is_tmp, double *att_sparse, const unsigned long long *items,
           const unsigned long long *item_count, unsigned long long item_cap, unsigned long long *fine, unsigned long long *fine_count)
{
    const unsigned long long n = min(*item_count, item_cap);
    const unsigned lane = threadIdx.x & 31u;
    unsigned long long *ticket = const_cast<unsigned long long *>(item_count) + (CNT_TICKET_ITEM - CNT_ITEMS);
    unsigned long long i_next = warp_ticket(ticket, 32ull, lane);
    while (i_next < n) {
        const unsigned long long i = i_next + lane;
        i_next = warp_ticket(ticket, 32ull, lane);
        if (i >= n) continue;       // (the lane rejoins the warp at the ticket shuffle)
        const unsigned long long it = items[i], wi = it >> 16;
        const int j = (int)(it & 0xffffull);
        const SolRec rec = worklist_load(worklist, wi);
        const int64_t row = sparse_is_tmp ? (int64_t)wi : rec.row;
        AttPlan plan;
        att_plan_rec(ice, rec, plan);
        const double s = __ldg(tb.fa + j);
        const bool leg0 = plan.turned && plan.u2 > plan.uT;
        const int n_slots = (leg0 ? 1 : 0) + (plan.u1 > plan.u2 ? 3 : 0);
        double acc = 0.0, acc2 = 0.0;
#pragma unroll 1
        for (int sl = 0; sl < n_slots; ++sl) {
            double lo, hi, mult;
            if (leg0 && sl == 0) { lo = plan.uT; hi = plan.u2; mult = 2.0; }
            else {
                const int k = sl - (leg0 ? 1 : 0);
                const double w0u = (plan.u1 - plan.u2) * (1.0 / 1.56);
                lo = plan.u2 + w0u * (k == 0 ? 0.0 : (k == 1 ? 1.0 : 1.4));
                hi = k == 2 ? plan.u1 : plan.u2 + w0u * (k == 0 ? 1.0 : 1.4);
                mult = 1.0;
            }
#pragma unroll 1
            for (int q = 0; q < NRMC_NQ / 2; ++q) {          // two nodes per trip: independent chains
                double za, wa, zb, wb;
                node_geometry_fast(ice, plan, lo, hi, c_glx[q], c_glw[q], za, wa);
                node_geometry_fast(ice, plan, lo, hi, c_glx[NRMC_NQ - 1 - q], c_glw[NRMC_NQ - 1 - q], zb, wb);
                const double Aa = NRMC_MAX(gl1_A(za), 100.0), Ab = NRMC_MAX(gl1_A(zb), 100.0);
                acc = fma(wa * mult, NRMC_RCP(NRMC_MAX(Aa - s, 1.0)), acc);        // attenuation.py:196, :252-255
                acc2 = fma(wb * mult, NRMC_RCP(NRMC_MAX(Ab - s, 1.0)), acc2);
            }
        }
        const double z_top = plan.turned ? NRMC_MIN(rec.zv, 0.0) : rec.z2;
        const double a_min = NRMC_MIN(NRMC_MAX(gl1_A(rec.z1), 100.0), NRMC_MAX(gl1_A(z_top), 100.0));
        acc += acc2;
        if (s > a_min - GL1_MARGIN && acc < 30.0) fine[atomicAdd(fine_count, 1ull)] = it;      // near the pole and still visible
        else att_sparse[row * (int64_t)tb.Fs + j] = exp_c_neg(-acc);
    }
}

// one near-pole item per WARP.  The pole of 1 / (A - s_f) is approached where A is smallest along the path -- its deep end, or its top
// for shallow paths (A has a minimum near -500 m) -- and is cut off by the 1 m floor, so the leg that touches that end is divided
// into GL1_FINE_GRADED sub-panels shrinking by GL1_FINE_RATIO towards it (the innermost ones are centimetres wide: below the scale
// 1 m / |dA/dz| on which the floored integrand varies), the other leg into GL1_FINE_UNIFORM equal ones; one sub-panel of 16 nodes
// per lane, one pass.  (First version: 32 - 64 uniform sub-panels per leg, the generic kernel's fine phase: 2 - 4 passes per lane.)
#define GL1_FINE_GRADED 22
#define GL1_FINE_UNIFORM 10
#define GL1_FINE_RATIO 0.6
__global__ void __launch_bounds__(128)
K_gl1_fine(IceParams ice, AttTables tb, WorkList worklist, int sparse_is_tmp, double *att_sparse, const unsigned long long *fine,
           const unsigned long long *fine_count)
{
    const unsigned lane = threadIdx.x & 31u;
    const unsigned long long n = *fine_count;
    // lane -> position of its sub-panel on the unit interval, graded towards 1:  1 - rho^k ... 1 - rho^(k+1), the last one reaching 1
    const int kg = (int)lane;
    const double g_lo = 1.0 - pow(GL1_FINE_RATIO, (double)kg), g_hi = kg == GL1_FINE_GRADED - 1 ? 1.0 : 1.0 - pow(GL1_FINE_RATIO, (double)(kg + 1));
    unsigned long long *ticket = const_cast<unsigned long long *>(fine_count) + (CNT_TICKET_FINE - CNT_FINE);
    unsigned long long i_next = warp_ticket(ticket, 1ull, lane);                // one item per warp and ticket
    while (i_next < n) {
        const unsigned long long i = i_next;
        i_next = warp_ticket(ticket, 1ull, lane);
        const unsigned long long it = fine[i], wi = it >> 16;
        const int j = (int)(it & 0xffffull);
        const SolRec rec = worklist_load(worklist, wi);
        const int64_t row = sparse_is_tmp ? (int64_t)wi : rec.row;
        AttPlan plan;
        att_plan_rec(ice, rec, plan);
        const double s = __ldg(tb.fa + j);
        const bool leg0 = plan.turned && plan.u2 > plan.uT, leg1 = plan.u1 > plan.u2;
        const double z_top = plan.turned ? NRMC_MIN(rec.zv, 0.0) : rec.z2;
        const bool deep_end = gl1_A(rec.z1) <= gl1_A(z_top);          // where A is smallest: u_1 (deep end of leg 1) or the top of the path
        // graded leg and the end it is graded towards; the top of the path is u_T (start of leg 0) for turned rays, u_2 (start of leg 1) else
        const int gleg = deep_end ? 1 : (leg0 ? 0 : 1);
        double a0 = 0.0, a1 = 0.0, mult = 0.0;
        if ((int)lane < GL1_FINE_GRADED) {
            if (gleg == 1 ? leg1 : leg0) {
                const double lo = gleg == 0 ? plan.uT : plan.u2, hi = gleg == 0 ? plan.u2 : plan.u1, wdt = hi - lo;
                if (deep_end) { a0 = lo + wdt * g_lo; a1 = lo + wdt * g_hi; }      // towards hi
                else { a0 = hi - wdt * g_hi; a1 = hi - wdt * g_lo; }               // towards lo
                mult = gleg == 0 ? 2.0 : 1.0;
            }
        } else {
            const int oleg = 1 - gleg, ku = (int)lane - GL1_FINE_GRADED;
            if (oleg == 1 ? leg1 : leg0) {
                const double lo = oleg == 0 ? plan.uT : plan.u2, hi = oleg == 0 ? plan.u2 : plan.u1, wsub = (hi - lo) * (1.0 / GL1_FINE_UNIFORM);
                a0 = lo + ku * wsub; a1 = ku == GL1_FINE_UNIFORM - 1 ? hi : lo + (ku + 1) * wsub;
                mult = oleg == 0 ? 2.0 : 1.0;
            }
        }
        double acc = 0.0, acc2 = 0.0;
        if (mult > 0.0 && a1 > a0) {
#pragma unroll 1
            for (int q = 0; q < NRMC_NQ / 2; ++q) {
                double za, wa, zb, wb;
                node_geometry_fast(ice, plan, a0, a1, c_glx[q], c_glw[q], za, wa);
                node_geometry_fast(ice, plan, a0, a1, c_glx[NRMC_NQ - 1 - q], c_glw[NRMC_NQ - 1 - q], zb, wb);
                const double Aa = NRMC_MAX(gl1_A(za), 100.0), Ab = NRMC_MAX(gl1_A(zb), 100.0);
                acc = fma(wa * mult, NRMC_RCP(NRMC_MAX(Aa - s, 1.0)), acc);
                acc2 = fma(wb * mult, NRMC_RCP(NRMC_MAX(Ab - s, 1.0)), acc2);
            }
        }
        acc += acc2;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) acc += __shfl_xor_sync(FULL_MASK, acc, d);
        if (lane == 0) att_sparse[row * (int64_t)tb.Fs + j] = exp_c_neg(-acc);
    }
}

#define EXPAND_F_SMEM 2048          // output bins whose interpolation tables fit the static shared memory of K_att_expand
__global__ void __launch_bounds__(256)
K_att_expand(AttTables tb, WorkList worklist, const unsigned long long *work_count, int sparse_is_tmp, const double *att_sparse,
             double *att_dense)
{
    // HBM-write bound (4 KB per row at F = 512): the interpolation tables sit in shared memory and every lane has four bins in flight
    __shared__ double s_it[EXPAND_F_SMEM];
    __shared__ int32_t s_ii[EXPAND_F_SMEM];
    const bool in_smem = tb.F <= EXPAND_F_SMEM;
    if (in_smem) for (int b = threadIdx.x; b < tb.F; b += blockDim.x) { s_it[b] = tb.it[b]; s_ii[b] = tb.ii[b]; }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    // (A/B on the B200, 105 GB of rows per cfg3 step: 4 bins in flight per lane 21.4 ms = 4.9 TB/s; 8 or 16 in flight, write-through /
    //  streaming stores, 16-byte stores: 28 - 31 ms; walking the rows in output order through an inverse map: 23.3 ms; rows staged in
    //  shared memory and handed to the copy engine as 4 KB TMA bulk stores (cp.async.bulk.global.shared::cta, 2 - 4 buffers per warp,
    //  4 - 8 warps per block): 43.3 ms for every variant -- one bulk store per 4 KB does not keep the engine busy;
    //  profiles/r2_ab_runs.log)
    const unsigned long long n_front = work_count[0], n_work = n_front + work_count[WL_BACK];
    for (unsigned long long w = (unsigned long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); w < n_work;
         w += (unsigned long long)gridDim.x * (blockDim.x >> 5)) {
        const unsigned long long wi = worklist_index(worklist, n_front, w);
        const int64_t row = worklist.row[wi];
        const double *src = att_sparse + (sparse_is_tmp ? (int64_t)wi : row) * tb.Fs;
        double *dst = att_dense + row * tb.F;
        auto bin = [&](int b) {
            const int i0 = in_smem ? s_ii[b] : __ldg(tb.ii + b);
            if (i0 < 0) return 1.0;
            const double f0 = src[i0], t = in_smem ? s_it[b] : __ldg(tb.it + b);
            return t != 0.0 ? (src[i0 + 1] - f0) * t + f0 : f0;      // t == 0: no right neighbour needed (Fs == 1, np.interp ends)
        };
        int b = lane;
#ifndef EXPAND_UNROLL
#define EXPAND_UNROLL 4
#endif
        for (; b + 32 * (EXPAND_UNROLL - 1) < tb.F; b += 32 * EXPAND_UNROLL) {
            double v[EXPAND_UNROLL];
#pragma unroll
            for (int u = 0; u < EXPAND_UNROLL; ++u) v[u] = bin(b + 32 * u);
#pragma unroll
            for (int u = 0; u < EXPAND_UNROLL; ++u) dst[b + 32 * u] = v[u];
        }
        for (; b < tb.F; b += 32) dst[b] = bin(b);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// compact (per-solution, CSR) output for host calls: exclusive scan of n_sol, then a gather of the existing rows
// ---------------------------------------------------------------------------------------------------------------
#define PACK_PAIRS 1024
#define PACK_THREADS 256
struct PackArrays { const unsigned char *src[N_OUT - 2]; unsigned char *dst[N_OUT - 2]; int32_t row_bytes[N_OUT - 2]; int32_t n; };

__global__ void __launch_bounds__(PACK_THREADS)
K_pack_count(const int32_t *n_sol, int64_t n_pairs, unsigned long long *block_sums)
{
    __shared__ int s_part[PACK_THREADS / 32];
    const int64_t p0 = (int64_t)blockIdx.x * PACK_PAIRS + 4 * threadIdx.x;
    int v = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) if (p0 + j < n_pairs) v += n_sol[p0 + j];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int w = 0; w < PACK_THREADS / 32; ++w) t += s_part[w];
        block_sums[blockIdx.x] = (unsigned long long)t;
    }
}

// single block: exclusive scan of the block sums in place; total -> block_sums[n_blocks]
__global__ void __launch_bounds__(1024)
K_pack_scan(unsigned long long *block_sums, int n_blocks, unsigned long long *base_dev)
{
    __shared__ unsigned long long s_warp[32];
    __shared__ unsigned long long s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int base = 0; base < n_blocks; base += 1024) {
        const int i = base + threadIdx.x;
        const unsigned long long v = i < n_blocks ? block_sums[i] : 0ull;
        unsigned long long incl = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const unsigned long long t = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += t; }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            unsigned long long w = s_warp[lane], wi = w;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const unsigned long long t = __shfl_up_sync(0xffffffffu, wi, d); if (lane >= d) wi += t; }
            s_warp[lane] = wi - w;     // exclusive over warps
        }
        __syncthreads();
        const unsigned long long carry = s_carry;
        if (i < n_blocks) block_sums[i] = carry + s_warp[warp] + incl - v;
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = carry + s_warp[warp] + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        block_sums[n_blocks] = s_carry;                       // rows of this chunk
        if (base_dev) {                                       // running row base of a multi-chunk device call
            const unsigned long long old = *base_dev;
            block_sums[n_blocks + 1] = old;
            *base_dev = old + s_carry;
        }
    }
}

__global__ void __launch_bounds__(PACK_THREADS)
K_pack(const int32_t *n_sol, int64_t n_pairs, int S, const unsigned long long *block_off, int64_t row_base, int base_from_scan,
       int64_t *sol_offset, PackArrays pa)
{
    if (base_from_scan) row_base = (int64_t)block_off[gridDim.x + 1];     // K_pack_scan left the device-side row base there
    __shared__ int s_off[PACK_PAIRS];
    __shared__ int s_n[PACK_PAIRS];
    __shared__ int s_part[PACK_THREADS / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t pb = (int64_t)blockIdx.x * PACK_PAIRS;
    int nloc[4], v = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) { const int64_t p = pb + 4 * threadIdx.x + j; nloc[j] = p < n_pairs ? n_sol[p] : 0; v += nloc[j]; }
    int incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += t; }
    if (lane == 31) s_part[warp] = incl;
    __syncthreads();
    int wbase = 0;
    for (int w = 0; w < warp; ++w) wbase += s_part[w];
    int run = wbase + incl - v;
#pragma unroll
    for (int j = 0; j < 4; ++j) { s_off[4 * threadIdx.x + j] = run; s_n[4 * threadIdx.x + j] = nloc[j]; run += nloc[j]; }
    __syncthreads();
    const int64_t row0 = row_base + (int64_t)block_off[blockIdx.x];
    for (int i = threadIdx.x; i < PACK_PAIRS && pb + i < n_pairs; i += PACK_THREADS) sol_offset[pb + i] = row0 + s_off[i];
    for (int i = warp; i < PACK_PAIRS && pb + i < n_pairs; i += PACK_THREADS / 32) {
        const int n = s_n[i];
        for (int sl = 0; sl < n; ++sl) {
            const int64_t src_row = (pb + i) * S + sl, dst_row = (int64_t)block_off[blockIdx.x] + s_off[i] + sl;   // chunk-local rows
            for (int a = 0; a < pa.n; ++a) {
                const int rb = pa.row_bytes[a];
                if ((rb & 7) == 0) {
                    const double *src = reinterpret_cast<const double *>(pa.src[a]) + src_row * (rb >> 3);
                    double *dst = reinterpret_cast<double *>(pa.dst[a]) + dst_row * (rb >> 3);
                    for (int j = lane; j < (rb >> 3); j += 32) dst[j] = src[j];
                } else {
                    for (int j = lane; j < rb; j += 32) pa.dst[a][dst_row * rb + j] = pa.src[a][src_row * rb + j];
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// propagation effects on spectra (analyticraytracing.py:2937-3033, in-ice branch): one block per solution row
// ---------------------------------------------------------------------------------------------------------------
struct cplx { double re, im; };
__device__ __forceinline__ cplx cmul(cplx a, cplx b) { return {a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re}; }
__device__ __forceinline__ cplx cdiv(cplx a, cplx b)
{
    const double d = b.re * b.re + b.im * b.im;
    return {(a.re * b.re + a.im * b.im) / d, (a.im * b.re - a.re * b.im) / d};
}

// Fresnel reflection coefficients for an interface n_1 -> n_2 at incidence angle a (geometryUtilities.py:211-263):
// r_p = conj((n^2 cos a - sqrt(n^2 - sin^2 a)) / (n^2 cos a + sqrt(...))), r_s = conj((cos a - sqrt(...)) / (cos a + sqrt(...))),
// n = n_2 / n_1, sqrt on the complex branch (scimath) beyond total internal reflection.
__device__ __forceinline__ void fresnel_r(double a, double n, cplx &rp, cplx &rs)
{
    const double ca = cos(a), sa = sin(a);
    const double rad = n * n - sa * sa;
    cplx sq = rad >= 0.0 ? cplx{sqrt(rad), 0.0} : cplx{0.0, sqrt(-rad)};
    cplx p = cdiv(cplx{n * n * ca - sq.re, -sq.im}, cplx{n * n * ca + sq.re, sq.im});
    cplx q = cdiv(cplx{ca - sq.re, -sq.im}, cplx{ca + sq.re, sq.im});
    rp = {p.re, -p.im};
    rs = {q.re, -q.im};
}

#define FX_THREADS 128
__global__ void __launch_bounds__(FX_THREADS)
K_apply_effects(nrmc_rt_effects fx, AttTables tb, int K1, double n_surface)
{
    __shared__ cplx s_c[2];      // total factor on eTheta and ePhi
    for (int64_t row = blockIdx.x; row < fx.n_rows; row += gridDim.x) {
        if (threadIdx.x == 0) {
            cplx ct = {1.0, 0.0}, cp = {1.0, 0.0}, last_p = {1.0, 0.0}, last_s = {1.0, 0.0};
            if (fx.reflection_angle) {
                for (int s = 0; s < K1; ++s) {
                    const double a = fx.reflection_angle[row * K1 + s];
                    if (a == a) {
                        cplx rp, rs;
                        fresnel_r(a, 1.0 / n_surface, rp, rs);
                        ct = cmul(ct, rp); cp = cmul(cp, rs);
                        last_p = rp; last_s = rs;
                    }
                }
            }
            // the field object keeps the coefficients of the last surface reflection (py:2993-2994 overwrite them per segment)
            if (fx.r_theta) { fx.r_theta[2 * row] = last_p.re; fx.r_theta[2 * row + 1] = last_p.im; }
            if (fx.r_phi) { fx.r_phi[2 * row] = last_s.re; fx.r_phi[2 * row + 1] = last_s.im; }
            const int k = fx.reflection ? fx.reflection[row] : 0;
            if (k > 0) {   // analyticraytracing.py:3002-3010
                const double amp = pow(fx.reflection_coefficient, (double)k);
                const double ph = fmod(k * fx.reflection_phase_shift, 2.0 * 3.14159265358979323846);
                const cplx b = {amp * cos(ph), amp * sin(ph)};
                ct = cmul(ct, b); cp = cmul(cp, b);
            }
            if (fx.focusing) { const double fc = fx.focusing[row]; ct.re *= fc; ct.im *= fc; cp.re *= fc; cp.im *= fc; }   // py:3012-3015
            s_c[0] = ct; s_c[1] = cp;
        }
        __syncthreads();
        const cplx ct = s_c[0], cp = s_c[1];
        cplx *spec = reinterpret_cast<cplx *>(fx.spectrum) + row * 3 * (int64_t)fx.n_freq;
        for (int b = threadIdx.x; b < fx.n_freq; b += FX_THREADS) {
            double att = 1.0;
            if (fx.attenuation) att = fx.attenuation[row * fx.n_freq + b];
            else if (fx.attenuation_sparse) {
                const int i0 = __ldg(tb.ii + b);
                if (i0 >= 0) {
                    const double *src = fx.attenuation_sparse + row * tb.Fs;
                    const double f0 = src[i0], t = __ldg(tb.it + b);
                    att = t != 0.0 ? (src[i0 + 1] - f0) * t + f0 : f0;      // a single integration frequency has no right neighbour
                }
            }
            cplx e0 = spec[b], e1 = spec[fx.n_freq + b], e2 = spec[2 * fx.n_freq + b];
            e0.re *= att; e0.im *= att;
            e1 = cmul(cplx{e1.re * att, e1.im * att}, ct);
            e2 = cmul(cplx{e2.re * att, e2.im * att}, cp);
            spec[b] = e0; spec[fx.n_freq + b] = e1; spec[2 * fx.n_freq + b] = e2;
        }
        __syncthreads();
    }
}

// Focusing factor of every solution of a trace result (py:2778-2888): thread per pair, exact d(launch angle)/d(receiver depth)
__global__ void __launch_bounds__(128)
K_focusing(IceParams ice, KInput in, nrmc_rt_focusing fo, int S)
{
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= in.n_pairs) return;
    const int n = fo.n_sol[p];
    const int64_t row0 = fo.sol_offset ? fo.sol_offset[p] : p * S;
    if (n > 0) {
        double x1, y1, z1, x2, y2, z2;
        load_pair(in, p, x1, y1, z1, x2, y2, z2);
        Frame2D f;
        make_frame(x1, y1, z1, x2, y2, z2, f);
        PairGeom g;
        make_pair_geom(ice, f.z1, f.z2, fmax(f.rho, 1e-12), g);
        for (int s = 0; s < n && s < S; ++s) {
            const int64_t q = row0 + s;
            const int k = fo.reflection ? fo.reflection[q] : 0, rcase = fo.reflection_case ? fo.reflection_case[q] : 1;
            fo.focusing[q] = focusing_factor(ice, g, f.swap, k, rcase > 0 ? rcase : 1, 1.0 / fo.C0[q], fo.path_length[q], fo.limit);
        }
    }
    if (!fo.sol_offset) for (int s = n < 0 ? 0 : n; s < S; ++s) fo.focusing[row0 + s] = NAN;
}

__global__ void K_att_length(IceParams ice, Gl3Table gl3, const double *z, const double *f, int64_t n, double *out)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    AttNode nd;
    att_node(ice.att_model, z[i], gl3, nd);
    double fa, fb;
    att_freq_consts(ice.att_model, f[i], fa, fb);
    const double inv = att_inv_length(ice.att_model, nd, fa, fb);
    out[i] = (z[i] > 0.0) ? INFINITY : 1.0 / inv;   // attenuation.py:256-257
}

// FP64 roofline denominator: 8 independent FMA chains per thread
__global__ void K_fp64_peak(double *out, int iters)
{
    double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double m = 1.0000001, c = 1e-7;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
            a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
        }
    }
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

// ---------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------
#define CK(call)                                                                                         \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess) {                                                                         \
            h->err = std::string(#call) + ": " + cudaGetErrorString(e_);                                 \
            return NRMC_ERR_CUDA;                                                                        \
        }                                                                                                \
    } while (0)

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t bytes)
    {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        cudaError_t e = cudaMalloc(&p, bytes);
        if (e == cudaSuccess) cap = bytes;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

struct Lane {               // one pipeline lane (scratch + timing events); lanes 0/1 own a stream (host-memory calls), lane DEV_LANE
                            // serves device-resident calls on the CALLER's stream (passed per call, never stored)
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t kev[3] = {nullptr, nullptr, nullptr};   // after K_classify, after K_hump, after the main attenuation kernel
    DevBuf in, out, work, fallback, sparse_tmp, rootq, humpq, packed, pack_sums, pack_off, modes, items;
    bool timed = false;
};

struct nrmc_rt_s {
    nrmc_rt_config cfg;
    IceParams ice;
    int S = 2, K1 = 1, n_sm = 148;
    std::string err;
    // frequencies
    std::vector<double> freq_out, freq_sparse;
    AttTables tb;
    Sp1Tables sp1;
    bool have_sp1 = false, have_gl1 = false, have_sep = false;
    int grid_att = 0, grid_sp1 = 0, grid_gl1 = 0, grid_sep = 0;      // resident blocks (occupancy x SMs) of the persistent attenuation kernels
    int grid_gl1_item = 0, grid_gl1_fine = 0;
    Gl1Tables gl1;
    int grid_small = 0, grid_small_noatt = 0;          // the same for the cooperative small-batch kernel (with / without the attenuation tables)
    void *small_host = nullptr;                        // pinned staging block of the small-batch host path (inputs and outputs in ONE copy each)
    size_t small_host_cap = 0;
    int grid_hump = 0, grid_roots = 0;   // the same for the persistent solver kernels
    int grid_hump_m = 0, grid_roots_m = 0;
    int64_t chunk_pairs = 0;             // 0: automatic; > 0: pairs per chunk (nrmc_rt_set_chunk_pairs)
    size_t smem_att = 0, smem_sp1 = 0, smem_gl1 = 0;
    DevBuf d_tables, d_gl3, d_sp1;
    bool have_freq = false;
    Lane lanes[N_LANES];
    cudaEvent_t dev_done = nullptr;   // recorded on the caller's stream after every device-resident call: the next call (possibly on
    bool dev_pending = false;         // another stream) and nrmc_rt_set_frequencies order themselves behind it
    DevBuf d_count;       // work-list counters (one block per lane) + the running row base
    DevBuf d_rmax;        // R_max table of the shadow-zone test
    RmaxTable rmax;
    DevBuf d_ant;         // antenna table for outer-product host calls
};

static void numpy_linspace(double a, double b, int n, std::vector<double> &out)
{
    if (n <= 0) return;
    if (n == 1) { out.push_back(a); return; }
    const double step = (b - a) / (n - 1);
    for (int i = 0; i < n - 1; ++i) out.push_back(a + i * step);
    out.push_back(b);
}

extern "C" {

const char *nrmc_rt_version(void) { return "nrmc_rt 0.1 (sm_100a)"; }

int nrmc_rt_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

int nrmc_rt_create(const nrmc_rt_config *cfg, nrmc_rt_t *out)
{
    if (!cfg || !out) return NRMC_ERR_INVALID_ARGUMENT;
    *out = nullptr;
    if (!(cfg->n_ice > 1.0) || !(cfg->delta_n > 0.0) || !(cfg->delta_n < cfg->n_ice) || !(cfg->z_0 > 0.0)) return NRMC_ERR_INVALID_ARGUMENT;
    if (cfg->n_reflections < 0 || cfg->n_reflections > NRMC_MAX_REFLECTIONS) return NRMC_ERR_UNSUPPORTED;
    if (cfg->attenuation_model < 0 || cfg->attenuation_model > 5) return NRMC_ERR_UNSUPPORTED;
    if (cfg->attenuation_model == NRMC_ATT_GL3 && (!cfg->gl3_table || cfg->gl3_rows < 2)) return NRMC_ERR_INVALID_ARGUMENT;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || cfg->device < 0 || cfg->device >= ndev) return NRMC_ERR_NO_DEVICE;
    nrmc_rt_s *h = new nrmc_rt_s();
    h->cfg = *cfg;
    IceParams &ice = h->ice;
    ice.n_ice = cfg->n_ice; ice.dn = cfg->delta_n; ice.z0 = cfg->z_0; ice.inv_z0 = 1.0 / cfg->z_0; ice.inv_dn = 1.0 / cfg->delta_n;
    ice.ns = cfg->n_ice - cfg->delta_n;
    int n_refl = cfg->n_reflections;
    if (n_refl > 0 && !(cfg->reflection_z == cfg->reflection_z)) n_refl = 0;   // propagation_base_class.py:128-133
    ice.n_refl = n_refl;
    ice.zr = n_refl > 0 ? cfg->reflection_z : -1e30;
    ice.gr = n_refl > 0 ? ice.dn * exp(ice.zr * ice.inv_z0) : 0.0;
    ice.nr = ice.n_ice - ice.gr;
    ice.att_model = cfg->attenuation_model;
    h->S = 2 + 4 * n_refl;
    h->K1 = n_refl + 1;
    memset(&h->tb, 0, sizeof(h->tb));
    if (cudaSetDevice(cfg->device) != cudaSuccess) { delete h; return NRMC_ERR_NO_DEVICE; }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, cfg->device) != cudaSuccess) { delete h; return NRMC_ERR_NO_DEVICE; }
    h->n_sm = prop.multiProcessorCount;
    for (int l = 0; l < N_LANES; ++l) {
        if (l != DEV_LANE && cudaStreamCreateWithFlags(&h->lanes[l].stream, cudaStreamNonBlocking) != cudaSuccess) { delete h; return NRMC_ERR_CUDA; }
        for (int e = 0; e < 6; ++e) cudaEventCreate(&h->lanes[l].ev[e]);
        for (int e = 0; e < 3; ++e) cudaEventCreate(&h->lanes[l].kev[e]);
    }
    if (cudaEventCreateWithFlags(&h->dev_done, cudaEventDisableTiming) != cudaSuccess) { delete h; return NRMC_ERR_CUDA; }
    if (h->d_count.reserve((CNT_ROWBASE + 16) * sizeof(unsigned long long)) != cudaSuccess) { delete h; return NRMC_ERR_CUDA; }
    {
        int nb = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, K_hump, HUMP_THREADS, 0) != cudaSuccess) { delete h; return NRMC_ERR_CUDA; }
        h->grid_hump = std::max(1, nb) * h->n_sm;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, K_roots<false>, ROOTS_THREADS, 0) != cudaSuccess) { delete h; return NRMC_ERR_CUDA; }
        h->grid_roots = std::max(1, nb) * h->n_sm;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, K_hump_m, HUMP_THREADS, 0) != cudaSuccess) { delete h; return NRMC_ERR_CUDA; }
        h->grid_hump_m = std::max(1, nb) * h->n_sm;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, K_roots_m, ROOTS_THREADS, 0) != cudaSuccess) { delete h; return NRMC_ERR_CUDA; }
        h->grid_roots_m = std::max(1, nb) * h->n_sm;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, K_small<false>, ATT_THREADS, 0) != cudaSuccess) { delete h; return NRMC_ERR_CUDA; }
        h->grid_small_noatt = std::max(1, nb) * h->n_sm;
    }
    h->rmax.t = nullptr; h->rmax.n = 0; h->rmax.dz = RMAX_DZ;
    {
        // R_max on a depth grid (used for the mode without bottom reflection): lets K_classify discard pairs far in the shadow zone without a maximum search
        if (h->d_rmax.reserve((size_t)RMAX_N * RMAX_N * sizeof(double)) != cudaSuccess) { delete h; return NRMC_ERR_CUDA; }
        K_rmax_table<<<(RMAX_N * RMAX_N + 127) / 128, 128>>>(ice, RMAX_N, RMAX_DZ, (double *)h->d_rmax.p);
        // R_max is monotone in both depths: a running maximum over the shallower corners keeps the table an upper bound
        // even where rounding noise of the search would dent it
        std::vector<double> t((size_t)RMAX_N * RMAX_N);
        if (cudaMemcpy(t.data(), h->d_rmax.p, t.size() * sizeof(double), cudaMemcpyDeviceToHost) != cudaSuccess) { delete h; return NRMC_ERR_CUDA; }
        for (int i1 = 0; i1 < RMAX_N; ++i1)
            for (int i2 = 0; i2 < RMAX_N; ++i2) {
                double v = t[(size_t)i1 * RMAX_N + i2];
                if (!(v == v)) v = INFINITY;
                if (i1 > 0) v = std::max(v, t[(size_t)(i1 - 1) * RMAX_N + i2]);
                if (i2 > 0) v = std::max(v, t[(size_t)i1 * RMAX_N + i2 - 1]);
                t[(size_t)i1 * RMAX_N + i2] = v;
            }
        if (cudaMemcpy(h->d_rmax.p, t.data(), t.size() * sizeof(double), cudaMemcpyHostToDevice) != cudaSuccess) { delete h; return NRMC_ERR_CUDA; }
        h->rmax.t = (const double *)h->d_rmax.p; h->rmax.n = RMAX_N;
    }
    if (cfg->attenuation_model == NRMC_ATT_GL3) {
        const size_t bytes = (size_t)cfg->gl3_rows * 3 * sizeof(double);
        if (h->d_gl3.reserve(bytes) != cudaSuccess ||
            cudaMemcpy(h->d_gl3.p, cfg->gl3_table, bytes, cudaMemcpyHostToDevice) != cudaSuccess) { delete h; return NRMC_ERR_CUDA; }
        h->tb.gl3.rows = (const double *)h->d_gl3.p;
        h->tb.gl3.n = cfg->gl3_rows;
    }
    *out = h;
    return NRMC_OK;
}

void nrmc_rt_destroy(nrmc_rt_t h)
{
    if (!h) return;
    cudaSetDevice(h->cfg.device);
    if (h->dev_pending) cudaEventSynchronize(h->dev_done);
    if (h->dev_done) cudaEventDestroy(h->dev_done);
    for (int l = 0; l < N_LANES; ++l) {
        if (h->lanes[l].stream) { cudaStreamSynchronize(h->lanes[l].stream); cudaStreamDestroy(h->lanes[l].stream); }
        for (int e = 0; e < 6; ++e) if (h->lanes[l].ev[e]) cudaEventDestroy(h->lanes[l].ev[e]);
        for (int e = 0; e < 3; ++e) if (h->lanes[l].kev[e]) cudaEventDestroy(h->lanes[l].kev[e]);
        h->lanes[l].in.release(); h->lanes[l].out.release(); h->lanes[l].work.release();
        h->lanes[l].fallback.release(); h->lanes[l].sparse_tmp.release(); h->lanes[l].rootq.release(); h->lanes[l].humpq.release();
        h->lanes[l].packed.release(); h->lanes[l].pack_sums.release(); h->lanes[l].pack_off.release(); h->lanes[l].modes.release(); h->lanes[l].items.release();
    }
    h->d_tables.release(); h->d_gl3.release(); h->d_sp1.release(); h->d_count.release(); h->d_ant.release(); h->d_rmax.release();
    if (h->small_host) cudaFreeHost(h->small_host);
    delete h;
}

const char *nrmc_rt_last_error(nrmc_rt_t h) { return h ? h->err.c_str() : "null handle"; }
int nrmc_rt_max_solutions(nrmc_rt_t h) { return h ? h->S : NRMC_ERR_INVALID_ARGUMENT; }

int nrmc_rt_set_frequencies(nrmc_rt_t h, const double *frequency, int32_t n, double max_detector_freq)
{
    if (!h || !frequency || n <= 0) return NRMC_ERR_INVALID_ARGUMENT;
    if (h->ice.att_model == 0) { h->err = "no attenuation model configured"; return NRMC_ERR_UNSUPPORTED; }
    cudaSetDevice(h->cfg.device);
    // the tables are rewritten in place: wait for device-resident calls still reading them on the caller's stream
    if (h->dev_pending) { CK(cudaEventSynchronize(h->dev_done)); h->dev_pending = false; }
    // --- __get_frequencies_for_attenuation, analyticraytracing.py:885-931 ---
    const int n_int = h->cfg.n_frequencies_integration > 0 ? h->cfg.n_frequencies_integration : 100;
    int n_nonnull = 0;
    double flo = INFINITY, fhi = -INFINITY;
    for (int i = 0; i < n; ++i) if (frequency[i] > 0) { ++n_nonnull; flo = std::min(flo, frequency[i]); fhi = std::max(fhi, frequency[i]); }
    if (n_nonnull == 0) { h->err = "no frequency above 0"; return NRMC_ERR_NO_FREQUENCIES; }
    std::vector<double> sp;
    int nf = std::min(n_int, n_nonnull);
    numpy_linspace(flo, fhi, nf, sp);
    if (nf < n_nonnull && max_detector_freq == max_detector_freq) {
        int n_tot = 0, n_above = 0;
        double tlo = INFINITY, thi = -INFINITY, alo = INFINITY, ahi = -INFINITY;
        for (int i = 0; i < n; ++i) {
            const bool det = frequency[i] <= max_detector_freq;
            if (det && frequency[i] > 0) { ++n_tot; tlo = std::min(tlo, frequency[i]); thi = std::max(thi, frequency[i]); }
            if (!det) { ++n_above; alo = std::min(alo, frequency[i]); ahi = std::max(ahi, frequency[i]); }
        }
        if (n_tot == 0) { h->err = "no frequency in (0, max_detector_freq]"; return NRMC_ERR_NO_FREQUENCIES; }
        nf = std::min(n_int, n_tot);
        sp.clear();
        numpy_linspace(tlo, thi, nf, sp);
        if (n_above > 1) numpy_linspace(alo, ahi, nf / 2, sp);
    }
    const int Fs = (int)sp.size();
    const int Fs_pad = (Fs + 1) & ~1, F_pad = (n + 3) & ~3;
    // --- np.interp tables (py:1077-1078): left index + weight per output bin; bins <= 0 Hz keep factor 1 ---
    std::vector<double> fa(Fs_pad, 0.0), fb(Fs_pad, 0.0), it(F_pad, 0.0);
    std::vector<int32_t> ii(F_pad, -1);
    for (int j = 0; j < Fs; ++j) att_freq_consts(h->ice.att_model, sp[j], fa[j], fb[j]);
    for (int b = 0; b < n; ++b) {
        const double x = frequency[b];
        if (!(x > 0)) { ii[b] = -1; continue; }
        if (Fs == 1) { ii[b] = 0; it[b] = 0.0; continue; }   // np.interp on one point: the value itself (the kernels do not touch i0+1 when t == 0)
        if (x <= sp[0]) { ii[b] = 0; it[b] = 0.0; }
        else if (x >= sp[Fs - 1]) { ii[b] = Fs - 2; it[b] = 1.0; }
        else {
            int lo = 0, hi = Fs - 1;
            while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (sp[mid] <= x) lo = mid; else hi = mid; }
            ii[b] = lo;
            it[b] = (x - sp[lo]) / (sp[hi] - sp[lo]);
        }
    }
    if (Fs == 1) { fa[1] = fa[0]; fb[1] = fb[0]; }               // Fs_pad == 2: the padding entry repeats the frequency
    const size_t bytes = (size_t)Fs_pad * 16 + (size_t)F_pad * 12 + 64;
    CK(h->d_tables.reserve(bytes));
    unsigned char *base = (unsigned char *)h->d_tables.p;
    double *d_fa = (double *)base, *d_fb = d_fa + Fs_pad, *d_it = d_fb + Fs_pad;
    int32_t *d_ii = (int32_t *)(d_it + F_pad);
    CK(cudaMemcpy(d_fa, fa.data(), Fs_pad * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_fb, fb.data(), Fs_pad * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_it, it.data(), F_pad * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_ii, ii.data(), F_pad * 4, cudaMemcpyHostToDevice));
    h->tb.fa = d_fa; h->tb.fb = d_fb; h->tb.it = d_it; h->tb.ii = d_ii;
    h->tb.Fs = Fs; h->tb.Fs_pad = Fs_pad; h->tb.F = n; h->tb.F_pad = F_pad;
    h->freq_out.assign(frequency, frequency + n);
    h->freq_sparse = sp;
    h->have_freq = true;
    // persistent-grid sizes: exactly the resident blocks, so that no partial second wave forms
    {
        const int nseg_max = h->ice.n_refl + 1;
        h->smem_att = (size_t)Fs_pad * 16 + (size_t)F_pad * 12 + (size_t)ATT_WARPS * (3 + nseg_max) * Fs_pad * 8;
        if (h->smem_att > 48 * 1024) {
            CK(cudaFuncSetAttribute(K_att<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_att));
            CK(cudaFuncSetAttribute(K_att<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_att));
        }
        int nb = 0;
        if (h->ice.att_model == NRMC_ATT_GL3) CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, K_att<true>, ATT_THREADS, h->smem_att));
        else CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, K_att<false>, ATT_THREADS, h->smem_att));
        h->grid_att = std::max(1, nb) * h->n_sm;
        if (h->smem_att > 48 * 1024) {
            CK(cudaFuncSetAttribute(K_small<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_att));
            CK(cudaFuncSetAttribute(K_small<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_att));
        }
        if (h->ice.att_model == NRMC_ATT_GL3) CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, K_small<true>, ATT_THREADS, h->smem_att));
        else CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, K_small<false>, ATT_THREADS, h->smem_att));
        h->grid_small = std::max(1, nb) * h->n_sm;
    }
    h->have_sep = false;
    if ((h->ice.att_model == NRMC_ATT_MB1 || h->ice.att_model == NRMC_ATT_GL2) && !getenv("NRMC_SEP_GENERIC")) {     // (the variable: A/B and tests)
        h->smem_gl1 = (size_t)Fs_pad * 8 + (size_t)(SP1_THREADS / 32) * 32 * GL1_ROW * sizeof(double);
        int nb = 0;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, K_att_sep, SP1_THREADS, h->smem_gl1));
        h->grid_sep = std::max(1, nb) * h->n_sm;
        h->have_sep = true;
    }
    h->have_gl1 = false;
    if (h->ice.att_model == NRMC_ATT_GL1 && h->ice.n_refl == 0 && !getenv("NRMC_GL1_GENERIC")) {     // (the variable: A/B and tests)
        h->smem_gl1 = (size_t)Fs_pad * 8 + (size_t)(SP1_THREADS / 32) * 32 * GL1_ROW * sizeof(double);
        int nb = 0;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, K_att_gl1, SP1_THREADS, h->smem_gl1));
        h->grid_gl1 = std::max(1, nb) * h->n_sm;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, K_gl1_item, 128, 0));
        h->grid_gl1_item = std::max(1, nb) * h->n_sm;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, K_gl1_fine, 128, 0));
        h->grid_gl1_fine = std::max(1, nb) * h->n_sm;
        {
            // the 75 MHz length A(z) (attenuation.py:99-128) on a 0.05 m grid: where it reaches its 100 m floor, where it turns, how steep it gets
            auto A = [](double z) { return (((( -3.63912864e-14 * z - 2.21040482e-10) * z - 3.50628312e-07) * z - 9.82378264e-05) * z + 6.87257150e-02) * z + 1.16052586e+03; };
            Gl1Tables &g = h->gl1;
            g.nzx = 0; g.pad = 0; g.z_min = -4000.0;
            const double dz = 0.05;
            for (double z = 0.0; z > -4000.0; z -= dz) if (A(z - dz) < 100.0 + 1.0) { g.z_min = z; break; }
            double smax = 0.0, dprev = 0.0;
            for (double z = g.z_min; z < 0.0; z += dz) {
                const double d = (A(z + dz) - A(z)) / dz;
                smax = std::max(smax, fabs(d));
                if (z > g.z_min && d * dprev < 0.0 && g.nzx < 4) g.zx[g.nzx++] = z;
                dprev = d;
            }
            for (int e = g.nzx; e < 4; ++e) g.zx[e] = 1.0;
            g.inv_dadz_max = 1.0 / (1.02 * smax);
        }
        h->have_gl1 = true;
    }
    h->have_sp1 = false;
    if (h->ice.att_model == NRMC_ATT_SP1 && h->ice.n_refl == 0) {
        // tables of the SP1 moment kernel: Chebyshev coefficients in the normalised temperature of 1/L(T, f_j) (attenuation.py:170-192)
        Sp1Tables &t = h->sp1;
        const double SB0[3] = {-6.74890, 0.026709, -0.000884}, SB1[3] = {-6.22121, -0.070927, -0.001773}, SB2[3] = {-4.09468, -0.002213, -0.000332};
        const double TC[4] = {1.83415e-09, -1.59061e-08, 0.00267687, -51.0696};          // temperature profile, attenuation.py:141-142
        auto temp = [&](double a) { return ((TC[0] * a + TC[1]) * a + TC[2]) * a + TC[3]; };
        const double T0 = temp(0.0), T1 = temp(SP1_DEPTH_MAX), Tmid = 0.5 * (T0 + T1), Thalf = 0.5 * (T1 - T0);
        for (int m = 0; m < 4; ++m) t.tau[m] = (TC[m] - (m == 3 ? Tmid : 0.0)) / Thalf;
        t.depth_max = SP1_DEPTH_MAX;
        std::vector<double> wk((size_t)Fs_pad * SP1_K, 0.0);
        bool series_ok = true;
        const int NC = 64, KX = SP1_KT + 6;
        for (int j = 0; j < Fs && series_ok; ++j) {
            const double w = log(sp[j]);
            const bool hi = !(sp[j] < 1.0);                                               // attenuation.py:180-185
            double coef[KX], g[NC];
            for (int m = 0; m < NC; ++m) {
                const double T = Tmid + Thalf * cos(M_PI * (m + 0.5) / NC);
                const double b0 = SB0[0] + T * (SB0[1] + T * SB0[2]), b1 = SB1[0] + T * (SB1[1] + T * SB1[2]), b2 = SB2[0] + T * (SB2[1] + T * SB2[2]);
                const double q = b1 + (hi ? (b2 - b1) / 1.1505720275988207 : (b1 - b0) / 9.210340371976182) * w;
                if (!(q < 0.0)) series_ok = false;                                        // 1 m floor (attenuation.py:252-255) within reach
                g[m] = exp(q);
            }
            for (int k = 0; k < KX; ++k) {
                double sum = 0.0;
                for (int m = 0; m < NC; ++m) sum += g[m] * cos(k * M_PI * (m + 0.5) / NC);
                coef[k] = (k == 0 ? 1.0 : 2.0) * sum / NC;
            }
            double tail = 0.0;
            for (int k = SP1_KT; k < KX; ++k) tail += fabs(coef[k]);
            if (!(tail <= SP1_TRUNCATION * fabs(coef[0]))) series_ok = false;            // this frequency set needs more moments
            for (int k = 0; k < SP1_KT; ++k) wk[(size_t)j * SP1_K + k] = coef[k];
        }
        const size_t b_wk = wk.size() * 8;
        CK(h->d_sp1.reserve(b_wk + 64));
        CK(cudaMemcpy(h->d_sp1.p, wk.data(), b_wk, cudaMemcpyHostToDevice));
        t.wk = (const double *)h->d_sp1.p;
        h->smem_sp1 = b_wk + (size_t)(SP1_THREADS / 32) * 32 * SP1_ROW * sizeof(double);
        if (h->smem_sp1 <= 200 * 1024 && series_ok) {
            if (h->smem_sp1 > 48 * 1024) CK(cudaFuncSetAttribute(K_att_sp1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_sp1));
            int nb = 0;
            CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, K_att_sp1, SP1_THREADS, h->smem_sp1));
            h->grid_sp1 = std::max(1, nb) * h->n_sm;
            h->have_sp1 = true;
        }
    }
    return Fs;
}

int nrmc_rt_get_sparse_frequencies(nrmc_rt_t h, double *out, int32_t capacity)
{
    if (!h || !h->have_freq) return NRMC_ERR_NO_FREQUENCIES;
    const int n = (int)h->freq_sparse.size();
    if (out) for (int i = 0; i < n && i < capacity; ++i) out[i] = h->freq_sparse[i];
    return n;
}

int nrmc_rt_set_chunk_pairs(nrmc_rt_t h, int64_t pairs)
{
    if (!h || pairs < 0) return NRMC_ERR_INVALID_ARGUMENT;
    h->chunk_pairs = pairs;
    return NRMC_OK;
}

int nrmc_rt_host_alloc(void **ptr, uint64_t bytes)
{
    if (!ptr) return NRMC_ERR_INVALID_ARGUMENT;
    return cudaHostAlloc(ptr, bytes, cudaHostAllocDefault) == cudaSuccess ? NRMC_OK : NRMC_ERR_CUDA;
}
int nrmc_rt_host_free(void *ptr) { return cudaFreeHost(ptr) == cudaSuccess ? NRMC_OK : NRMC_ERR_CUDA; }

// ---- result gather over NVLink peer memory: CUDA IPC handles of plain cudaMalloc blocks --------------------------------------
int nrmc_rt_peer_alloc(int32_t device, uint64_t bytes, void **dev_ptr, unsigned char *handle)
{
    if (!dev_ptr || !handle || bytes == 0) return NRMC_ERR_INVALID_ARGUMENT;
    static_assert(sizeof(cudaIpcMemHandle_t) == NRMC_PEER_HANDLE_BYTES, "handle size");
    if (cudaSetDevice(device) != cudaSuccess) return NRMC_ERR_NO_DEVICE;
    void *p = nullptr;
    if (cudaMalloc(&p, bytes) != cudaSuccess) { cudaGetLastError(); return NRMC_ERR_CUDA; }
    cudaIpcMemHandle_t hd;
    if (cudaIpcGetMemHandle(&hd, p) != cudaSuccess) { cudaGetLastError(); cudaFree(p); return NRMC_ERR_CUDA; }
    memcpy(handle, &hd, sizeof(hd));
    *dev_ptr = p;
    return NRMC_OK;
}
int nrmc_rt_peer_open(int32_t device, const unsigned char *handle, void **dev_ptr)
{
    if (!dev_ptr || !handle) return NRMC_ERR_INVALID_ARGUMENT;
    if (cudaSetDevice(device) != cudaSuccess) return NRMC_ERR_NO_DEVICE;
    cudaIpcMemHandle_t hd;
    memcpy(&hd, handle, sizeof(hd));
    void *p = nullptr;
    if (cudaIpcOpenMemHandle(&p, hd, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); return NRMC_ERR_CUDA; }
    *dev_ptr = p;
    return NRMC_OK;
}
int nrmc_rt_peer_close(void *dev_ptr) { return cudaIpcCloseMemHandle(dev_ptr) == cudaSuccess ? NRMC_OK : NRMC_ERR_CUDA; }
int nrmc_rt_peer_free(void *dev_ptr) { return cudaFree(dev_ptr) == cudaSuccess ? NRMC_OK : NRMC_ERR_CUDA; }
int nrmc_rt_copy_async(void *dst, const void *src, uint64_t bytes, void *stream)
{
    if (bytes == 0) return NRMC_OK;
    if (!dst || !src) return NRMC_ERR_INVALID_ARGUMENT;
    return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, (cudaStream_t)stream) == cudaSuccess ? NRMC_OK : NRMC_ERR_CUDA;
}

}  // extern "C"

// enqueue the kernels for one chunk of pairs whose inputs/outputs are device resident
// compact layout inside the binned solver: rows are assigned by a scan of n_sol between K_hump and K_roots
struct CompactCtx {
    bool on = false;
    int64_t *sol_offset = nullptr;          // [n_pairs of the chunk] device, receives GLOBAL rows
    unsigned long long *base_dev = nullptr; // device-side running row base (device-resident calls), or
    int64_t base_host = 0;                  // the row base known on the host (host calls)
    int64_t row_limit = 0;                  // rows the caller's per-slot arrays hold
};

static int launch_chunk(nrmc_rt_s *h, Lane &ln, int lane_id, cudaStream_t st, const KInput &kin, const TraceOutputs &to_in, double *att_sparse,
                        double *att_dense, int *n_launches, const CompactCtx &cc = CompactCtx())
{
    TraceOutputs to = to_in;
    to.row_offset = cc.on ? cc.sol_offset : nullptr;
    to.row_limit = cc.row_limit;
    const bool want_att = (att_sparse || att_dense);
    unsigned long long *cnt = (unsigned long long *)h->d_count.p + CNT_STRIDE * lane_id;
    unsigned long long *d_count = cnt + CNT_WORK;
    const unsigned long long work_cap = (unsigned long long)kin.n_pairs * h->S;
    WorkList wl;
    memset(&wl, 0, sizeof(wl));
    if (want_att) {
        CK(ln.work.reserve(worklist_bytes(work_cap)));
        wl = carve_worklist(ln.work.p, work_cap);
    }
    CK(cudaMemsetAsync(cnt, 0, CNT_STRIDE * sizeof(unsigned long long), st));
    if (ln.timed) cudaEventRecord(ln.ev[0], st);
    const int M = 1 + 2 * h->ice.n_refl;
    unsigned long long *d_roots = cnt + CNT_ROOTS, *d_humps = cnt + CNT_HUMPS;
    CK(ln.rootq.reserve(rootq_bytes((size_t)kin.n_pairs * M)));
    CK(ln.humpq.reserve(humpq_bytes((size_t)kin.n_pairs * M)));
    const RootQ rootq = carve_rootq(ln.rootq.p, (size_t)kin.n_pairs * M);
    const HumpQ humpq = carve_humpq(ln.humpq.p, (size_t)kin.n_pairs * M);
    AttFill af;
    af.sparse = att_sparse; af.dense = att_dense; af.Fs = h->tb.Fs; af.F = h->tb.F;
    auto row_scan = [&]() -> int {
        // compact layout: exclusive scan of n_sol -> row of every pair's first solution (global rows)
        const int nblk = (int)((kin.n_pairs + PACK_PAIRS - 1) / PACK_PAIRS);
        CK(ln.pack_sums.reserve((size_t)(nblk + 2) * sizeof(unsigned long long)));
        unsigned long long *sums = (unsigned long long *)ln.pack_sums.p;
        PackArrays none;
        none.n = 0;
        K_pack_count<<<nblk, PACK_THREADS, 0, st>>>(to.n_sol, kin.n_pairs, sums);
        K_pack_scan<<<1, 1024, 0, st>>>(sums, nblk, cc.base_dev);
        K_pack<<<nblk, PACK_THREADS, 0, st>>>(to.n_sol, kin.n_pairs, h->S, sums, cc.base_host, cc.base_dev ? 1 : 0, cc.sol_offset, none);
        *n_launches += 3;
        return NRMC_OK;
    };
    if (h->ice.n_refl == 0) {
        // binned solver: classify -> hump search -> roots, re-packed through queues
        const int64_t blocks = (kin.n_pairs + CLASSIFY_THREADS - 1) / CLASSIFY_THREADS;
        K_classify<<<(unsigned)blocks, CLASSIFY_THREADS, 0, st>>>(h->ice, kin, to, af, h->rmax, rootq, d_roots, humpq, d_humps);
        if (ln.timed) cudaEventRecord(ln.kev[0], st);
        K_hump<<<h->grid_hump, HUMP_THREADS, 0, st>>>(h->ice, kin, to, af, humpq, d_humps, rootq, d_roots);
        if (cc.on) { const int rc = row_scan(); if (rc != NRMC_OK) return rc; }
        if (ln.timed) cudaEventRecord(ln.kev[1], st);
        if (kin.sx)
            K_roots<true><<<h->grid_roots, ROOTS_THREADS, 0, st>>>(h->ice, kin, to, af, rootq, d_roots, wl, d_count);
        else
            K_roots<false><<<h->grid_roots, ROOTS_THREADS, 0, st>>>(h->ice, kin, to, af, rootq, d_roots, wl, d_count);
        *n_launches += 3;
        if (wl.beta) { K_unpack_work<<<1, 1, 0, st>>>(cnt); ++*n_launches; }
    } else {
        // bottom reflections: the same pipeline over (pair, mode) work items
        CK(ln.modes.reserve((size_t)kin.n_pairs * M));
        int8_t *modes = (int8_t *)ln.modes.p;
#ifdef CLASSIFY_M_PAIR_MAJOR
        const int64_t items = kin.n_pairs * M;
#else
        const int64_t items = ((kin.n_pairs + 31) / 32) * 32 * M;       // tiles of 32 pairs x M modes
#endif
        K_classify_m<<<(unsigned)((items + CLASSIFY_THREADS - 1) / CLASSIFY_THREADS), CLASSIFY_THREADS, 0, st>>>(
            h->ice, kin, to, h->rmax, M, modes, rootq, d_roots, humpq, d_humps);
        if (ln.timed) cudaEventRecord(ln.kev[0], st);
        K_hump_m<<<h->grid_hump_m, HUMP_THREADS, 0, st>>>(h->ice, kin, M, modes, humpq, d_humps, rootq, d_roots);
        K_slots_m<<<(unsigned)((kin.n_pairs + 255) / 256), 256, 0, st>>>(kin, to, af, M, h->S, h->K1, modes);
        if (cc.on) { const int rc = row_scan(); if (rc != NRMC_OK) return rc; }
        if (ln.timed) cudaEventRecord(ln.kev[1], st);
        K_roots_m<<<h->grid_roots_m, ROOTS_THREADS, 0, st>>>(h->ice, kin, to, af, M, h->S, h->K1, modes, rootq, d_roots, wl, d_count);
        *n_launches += 4;
    }
    if (ln.timed) cudaEventRecord(ln.ev[1], st);
    if (want_att) {
        const AttTables &tb = h->tb;
        const int nseg_max = h->ice.n_refl + 1;
        // separable models: the thread-per-solution kernel forms the PRODUCT of the segments' factors, which is all a sparse output
        // needs; a dense output of a bottom-reflected path interpolates every segment on its own (generic kernel)
        const bool use_sep = h->have_sep && (h->ice.n_refl == 0 || att_dense == nullptr);
        if (h->have_sp1 || h->have_gl1 || use_sep) {
            // thread-per-solution kernel (SP1 moments / GL1) -> sparse factors; the solutions it hands back -> generic kernel;
            // dense = interp(sparse)
            unsigned long long *d_fb = cnt + CNT_FALLBACK;
            CK(ln.fallback.reserve(worklist_bytes(work_cap)));
            const WorkList fb = carve_worklist(ln.fallback.p, work_cap);
            double *sparse = att_sparse;
            const int sparse_is_tmp = sparse ? 0 : 1;
            if (!sparse) {
                CK(ln.sparse_tmp.reserve((size_t)kin.n_pairs * h->S * tb.Fs * sizeof(double)));
                sparse = (double *)ln.sparse_tmp.p;
            }
            if (h->have_gl1) {
                // moments + closed-form series; the hard (solution, frequency) items go through two item kernels
                const unsigned long long item_cap = 2ull * work_cap;
                CK(ln.items.reserve((size_t)item_cap * 2 * sizeof(unsigned long long)));
                unsigned long long *items = (unsigned long long *)ln.items.p, *fine = items + item_cap;
                K_att_gl1<<<h->grid_gl1, SP1_THREADS, h->smem_gl1, st>>>(h->ice, tb, h->gl1, wl, d_count, sparse_is_tmp, sparse, fb, d_fb, items,
                                                                         cnt + CNT_ITEMS, item_cap);
                K_gl1_item<<<h->grid_gl1_item, 128, 0, st>>>(h->ice, tb, wl, sparse_is_tmp, sparse, items, cnt + CNT_ITEMS, item_cap, fine, cnt + CNT_FINE);
                K_gl1_fine<<<h->grid_gl1_fine, 128, 0, st>>>(h->ice, tb, wl, sparse_is_tmp, sparse, fine, cnt + CNT_FINE);
                *n_launches += 2;
            } else if (use_sep)
                K_att_sep<<<h->grid_sep, SP1_THREADS, h->smem_gl1, st>>>(h->ice, tb, wl, d_count, sparse_is_tmp, sparse, fb, d_fb);
            else
                K_att_sp1<<<h->grid_sp1, SP1_THREADS, h->smem_sp1, st>>>(h->ice, kin, tb, h->sp1, wl, d_count, sparse_is_tmp, sparse, fb, d_fb);
            if (ln.timed) cudaEventRecord(ln.kev[2], st);
            K_att<false><<<h->grid_att, ATT_THREADS, h->smem_att, st>>>(h->ice, kin, tb, fb, d_fb, cnt + CNT_TICKET_KATT, nseg_max, sparse, nullptr);
            *n_launches += 2;
            if (att_dense) {
                K_att_expand<<<h->n_sm * 8, 256, 0, st>>>(tb, wl, d_count, sparse_is_tmp, sparse, att_dense);
                ++*n_launches;
            }
        } else {
            if (h->ice.att_model == NRMC_ATT_GL3)
                K_att<true><<<h->grid_att, ATT_THREADS, h->smem_att, st>>>(h->ice, kin, tb, wl, d_count, cnt + CNT_TICKET_KATT, nseg_max, att_sparse, att_dense);
            else
                K_att<false><<<h->grid_att, ATT_THREADS, h->smem_att, st>>>(h->ice, kin, tb, wl, d_count, cnt + CNT_TICKET_KATT, nseg_max, att_sparse, att_dense);
            ++*n_launches;
            if (ln.timed) cudaEventRecord(ln.kev[2], st);
        }
    } else if (ln.timed) cudaEventRecord(ln.kev[2], st);
    if (ln.timed) cudaEventRecord(ln.ev[2], st);
    CK(cudaGetLastError());
    return NRMC_OK;
}

// adds the CUDA-event durations of the chunk last run on this lane: ms[0] solver, ms[1] attenuation, ms[2..6] per kernel
static void accumulate_lane_times(Lane &ln, float *ms)
{
    cudaEventSynchronize(ln.ev[2]);
    float t = 0;
    cudaEventElapsedTime(&t, ln.ev[0], ln.ev[1]); ms[0] += t;
    cudaEventElapsedTime(&t, ln.ev[1], ln.ev[2]); ms[1] += t;
    cudaEventElapsedTime(&t, ln.ev[0], ln.kev[0]); ms[2] += t;    // K_classify / K_classify_m
    cudaEventElapsedTime(&t, ln.kev[0], ln.kev[1]); ms[3] += t;   // K_hump
    cudaEventElapsedTime(&t, ln.kev[1], ln.ev[1]); ms[4] += t;    // K_roots
    cudaEventElapsedTime(&t, ln.ev[1], ln.kev[2]); ms[5] += t;    // main attenuation kernel (+ NaN fill for bottom reflections)
    cudaEventElapsedTime(&t, ln.kev[2], ln.ev[2]); ms[6] += t;    // fallback / dense expansion kernels
}

static void store_times(nrmc_rt_stats *stats, const float *ms)
{
    stats->ms_solve = ms[0]; stats->ms_attenuation = ms[1];
    for (int i = 0; i < 5; ++i) stats->ms_kernel[i] = ms[2 + i];
}

struct OutLayout {           // byte offsets of every output inside one contiguous per-chunk device block
    size_t off[N_OUT];
    size_t elem[N_OUT];         // bytes per pair
    size_t total;
};

static void *out_ptr(const nrmc_rt_output *o, int i)
{
    switch (i) {
    case 0: return o->n_sol; case 1: return o->status; case 2: return o->solution_type; case 3: return o->reflection;
    case 4: return o->reflection_case; case 5: return o->C0; case 6: return o->C1; case 7: return o->path_length;
    case 8: return o->travel_time; case 9: return o->launch_vector; case 10: return o->receive_vector;
    case 11: return o->reflection_angle; case 12: return o->attenuation_sparse; case 13: return o->attenuation;
    case 14: return o->viewing_angle;
    }
    return nullptr;
}

// one cooperative launch for a small batch in the padded layout (device pointers; scratch of lane `lane_id`)
static int launch_small(nrmc_rt_s *h, Lane &ln, int lane_id, cudaStream_t st, const KInput &kin, const TraceOutputs &to, double *att_sparse,
                        double *att_dense, int *n_launches)
{
    const bool want_att = att_sparse || att_dense;
    unsigned long long *cnt = (unsigned long long *)h->d_count.p + CNT_STRIDE * lane_id;
    const unsigned long long work_cap = (unsigned long long)kin.n_pairs * h->S;
    WorkList wl;
    memset(&wl, 0, sizeof(wl));
    if (want_att) {
        CK(ln.work.reserve(worklist_bytes(work_cap)));
        wl = carve_worklist(ln.work.p, work_cap);
        CK(cudaMemsetAsync(cnt, 0, CNT_STRIDE * sizeof(unsigned long long), st));
    }
    if (ln.timed) cudaEventRecord(ln.ev[0], st);
    AttFill af;
    af.sparse = att_sparse; af.dense = att_dense; af.Fs = h->tb.Fs; af.F = h->tb.F;
    IceParams ice = h->ice;
    KInput k = kin;
    TraceOutputs t = to;
    t.row_offset = nullptr; t.row_limit = 0;
    AttTables tb = h->tb;
    unsigned long long *d_count = cnt + CNT_WORK;
    int nseg_max = h->ice.n_refl + 1;
    const int64_t want_blocks = std::max<int64_t>((kin.n_pairs + ATT_THREADS - 1) / ATT_THREADS, want_att ? ((int64_t)work_cap + ATT_WARPS - 1) / ATT_WARPS : 1);
    const int grid = (int)std::min<int64_t>(want_att ? h->grid_small : h->grid_small_noatt, want_blocks);
    void *args[] = {&ice, &k, &t, &af, &tb, &wl, &d_count, &nseg_max, &att_sparse, &att_dense};
    const void *fn = h->ice.att_model == NRMC_ATT_GL3 ? (const void *)K_small<true> : (const void *)K_small<false>;
    CK(cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(ATT_THREADS), args, want_att ? h->smem_att : 0, st));
    ++*n_launches;
    if (ln.timed) { cudaEventRecord(ln.ev[1], st); cudaEventRecord(ln.kev[0], st); cudaEventRecord(ln.kev[1], st); cudaEventRecord(ln.kev[2], st); cudaEventRecord(ln.ev[2], st); }
    return NRMC_OK;
}

static void accumulate_lane_times(Lane &ln, float *ms);
static void store_times(nrmc_rt_stats *stats, const float *ms);
static void *out_ptr(const nrmc_rt_output *o, int i);

// Host-memory call on a small batch: ONE pinned staging block carries all inputs to the device in one copy and all outputs back
// in one copy (a scalar call of the Python class spent 230 us in ~20 separate copies, launches and synchronisations)
#define NRMC_SMALL_STAGE_MAX ((size_t)8 << 20)
static int trace_small_host(nrmc_rt_s *h, const nrmc_rt_input *in, const nrmc_rt_output *out, nrmc_rt_stats *stats, int64_t N, const size_t *elem)
{
    Lane &ln = h->lanes[0];
    cudaStream_t st = ln.stream;
    const int64_t nv = in->n_vertices, na = in->n_antennas;
    const size_t n_in = (size_t)(3 * nv + 3 * na + (in->sx ? 3 * nv : 0));
    const size_t in_bytes = (n_in * 8 + 15) & ~(size_t)15;
    bool want[N_OUT];
    size_t off[N_OUT], total = 0;
    for (int i = 0; i < N_OUT; ++i) {
        want[i] = out_ptr(out, i) != nullptr;
        off[i] = total;
        if (want[i]) total += ((size_t)N * elem[i] + 15) & ~(size_t)15;
    }
    const bool staged = total <= NRMC_SMALL_STAGE_MAX;
    const size_t need = in_bytes + (staged ? total : 0);
    if (need > h->small_host_cap) {
        if (h->small_host) cudaFreeHost(h->small_host);
        h->small_host = nullptr; h->small_host_cap = 0;
        const size_t cap = std::max(need, (size_t)1 << 16);
        CK(cudaHostAlloc(&h->small_host, cap, cudaHostAllocDefault));
        h->small_host_cap = cap;
    }
    double *hin = (double *)h->small_host;
    memcpy(hin, in->vx, nv * 8); memcpy(hin + nv, in->vy, nv * 8); memcpy(hin + 2 * nv, in->vz, nv * 8);
    memcpy(hin + 3 * nv, in->ax, na * 8); memcpy(hin + 3 * nv + na, in->ay, na * 8); memcpy(hin + 3 * nv + 2 * na, in->az, na * 8);
    if (in->sx) { double *hs = hin + 3 * nv + 3 * na; memcpy(hs, in->sx, nv * 8); memcpy(hs + nv, in->sy, nv * 8); memcpy(hs + 2 * nv, in->sz, nv * 8); }
    CK(ln.in.reserve(in_bytes));
    CK(ln.out.reserve(std::max<size_t>(total, 16)));
    cudaEvent_t e0 = ln.ev[3], e1 = ln.ev[4];
    if (stats) cudaEventRecord(e0, st);
    CK(cudaMemcpyAsync(ln.in.p, hin, n_in * 8, cudaMemcpyHostToDevice, st));
    double *din = (double *)ln.in.p;
    KInput kin;
    kin.outer = in->outer; kin.n_antennas = na; kin.n_pairs = N; kin.delta_C_cut = in->delta_C_cut;
    kin.vx = din; kin.vy = din + nv; kin.vz = din + 2 * nv; kin.ax = din + 3 * nv; kin.ay = din + 3 * nv + na; kin.az = din + 3 * nv + 2 * na;
    kin.sx = kin.sy = kin.sz = nullptr;
    if (in->sx) { kin.sx = din + 3 * nv + 3 * na; kin.sy = kin.sx + nv; kin.sz = kin.sx + 2 * nv; }
    unsigned char *dout = (unsigned char *)ln.out.p;
    auto dp = [&](int i) -> void * { return want[i] ? (void *)(dout + off[i]) : nullptr; };
    TraceOutputs to;
    to.n_sol = (int32_t *)dp(0); to.status = (int32_t *)dp(1); to.type = (int8_t *)dp(2); to.reflection = (int8_t *)dp(3);
    to.reflection_case = (int8_t *)dp(4); to.C0 = (double *)dp(5); to.C1 = (double *)dp(6); to.path_length = (double *)dp(7);
    to.travel_time = (double *)dp(8); to.launch = (double *)dp(9); to.receive = (double *)dp(10); to.reflection_angle = (double *)dp(11);
    to.viewing_angle = (double *)dp(14);
    to.row_offset = nullptr; to.row_limit = 0;
    ln.timed = (stats != nullptr);
    int n_launches = 0;
    const int rc = launch_small(h, ln, 0, st, kin, to, (double *)dp(12), (double *)dp(13), &n_launches);
    if (rc != NRMC_OK) { ln.timed = false; return rc; }
    int64_t d2h = 0;
    unsigned char *hout = (unsigned char *)h->small_host + in_bytes;
    if (staged) {
        if (total) CK(cudaMemcpyAsync(hout, dout, total, cudaMemcpyDeviceToHost, st));
        d2h = (int64_t)total;
    } else {
        for (int i = 0; i < N_OUT; ++i)
            if (want[i]) { CK(cudaMemcpyAsync(out_ptr(out, i), dout + off[i], (size_t)N * elem[i], cudaMemcpyDeviceToHost, st)); d2h += (int64_t)N * elem[i]; }
    }
    if (stats) cudaEventRecord(e1, st);
    CK(cudaStreamSynchronize(st));
    if (staged) for (int i = 0; i < N_OUT; ++i) if (want[i]) memcpy(out_ptr(out, i), hout + off[i], (size_t)N * elem[i]);
    if (stats) {
        float ms[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        accumulate_lane_times(ln, ms);
        cudaEventElapsedTime(&stats->ms_total, e0, e1);
        store_times(stats, ms);
        stats->n_pairs = N; stats->n_launches = n_launches; stats->n_chunks = 1;
        stats->h2d_bytes = (int64_t)n_in * 8; stats->d2h_bytes = d2h;
        if (out->n_sol) { int64_t n = 0; for (int64_t i = 0; i < N; ++i) n += out->n_sol[i]; stats->n_solutions = n; }
    }
    ln.timed = false;
    return NRMC_OK;
}

extern "C" int nrmc_rt_trace(nrmc_rt_t h, const nrmc_rt_input *in, const nrmc_rt_output *out, void *stream, nrmc_rt_stats *stats)
{
    if (!h || !in || !out) return NRMC_ERR_INVALID_ARGUMENT;
    if (in->n_vertices < 0 || in->n_antennas < 0) return NRMC_ERR_INVALID_ARGUMENT;
    if (!in->outer && in->n_antennas != in->n_vertices) { h->err = "pair mode needs n_antennas == n_vertices"; return NRMC_ERR_INVALID_ARGUMENT; }
    const int64_t N = in->outer ? in->n_vertices * in->n_antennas : in->n_vertices;
    if (stats) memset(stats, 0, sizeof(*stats));
    if (N == 0) { if (out->compact && out->sol_offset) out->sol_offset[0] = 0; return NRMC_OK; }
    if (!in->vx || !in->vy || !in->vz || !in->ax || !in->ay || !in->az) return NRMC_ERR_INVALID_ARGUMENT;
    if (in->sx && (!in->sy || !in->sz || !(in->delta_C_cut >= 0.0))) { h->err = "the viewing-angle cut needs sx, sy, sz and delta_C_cut >= 0"; return NRMC_ERR_INVALID_ARGUMENT; }
    const bool want_att = out->attenuation_sparse || out->attenuation;
    if (want_att && h->ice.att_model == 0) { h->err = "attenuation requested but no attenuation model configured"; return NRMC_ERR_UNSUPPORTED; }
    if (want_att && !h->have_freq) { h->err = "attenuation requested before nrmc_rt_set_frequencies"; return NRMC_ERR_NO_FREQUENCIES; }
    const bool compact = out->compact != 0;
    if (compact && (!out->sol_offset || out->row_capacity < 0)) { h->err = "compact output needs sol_offset[N+1] and row_capacity"; return NRMC_ERR_INVALID_ARGUMENT; }
    CK(cudaSetDevice(h->cfg.device));
    const int S = h->S, K1 = h->K1, Fs = h->tb.Fs, F = h->tb.F;
    const size_t elem[N_OUT] = {4, 4, (size_t)S, (size_t)S, (size_t)S, 8u * S, 8u * S, 8u * S, 8u * S, 24u * S, 24u * S,
                                8u * S * K1, 8u * (size_t)S * Fs, 8u * (size_t)S * F, 8u * S};
    int n_launches = 0;

    if (in->memory == NRMC_MEMORY_DEVICE) {
        // device-resident: the caller's pointers are used directly; chunked only to bound the work-list scratch
        // (own scratch lane: a host-memory call that follows on the library's streams shares nothing with it; a second
        //  device-resident call, possibly on another stream, is ordered behind this one through dev_done)
        Lane &ln = h->lanes[DEV_LANE];
        cudaStream_t user = (cudaStream_t)stream;
        if (h->dev_pending) CK(cudaStreamWaitEvent(user, h->dev_done, 0));
        ln.timed = false;
        cudaEvent_t e0 = ln.ev[3], e1 = ln.ev[4];
        if (stats) cudaEventRecord(e0, user);
        int64_t chunk = std::min<int64_t>(N, h->chunk_pairs > 0 ? h->chunk_pairs : ((int64_t)1 << 24) / (1 + 2 * h->ice.n_refl));     // bounds the queue / work-list scratch
        if (in->outer && chunk < N) chunk = std::max<int64_t>(in->n_antennas, (chunk / in->n_antennas) * in->n_antennas);
        float ms[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        int rc = NRMC_OK, n_chunks = 0;
        unsigned long long *row_base_dev = (unsigned long long *)h->d_count.p + CNT_ROWBASE;
        if (compact) K_set_u64<<<1, 1, 0, user>>>(row_base_dev, (unsigned long long)out->row_base);   // first row of this call's solutions
        const bool small = !compact && h->chunk_pairs == 0 && N <= NRMC_SMALL_PAIRS && !getenv("NRMC_NO_SMALL_PATH");
        for (int64_t p0 = 0; p0 < N && rc == NRMC_OK; p0 += chunk, ++n_chunks) {
            const int64_t np = std::min(chunk, N - p0);
            const int64_t ps = compact ? 0 : p0 * S;        // first row of the chunk in the per-slot arrays (compact: rows are global)
            KInput kin;
            kin.outer = in->outer; kin.n_antennas = in->n_antennas; kin.n_pairs = np;
            kin.sx = kin.sy = kin.sz = nullptr; kin.delta_C_cut = in->delta_C_cut;
            if (in->sx) {
                const int64_t v0s = in->outer ? p0 / in->n_antennas : p0;
                kin.sx = in->sx + v0s; kin.sy = in->sy + v0s; kin.sz = in->sz + v0s;
            }
            if (in->outer) {
                // chunk boundaries must fall on whole vertices
                if (p0 % in->n_antennas != 0) { rc = NRMC_ERR_INVALID_ARGUMENT; break; }
                const int64_t v0 = p0 / in->n_antennas;
                kin.vx = in->vx + v0; kin.vy = in->vy + v0; kin.vz = in->vz + v0;
                kin.ax = in->ax; kin.ay = in->ay; kin.az = in->az;
            } else {
                kin.vx = in->vx + p0; kin.vy = in->vy + p0; kin.vz = in->vz + p0;
                kin.ax = in->ax + p0; kin.ay = in->ay + p0; kin.az = in->az + p0;
            }
            TraceOutputs to;
            to.n_sol = out->n_sol ? out->n_sol + p0 : nullptr;
            to.status = out->status ? out->status + p0 : nullptr;
            to.type = out->solution_type ? out->solution_type + ps : nullptr;
            to.reflection = out->reflection ? out->reflection + ps : nullptr;
            to.reflection_case = out->reflection_case ? out->reflection_case + ps : nullptr;
            to.C0 = out->C0 ? out->C0 + ps : nullptr;
            to.C1 = out->C1 ? out->C1 + ps : nullptr;
            to.path_length = out->path_length ? out->path_length + ps : nullptr;
            to.travel_time = out->travel_time ? out->travel_time + ps : nullptr;
            to.launch = out->launch_vector ? out->launch_vector + ps * 3 : nullptr;
            to.receive = out->receive_vector ? out->receive_vector + ps * 3 : nullptr;
            to.reflection_angle = out->reflection_angle ? out->reflection_angle + ps * K1 : nullptr;
            to.viewing_angle = out->viewing_angle ? out->viewing_angle + ps : nullptr;
            to.row_offset = nullptr;
            if ((want_att || compact) && !to.n_sol) {   // the fill kernel / the row scan need n_sol
                CK(ln.out.reserve((size_t)np * 4));
                to.n_sol = (int32_t *)ln.out.p;
            }
            ln.timed = (stats != nullptr);
            CompactCtx cc;
            if (compact) { cc.on = true; cc.sol_offset = out->sol_offset + p0; cc.base_dev = row_base_dev; cc.row_limit = out->row_base + out->row_capacity; }
            if (small)
                rc = launch_small(h, ln, DEV_LANE, user, kin, to, out->attenuation_sparse, out->attenuation, &n_launches);
            else
                rc = launch_chunk(h, ln, DEV_LANE, user, kin, to, out->attenuation_sparse ? out->attenuation_sparse + ps * Fs : nullptr,
                                  out->attenuation ? out->attenuation + ps * F : nullptr, &n_launches, cc);
            if (rc == NRMC_OK && stats) {
                accumulate_lane_times(ln, ms);
                if (want_att) {
                    unsigned long long cnt2[2] = {0, 0};
                    cudaMemcpy(cnt2, (unsigned long long *)h->d_count.p + CNT_STRIDE * DEV_LANE, sizeof(cnt2), cudaMemcpyDeviceToHost);
                    stats->n_solutions += (int64_t)(cnt2[0] + cnt2[WL_BACK]);
                }
            }
        }
        if (compact && rc == NRMC_OK)      // sol_offset[N] = number of rows
            CK(cudaMemcpyAsync(out->sol_offset + N, row_base_dev, sizeof(int64_t), cudaMemcpyDeviceToDevice, user));
        if (stats && rc == NRMC_OK) {
            cudaEventRecord(e1, user);
            CK(cudaEventSynchronize(e1));
            cudaEventElapsedTime(&stats->ms_total, e0, e1);
            store_times(stats, ms);
            stats->n_pairs = N; stats->n_launches = n_launches; stats->n_chunks = n_chunks;
            if (compact) {
                unsigned long long rows = 0;
                cudaMemcpy(&rows, row_base_dev, sizeof(rows), cudaMemcpyDeviceToHost);
                rows -= (unsigned long long)out->row_base;
                if ((int64_t)rows > out->row_capacity) { h->err = "compact output: row_capacity exceeded"; rc = NRMC_ERR_CAPACITY; }
                if (!want_att) stats->n_solutions = (int64_t)rows;
            }
        }
        cudaEventRecord(h->dev_done, user);
        h->dev_pending = true;
        ln.timed = false;
        return rc;
    }

    // ---------------- host memory, small batch in the padded layout: one staged copy each way, one cooperative launch --------
    if (!compact && h->chunk_pairs == 0 && N <= NRMC_SMALL_PAIRS && !getenv("NRMC_NO_SMALL_PATH")) return trace_small_host(h, in, out, stats, N, elem);
    // ---------------- host memory: chunked, two lanes (streams) so copies of one chunk overlap kernels of the other --------
    const int64_t na = in->n_antennas;
    size_t per_pair = 0;
    bool want[N_OUT];
    for (int i = 0; i < N_OUT; ++i) { want[i] = out_ptr(out, i) != nullptr; }
    const bool need_nsol_dev = want_att || want[0] || compact;
    for (int i = 0; i < N_OUT; ++i) if (want[i] || (i == 0 && need_nsol_dev)) per_pair += elem[i] + 16;
    const bool direct = compact;      // the binned solver assigns the compact rows itself (scan of n_sol before K_roots)
    per_pair += (size_t)h->S * 2 * WORKLIST_BYTES_PER_ENTRY + 48;
    per_pair += (size_t)(1 + 2 * h->ice.n_refl) * (ROOTQ_BYTES_PER_ENTRY + HUMPQ_BYTES_PER_ENTRY + 1);
    int64_t chunk = (int64_t)((size_t)1536 * 1024 * 1024 / per_pair);   // ~1.5 GB of device scratch per lane
    chunk = std::max<int64_t>(chunk, 1024);
    if (h->chunk_pairs > 0) chunk = h->chunk_pairs;
    if (in->outer) chunk = std::max<int64_t>(na, (chunk / na) * na);
    chunk = std::min(chunk, N);
    if (h->chunk_pairs <= 0 && N > chunk && N < 2 * chunk) { chunk = (N + 1) / 2; if (in->outer) chunk = ((chunk + na - 1) / na) * na; }
    const double *d_ax = nullptr, *d_ay = nullptr, *d_az = nullptr;
    int64_t h2d = 0, d2h = 0;
    if (in->outer) {
        CK(h->d_ant.reserve((size_t)na * 24));
        double *a = (double *)h->d_ant.p;
        CK(cudaMemcpyAsync(a, in->ax, na * 8, cudaMemcpyHostToDevice, h->lanes[0].stream));
        CK(cudaMemcpyAsync(a + na, in->ay, na * 8, cudaMemcpyHostToDevice, h->lanes[0].stream));
        CK(cudaMemcpyAsync(a + 2 * na, in->az, na * 8, cudaMemcpyHostToDevice, h->lanes[0].stream));
        CK(cudaStreamSynchronize(h->lanes[0].stream));
        d_ax = a; d_ay = a + na; d_az = a + 2 * na;
        h2d += na * 24;
    }
    cudaEvent_t e0 = h->lanes[0].ev[3], e1 = h->lanes[0].ev[4];
    cudaEventRecord(e0, h->lanes[0].stream);
    float ms[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int n_chunks = 0;
    int64_t n_solutions = 0, row_base = 0;
    for (int64_t p0 = 0; p0 < N; p0 += chunk, ++n_chunks) {
        const int lid = n_chunks & 1;
        Lane &ln = h->lanes[lid];
        const int64_t np = std::min(chunk, N - p0);
        if (n_chunks >= 2 && stats && ln.timed) accumulate_lane_times(ln, ms);   // the chunk that used this lane before
        // inputs
        KInput kin;
        kin.outer = in->outer; kin.n_antennas = na; kin.n_pairs = np;
        kin.sx = kin.sy = kin.sz = nullptr; kin.delta_C_cut = in->delta_C_cut;
        const int64_t nv = in->outer ? np / na : np, v0 = in->outer ? p0 / na : p0;
        const size_t in_bytes = (size_t)nv * 24 + (in->outer ? 0 : (size_t)np * 24) + (in->sx ? (size_t)nv * 24 : 0);
        CK(ln.in.reserve(in_bytes));
        double *din = (double *)ln.in.p;
        CK(cudaMemcpyAsync(din, in->vx + v0, nv * 8, cudaMemcpyHostToDevice, ln.stream));
        CK(cudaMemcpyAsync(din + nv, in->vy + v0, nv * 8, cudaMemcpyHostToDevice, ln.stream));
        CK(cudaMemcpyAsync(din + 2 * nv, in->vz + v0, nv * 8, cudaMemcpyHostToDevice, ln.stream));
        kin.vx = din; kin.vy = din + nv; kin.vz = din + 2 * nv;
        h2d += nv * 24;
        if (in->sx) {
            double *ds = din + 3 * nv + (in->outer ? 0 : 3 * np);
            CK(cudaMemcpyAsync(ds, in->sx + v0, nv * 8, cudaMemcpyHostToDevice, ln.stream));
            CK(cudaMemcpyAsync(ds + nv, in->sy + v0, nv * 8, cudaMemcpyHostToDevice, ln.stream));
            CK(cudaMemcpyAsync(ds + 2 * nv, in->sz + v0, nv * 8, cudaMemcpyHostToDevice, ln.stream));
            kin.sx = ds; kin.sy = ds + nv; kin.sz = ds + 2 * nv;
            h2d += nv * 24;
        }
        if (in->outer) { kin.ax = d_ax; kin.ay = d_ay; kin.az = d_az; }
        else {
            double *da = din + 3 * nv;
            CK(cudaMemcpyAsync(da, in->ax + p0, np * 8, cudaMemcpyHostToDevice, ln.stream));
            CK(cudaMemcpyAsync(da + np, in->ay + p0, np * 8, cudaMemcpyHostToDevice, ln.stream));
            CK(cudaMemcpyAsync(da + 2 * np, in->az + p0, np * 8, cudaMemcpyHostToDevice, ln.stream));
            kin.ax = da; kin.ay = da + np; kin.az = da + 2 * np;
            h2d += np * 24;
        }
        // outputs: one device block per lane, every array 16-byte aligned
        size_t off[N_OUT], total = 0;
        for (int i = 0; i < N_OUT; ++i) {
            off[i] = total;
            if (want[i] || (i == 0 && need_nsol_dev)) total += ((size_t)np * elem[i] + 15) & ~(size_t)15;
        }
        CK(ln.out.reserve(total));
        unsigned char *dout = (unsigned char *)ln.out.p;
        auto dp = [&](int i) -> void * { return (want[i] || (i == 0 && need_nsol_dev)) ? (void *)(dout + off[i]) : nullptr; };
        // per-slot arrays as the kernels address them: in direct compact mode rows are GLOBAL, the chunk's block holds the rows
        // [row_base, row_base + rows), so the base pointers are shifted back by row_base rows
        auto dps = [&](int i) -> void * {
            unsigned char *q = (unsigned char *)dp(i);
            return (q && direct) ? (void *)(q - (size_t)row_base * (elem[i] / S)) : (void *)q;
        };
        TraceOutputs to;
        to.n_sol = (int32_t *)dp(0); to.status = (int32_t *)dp(1); to.type = (int8_t *)dps(2); to.reflection = (int8_t *)dps(3);
        to.reflection_case = (int8_t *)dps(4); to.C0 = (double *)dps(5); to.C1 = (double *)dps(6); to.path_length = (double *)dps(7);
        to.travel_time = (double *)dps(8); to.launch = (double *)dps(9); to.receive = (double *)dps(10);
        to.reflection_angle = (double *)dps(11);
        to.viewing_angle = (double *)dps(14);
        to.row_offset = nullptr; to.row_limit = 0;
        ln.timed = (stats != nullptr);
        CompactCtx cc;
        if (direct) {
            CK(ln.pack_off.reserve((size_t)np * sizeof(int64_t)));
            cc.on = true; cc.sol_offset = (int64_t *)ln.pack_off.p; cc.base_host = row_base;
            cc.row_limit = std::min<int64_t>(out->row_capacity, row_base + np * S);
        }
        int rc = launch_chunk(h, ln, lid, ln.stream, kin, to, (double *)dps(12), (double *)dps(13), &n_launches, cc);
        if (rc != NRMC_OK) return rc;
        if (direct) {
            const int nblk = (int)((np + PACK_PAIRS - 1) / PACK_PAIRS);
            unsigned long long rows = 0;
            CK(cudaMemcpyAsync(&rows, (unsigned long long *)ln.pack_sums.p + nblk, sizeof(rows), cudaMemcpyDeviceToHost, ln.stream));
            CK(cudaStreamSynchronize(ln.stream));      // the row count sizes the copies; the other lane's copies keep the link busy meanwhile
            d2h += sizeof(rows);
            if (row_base + (int64_t)rows > out->row_capacity) {
                h->err = "compact output: row_capacity exceeded";
                for (int l = 0; l < 2; ++l) cudaStreamSynchronize(h->lanes[l].stream);
                return NRMC_ERR_CAPACITY;
            }
            for (int i = 0; i < 2; ++i) {
                if (!want[i]) continue;
                CK(cudaMemcpyAsync((unsigned char *)out_ptr(out, i) + (size_t)p0 * elem[i], dout + off[i], (size_t)np * elem[i], cudaMemcpyDeviceToHost, ln.stream));
                d2h += (size_t)np * elem[i];
            }
            CK(cudaMemcpyAsync(out->sol_offset + p0, ln.pack_off.p, (size_t)np * sizeof(int64_t), cudaMemcpyDeviceToHost, ln.stream));
            d2h += (size_t)np * sizeof(int64_t);
            for (int i = 2; i < N_OUT; ++i) {
                if (!want[i] || rows == 0) continue;
                const size_t rb = elem[i] / S;
                CK(cudaMemcpyAsync((unsigned char *)out_ptr(out, i) + (size_t)row_base * rb, dout + off[i], (size_t)rows * rb, cudaMemcpyDeviceToHost, ln.stream));
                d2h += (size_t)rows * rb;
            }
            row_base += (int64_t)rows;
        } else if (!compact) {
            for (int i = 0; i < N_OUT; ++i) {
                if (!want[i]) continue;
                const size_t bytes = (size_t)np * elem[i];
                CK(cudaMemcpyAsync((unsigned char *)out_ptr(out, i) + (size_t)p0 * elem[i], dout + off[i], bytes, cudaMemcpyDeviceToHost, ln.stream));
                d2h += bytes;
            }
        }
    }
    for (int l = 0; l < 2; ++l) {
        Lane &ln = h->lanes[l];
        CK(cudaStreamSynchronize(ln.stream));
        if (stats && ln.timed && l < n_chunks) accumulate_lane_times(ln, ms);
        ln.timed = false;
    }
    cudaEventRecord(e1, h->lanes[0].stream);
    CK(cudaEventSynchronize(e1));
    if (compact) out->sol_offset[N] = row_base;
    if (stats) {
        cudaEventElapsedTime(&stats->ms_total, e0, e1);
        store_times(stats, ms);
        stats->n_pairs = N; stats->n_launches = n_launches; stats->n_chunks = n_chunks;
        stats->h2d_bytes = h2d; stats->d2h_bytes = d2h;
        if (compact) stats->n_solutions = row_base;
        else if (out->n_sol) { for (int64_t i = 0; i < N; ++i) n_solutions += out->n_sol[i]; stats->n_solutions = n_solutions; }
    }
    return NRMC_OK;
}

extern "C" int nrmc_rt_apply_propagation_effects(nrmc_rt_t h, const nrmc_rt_effects *fx, void *stream)
{
    if (!h || !fx || fx->n_rows < 0 || fx->n_freq <= 0) return NRMC_ERR_INVALID_ARGUMENT;
    if (fx->n_rows == 0) return NRMC_OK;
    if (!fx->spectrum) return NRMC_ERR_INVALID_ARGUMENT;
    if (!fx->attenuation && fx->attenuation_sparse) {
        if (!h->have_freq) { h->err = "sparse attenuation needs nrmc_rt_set_frequencies"; return NRMC_ERR_NO_FREQUENCIES; }
        if (fx->n_freq != h->tb.F) { h->err = "n_freq differs from the frequency vector of nrmc_rt_set_frequencies"; return NRMC_ERR_INVALID_ARGUMENT; }
        if (h->ice.n_refl > 0) {
            h->err = "with bottom reflections the factors of the path segments are interpolated separately (py:1077-1086): pass the dense attenuation";
            return NRMC_ERR_UNSUPPORTED;
        }
    }
    CK(cudaSetDevice(h->cfg.device));
    const double n_surface = h->ice.n_ice - h->ice.dn * exp(-0.01 * h->ice.inv_z0);    // n(z = -1 cm), py:2990
    const int64_t blocks = std::min<int64_t>(fx->n_rows, (int64_t)h->n_sm * 16);
    K_apply_effects<<<(unsigned)blocks, FX_THREADS, 0, (cudaStream_t)stream>>>(*fx, h->tb, h->K1, n_surface);
    CK(cudaGetLastError());
    return NRMC_OK;
}

extern "C" int nrmc_rt_focusing_factor(nrmc_rt_t h, const nrmc_rt_input *in, const nrmc_rt_focusing *fo, void *stream)
{
    if (!h || !in || !fo) return NRMC_ERR_INVALID_ARGUMENT;
    if (in->memory != NRMC_MEMORY_DEVICE) { h->err = "nrmc_rt_focusing_factor takes device pointers"; return NRMC_ERR_INVALID_ARGUMENT; }
    if (in->n_vertices < 0 || in->n_antennas < 0 || (!in->outer && in->n_antennas != in->n_vertices)) return NRMC_ERR_INVALID_ARGUMENT;
    const int64_t N = in->outer ? in->n_vertices * in->n_antennas : in->n_vertices;
    if (N == 0) return NRMC_OK;
    if (!in->vx || !in->vy || !in->vz || !in->ax || !in->ay || !in->az || !fo->n_sol || !fo->C0 || !fo->path_length || !fo->focusing ||
        !(fo->limit > 0.0))
        return NRMC_ERR_INVALID_ARGUMENT;
    CK(cudaSetDevice(h->cfg.device));
    KInput kin;
    kin.vx = in->vx; kin.vy = in->vy; kin.vz = in->vz; kin.ax = in->ax; kin.ay = in->ay; kin.az = in->az;
    kin.n_pairs = N; kin.n_antennas = in->outer ? in->n_antennas : 1; kin.outer = in->outer;
    kin.sx = kin.sy = kin.sz = nullptr; kin.delta_C_cut = 0.0;
    K_focusing<<<(unsigned)((N + 127) / 128), 128, 0, (cudaStream_t)stream>>>(h->ice, kin, *fo, h->S);
    CK(cudaGetLastError());
    return NRMC_OK;
}

extern "C" int nrmc_rt_attenuation_length(nrmc_rt_t h, const double *z, const double *frequency, int64_t n, double *out_host)
{
    if (!h || !z || !frequency || !out_host || n < 0) return NRMC_ERR_INVALID_ARGUMENT;
    if (h->ice.att_model == 0) return NRMC_ERR_UNSUPPORTED;
    if (n == 0) return NRMC_OK;
    CK(cudaSetDevice(h->cfg.device));
    DevBuf b;
    CK(b.reserve((size_t)n * 24));
    double *dz = (double *)b.p, *df = dz + n, *dout = df + n;
    cudaStream_t st = h->lanes[0].stream;
    CK(cudaMemcpyAsync(dz, z, n * 8, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(df, frequency, n * 8, cudaMemcpyHostToDevice, st));
    K_att_length<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(h->ice, h->tb.gl3, dz, df, n, dout);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out_host, dout, n * 8, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    b.release();
    return NRMC_OK;
}

extern "C" int nrmc_rt_measure_fp64_peak(int32_t device, double seconds, double *tflops, double *sm_clock_mhz)
{
    if (!tflops) return NRMC_ERR_INVALID_ARGUMENT;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) return NRMC_ERR_NO_DEVICE;
    if (cudaSetDevice(device) != cudaSuccess) return NRMC_ERR_NO_DEVICE;
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, device);
    const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 4096;
    double *d = nullptr;
    if (cudaMalloc(&d, (size_t)blocks * threads * 8) != cudaSuccess) return NRMC_ERR_CUDA;
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    K_fp64_peak<<<blocks, threads>>>(d, 64);
    cudaDeviceSynchronize();
    double best = 0, elapsed = 0;
    while (elapsed < seconds * 1e3) {
        cudaEventRecord(a);
        K_fp64_peak<<<blocks, threads>>>(d, iters);
        cudaEventRecord(b);
        if (cudaEventSynchronize(b) != cudaSuccess) { cudaFree(d); return NRMC_ERR_CUDA; }
        float ms = 0;
        cudaEventElapsedTime(&ms, a, b);
        elapsed += ms;
        const double flops = 2.0 * 64.0 * iters * (double)blocks * threads;
        best = std::max(best, flops / (ms * 1e-3) / 1e12);
    }
    *tflops = best;
    if (sm_clock_mhz) *sm_clock_mhz = prop.clockRate / 1e3;
    cudaEventDestroy(a); cudaEventDestroy(b);
    cudaFree(d);
    return NRMC_OK;
}
