/*
 * nrmc_rt.h -- C ABI of the B200-native batched analytic ray tracer (libnrmc_rt.so).
 *
 * Drop-in boundary.  The reference's only native FFI for this path is the Cython shim
 *   NuRadioMC/SignalProp/CPPAnalyticRayTracing/wrapper.pyx:8-33
 *     find_solutions(x1, x2, n_ice, delta_n, z_0, reflection, reflection_case, ice_reflection) -> list[dict]
 *     get_attenuation_along_path(x1, x2, C0, frequency, n_ice, delta_n, z_0, model) -> float
 *     get_attenuation_length(z, frequency, model) -> float
 * over analytic_raytracing.cpp:877-896 (find_solutions2) / :365 (get_attenuation_along_path): scalar, one pair and
 * one frequency per call.  This header replaces that surface with a batched one: one call traces N (vertex, antenna)
 * pairs through everything the Python class needs for
 *   ray_tracing.set_start_and_end_point / find_solutions / get_solution_type / get_launch_vector /
 *   get_receive_vector / get_reflection_angle / get_path_length / get_travel_time / get_attenuation
 *   (NuRadioMC/SignalProp/analyticraytracing.py:2057-2146, 2560-2776).
 * Plain pointers and sizes only; the caller owns every buffer; status codes instead of exceptions.
 * Units are NuRadioMC's (NuRadioReco/utilities/units.py): metre, nanosecond, GHz, radian.
 */
#ifndef NRMC_RT_H
#define NRMC_RT_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NRMC_OK 0
#define NRMC_ERR_INVALID_ARGUMENT (-1)
#define NRMC_ERR_CUDA (-2)
#define NRMC_ERR_NO_DEVICE (-3)
#define NRMC_ERR_UNSUPPORTED (-4)
#define NRMC_ERR_NO_FREQUENCIES (-5)
#define NRMC_ERR_CAPACITY (-6)

/* attenuation model integers: NuRadioMC/utilities/attenuation.py:14 */
#define NRMC_ATT_NONE 0
#define NRMC_ATT_SP1 1
#define NRMC_ATT_GL1 2
#define NRMC_ATT_MB1 3
#define NRMC_ATT_GL2 4
#define NRMC_ATT_GL3 5

/* per-pair status bits (nrmc_rt_output.status) */
#define NRMC_PAIR_IN_AIR 1           /* a point lies above the surface: the reference finds no solution (py:1445-1448) */
#define NRMC_PAIR_BELOW_REFLECTOR 2  /* a point lies below the reflective layer: the reference raises AttributeError
                                        (propagation_base_class.py:156-161) */
#define NRMC_PAIR_NONFINITE 4        /* NaN / inf coordinate */

#define NRMC_MEMORY_HOST 0
#define NRMC_MEMORY_DEVICE 1

typedef struct nrmc_rt_s *nrmc_rt_t;

/* Replaces the (medium, attenuation_model, n_frequencies_integration, n_reflections) constructor arguments of
 * ray_tracing (analyticraytracing.py:1938-1941) and the IceModelSimple parameters (medium_base.py:206-252). */
typedef struct {
    double n_ice, delta_n, z_0;
    double reflection_z;               /* depth of the reflective bottom layer [m] (medium_base.py:47-66); NaN if none */
    int32_t attenuation_model;         /* NRMC_ATT_* */
    int32_t n_reflections;             /* 0..4 */
    int32_t n_frequencies_integration; /* propagation_base_class.py:118 default 100 */
    int32_t device;                    /* CUDA device ordinal */
    const double *gl3_table;           /* GL3 only: rows of (depth, slope, offset), attenuation.py:16-33 */
    int32_t gl3_rows;
    int32_t reserved;
} nrmc_rt_config;

/* N pairs.  outer == 0: pair i = (vertex i, antenna i), n_antennas == n_vertices.
 *           outer == 1: pair i = (vertex i / n_antennas, antenna i % n_antennas), N = n_vertices * n_antennas. */
typedef struct {
    int64_t n_vertices;
    const double *vx, *vy, *vz;        /* SoA start points (the reference's x1), metres */
    int64_t n_antennas;
    const double *ax, *ay, *az;        /* SoA end points (the reference's x2) */
    int32_t outer;
    int32_t memory;                    /* NRMC_MEMORY_HOST or NRMC_MEMORY_DEVICE: where ALL input and output pointers live */
    /* Optional viewing-angle cut of the simulation loop (NuRadioMC/simulation/simulation.py:175-208).  sx, sy, sz: SoA
     * [n_vertices] propagation direction of the shower at each vertex (-shower axis); NULL: no cut.  A solution whose
     * launch vector makes an angle with it that differs from the Cherenkov angle arccos(1/n(vertex)) by more than
     * delta_C_cut [rad] keeps its geometric outputs but gets NO attenuation (NaN rows): the reference skips it. */
    const double *sx, *sy, *sz;
    double delta_C_cut;
} nrmc_rt_input;

/* SoA outputs, pair-major, S = nrmc_rt_max_solutions() slots per pair ordered as the reference orders
 * ray_tracing._results (mode (reflection, case) first, then C0 ascending; py:2122-2125, :1547).  Unused slots hold
 * 0 / NaN.  Any pointer may be NULL (that output is skipped). */
typedef struct {
    int32_t *n_sol;            /* [N]      get_number_of_solutions()                                  */
    int32_t *status;           /* [N]      NRMC_PAIR_* bits                                            */
    int8_t *solution_type;     /* [N,S]    get_solution_type(iS): 1 direct, 2 refracted, 3 reflected   */
    int8_t *reflection;        /* [N,S]    get_results()[iS]['reflection']                             */
    int8_t *reflection_case;   /* [N,S]    get_results()[iS]['reflection_case']                        */
    double *C0, *C1;           /* [N,S]    get_results()[iS]['C0'|'C1']                                */
    double *path_length;       /* [N,S]    get_path_length(iS)   (analytic)                            */
    double *travel_time;       /* [N,S]    get_travel_time(iS)   (analytic)                            */
    double *launch_vector;     /* [N,S,3]  get_launch_vector(iS)                                       */
    double *receive_vector;    /* [N,S,3]  get_receive_vector(iS)                                      */
    double *reflection_angle;  /* [N,S,n_reflections+1]  get_reflection_angle(iS), NaN = None          */
    double *attenuation_sparse;/* [N,S,Fs] attenuation factors at the Fs integration frequencies       */
    double *attenuation;       /* [N,S,F]  get_attenuation(iS, frequency, max_detector_freq)           */
    double *viewing_angle;     /* [N,S]    angle(shower direction, launch vector) (simulation.py:191); NaN without sx/sy/sz */
    /* Compact (per-solution) layout.  compact != 0: every [N,S,...] array above is written as
     * [n_rows,...] with one row per EXISTING solution: the rows of pair i are sol_offset[i] .. sol_offset[i+1]-1, in slot
     * order (CSR; sol_offset[N] = n_rows = sum of n_sol).  Empty slots are neither stored nor copied over PCIe.
     * Device-resident calls: sol_offset is a device array too; rows beyond row_capacity are dropped and reported as
     * NRMC_ERR_CAPACITY when the call synchronises (stats != NULL).
     * row_capacity = rows the per-solution arrays can hold (N*S always suffices); NRMC_ERR_CAPACITY if exceeded. */
    int32_t compact;
    int32_t reserved;
    int64_t *sol_offset;       /* [N+1]    compact only (required then)                                */
    int64_t row_capacity;
    /* Device-resident compact calls only: row of this call's first solution.  The per-solution arrays are then addressed as
     * array[row_base + local row] and sol_offset holds those global rows -- the caller passes the BASE pointers of arrays that
     * several calls (or several GPUs, through peer mappings: nrmc_rt_peer_open) fill side by side, each in its own segment
     * [row_base, row_base + row_capacity).  0 for a stand-alone call. */
    int64_t row_base;
} nrmc_rt_output;

typedef struct {
    int64_t n_pairs, n_solutions;
    float ms_solve;            /* device time of the root-finding + properties kernel(s), CUDA events   */
    float ms_attenuation;      /* device time of the attenuation kernel(s)                              */
    float ms_total;            /* device time of the whole call on the library's stream (incl. copies)  */
    int32_t n_launches;        /* kernels launched by this call                                         */
    int32_t n_chunks;
    int64_t h2d_bytes, d2h_bytes;
    float ms_kernel[6];        /* device time per kernel: [0] K_classify, [1] K_hump (+ row scan / slot offsets), [2] K_roots,
                                  [3] main attenuation kernel, [4] fallback / dense-expansion kernels, [5] reserved      */
} nrmc_rt_stats;

int nrmc_rt_create(const nrmc_rt_config *cfg, nrmc_rt_t *out);
void nrmc_rt_destroy(nrmc_rt_t h);
const char *nrmc_rt_last_error(nrmc_rt_t h);
int nrmc_rt_max_solutions(nrmc_rt_t h);       /* 2 + 4 n_reflections, propagation_base_class.py:424-429 */

/* The frequency vector of get_attenuation(iS, frequency, max_detector_freq) (max_detector_freq = NaN for None).
 * Builds the sparse integration frequencies exactly as __get_frequencies_for_attenuation (py:885-931) and the
 * np.interp tables (py:1075-1078).  Returns Fs (>0) or an error code. */
int nrmc_rt_set_frequencies(nrmc_rt_t h, const double *frequency, int32_t n, double max_detector_freq);
int nrmc_rt_get_sparse_frequencies(nrmc_rt_t h, double *out, int32_t capacity);

/* Traces all pairs.  With NRMC_MEMORY_HOST the call is synchronous and copies inputs/outputs itself (chunked and
 * overlapped on two streams); with NRMC_MEMORY_DEVICE the work is enqueued on `stream` (a cudaStream_t, may be 0)
 * and the call returns after enqueueing unless `stats` is non-NULL (then it synchronises to fill it). */
int nrmc_rt_trace(nrmc_rt_t h, const nrmc_rt_input *in, const nrmc_rt_output *out, void *stream, nrmc_rt_stats *stats);

/* Propagation effects on electric-field spectra, the step after the ray trace
 * (ray_tracing.apply_propagation_effects, analyticraytracing.py:2937-3033, in-ice branch; Fresnel coefficients
 * NuRadioReco/utilities/geometryUtilities.py:211-263).  One row = one ray-tracing solution with a spectrum of 3 components
 * (eR, eTheta, ePhi) x F complex bins:
 *   all components          *= attenuation factor of the row (if given)
 *   eTheta, ePhi            *= r_p, r_s of every surface reflection of the path (reflection_angle[row, 0..K1-1], NaN = none)
 *                              with n_1 = n(z = -1 cm), n_2 = 1 (complex beyond total internal reflection)
 *   eTheta, ePhi            *= (reflection_coefficient * exp(i * phase))^k for k = reflection[row] bottom reflections.
 * Attenuation comes from `attenuation` ([n_rows, F], already on the spectrum's frequency grid) or, for single-segment
 * paths, is interpolated on the fly from `attenuation_sparse` ([n_rows, Fs]) with the tables of nrmc_rt_set_frequencies
 * (F must then equal the length of that frequency vector): the dense factors are never materialised.
 * All pointers are DEVICE pointers; the work is enqueued on `stream`.  Optional outputs r_theta / r_phi
 * ([n_rows] complex, interleaved re/im): the coefficients the reference stores on the field object (:2993-2994). */
typedef struct {
    int64_t n_rows;
    int32_t n_freq;                    /* F */
    int32_t reserved;
    double *spectrum;                  /* [n_rows, 3, F] complex128 (interleaved re, im), modified in place */
    const double *attenuation;         /* [n_rows, F] or NULL */
    const double *attenuation_sparse;  /* [n_rows, Fs] or NULL */
    const double *reflection_angle;    /* [n_rows, K1], K1 = n_reflections + 1; NULL: no surface reflection anywhere */
    const int8_t *reflection;          /* [n_rows] number of bottom reflections; NULL: none */
    double reflection_coefficient;     /* medium.reflection_coefficient (medium_base.py:61) */
    double reflection_phase_shift;     /* medium.reflection_phase_shift [rad] */
    double *r_theta, *r_phi;           /* [n_rows] complex, or NULL */
    const double *focusing;            /* [n_rows] focusing factor on eTheta, ePhi (py:3012-3015: spec[1:] *= focusing), or NULL */
} nrmc_rt_effects;
int nrmc_rt_apply_propagation_effects(nrmc_rt_t h, const nrmc_rt_effects *fx, void *stream);

/* Signal focusing of every solution (ray_tracing.get_focusing, analyticraytracing.py:2778-2888, numerical branch; the
 * reference's analytic branch raises AttributeError at :831 and is not reproduced).  The reference re-traces the pair with the
 * receiver moved by dz = -1 cm and differences the launch angle; here d(launch angle)/d(receiver depth) is the exact
 * derivative from the closed-form dR/dbeta (the dz -> 0 limit; 2e-3 from the 1 cm difference quotient, inside the 3e-3 noise of
 * the reference's own root finding):
 *   focusing = min(limit, sqrt(D / sin(rec) |d launch / dz|) sqrt(D sin(launch) / rho)) * sqrt(n_emitter / n_receiver).
 * `in` as in nrmc_rt_trace (memory must be NRMC_MEMORY_DEVICE); the solutions are read from a result of nrmc_rt_trace over the
 * same input: padded [N,S] arrays (sol_offset NULL) or compact rows.  The receiver is the antenna (the reference's X2). */
typedef struct {
    const int32_t *n_sol;              /* [N] */
    const double *C0;                  /* [N,S] or [n_rows] */
    const int8_t *reflection;          /* [N,S] or [n_rows]; NULL: no bottom reflections */
    const int8_t *reflection_case;     /* [N,S] or [n_rows]; NULL: case 1 */
    const double *path_length;         /* [N,S] or [n_rows] */
    const int64_t *sol_offset;         /* [N+1] compact layout, or NULL */
    double limit;                      /* config['propagation']['focusing_limit'] (default 2) */
    double *focusing;                  /* out: [N,S] (NaN in empty slots) or [n_rows] */
} nrmc_rt_focusing;
int nrmc_rt_focusing_factor(nrmc_rt_t h, const nrmc_rt_input *in, const nrmc_rt_focusing *fo, void *stream);

/* Pairs per internal chunk (device scratch and the host-call pipeline are sized per chunk).  0 = automatic (2^24 pairs for
 * device-resident calls, ~1.5 GB of scratch per stream for host calls).  A tuning / testing knob: results do not depend on it. */
int nrmc_rt_set_chunk_pairs(nrmc_rt_t h, int64_t pairs);

/* Result gather over NVLink peer memory (replaces the reference's file-level merge of per-job outputs,
 * NuRadioMC/utilities/merge_hdf5.py / runner.py:42-99, for one process per GPU on an NVSwitch node).  The gathering rank allocates the
 * arrays with nrmc_rt_peer_alloc and hands the 64-byte handles to the other processes (any transport: torch.distributed, MPI, a
 * file); they map them with nrmc_rt_peer_open and pass the mapped BASE pointers, their own row_base and their own slice of the
 * per-pair arrays as the outputs of a device-resident compact nrmc_rt_trace: the kernels then store every result row straight
 * into the gathering GPU's HBM while they compute (no staging copy, no host involvement, no collective call).  The writer
 * synchronises its stream, then any inter-process barrier makes the rows visible to the owner. */
#define NRMC_PEER_HANDLE_BYTES 64
int nrmc_rt_peer_alloc(int32_t device, uint64_t bytes, void **dev_ptr, unsigned char *handle /* [64] out */);
int nrmc_rt_peer_open(int32_t device, const unsigned char *handle /* [64] */, void **dev_ptr);
int nrmc_rt_peer_close(void *dev_ptr);
int nrmc_rt_peer_free(void *dev_ptr);
/* cudaMemcpyAsync between any two device / peer-mapped / pinned pointers on `stream` (the per-pair arrays of a gather) */
int nrmc_rt_copy_async(void *dst, const void *src, uint64_t bytes, void *stream);

/* pinned host memory for fast NRMC_MEMORY_HOST transfers */
int nrmc_rt_host_alloc(void **ptr, uint64_t bytes);
int nrmc_rt_host_free(void *ptr);

/* attenuation length L(z, f) [m] on the device for n points (replaces wrapper.pyx:32-33 get_attenuation_length) */
int nrmc_rt_attenuation_length(nrmc_rt_t h, const double *z, const double *frequency, int64_t n, double *out_host);

/* measured FP64 FMA peak of the device [TFLOP/s] (independent DFMA chains on all SMs for >= `seconds`) */
int nrmc_rt_measure_fp64_peak(int32_t device, double seconds, double *tflops, double *sm_clock_mhz);

int nrmc_rt_device_count(void);
const char *nrmc_rt_version(void);

#ifdef __cplusplus
}
#endif
#endif /* NRMC_RT_H */
