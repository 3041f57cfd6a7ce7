/* cf_b200.h -- C ABI of the B200 Monte-Carlo path engine.
 *
 * Drop-in boundary for the data-parallel hot path of asavine/CompFinance: the six template
 * algorithms of mcBase.h (mcSimul :267, mcParallelSimul :314, mcSimulAAD :429,
 * mcParallelSimulAAD :566, mcSimulAADMulti :776, mcParallelSimulAADMulti :859) together with the
 * RNG / model / product virtuals they call.  Everything that is path-INDEPENDENT (timelines,
 * Model<T>::init tables, parameter chain rule = the part of the tape before tape.mark(),
 * mcBase.h:559) stays on the host (compfinance_b200/host/ *.h mirrors the reference classes and
 * flattens them into the POD structs below); everything O(paths) runs on the GPU.
 *
 * Conventions: plain pointers and sizes, caller owns every buffer, all functions return 0 on
 * success and non-zero on failure with the message available from cf_last_error() (the C++ shim
 * rethrows it as std::runtime_error, the reference's error convention, mcBase.h:273).
 * There is NO CPU fallback: every entry point fails if no CUDA device is usable.
 */
#ifndef CF_B200_H
#define CF_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------------------------------
 * Random number generators (RNG interface, mcBase.h:228-246)
 * ---------------------------------------------------------------------------------------- */
enum { CF_RNG_SOBOL = 0,     /* sobol.h:31  (Joe-Kuo old 1111 direction numbers, Gray code) */
       CF_RNG_MRG32K3A = 1   /* mrg32k3a.h:23 (L'Ecuyer, antithetic pairing) */ };

typedef struct cf_rng {
    int32_t  kind;
    uint32_t seed1;   /* mrg32k3a a (default 12345), ignored by Sobol */
    uint32_t seed2;   /* mrg32k3a b (default 12346) */
} cf_rng;

/* ------------------------------------------------------------------------------------------
 * Models: flat image of what Model<T>::allocate/init leave behind (the "tables")
 * ---------------------------------------------------------------------------------------- */
enum { CF_MODEL_BS = 0,        /* mcMdlBS.h:25      */
       CF_MODEL_DUPIRE = 1,    /* mcMdlDupire.h:28  */
       CF_MODEL_DISPLACED = 2  /* mcMdlMultiDisplaced.h */ };

typedef struct cf_model {
    int32_t kind;
    int32_t n_assets;      /* 1 for BS / Dupire */
    int32_t n_steps;       /* time steps on the simulation timeline; simDim = n_steps * n_assets */
    int32_t n_events;      /* size of the product timeline (samples per path) */
    /* [n_steps + 1] 1 when simulation-timeline point i is an event date (Dupire myCommonSteps,
     * mcMdlDupire.h:179-184; BS: point 0 = myTodayOnTimeline, all others 1). */
    const uint8_t* is_event;

    double spot;           /* single-asset S0 (parameter leaf) */

    /* Black-Scholes tables, mcMdlBS.h:203-276 */
    const double* bs_drifts;       /* [n_steps]  (mu -/+ vol^2/2) dt                 */
    const double* bs_stds;         /* [n_steps]  vol sqrt(dt)                          */
    /* Per event date, what the product's defline asks for (first forward / first discount only:
     * every single-asset product of mcPrd.h in scope reads forwards[0][0], discounts[0], numeraire).
     * NULL = the Sample defaults (numeraire 1, discount 1, forward factor 1; mcBase.h:91-99). */
    const double* numeraires;      /* [n_events] */
    const double* fwd_factors;     /* [n_events] */
    const double* discounts;       /* [n_events] */
    const double* libors;          /* [n_events] first libor of each event date (ContingentBond: libor(T_e, T_e+1),
                                      mcMdlBS.h:262-276); NULL = the Sample default 0 */

    /* Dupire tables, mcMdlDupire.h:195-217 */
    int32_t       n_knots;         /* spot knots of the local-vol surface */
    const double* log_spots;       /* [n_knots] */
    const double* interp_vols;     /* [n_steps][n_knots]  sqrt(dt_i) * vol(spot_j, t_i) */
    /* Optional: Dupire::init() as a linear map (mcMdlDupire.h:202-216: time interpolation x sqrt(dt))
     *   interp_vols[i][j] = time_w1[i] * vols[j][time_col1[i]] + time_w2[i] * vols[j][time_col2[i]]
     * When all four arrays are given (n_times > 0) cf_run_aad folds this map into the accumulation and
     * returns the adjoints of vols ([n_knots][n_times], the parameter order of mcMdlDupire.h:116-121)
     * instead of the adjoints of interp_vols.  The host mirror derives the map from its own tape of init(). */
    int32_t        n_times;
    const int32_t* time_col1;      /* [n_steps] */
    const int32_t* time_col2;      /* [n_steps] (may equal time_col1 with weight 0) */
    const double*  time_w1;        /* [n_steps] */
    const double*  time_w2;        /* [n_steps] */

    /* Multi-asset displaced model tables, mcMdlMultiDisplaced.h:474-606 */
    const double*  dlm_spots;      /* [n_assets] */
    const double*  dlm_chol;       /* [n_assets][n_assets] lower */
    const double*  dlm_alphas;     /* [n_assets] */
    const int32_t* dlm_dynamics;   /* [n_assets] 0 lognormal 1 normal 2 surnormal 3 subnormal */
    const double*  dlm_dyn_fwd;    /* [n_steps][n_assets] */
    const double*  dlm_drifts;     /* [n_steps][n_assets] */
    const double*  dlm_stds;       /* [n_steps][n_assets] */
    const double*  dlm_fwd_factors;/* [n_events][n_assets] first forward maturity per asset */
} cf_model;

/* ------------------------------------------------------------------------------------------
 * Products: constants of Product<T>::payoffs
 * ---------------------------------------------------------------------------------------- */
enum { CF_PRODUCT_EUROPEAN = 0,   /* mcPrd.h:29  */
       CF_PRODUCT_UOC = 1,        /* mcPrd.h:128 */
       CF_PRODUCT_EUROPEANS = 2,  /* mcPrd.h:290 */
       CF_PRODUCT_BASKETS = 3,    /* mcPrdMulti.h:181 */
       CF_PRODUCT_AUTOCALL = 4,   /* mcPrdMulti.h:289 */
       CF_PRODUCT_MULTISTATS = 5, /* mcPrdMulti.h:11  */
       CF_PRODUCT_CONTINGENT = 6  /* mcPrd.h:404 ContingentBond (Black-Scholes only) */ };

typedef struct cf_product {
    int32_t kind;
    int32_t n_events;
    int32_t n_payoffs;
    int32_t is_put;          /* UOC myCallPut */
    double  strike;          /* European / UOC / Autocall */
    double  barrier;         /* UOC barrier, Autocall KO */
    double  smooth;          /* UOC: ABSOLUTE half-width double(S(t0) * smoothFactor), mcPrd.h:247;
                                Autocall: max(smooth, EPS), mcPrdMulti.h:318;
                                ContingentBond: ABSOLUTE half-width double(S(t0) * smoothFactor), mcPrd.h:531 */
    double  coupon;          /* Autocall; ContingentBond */
    /* Europeans: strikes of event e are strikes[strike_offsets[e] .. strike_offsets[e+1]) */
    const int32_t* strike_offsets;  /* [n_events + 1] */
    const double*  strikes;         /* Europeans / Baskets */
    const double*  weights;         /* Baskets weights [n_assets]; Autocall refs [n_assets] */
    const double*  event_dt;        /* Autocall: accrual period per event [n_events];
                                       ContingentBond: coverage of the period STARTING at event e, [n_events - 1] */
} cf_product;

/* ------------------------------------------------------------------------------------------
 * Context
 * ---------------------------------------------------------------------------------------- */
/* Open the context on the CUDA devices device_ids[0 .. n_devices) (1 to 16, distinct).
 * The reference's entry points run over a process-wide pool of worker threads started once
 * (ThreadPool::getInstance()->start, xlExport.cpp:1605); here the workers are GPUs.  With one device the calling
 * thread drives it.  With several (they must have peer access: NVLink) every run through cf_run_* and cf_plan_run_*
 * is sharded over them in disjoint skip-ahead blocks of paths (cf_shard_range), one host thread per device issuing
 * the launches at the same time, and the sum over devices is part of the final reduction kernels (peer memory);
 * results are identical on every device and read from the first.  Without cf_init the first call opens a context
 * on the current CUDA device. */
int cf_init(int n_devices, const int* device_ids);
int cf_shutdown(void);
int cf_device_count(void);   /* devices of the open context, 0 before the first call */
/* Bumped whenever the context is closed (cf_init again, cf_shutdown): plans created before are dead (their calls fail,
 * cf_plan_destroy stays valid); callers that cache plans compare this number. */
int cf_context_generation(void);
const char* cf_last_error(void);
/* Number of kernels launched by this library since cf_init (for bench accounting). */
uint64_t cf_launch_count(void);

/* Number of doubles in the table-adjoint vector of (model, product):
 *   BS      : 1 (spot) + n_steps (drifts) + n_steps (stds) + 4 * n_events (numeraire, fwd factor, discount, libor)
 *   Dupire  : 1 (spot) + n_steps * n_knots (interp_vols, step-major), or, when the time map is
 *             given, 1 (spot) + n_knots * n_times (vols, spot-major)
 *   Displaced: spots [A] | alphas [A] | chol [A][A] (lower) | dyn_fwd [D][A] | drifts [D][A] | stds [D][A] |
 *             numeraires [E] | fwd_factors [E][A]      (A assets, D steps, E events; cf_dlm.cuh) */
size_t cf_table_adjoint_size(const cf_model* mdl, const cf_product* prd);

/* ------------------------------------------------------------------------------------------
 * One-shot runs with HOST buffers (tables are uploaded, results downloaded inside the call)
 * ---------------------------------------------------------------------------------------- */
/* Replaces mcSimul / mcParallelSimul (mcBase.h:267, 314) for paths [first_path, first_path+n_paths).
 *   payoff_sums     [n_payoffs]           sum over paths of each payoff (main.h:69-74 divides by N)
 *   per_path_payoffs[n_paths][n_payoffs]  optional (NULL to skip): the reference's result matrix */
int cf_run_value(const cf_model* mdl, const cf_product* prd, const cf_rng* rng,
                 uint64_t first_path, uint64_t n_paths,
                 double* payoff_sums, double* per_path_payoffs);

/* Replaces mcSimulAAD / mcParallelSimulAAD (mcBase.h:429, 566) with the aggregator
 * sum_k payoff_weights[k] * payoff[k] (main.h:135 and main.h:210-213 are both of this form).
 *   payoff_sums     [n_payoffs]
 *   agg_sum         [1]              sum over paths of the aggregate
 *   table_adjoints  [cf_table_adjoint_size]  sum over paths of d aggregate / d table (NOT divided by N)
 *   per_path_payoffs / per_path_agg optional */
int cf_run_aad(const cf_model* mdl, const cf_product* prd, const cf_rng* rng,
               uint64_t first_path, uint64_t n_paths, const double* payoff_weights,
               double* payoff_sums, double* agg_sum, double* table_adjoints,
               double* per_path_payoffs, double* per_path_agg);

/* Replaces mcSimulAADMulti / mcParallelSimulAADMulti (mcBase.h:776, 859): one adjoint vector per payoff.
 *   payoff_sums [n_payoffs]
 *   risk_tables [cf_table_adjoint_size][n_payoffs]  sum over paths of d payoff[k] / d table (NOT divided by N),
 *               table-major like the reference's matrix risks(nParam, nPay) (mcBase.h:764-770)
 * Dupire (with the time map) x Europeans runs one sweep per maturity accumulated by strike class
 * (cf_multi.cuh; fixed-point integer accumulation: bit-reproducible); other pairs run one aggregate sweep per payoff. */
int cf_run_aad_multi(const cf_model* mdl, const cf_product* prd, const cf_rng* rng,
                     uint64_t first_path, uint64_t n_paths, double* payoff_sums, double* risk_tables);

/* ------------------------------------------------------------------------------------------
 * Resident plans: tables stay in HBM, results stay on the device (multi-GPU: the caller
 * all-reduces d_out over NCCL and downloads once)
 * ---------------------------------------------------------------------------------------- */
typedef struct cf_plan cf_plan;
int  cf_plan_create(const cf_model* mdl, const cf_product* prd, const cf_rng* rng, cf_plan** out);
void cf_plan_destroy(cf_plan* plan);
/* d_out (device): [n_payoffs] payoff sums.  stream: cudaStream_t (NULL = default stream). */
int cf_plan_launch_value(cf_plan* plan, uint64_t first_path, uint64_t n_paths, double* d_out, void* stream);
/* d_out (device): [n_payoffs] payoff sums, [1] aggregate sum, [table_adjoint_size] adjoints. */
int cf_plan_launch_aad(cf_plan* plan, const double* payoff_weights, uint64_t first_path, uint64_t n_paths,
                       double* d_out, void* stream);
size_t cf_plan_out_size(const cf_plan* plan, int aad);   /* doubles in d_out */
/* Average duration in ms of the dominant (path) kernel over the launches since the last call
 * (CUDA events recorded on the launch stream), and the number of launches averaged. */
int cf_plan_kernel_ms(cf_plan* plan, double* avg_ms, int* n_launches);
/* Diagnostics (CF_DEBUG_TIMES=1 in the environment): globaltimer stamps in ns of the phases of the last Dupire fast-path
 * launch, out[3][1024][8]: [0] forward, [1] reverse (per block: entry, tables staged, dependency met, live paths
 * compacted, sweep done, block sum, tables combined), [2][block][0] = live paths of the block. */
int cf_plan_debug_times(cf_plan* plan, unsigned long long* out);

/* ------------------------------------------------------------------------------------------
 * RNG kernels exposed for bit-exact parity tests (Sobol::next/skipTo sobol.h:77-151,
 * mrg32k3a::nextNumber/skipTo mrg32k3a.h:55-81, 192-238, invNormalCdf gaussians.h:47-87)
 * ---------------------------------------------------------------------------------------- */
/* Sobol integer states of paths [first_path, first_path+n_paths): out[n_paths][dim] uint32 */
int cf_sobol_states(int dim, uint64_t first_path, uint64_t n_paths, uint32_t* out);
/* Direction number jkDir[bit][dim] regenerated from the Joe-Kuo initialisers (host, no GPU) */
uint32_t cf_sobol_direction_number(int bit, int dim);
int cf_sobol_max_dim(void);
/* Uniforms (gaussian = 0) or Gaussians (gaussian = 1): out[n_paths][dim] doubles */
int cf_rng_draw(const cf_rng* rng, int dim, uint64_t first_path, uint64_t n_paths, int gaussian, double* out);
/* mrg32k3a integer numerators (x - y mod m1) of the stream positions used by the paths:
 * out[n_paths][dim] uint32 (odd paths repeat their even partner, as the reference caches them) */
int cf_mrg_numerators(const cf_rng* rng, int dim, uint64_t first_path, uint64_t n_paths, uint32_t* out);
int cf_inv_normal(const double* p, double* out, uint64_t n);
/* Device self-test: the number of 32-bit numerators z for which the path kernels' lean quotient z / (m1 + 1)
 * differs from the IEEE division (must be 0: the uniforms of mrg32k3a are bit-exact) */
int cf_selftest_mrg_uniform(uint64_t* mismatches);


/* Host-buffer runs of a resident plan (same outputs as cf_run_value / cf_run_aad without re-uploading the tables);
 * sharded over the devices of the context / the processes of the communicator like the one-shot runs. */
int cf_plan_run_value(cf_plan* plan, uint64_t first_path, uint64_t n_paths, double* payoff_sums, double* per_path_payoffs);
int cf_plan_run_aad(cf_plan* plan, const double* payoff_weights, uint64_t first_path, uint64_t n_paths,
                    double* payoff_sums, double* agg_sum, double* table_adjoints,
                    double* per_path_payoffs, double* per_path_agg);
int cf_plan_run_aad_multi(cf_plan* plan, uint64_t first_path, uint64_t n_paths, double* payoff_sums, double* risk_tables);
/* Device time in ms (CUDA events on the launch stream) of the path kernels of the last host-buffer run on the first
 * device of the context: the dominant kernel of a call made through the host API, for bench accounting. */
double cf_last_run_kernel_ms(void);

/* ------------------------------------------------------------------------------------------
 * Multi-GPU with one process per GPU.  mcBase.h has no counterpart: its workers share one address space and add
 * their risks in a loop (mcBase.h:737-746, multi: 976-984).  The participants' sum of the result vector -- the
 * path's only exchange step -- is done over peer memory inside the final reduction kernels (cf_comm.cuh: every double
 * travels as two {half, epoch} words, receivers poll their own memory), not by a collective call:
 *   1. every process: cf_init(1, {its device}); cf_comm_create(world, rank, capacity, handle) allocates its receive
 *      block (2 * world * capacity slots of 16 bytes) and returns a CUDA IPC handle of CF_COMM_HANDLE_BYTES bytes;
 *   2. the launcher gathers the handles in rank order (any transport: torch.distributed, MPI, a file);
 *   3. every process: cf_comm_connect(all handles).
 * From then on every run sums over the participants: cf_run_* / cf_plan_run_* take the WHOLE path range and run this
 * process's shard of it (cf_shard_range); cf_plan_launch_* run the caller's own range.  Every participant must issue
 * the same sequence of runs; all end with bit-identical sums.  capacity >= the longest result vector
 * (cf_plan_out_size; n_payoffs * (1 + cf_table_adjoint_size) for cf_run_aad_multi).
 * cf_comm_enable(0) keeps the communicator but makes launches local again (the caller reduces, e.g. with NCCL).
 * A participant that never shows up makes the others return NaN after ~10 s and cf_comm_status() non-zero
 * (the next call fails with that message) instead of hanging.
 * ---------------------------------------------------------------------------------------- */
#define CF_COMM_HANDLE_BYTES 64
int cf_comm_create(int world, int rank, size_t capacity_doubles, void* handle_out);
int cf_comm_connect(const void* handles /* [world][CF_COMM_HANDLE_BYTES] */);
int cf_comm_enable(int on);
int cf_comm_destroy(void);
int cf_comm_info(int* world, int* rank, int* enabled);
int cf_comm_status(void);
/* Shard of participant `rank` of `world` over n_paths paths: boundaries on multiples of 256 paths (one Sobol window)
 * and even (an antithetic pair of mrg32k3a is never split, as the reference's 64-path batches guarantee,
 * mcBase.h:312); the last participant takes the remainder. */
int cf_shard_range(uint64_t n_paths, int rank, int world, uint64_t* first, uint64_t* count);

/* ------------------------------------------------------------------------------------------
 * Measurement helpers (bench.py): scalar FP64 DFMA peak of the bound device in TFLOP/s
 * (2 flops per DFMA, 8 independent chains per thread, best of 4 timed launches) -- the roofline
 * denominator of this FP64-pipe-bound path; and the SM count.
 * ---------------------------------------------------------------------------------------- */
int cf_measure_fp64_peak(double* tflops, double* ms);
int cf_device_sm_count(void);

#ifdef __cplusplus
}
#endif
#endif /* CF_B200_H */
