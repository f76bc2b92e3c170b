/*
 * b200pf.h — C ABI of the B200-native batched power-flow engine (libb200pf.so).
 *
 * This is the drop-in boundary for ONE path of Grid2Op: what
 * grid2op.Backend.PandaPowerBackend delegates to pandapower in
 *   runpf()                     reference grid2op/Backend/pandaPowerBackend.py:1220-1255
 *     -> pp.runpp(init="dc", max_iteration=10, check_connectivity=False)   pPB:1097-1105
 *     -> pp.rundcpp(check_connectivity=True)                               pPB:1090
 *   _fetch_data_pf_converged()  pPB:1122-1218 (+ _gens_info :1526, _loads_info :1549,
 *                               _storages_info :1621, shunt_info :1596)
 * for a *batch* of independent grid instances that share one static grid description.
 *
 * Plain C: pointers and sizes only, no torch types.  Every function returns 0 on success or a
 * negative B200PF_E_* code; b200pf_last_error() gives the message.  A handle is not thread safe;
 * different handles are independent (one CUDA stream each).  No CPU fallback exists: with no
 * CUDA device b200pf_create() fails with B200PF_E_CUDA.
 * Ordering: every call that takes HOST data which later launches read (grid description, series bind / set_topo /
 * reset_instances, static injections, thermal limits) copies it on the handle's stream and returns when the copy has
 * arrived (the caller may reuse its buffer at once, and launches on any stream of the handle — chunk and group launches
 * run on internal streams — see the data); nothing is issued on the legacy default stream.
 *
 * Data layout (all row-major, instance-major so that one warp/CTA reads one contiguous record):
 *   topo   int8  [batch][n_topo_in]   n_topo_in = dim_topo + n_shunt + n_hidden
 *            [0, dim_topo)            local busbar of every element, grid2op topo_vect order,
 *                                     -1 = disconnected, 1..n_busbar        (pPB:1489-1524)
 *            [dim_topo, +n_shunt)     local busbar of every shunt (-1 = off) (pPB:1609-1611)
 *            [.., +n_hidden)          local busbar of every hidden reference unit (ext_grid kept
 *                                     next to the generator the reference creates, pPB:394-453)
 *   inj    f64   [batch][n_inj]       n_inj = n_gen + n_unit + 2 n_load + n_storage + 2 n_shunt
 *            gen_p[n_gen] MW | unit_vm[n_unit] p.u. (hidden units first) | load_p | load_q |
 *            storage_p | shunt_p | shunt_q        (what pPB:926-969 writes into its tables)
 *   out    f32   [batch][n_out]       n_out = 10 n_line + 4 n_unit + 2 n_load + n_storage + 3 n_shunt
 *            p_or q_or v_or a_or theta_or p_ex q_ex v_ex a_ex theta_ex   (each n_line)
 *            unit_p unit_q unit_v unit_theta (each n_unit, hidden units first)
 *            load_v load_theta (n_load) | storage_v (n_storage) | shunt_p shunt_q shunt_v (n_shunt)
 *            units: MW, MVAr, kV, A, degrees — the float32 values PandaPowerBackend returns.
 *   status int32 [batch]              B200PF_ST_*
 *   iters  int32 [batch]              Newton iterations used (0 in DC mode)
 *   busv   f64   [batch][2*n_slot]    optional: |V| p.u. then angle rad per bus slot
 *                                     (slot = sub + (busbar-1)*n_sub, NaN when unused)
 */
#ifndef B200PF_H
#define B200PF_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200PF_ABI_VERSION 2

enum {
    B200PF_OK = 0,
    B200PF_E_ARG = -1,      /* bad argument */
    B200PF_E_CUDA = -2,     /* CUDA runtime error / no device */
    B200PF_E_CAPACITY = -3, /* grid does not fit the on-chip budget of this build */
    B200PF_E_STATE = -4     /* call sequence error */
};

/* per-instance solver status (what runpf() turns into (bool, exception), pPB:1248-1255) */
enum {
    B200PF_ST_CONVERGED = 0,
    B200PF_ST_DIVERGED = 1,   /* Newton did not reach the tolerance / singular matrix / NaN */
    B200PF_ST_UNSUPPLIED = 2, /* an in-service bus is not connected to a reference bus */
    B200PF_ST_NO_REF = 3,     /* no reference unit in service */
    B200PF_ST_TOO_LARGE = 4,  /* more active buses than the launch was sized for (nb_cap) */
    B200PF_ST_DONE = 5        /* series mode with protections: the instance diverged at an earlier step */
};

/* Static grid description (host pointers, copied at create).  Element order = the reference's
 * (lines then trafos = "lines"; pPB:481-483).  Sub ids are substation indices 0..n_sub-1. */
typedef struct b200pf_grid_desc {
    int32_t abi_version; /* B200PF_ABI_VERSION */
    int32_t n_sub, n_busbar;
    int32_t n_line, n_gen, n_hidden, n_load, n_storage, n_shunt, dim_topo;
    double sn_mva;
    /* lines (n_line) */
    const int32_t *line_or_sub, *line_ex_sub, *line_or_pos, *line_ex_pos;
    const double *line_y;     /* [n_line][8]: yff.re yff.im yft.re yft.im ytf.re ytf.im ytt.re ytt.im (p.u.) */
    const double *line_bdc;   /* [n_line] 1/(x*ratio) */
    const double *line_pshift;/* [n_line] DC phase-shift injection  b*(-shift_rad) */
    const float *line_or_vn, *line_ex_vn; /* kV (float32 like pPB:796-801) */
    /* units (n_unit = n_hidden + n_gen; hidden reference units first, like the ppc gen table) */
    const int32_t *unit_sub;  /* [n_unit] */
    const int32_t *unit_pos;  /* [n_unit] topo_vect position; hidden unit h: dim_topo + n_shunt + h */
    const int32_t *unit_is_ref; /* [n_unit] ext_grid or gen.slack */
    const double *unit_qmin, *unit_qmax; /* [n_unit] MVAr (reactive sharing, PYPOWER pfsoln) */
    const float *unit_vn;     /* [n_unit] kV */
    /* loads / storages / shunts */
    const int32_t *load_sub, *load_pos;        const float *load_vn;
    const int32_t *storage_sub, *storage_pos;  const float *storage_vn;  const double *storage_q;
    const int32_t *shunt_sub;                  const float *shunt_vn;    const double *shunt_vratio; /* (vn_bus/vn_shunt)^2*step */
    /* [n_sub] position of every substation in a bandwidth-reducing order of the substation graph (e.g.
     * reverse Cuthill-McKee; identity is valid): the CTA-per-instance kernel numbers buses in this order so
     * that the Jacobian is banded and solves it with a band-limited LU.  NULL = identity. */
    const int32_t *sub_rank;
} b200pf_grid_desc;

typedef struct b200pf_handle b200pf_handle;

const char *b200pf_last_error(void);
int b200pf_abi_version(void);
int b200pf_device_count(void);

/* create / destroy.  max_batch sizes the device and pinned staging buffers. */
int b200pf_create(const b200pf_grid_desc *grid, int max_batch, int device, b200pf_handle **out);
int b200pf_destroy(b200pf_handle *h);

/* Active buses (bus slots with at least one connected element) of n topology records int8 [n][n_topo_in]: per_instance (may be
 * NULL) int32 [n], *max_out their maximum = the tight nb_cap of a launch on these records (the bound the callers of the reference's
 * backend never need because pandapower re-derives the bus set per call, pPB:1097-1105 -> pandapower pd2ppc).  Pure host function
 * of the grid description: no device, no handle. */
int b200pf_grid_max_active_buses(const b200pf_grid_desc *grid, int n, const int8_t *topo, int32_t *per_instance, int32_t *max_out);

/* sizes derived from the grid description */
int b200pf_sizes(const b200pf_handle *h, int *n_topo_in, int *n_inj, int *n_out, int *n_slot);

/* One batched power flow from HOST buffers: H2D copies, kernel, D2H copies, synchronised on return.
 * busv may be NULL.  nb_cap = upper bound of active buses per instance the launch is sized for
 * (<= 0: all n_sub*n_busbar slots, or the largest size that fits on chip).
 * tol_mva: 1e-8 reproduces pandapower's default tolerance_mva; max_iter 10 = pPB:124. */
int b200pf_run_host(b200pf_handle *h, int batch, const int8_t *topo, const double *inj, int is_dc,
                    int max_iter, double tol_mva, int nb_cap, float *out, int32_t *status,
                    int32_t *iters, double *busv);

/* Same with DEVICE pointers, asynchronous on the handle's stream (b200pf_sync to wait). */
int b200pf_run_device(b200pf_handle *h, int batch, const int8_t *d_topo, const double *d_inj,
                      int is_dc, int max_iter, double tol_mva, int nb_cap, float *d_out,
                      int32_t *d_status, int32_t *d_iters, double *d_busv);

/* b200pf_run_device plus a HOST copy of the same topology records: with it the engine can use the planned sparse kernel
 * (its topology plans are looked up / built on the host) for inputs that already live in device memory, e.g. injections
 * written by another kernel or a torch model.  The per-instance plan ids travel on the handle's stream before the launch. */
int b200pf_run_device_topo(b200pf_handle *h, int batch, const int8_t *host_topo, const int8_t *d_topo, const double *d_inj,
                           int is_dc, int max_iter, double tol_mva, int nb_cap, float *d_out, int32_t *d_status,
                           int32_t *d_iters, double *d_busv);

/* Batched time-series stepping (the DoNothing subset of BaseEnv.step, reference
 * grid2op/Environment/baseEnv.py:3778-3872 with NO_OVERFLOW_DISCONNECTION): the chronics
 * (grid2op/Chronics/gridStateFromFile.py:749-808 rows: load_p, load_q, prod_p, prod_v[kV] in
 * BACKEND element order) live in HBM as float32 [n_scen][n_rows][2 n_load + 2 n_gen]; instance i
 * replays scenario scen[i] from row t0[i], wrapping around.  Every call advances all instances by
 * one row, solves, and writes out/status/iters plus rho = a_or / thermal_limit (f32 [batch][n_line]).
 * topo is device state owned by the handle (set with b200pf_series_set_topo). */
int b200pf_series_bind(b200pf_handle *h, const float *chron_host, int n_scen, int n_rows,
                       const int32_t *scen, const int32_t *t0, int batch,
                       const double *static_inj /* [n_inj] defaults for storage / shunt / hidden vm */,
                       const float *thermal_limit_a /* [n_line] */);
int b200pf_series_set_topo(b200pf_handle *h, const int8_t *topo /* [batch][n_topo_in] host */);
int b200pf_series_step(b200pf_handle *h, int is_dc, int max_iter, double tol_mva, int nb_cap);
/* Protections / cascading failure inside the series step (reference Backend.next_grid_state,
 * grid2op/Backend/backend.py:1433-1521, and the counters of BaseEnv._aux_register_env_converged,
 * grid2op/Environment/baseEnv.py:3361-3370): lines whose flow exceeds hard_thr * limit, or that stayed above
 * soft_thr * limit for more than max_allowed consecutive steps, are disconnected and the flow is re-solved
 * until no line trips; the outages persist in the device topology; an instance whose flow diverges is
 * "done" (B200PF_ST_DONE from then on).  enabled = 0 restores NO_OVERFLOW_DISCONNECTION.
 * Two implementations: grids that fit the warp-per-instance kernel (<= 32 lines / bus slots) cascade entirely on the
 * device, inside one launch; all other grids (and every grid under kernel policy 2) run the planned kernel, which reports
 * the lines that trip — a tripped line is a new topology, i.e. a new plan — and the library re-plans those instances on
 * the host and launches the next cascade round for them (b200pf_series_step is then synchronous: one 4-byte read-back
 * per round). */
int b200pf_series_protections(b200pf_handle *h, int enabled, float hard_thr, float soft_thr, int max_allowed);
/* the NEXT series step is an environment reset step: no soft-overflow counting (backend.py:1488-1490) */
int b200pf_series_next_is_reset(b200pf_handle *h);
/* host copies of the protection state after the last step (any pointer may be NULL):
 * counters / timestep_overflow / disc_lines int32 [batch][n_line], done int32 [batch] */
int b200pf_series_fetch_state(b200pf_handle *h, int32_t *protection_counter, int32_t *timestep_overflow,
                              int32_t *disc_lines, int32_t *done);
/* device pointers of the resident results of the last series step (valid until destroy) */
int b200pf_series_results(b200pf_handle *h, float **d_out, int32_t **d_status, int32_t **d_iters,
                          float **d_rho, int32_t **d_t);
/* copy results of the last series step to host buffers (any pointer may be NULL) */
int b200pf_series_fetch(b200pf_handle *h, float *out, int32_t *status, int32_t *iters, float *rho);

/* Pinned host staging buffers owned by the handle (sized for max_batch): a caller that fills topo /
 * inj in place and reads out / status / iters in place avoids one host copy in each direction. */
int b200pf_staging(b200pf_handle *h, int8_t **topo, double **inj, float **out, int32_t **status,
                   int32_t **iters, double **busv);
/* b200pf_run_host on the staging buffers (H2D, kernel, D2H, synchronised on return) */
int b200pf_run_staged(b200pf_handle *h, int batch, int is_dc, int max_iter, double tol_mva, int nb_cap,
                      int want_busv);
/* let the series results land in caller-owned DEVICE buffers (e.g. torch tensors); NULL keeps the
 * handle's own buffer for that array */
int b200pf_series_bind_outputs(b200pf_handle *h, float *d_out, int32_t *d_status, int32_t *d_iters,
                               float *d_rho);
/* N-1 contingency sweep (the reference's pattern: copy_public(); _disconnect_line(i); runpf();
 * get_relative_flow() for every line i - examples/backend_dependant_code/_obs_with_n1.py:111-125, and
 * obs.simulate, grid2op/Observation/baseObservation.py:3365): batch base states x n_line single-line
 * outages = batch*n_line independent AC power flows in ONE launch sharing the base records.
 *   rho    f32   [batch][n_line outages][n_line]   a_or / thermal_limit (NaN rows where the flow diverged)
 *   status int32 [batch][n_line]
 * Host pointers; thermal limits from b200pf_set_thermal_limit.  batch*n_line must be <= max_batch. */
int b200pf_set_thermal_limit(b200pf_handle *h, const float *thermal_limit_a /* [n_line] */);
int b200pf_n1_host(b200pf_handle *h, int batch, const int8_t *topo, const double *inj, int max_iter,
                   double tol_mva, int nb_cap, float *rho, int32_t *status);
/* "rows" entry point: like b200pf_run_staged but the injections of this step come as the float32
 * chronics ROW of every instance (load_p, load_q, prod_p, prod_v[kV]; 2 n_load + 2 n_gen values, backend
 * element order) written by the caller into the pinned rows staging buffer (b200pf_rows_staging);
 * storage / shunt / hidden-unit values come from static_inj (b200pf_set_static_inj).  Less than half
 * the host->device bytes of the f64 record, and no host-side float64 packing. */
int b200pf_set_static_inj(b200pf_handle *h, const double *static_inj /* [n_inj] */);
int b200pf_rows_staging(b200pf_handle *h, float **rows);
int b200pf_run_rows_staged(b200pf_handle *h, int batch, int is_dc, int max_iter, double tol_mva, int nb_cap);
/* Pipelined variant of b200pf_run_rows_staged: the caller fills the pinned rows / topo staging buffers chunk by
 * chunk and launches every chunk as soon as it is ready (asynchronous: H2D, kernel, D2H of that chunk on one of
 * a few internal streams), so host-side preparation of chunk c+1 overlaps the device work of chunk c;
 * b200pf_rows_chunk_wait returns when every launched chunk has landed in the pinned out/status/iters buffers. */
int b200pf_rows_chunk_launch(b200pf_handle *h, int first, int count, int is_dc, int max_iter, double tol_mva, int nb_cap);
int b200pf_rows_chunk_wait(b200pf_handle *h);
/* same, but the rows of the chunk are read from caller-provided PINNED host memory (b200pf_pinned_alloc) instead of
 * the rows staging buffer: lets a driver pre-collate the time series once (rows of all instances for step k
 * contiguous) and feed every step without touching the data on the host. */
int b200pf_rows_chunk_launch_from(b200pf_handle *h, int first, int count, const float *pinned_rows, int is_dc,
                                  int max_iter, double tol_mva, int nb_cap);
/* Group-pipelined stepping (asynchronous vectorised environments): the batch is cut into up to B200PF_MAX_GROUPS
 * groups of instances [first, first+count); each group owns a stream and an event.  launch enqueues H2D (topology
 * records from the pinned staging buffer, rows from pinned_rows or, when NULL, from the pinned rows staging buffer)
 * -> fused kernel -> D2H into the pinned out/status/iters staging buffers and returns at once; wait blocks until that
 * group's results have landed.  While the caller consumes group g (the agent of reference
 * grid2op/Environment/baseEnv.py:3778 `step` acting on its observations) the other groups are in flight, so PCIe
 * transfers of one group overlap the kernels of the others.  A group must be waited for before it is launched again. */
#define B200PF_MAX_GROUPS 16
int b200pf_rows_group_launch(b200pf_handle *h, int group, int first, int count, const float *pinned_rows, int is_dc,
                             int max_iter, double tol_mva, int nb_cap);
int b200pf_rows_group_wait(b200pf_handle *h, int group);
/* flags bit 0 (B200PF_GROUP_DIRECT_OUT): the kernel stores result records / status / iters straight into the pinned
 * staging buffers (device-addressable host memory, posted PCIe writes that overlap the solve of the other instances)
 * instead of a copy-engine D2H after the kernel.  Only while no group is in flight. */
#define B200PF_GROUP_DIRECT_OUT 1
/* flags bit 1 (B200PF_GROUP_ZEROCOPY_IN): the kernel reads the chronics rows (and, for the pivoting kernels, the topology
 * records) of its instances straight from the pinned host buffers — a few hundred bytes per instance, read once — instead
 * of two copy-engine H2D transfers per launch; bit 2 (B200PF_GROUP_DIRECT_STATUS): status / iteration counts (8 bytes per
 * instance) are stored by the kernel straight into the pinned staging buffers, the result records still travel by the copy
 * engine.  Both cut the number of driver calls per group launch (7 -> 3), which bounds the host-side stepping rate. */
#define B200PF_GROUP_ZEROCOPY_IN 2
#define B200PF_GROUP_DIRECT_STATUS 4
int b200pf_rows_group_config(b200pf_handle *h, int flags);
int b200pf_pinned_alloc(size_t bytes, void **ptr);
int b200pf_pinned_free(void *ptr);
/* run all work of this handle on the caller's stream (cudaStream_t as integer; 0 = the handle's own) */
int b200pf_set_stream(b200pf_handle *h, uint64_t stream);

/* Kernel selection.  The engine owns three kernels for the same path (reference pPB:1090 / 1097-1105):
 *   - the PLANNED SPARSE kernel (csrc/b200pf_sparse.cuh): everything that depends on the topology only (bus fusion, bus
 *     types, connectivity check, elimination order, filled Jacobian pattern, factorisation schedule, factorised DC matrix)
 *     is computed once per distinct topology vector on the host and cached ("plan"); the device runs the numeric part.
 *     Needs a host copy of the topology (every entry point except b200pf_run_device has one) and no device-side
 *     protections;
 *   - the warp-per-instance and CTA-per-instance kernels with on-device topology discovery and partial pivoting.
 * policy 0 (default): planned kernel whenever it applies, unless one call would have to build more than a few hundred
 * new plans; 1: pivoting kernels only; 2: planned kernel whenever it applies.  Env B200PF_PLAN_POLICY presets it. */
int b200pf_set_kernel_policy(b200pf_handle *h, int policy);
/* number of cached plans, their bytes, and the kernel of the last launch (1 warp/pivoting, 2 CTA/pivoting, 3 planned sparse) */
int b200pf_plan_stats(const b200pf_handle *h, int64_t *n_plans, int64_t *plan_bytes, int *last_kernel);
/* out4 = { topology rows looked up in the plan cache, of which found, plans built, cache resets (the cache is append-only and
 * starts over when it is full: 32768 plans / 1 GiB) } since b200pf_create.  New plans of one call are built in parallel on the
 * host threads (env B200PF_PLAN_THREADS caps them). */
int b200pf_plan_counters(const b200pf_handle *h, int64_t *out4);

/* Multi-GPU result collection (instances shard across GPUs, the step results go to the agent's rank).  Instead
 * of one collective launch per 60-microsecond step, the agent's rank allocates the result buffer with b200pf_device_alloc
 * and exports it (CUDA IPC); every other rank (one process per GPU) opens the handle — the buffer is then peer-mapped over
 * NVLink — and binds a slice of it as its `rho` output (b200pf_series_bind_outputs): the kernel's own result stores land in
 * the agent's HBM, fused with the solve, no extra launch.  Completion is signalled by whatever orders the streams (a tiny
 * NCCL all-reduce every K steps in bench.py).  handle64: 64 bytes (cudaIpcMemHandle_t). */
/* env.reset() of single instances of a bound series (reference Environment.reset, environment.py; BaseEnv counters
 * baseEnv.py:3352-3393): done flag, protection counters, overflow counters and disc_lines back to a fresh episode; t_new (may be
 * NULL): the chronics row the instance's next solve uses; topo_rows (may be NULL): int8 [n][n_topo_in], its topology. */
int b200pf_series_reset_instances(b200pf_handle *h, int n, const int32_t *idx, const int32_t *t_new, const int8_t *topo_rows);
/* b200pf_series_bind_flag: after every b200pf_series_step the device stores the number of steps done so far to *d_flag (behind
 * all the step's kernels, system-wide visible) — d_flag may point into the agent rank's buffer, next to the results: the agent
 * learns that a rank's step k has fully arrived by reading a 4-byte word, no collective involved.  NULL unbinds. */
int b200pf_series_bind_flag(b200pf_handle *h, int32_t *d_flag);
int b200pf_device_alloc(size_t bytes, void **d_ptr);
int b200pf_device_free(void *d_ptr);
int b200pf_device_read(const void *d_src, void *host_dst, size_t bytes);   /* synchronous D2H copy (checks) */
int b200pf_ipc_export(const void *d_ptr, unsigned char *handle64);
int b200pf_ipc_open(const unsigned char *handle64, void **d_ptr);
int b200pf_ipc_close(void *d_ptr);

/* Safety net of the planned kernel.  The planned kernel eliminates without pivoting on an fp32 Jacobian; the reference's
 * solver (pp.runpp, pPB:1097-1105) pivots in fp64, and a failed power flow is a game over (pPB:1241-1255,
 * grid2op/Environment/baseEnv.py:3523-3524).  Every planned launch is therefore followed, on the same stream and inside the
 * same C-ABI call, by a launch that re-solves with partial pivoting (fp64 Jacobian where the workspace allows) the instances
 * the planned kernel left as DIVERGED; a status only leaves the library after that.  b200pf_set_debug is a TEST knob:
 * planned_div_mod > 0 makes the planned kernel give up on every instance with (index % planned_div_mod == 0), which
 * exercises the re-solve path; redo_enabled = 0 switches the safety net off (measurements).  Env B200PF_NO_REDO=1 presets
 * the latter.  b200pf_redo_launch_count: safety-net launches so far (they are not part of b200pf_launch_count). */
int b200pf_set_debug(b200pf_handle *h, int planned_div_mod, int redo_enabled);
int64_t b200pf_redo_launch_count(const b200pf_handle *h);

int b200pf_sync(b200pf_handle *h);
/* cudaStream_t of the handle as an integer (for event timing by the caller) */
uint64_t b200pf_stream(b200pf_handle *h);
/* number of kernels launched by this handle so far */
int64_t b200pf_launch_count(const b200pf_handle *h);
/* dynamic shared memory bytes and threads-per-instance the last launch used (diagnostics) */
int b200pf_last_launch_info(const b200pf_handle *h, int *smem_bytes, int *threads_per_instance,
                            int *grid_blocks, int *block_threads);

#ifdef __cplusplus
}
#endif
#endif /* B200PF_H */
