/*
 * mdevaluate_b200 -- C ABI of the B200 (sm_100a) shifted-correlation path.
 *
 * The reference (mdevaluate, pure Python) has no FFI: its boundary for this path
 * is three Python call protocols.  This header is the
 * drop-in boundary underneath them: every entry point below replaces one NumPy
 * loop of the reference and is what a maintainer would bind with ctypes (see
 * INTEGRATION.md).  Rules of the ABI:
 *
 *   - plain `extern "C"`, pointers + sizes only, no C++/torch types;
 *   - every function returns an `mdb_status` (0 = ok); nothing throws; the text
 *     of the last error of the calling thread is `mdb_last_error()`;
 *   - pointers documented as DEVICE are CUDA device pointers owned by the caller
 *     and stay valid until the stream has drained; HOST pointers are read before
 *     the call returns;
 *   - `stream` is a `cudaStream_t` passed as `void*` (NULL = default stream); all
 *     work is enqueued on it and the call does not synchronise unless stated;
 *   - no hidden global state: calls are re-entrant per stream, nothing is read from the environment; every
 *     switch is a field of an argument struct (`mdb_tuning`), 0 = the measured default.
 *
 * Reference lines are relative to /root/reference/mdevaluate/.
 */
#ifndef MDEVALUATE_B200_H
#define MDEVALUATE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MDB_ABI_VERSION 2

typedef enum mdb_status {
    MDB_OK = 0,
    MDB_ERR_ARGUMENT = 1,     /* NULL / negative size / misaligned / inconsistent arguments */
    MDB_ERR_CUDA = 2,         /* a CUDA runtime call failed (text in mdb_last_error) */
    MDB_ERR_UNSUPPORTED = 3,  /* valid request this build has no kernel for */
    MDB_ERR_WORKSPACE = 4,    /* workspace missing or too small */
    MDB_ERR_TIMEOUT = 5       /* a peer GPU did not deliver its partial sums in time (mdb_comm) */
} mdb_status;

/* image counts are stored as uint8 = count + MDB_IMAGE_BIAS */
#define MDB_IMAGE_BIAS 128

/* Image-count storage.  The jump scan writes MDB_IMAGE_BYTES (one excess-128 byte per component).  When no count
 * leaves [-1, 1] -- the rule for liquids whose atoms travel less than a box length -- mdb_images_pack re-codes the
 * rows as MDB_IMAGE_PACKED: one 32-bit word per group of 4 atoms, byte a of the word = atom a of the group,
 * bits 0-1 / 2-3 / 4-5 = n_x+1, n_y+1, n_z+1, bits 6-7 zero (1 byte instead of 3 per atom and frame: the
 * correlation kernels then move 13 instead of 15 bytes per evaluation).  Counts beyond +-127 are kept as int16
 * (MDB_IMAGE_I16, direct-load kernel). */
#define MDB_IMAGE_BYTES  1
#define MDB_IMAGE_PACKED 2
#define MDB_IMAGE_I16    3

/* Explicit switches of the kernels; 0 everywhere = the measured defaults. */
typedef struct mdb_tuning {
    int32_t ring_stages;     /* depth of the bulk-copy ring (3..8); 0 = what fits                             */
    int32_t atom_blocks;     /* atom-block split of a correlation launch; 0 = wave model                      */
    int32_t isf_float64;     /* 1: ISF displacements through the float64 chain (default float32, see DESIGN)  */
    int32_t hist_float64;    /* 1: van Hove (and fused van Hove + alpha_2) displacements in float64           */
    int32_t scan_direct;     /* 1: register-load variant of the jump scan instead of the bulk-copy ring       */
    int32_t dense_stage_cta; /* 1: CTA-wide staging loop of the float64 dense kernel instead of warp-per-row  */
    int32_t dense_float64;   /* 1: dense schedules through the float64 tile (default: float32 offsets, v2)    */
    int32_t variant;         /* measurement switches, 0 = the product path.  bit 0: dense kernel, general origin loop
                              * even when every frame is an origin (no rotating row window); bit 1: dense kernel,
                              * float64 staging even when the box lengths are float32 numbers; bit 2: float32-
                              * displacement kernels without the no-image-change shortcut                      */
} mdb_tuning;

/* coordinate element type */
#define MDB_F32 0
#define MDB_F64 1

/* what one correlation launch accumulates per (origin, lag) */
#define MDB_SUM_R2   1   /* sum_i |dr_i|^2              -> msd            correlation.py:203-208 */
#define MDB_SUM_R4   2   /* sum_i |dr_i|^4              -> non_gaussian   correlation.py:355-362 */
#define MDB_SUM_SINC 4   /* sum_i sin(q r_i)/(q r_i)    -> isf            correlation.py:211-220 */
#define MDB_HIST     8   /* histogram of r_i            -> van_hove_self  correlation.py:239-248 */
#define MDB_SUM_LEGENDRE 16 /* sum_i P_l(u_i(0).u_i(t)), alone -> rotational_autocorrelation correlation.py:223-236;
                               the rows hold vectors, the sum lands in slot 0 of `sums` */

/* histogram binning rule (utils.py:328-338) */
#define MDB_BINS_UNIFORM 0  /* bins.min()==0 and equal widths: bin = floor_divide(r, dx), r >= n_bins*dx dropped */
#define MDB_BINS_EDGES   1  /* np.histogram(r, bins): half-open bins, last bin closed, outside dropped        */

const char *mdb_last_error(void);
int mdb_abi_version(void);

/* Number of SMs / compute capability of the current device. */
int mdb_device_info(int *sm_count, int *cc_major, int *cc_minor);

/* ------------------------------------------------------------------------------------------
 * Correlation over many origins and lags.
 *
 * Replaces the S x L Python loop of `shifted_correlation` (correlation.py:147-149,190-199)
 * together with the per-pair kernels msd / isf / van_hove_self / non_gaussian
 * (correlation.py:203-220,239-248,355-362) for one shard of atoms.
 *
 * Storage: `coords` is a row-major block of frame rows, row r = `coords + r*pitch` holding
 * `3*n_atoms` interleaved xyz values (float32 as delivered by a trajectory reader, or float64
 * for already unwrapped frames).  With `images` != NULL every row also has cumulative periodic
 * image counts n, one byte per component stored excess-128 (uint8 = n + MDB_IMAGE_BIAS; row r at
 * `images + r*img_pitch`, 3*n_atoms bytes, as written by mdb_nojump_scan) and a box diagonal
 * `box[3*r .. 3*r+2]` (or the single `box_const` when the box never changes); the kernels then use the unwrapped position  x - n*box  evaluated in
 * float64, which is exactly the float64 frame `pbc.nojump` returns (pbc.py:235-247).
 *
 * `rows[s*n_lags + l]` is the storage row of frame  start_frames[s] + idx[l]; entry l == 0 is
 * the origin of segment s (correlation.py:21-24: the origin is correlated with itself first).
 *
 * Output: `sums[(s*n_lags + l)*3 + {0,1,2}]` = sum over the shard's atoms of r^2, r^4 and
 * sinc(q r) (unrequested slots are 0); `hist[(s*n_lags + l)*n_bins + b]` = integer counts.
 * Division by N, the alpha_2 ratio, the average over origins and the variance for chi4 are
 * done by the caller AFTER shards have been summed (correlation.py:194-199 averages values
 * per origin, not moments).  Results are bit-reproducible run to run.
 * ------------------------------------------------------------------------------------------ */
typedef struct mdb_corr_args {
    const void *coords;          /* DEVICE [n_rows][pitch] of coord_type                        */
    int64_t pitch;               /* elements per row, >= 3*n_atoms                              */
    int32_t coord_type;          /* MDB_F32 or MDB_F64                                           */
    const uint8_t *images;       /* DEVICE [n_rows][img_pitch] excess-128, or NULL (raw / unwrapped) */
    int64_t img_pitch;           /* bytes per image row, >= 3*n_atoms                            */
    const double *box;           /* DEVICE [n_rows][3] box diagonals per row, or NULL: box_const  */
    double box_const[3];         /* the one box diagonal of every row when box == NULL (with images) */
    int32_t n_atoms;             /* atoms of this shard                                          */
    int32_t n_origins;           /* S                                                            */
    int32_t n_lags;              /* L                                                            */
    const int32_t *rows;         /* DEVICE [S*L] storage row per (origin, lag)                   */
    const uint8_t *select;       /* DEVICE [S][sel_pitch] nonzero = atom takes part for origin s, or NULL (all atoms);
                                    the sub-ensemble of subensemble_correlation, correlation.py:27-35 */
    int64_t sel_pitch;           /* bytes per origin, >= n_atoms, multiple of 4                  */
    int32_t flags;               /* MDB_SUM_* | MDB_HIST                                         */
    double q;                    /* scattering vector length for MDB_SUM_SINC                    */
    int32_t legendre_order;      /* l of P_l for MDB_SUM_LEGENDRE (0..32)                        */
    int32_t n_bins;              /* B, for MDB_HIST                                              */
    int32_t bin_mode;            /* MDB_BINS_UNIFORM / MDB_BINS_EDGES                            */
    double dx;                   /* uniform bin width (np.diff(bins).mean())                     */
    const double *edges;         /* DEVICE [B+1] bin edges for MDB_BINS_EDGES                    */
    double *sums;                /* DEVICE out [S*L*3]                                           */
    unsigned long long *hist;    /* DEVICE out [S*L*B] (zeroed by the call); NULL without HIST   */
    void *workspace;             /* DEVICE scratch of mdb_corr_workspace_bytes()                 */
    size_t workspace_bytes;
    const uint32_t *packed;      /* DEVICE [n_rows][packed_pitch] MDB_IMAGE_PACKED rows, or NULL; preferred over
                                    `images` when both are given                                   */
    int64_t packed_pitch;        /* 32-bit words per packed row, >= ceil(n_atoms/4)              */
    const int16_t *images16;     /* DEVICE [n_rows][img16_pitch] MDB_IMAGE_I16 rows, or NULL      */
    int64_t img16_pitch;         /* int16 elements per row, >= 3*n_atoms                         */
    mdb_tuning tuning;
    int32_t n_segments;          /* 0 / 1: `select` is a mask.  > 1 (MDB_SUM_SINC alone): byte j of `select` names the
                                    sub-ensemble 1..n_segments atom j belongs to for that origin (0 = none); ONE pass
                                    accumulates all of them (multi_subensemble_correlation, correlation.py:38-81):
                                    sums[((s*n_lags + l)*n_segments + k)*3 + c]                     */
} mdb_corr_args;

int mdb_corr_workspace_bytes(const mdb_corr_args *args, size_t *bytes);
int mdb_shifted_corr(const mdb_corr_args *args, void *stream);
/* Number of kernel launches the last successful mdb_shifted_corr of this thread enqueued. */
int mdb_corr_last_launches(void);
/* Measurement aid: with profiling enabled (per calling thread) mdb_shifted_corr brackets its
 * dominant kernel -- the correlation kernel, not the partial-sum reduction -- with CUDA events on
 * the launch stream; mdb_corr_last_kernel_ms waits for that kernel and returns its duration. */
int mdb_corr_profile(int enable);
int mdb_corr_last_kernel_ms(float *ms);

/* ------------------------------------------------------------------------------------------
 * Repeated calls: a PLAN owns everything one `shifted_correlation` call needs besides the frame rows --
 * the atom-block workspace, the per-(origin, lag) sums and histogram on the device, and a page-locked
 * host block the results land in -- so that one call is two kernel launches and no allocation:
 *
 *   corr kernel  ->  finish kernel: adds the atom-block partials in fixed order, (with a communicator)
 *                    exchanges them with the peer GPUs and adds the shards in rank order, writes the sums
 *                    to the device buffer and, through the mapped host block, straight into host memory,
 *                    then raises a sequence flag the host is polling.
 *
 * The caller keeps ownership of coords / images / packed / box / rows / select / edges (DEVICE, they must
 * stay valid while the plan is used; `sums`, `hist`, `workspace` of `args` are ignored).  What the reference
 * does per call in Python -- the S x L loop and the per-frame function calls, correlation.py:147-149,190-199 --
 * is mdb_corr_plan_run.
 * ------------------------------------------------------------------------------------------ */
typedef struct mdb_corr_plan mdb_corr_plan;
typedef struct mdb_comm mdb_comm;

int mdb_corr_plan_create(const mdb_corr_args *args, mdb_comm *comm /* or NULL */, mdb_corr_plan **plan);
/* Enqueue both kernels on `stream` and wait for the results (spins on the host flag; no cudaStreamSynchronize on
 * the fast path).  With `time_slot` >= 0 the correlation kernel is bracketed by the CUDA event pair of that slot
 * (0..MDB_PLAN_TIME_SLOTS-1); mdb_corr_plan_kernel_ms reads a slot afterwards. */
#define MDB_PLAN_TIME_SLOTS 1024
int mdb_corr_plan_run(mdb_corr_plan *plan, void *stream, int32_t time_slot);
/* The two halves of mdb_corr_plan_run for callers that overlap host work with the kernels. */
int mdb_corr_plan_launch(mdb_corr_plan *plan, void *stream, int32_t time_slot);
int mdb_corr_plan_wait(mdb_corr_plan *plan);
int mdb_corr_plan_kernel_ms(mdb_corr_plan *plan, int32_t time_slot, float *ms);
/* How the host learns that a run is complete: 0 (default) the finish kernel writes the sums through the mapped host
 * block and raises a flag the host polls; 1 the same writes, cudaStreamSynchronize instead of polling; 2 the sums
 * stay on the device and are copied with cudaMemcpyAsync + cudaStreamSynchronize. */
int mdb_corr_plan_set_wait(mdb_corr_plan *plan, int32_t mode);
/* HOST pointers into the plan's page-locked result block, valid until mdb_corr_plan_destroy: sums [S*L*3]
 * (summed over the shards of `comm`), hist [S*L*B] or NULL.  Overwritten by the next run. */
const double *mdb_corr_plan_sums(const mdb_corr_plan *plan);
const unsigned long long *mdb_corr_plan_hist(const mdb_corr_plan *plan);
/* DEVICE copies of the same (e.g. to feed another kernel without a round trip). */
const double *mdb_corr_plan_sums_device(const mdb_corr_plan *plan);
int mdb_corr_plan_launches(const mdb_corr_plan *plan);      /* kernels one run enqueues */
/* Where the time of the plan's last completed run went, in ns.  out[0..2]: the host's steady clock when
 * mdb_corr_plan_launch was entered, when it returned and when the wait saw the result.  out[3..5]: the GPU's
 * %globaltimer when the first CTA of the correlation kernel started, when the finish kernel started and just before it
 * raised the host flag.  The clocks have different origins; mdb_clock_pair relates them (one tiny kernel that stores
 * %globaltimer into mapped host memory: *gpu_ns lies between *host_before_ns and *host_seen_ns). */
int mdb_corr_plan_timeline(const mdb_corr_plan *plan, unsigned long long *out6);
int mdb_clock_pair(void *stream, unsigned long long *gpu_ns, unsigned long long *host_before_ns,
                   unsigned long long *host_seen_ns);
int mdb_corr_plan_destroy(mdb_corr_plan *plan);

/* ------------------------------------------------------------------------------------------
 * Multi-GPU (one process per GPU, atoms sharded by contiguous blocks).  The one
 * collective of the path -- the sum of the small per-(origin, lag) partials over the shards -- is done by
 * the finish kernel itself over NVLink peer memory: every rank owns an inbox (device memory exported with
 * CUDA IPC), the finish kernel of rank r stores its partials into slot r of EVERY rank's inbox, raises the
 * per-slot sequence flags, waits for the flags of its own inbox and adds the slots in rank order (every
 * rank gets bit-identical results).  Inboxes are double-buffered by sequence parity.  A peer that does not
 * deliver within `timeout_ms` makes the run fail with MDB_ERR_TIMEOUT instead of hanging.
 *
 *   mdb_comm_create(rank, world, capacity, &comm);  mdb_comm_handle(comm, my_handle);
 *   -- exchange the MDB_COMM_HANDLE_BYTES of all ranks by any means (torch.distributed.all_gather) --
 *   mdb_comm_connect(comm, all_handles);
 * ------------------------------------------------------------------------------------------ */
#define MDB_COMM_HANDLE_BYTES 64
#define MDB_COMM_MAX_RANKS 16
int mdb_comm_create(int32_t rank, int32_t world, size_t capacity_words /* 8-byte words per rank slot */,
                    int32_t timeout_ms, mdb_comm **comm);
int mdb_comm_handle(mdb_comm *comm, void *handle_out /* HOST MDB_COMM_HANDLE_BYTES */);
int mdb_comm_connect(mdb_comm *comm, const void *handles /* HOST world * MDB_COMM_HANDLE_BYTES, by rank */);
int mdb_comm_destroy(mdb_comm *comm);

/* ------------------------------------------------------------------------------------------
 * nojump, producer side: periodic-jump detection + running image count.
 *
 * Replaces `reader.parse_jumps` (reader.py:165-186: J_t = round((x_t - x_{t-1}) / box_0) as
 * int8, round-half-even, float32 arithmetic when positions and box are float32) and the prefix
 * sum `m[i0:step+1].sum(axis=0)` inside `pbc.nojump` (pbc.py:235-237).  A trajectory may be
 * passed in consecutive time segments: `prev` / `cum` carry the last frame and the running
 * counts from one call to the next.
 * ------------------------------------------------------------------------------------------ */
typedef struct mdb_nojump_args {
    const float *coords;         /* DEVICE [n_frames][pitch] wrapped positions of this segment   */
    int64_t pitch;               /* floats per row, >= 3*n_atoms                                 */
    int32_t n_frames;
    int32_t n_atoms;
    double box0[3];              /* box diagonal of trajectory frame 0 (reader.py:167)           */
    int32_t box_is_f32;          /* 1: divide in float32 (float32 box), 0: in float64            */
    int32_t first_segment;       /* 1: row 0 is trajectory frame 0 (compared with whole0/itself) */
    const float *whole0;         /* DEVICE [3*n_atoms] frame0.whole, or NULL (= frame 0)         */
    float *prev;                 /* DEVICE [3*n_atoms] in/out carry: last frame seen             */
    int32_t *cum;                /* DEVICE [3*n_atoms] in/out carry: running image counts        */
    const int32_t *out_row;      /* DEVICE [n_frames] destination row per frame, <0 = drop; NULL = identity */
    uint8_t *images;             /* DEVICE out, cumulative counts (excess-128) per kept frame, or NULL */
    int64_t img_pitch;           /* bytes per row of images / jumps                              */
    int8_t *jumps;               /* DEVICE out, per-step jumps J_t per kept frame, or NULL       */
    float *coords_out;           /* DEVICE out, copy of kept frames (compaction), or NULL        */
    int64_t pitch_out;           /* floats per row of coords_out                                 */
    int32_t *overflow;           /* DEVICE int, set to 1 if a running count left [-128, 127]     */
    int16_t *images16;           /* DEVICE out, cumulative counts as int16 per kept frame (MDB_IMAGE_I16), or NULL;
                                    the fallback when `overflow` was raised: counts up to +-32767 */
    int64_t img16_pitch;         /* int16 elements per row of images16                           */
    mdb_tuning tuning;
} mdb_nojump_args;

int mdb_nojump_scan(const mdb_nojump_args *args, void *stream);

/* nojump from STORED jump matrices (the reference's `.{traj}.nojump.npz` cache: three CSR int8 matrices
 * J_d of shape frames x atoms, reader.py:189-241).  The caller re-sorts them by column into ONE compressed
 * list over the 3*n_atoms interleaved components (column c = 3*atom + d): the stored jumps of column c are
 * entries indptr[c] .. indptr[c+1]-1 of `frames` (ascending) / `jumps`.  Writes the running image count
 * sum_{t <= keep[r]} J[t, c]  -- `m[:step+1].sum(axis=0)` of pbc.py:235-237 -- of every kept frame in the
 * excess-128 byte format the correlation kernels read.  With the matrices at hand only the frames a
 * schedule touches have to be read from the trajectory (no scan over all frames). */
typedef struct mdb_jumpmat_args {
    const int64_t *indptr;       /* DEVICE [3*n_atoms + 1] offsets per component column          */
    const int32_t *frames;       /* DEVICE [nnz] frame of every stored jump, ascending per column */
    const int8_t *jumps;         /* DEVICE [nnz] the jump J[frame, column]                        */
    int64_t nnz;
    int32_t n_atoms;
    const int32_t *keep;         /* DEVICE [n_keep] ascending frames to produce rows for; NULL = 0..n_keep-1 */
    int32_t n_keep;
    uint8_t *images;             /* DEVICE out [n_keep][img_pitch] excess-128                     */
    int64_t img_pitch;           /* bytes per row, >= 3*n_atoms                                   */
    int32_t *overflow;           /* DEVICE int, set to 1 if a count left [-128, 127]; or NULL     */
    int16_t *images16;           /* DEVICE out [n_keep][img16_pitch] plain int16 counts instead of `images` (pass
                                    images = NULL): the fallback when `overflow` was raised       */
    int64_t img16_pitch;
} mdb_jumpmat_args;

int mdb_nojump_images_from_jumps(const mdb_jumpmat_args *args, void *stream);

/* MDB_IMAGE_BYTES rows -> MDB_IMAGE_PACKED rows (see above).  `max_abs` (DEVICE int, zeroed by the caller) receives
 * the largest |count| seen: the packed rows are valid iff it ends up <= 1. */
int mdb_images_pack(const uint8_t *images, int64_t img_pitch, int64_t n_rows, int32_t n_atoms, uint32_t *packed,
                    int64_t packed_pitch, int32_t *max_abs, void *stream);

/* nojump, consumer side (pbc.py:235-247): out = x - n*box in float64 for one frame row. */
int mdb_nojump_apply(const float *coords_row, const uint8_t *images_row, const double box[3],
                     int32_t n_atoms, double *out, void *stream);
/* The same for a row in any image format (`image_format` = MDB_IMAGE_BYTES / _PACKED / _I16). */
int mdb_nojump_apply_format(const float *coords_row, const void *images_row, int32_t image_format, const double box[3],
                            int32_t n_atoms, double *out, void *stream);

/* Trajectory ingestion (the reference pulls frames one by one through its reader, coordinates.py:257-271):
 * `n_rows` rows of `width_bytes` from page-locked host memory (row stride `src_pitch_bytes`) into device rows
 * (stride `dst_pitch_bytes`) as ONE strided DMA on `stream` -- e.g. the columns of an atom slab of a
 * (frames x 3N) float32 trajectory.  Asynchronous when the source is page-locked. */
int mdb_copy_rows_h2d(const void *src_host, int64_t src_pitch_bytes, void *dst_dev, int64_t dst_pitch_bytes,
                      int64_t width_bytes, int64_t n_rows, void *stream);

/* ------------------------------------------------------------------------------------------
 * Vectors between atoms for the frames a schedule touches (replaces `coordinates.vectors`, coordinates.py:524-563,
 * mapped over the frames by CoordinatesMap, :332-393): the input rows of `rotational_autocorrelation`
 * (MDB_SUM_LEGENDRE) are built on the device from the resident frame rows.
 *   out[r][v] = pbc_diff( mean_i x[atoms_a[i][v]], mean_j x[atoms_b[j][v]], pbc_box ) [ / norm ]
 * in the reference's arithmetic, operation by operation: float64 with images (the nojump frame x - n*box), float32
 * for raw frames.
 * ------------------------------------------------------------------------------------------ */
typedef struct mdb_vectors_args {
    const float *coords;         /* DEVICE [n_store_rows][pitch] float32 frame rows                */
    int64_t pitch;
    const void *images;          /* DEVICE image rows in `image_format`, or NULL (raw frames)      */
    int32_t image_format;        /* MDB_IMAGE_BYTES / _PACKED / _I16                               */
    int64_t img_pitch;           /* elements of that format per row                                */
    const double *box;           /* DEVICE [n_store_rows][3] or NULL: box_const (for the unwrap)   */
    double box_const[3];
    const int32_t *rows;         /* DEVICE [n_rows] storage rows to transform, or NULL (0..n_rows-1) */
    int32_t n_rows;
    const int32_t *atoms_a;      /* DEVICE [m_a][n_vec] atom indices; the mean over m_a is taken   */
    int32_t m_a;
    const int32_t *atoms_b;      /* DEVICE [m_b][n_vec]                                            */
    int32_t m_b;
    int32_t n_vec;
    int32_t normed;              /* divide by the norm                                             */
    int32_t use_pbc;             /* 1: minimum image with pbc_box (pbc.py:44-56)                   */
    double pbc_box[3];
    int32_t out_type;            /* MDB_F32 (raw frames) or MDB_F64 (required with images)          */
    void *out;                   /* DEVICE out [n_rows][out_pitch] of out_type, 3*n_vec used        */
    int64_t out_pitch;
} mdb_vectors_args;

int mdb_vectors(const mdb_vectors_args *args, void *stream);

/* ------------------------------------------------------------------------------------------
 * Dense schedules: every origin of an arithmetic progression against a contiguous range of lags,
 * averaged over the origins (BASELINE config 5: `index_distribution = lambda a, b: np.arange(a, b)`,
 * `average=True`).  Pair (s, l) correlates frame  first_origin + s*origin_step  with the frame
 * `first_lag + l` later.  Blocks of origins x lags are evaluated from frame rows held in shared
 * memory (unwrapped once per row, reused by every pair of the block), so the kernel is bound by
 * arithmetic instead of HBM.  Output `sums[l*2 + {0,1}]` = sum over origins and atoms of r^2 and
 * sinc(q r) (the caller divides by n_origins * N).  Supported flags: MDB_SUM_R2 | MDB_SUM_SINC.
 * ------------------------------------------------------------------------------------------ */
typedef struct mdb_dense_args {
    const float *coords;         /* DEVICE [n_rows][pitch] float32                              */
    int64_t pitch;
    const uint8_t *images;       /* DEVICE excess-128 image counts or NULL (raw)                */
    int64_t img_pitch;
    const double *box;           /* DEVICE [n_rows][3] or NULL: box_const                        */
    double box_const[3];
    const int32_t *frame_rows;   /* DEVICE [n_frames] storage row of every frame, or NULL (row = frame) */
    int32_t n_atoms;
    int32_t first_origin;        /* frame of origin 0                                            */
    int32_t origin_step;         /* frames between consecutive origins (1..)                     */
    int32_t n_origins;
    int32_t first_lag;           /* lag of result 0 (frames)                                     */
    int32_t n_lags;              /* result l is lag first_lag + l                                */
    int32_t flags;
    double q;
    double *sums;                /* DEVICE out [n_lags*2]                                        */
    void *workspace;             /* DEVICE scratch of mdb_dense_workspace_bytes()                */
    size_t workspace_bytes;
    const uint32_t *packed;      /* DEVICE MDB_IMAGE_PACKED rows, or NULL; preferred over `images` */
    int64_t packed_pitch;
    mdb_tuning tuning;
} mdb_dense_args;

int mdb_dense_workspace_bytes(const mdb_dense_args *args, size_t *bytes);
int mdb_dense_corr(const mdb_dense_args *args, void *stream);

/* ------------------------------------------------------------------------------------------------
 * Overlap with a reference configuration.  Replaces correlation.overlap (correlation.py:288-312) evaluated through
 * shifted_correlation on a coordinates.CoordinatesKDTree (coordinates.py:415-455): for every (origin s, lag l)
 *   counts[s*L + l] = #{ atoms i of frame rows[s*L + l] : some tree atom j with |x_i - y_j| <= radius }
 * where the tree of origin s holds the atoms of row ref_rows[s] of the tree's trajectory that ref_select marks
 * (all when NULL).  The caller divides by the tree size (tree_sizes).  Distances as scipy's cKDTree computes them
 * (float64, no fused multiply-adds); boxsize[d] > 0 makes dimension d periodic like cKDTree(boxsize=...): query
 * points are wrapped, tree atoms must lie in [0, boxsize) (out_of_box reports a violation; scipy raises ValueError).
 * Frames: float32 or float64 rows, raw or with image counts in any MDB_IMAGE_* format (nojump frames x - n*box).
 * ------------------------------------------------------------------------------------------------ */
typedef struct mdb_overlap_args {
    const void *ref_coords;      /* DEVICE rows of the tree's trajectory                         */
    int64_t ref_pitch;
    int32_t ref_type;            /* MDB_F32 / MDB_F64                                            */
    int32_t n_ref;               /* atoms per tree row                                           */
    const void *ref_images;      /* DEVICE image counts of those rows, or NULL (raw frames)      */
    int32_t ref_image_format;    /* MDB_IMAGE_*                                                  */
    int32_t reserved0;
    int64_t ref_img_pitch;       /* elements of that format per row                              */
    const double *ref_box;       /* DEVICE [rows][3] or NULL: ref_box_const                      */
    double ref_box_const[3];
    const int32_t *ref_rows;     /* DEVICE [n_origins] tree row of every origin                  */
    const uint8_t *ref_select;   /* DEVICE [n_origins][ref_sel_pitch] 1 = atom is in the tree; or NULL */
    int64_t ref_sel_pitch;
    const void *coords;          /* DEVICE rows of the correlated trajectory                     */
    int64_t pitch;
    int32_t coord_type;
    int32_t n_atoms;
    const void *images;
    int32_t image_format;
    int32_t file_order;          /* 1: query the atoms in file order even when the tree holds all of them
                                  * (default: in the order of the origin's sorted tree, which coalesces)   */
    int64_t img_pitch;
    const double *box;
    double box_const[3];
    const int32_t *rows;         /* DEVICE [n_origins*n_lags]                                    */
    int32_t n_origins, n_lags;
    double radius;
    double boxsize[3];           /* periodic tree, or zeros                                      */
    unsigned long long *counts;  /* DEVICE out [n_origins*n_lags]                                */
    int32_t *tree_sizes;         /* DEVICE out [n_origins], or NULL                              */
    double *bounds;              /* DEVICE out [n_origins][6] min xyz, max xyz of every tree, or NULL */
    int32_t *out_of_box;         /* DEVICE out [1]: 1 if a periodic tree has atoms outside the box, or NULL */
    void *workspace;             /* DEVICE scratch of mdb_overlap_workspace_bytes()              */
    size_t workspace_bytes;
} mdb_overlap_args;

int mdb_overlap_workspace_bytes(const mdb_overlap_args *args, size_t *bytes);
int mdb_overlap(const mdb_overlap_args *args, void *stream);
int mdb_overlap_launches(void);                  /* kernels one mdb_overlap call enqueues */

#ifdef __cplusplus
}
#endif
#endif /* MDEVALUATE_B200_H */
