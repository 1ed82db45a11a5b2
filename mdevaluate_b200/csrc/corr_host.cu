// Host side of the correlation path: launch planning, the C ABI (mdb_shifted_corr, mdb_corr_plan_*) and the
// finish kernel that turns the atom-block partials of a launch into the per-(origin, lag) sums -- summed over the
// GPUs of a communicator through NVLink peer memory -- and hands them to the host through a mapped, page-locked
// result block.
#include <atomic>
#include <chrono>
#include <new>

#include "corr_core.cuh"

namespace mdb {
namespace {

constexpr size_t kSmemBudgetPerCta = (MDB_BULK_CTAS_PER_SM == 3 ? 72 : 110) * 1024;   // CTAs per SM share 227 KiB

int make_plan(const mdb_corr_args *a, Plan *plan) {
    if (!a) return fail(MDB_ERR_ARGUMENT, "args is NULL");
    if (a->n_atoms <= 0 || a->n_origins <= 0 || a->n_lags <= 0)
        return fail(MDB_ERR_ARGUMENT, "n_atoms, n_origins and n_lags must be positive");
    if (!(a->flags & (MDB_SUM_R2 | MDB_SUM_R4 | MDB_SUM_SINC | MDB_HIST | MDB_SUM_LEGENDRE)))
        return fail(MDB_ERR_ARGUMENT, "flags selects nothing to compute");
    const bool any_images = a->images || a->packed || a->images16;
    if ((a->flags & MDB_SUM_LEGENDRE) && (a->flags != MDB_SUM_LEGENDRE || any_images || a->legendre_order < 0 ||
                                          a->legendre_order > 32))
        return fail(MDB_ERR_ARGUMENT, "MDB_SUM_LEGENDRE stands alone, on raw vector rows, with 0 <= order <= 32");
    if (a->pitch < 3ll * a->n_atoms) return fail(MDB_ERR_ARGUMENT, "pitch < 3*n_atoms");
    plan->n_groups = (a->n_atoms + 3) / 4;
    plan->n_chunks = (a->n_lags + kLagChunk - 1) / kLagChunk;

    const size_t esz = a->coord_type == MDB_F32 ? 4 : 8;
    const bool coords_vec = (reinterpret_cast<uintptr_t>(a->coords) % 16 == 0) && ((a->pitch * esz) % 16 == 0);
    // bulk copies need 16-byte aligned sources and sizes: float32 rows with pitch % 4 == 0 and, for the image rows,
    // a 16-byte aligned pitch (engine.row_pitch pads rows to 128 bytes, so stores built by the package qualify)
    const bool bulk_coords = a->coord_type == MDB_F32 && coords_vec;
    const bool packed_ok = a->packed && (reinterpret_cast<uintptr_t>(a->packed) % 16 == 0) && (a->packed_pitch % 4 == 0) &&
                           a->packed_pitch >= plan->n_groups;
    const bool bytes_ok = a->images && (reinterpret_cast<uintptr_t>(a->images) % 16 == 0) && (a->img_pitch % 16 == 0);
    plan->img = IMG_NONE;
    if (any_images) {
        if (bulk_coords && packed_ok) plan->img = IMG_PACKED;
        else if (a->images) plan->img = IMG_BYTES;
        else if (a->images16) plan->img = IMG_I16;
        else return fail(MDB_ERR_UNSUPPORTED, "packed image counts need float32 rows and 16-byte aligned pitches "
                                              "(pass the byte rows as well)");
    }
    plan->bulk = bulk_coords && (plan->img == IMG_NONE || plan->img == IMG_PACKED || (plan->img == IMG_BYTES && bytes_ok));
    plan->vec = coords_vec;
    if (plan->img == IMG_BYTES)
        plan->vec = plan->vec && (reinterpret_cast<uintptr_t>(a->images) % 4 == 0) && (a->img_pitch % 4 == 0);

    size_t hist_smem = 0;
    if (a->flags & MDB_HIST) {
        if (a->n_bins <= 0) return fail(MDB_ERR_ARGUMENT, "n_bins must be positive with MDB_HIST");
        hist_smem = (size_t)kLagChunk * a->n_bins * sizeof(unsigned int);
        if (a->bin_mode == MDB_BINS_EDGES) hist_smem += (size_t)(a->n_bins + 1) * sizeof(double);
        hist_smem += kHistTailBytes;
    }
    const int kf = kernel_flags(a->flags);
    const int nv_kernel = ((kf & (MDB_SUM_R2 | MDB_SUM_LEGENDRE)) ? 1 : 0) + ((kf & MDB_SUM_R4) ? 1 : 0) +
                          ((kf & MDB_SUM_SINC) ? 1 : 0);
    const int n_seg = a->n_segments > 1 ? a->n_segments : 1;
    if (n_seg > 1) {
        if (a->flags != MDB_SUM_SINC || !a->select || n_seg > 63)
            return fail(MDB_ERR_UNSUPPORTED, "n_segments > 1 goes with MDB_SUM_SINC alone, a select table and at most 63 "
                                             "sub-ensembles");
        if (!bulk_coords || !any_images || a->box || a->tuning.isf_float64)
            return fail(MDB_ERR_UNSUPPORTED, "several sub-ensembles per pass need the float32-displacement ISF kernel "
                                             "(nojump rows with one constant box, aligned float32 rows)");
        for (int d = 0; d < 3; ++d)
            if ((double)(float)a->box_const[d] != a->box_const[d])
                return fail(MDB_ERR_UNSUPPORTED, "several sub-ensembles per pass need a float32-representable box");
    }
    const size_t acc_smem = (size_t)nv_kernel * kLagChunk * kThreads * sizeof(double) * n_seg;
    for (int attempt = 0; attempt < 2; ++attempt) {
        const size_t stage_bytes = kCoordTileBytes + img_tile_bytes(plan->img);
        // ring depth: what fits next to the accumulators / histogram with two CTAs per SM, 3..8 stages
        plan->stages = 0;
        if (acc_smem + hist_smem + 3 * stage_bytes <= kSmemBudgetPerCta)
            plan->stages = (int)((kSmemBudgetPerCta - acc_smem - hist_smem) / stage_bytes);
        else if (acc_smem + hist_smem + 3 * stage_bytes <= 220 * 1024)
            plan->stages = (int)((220 * 1024 - acc_smem - hist_smem) / stage_bytes);   // one CTA per SM
        if (plan->stages > kMaxStages) plan->stages = kMaxStages;
        if (a->tuning.ring_stages >= 2 && a->tuning.ring_stages <= plan->stages) plan->stages = a->tuning.ring_stages;
        if (plan->stages >= 3 || !plan->bulk) break;
        if (n_seg > 1)
            return fail(MDB_ERR_UNSUPPORTED, "%d sub-ensembles per pass do not fit into shared memory", n_seg);
        // huge histograms: fall back to the direct-load kernel, which reads byte (or int16) counts
        plan->bulk = false;
        if (plan->img == IMG_PACKED) {
            if (a->images) plan->img = IMG_BYTES;
            else return fail(MDB_ERR_UNSUPPORTED, "n_bins=%d leaves no room for the bulk-copy ring and the direct-load "
                                                  "kernel needs the byte image rows", a->n_bins);
        }
    }
    if (n_seg > 1 && !plan->bulk)
        return fail(MDB_ERR_UNSUPPORTED, "several sub-ensembles per pass need the bulk-copy kernel (aligned rows)");
    const size_t bulk_smem = plan->stages * (size_t)(kCoordTileBytes + img_tile_bytes(plan->img)) + acc_smem + hist_smem;
    plan->smem = plan->bulk ? bulk_smem : hist_smem;
    if (plan->smem > 220 * 1024)
        return fail(MDB_ERR_UNSUPPORTED, "n_bins=%d needs %zu bytes of shared memory (limit 220 KiB)", a->n_bins,
                    plan->smem);

    int sms = 148;
    int dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    // Atom blocks: every CTA gets a whole number of 1024-atom tiles.  All CTAs cost about the same, so
    // the launch runs in ceil(CTAs / resident slots) waves of (tiles per CTA + fixed overhead): pick
    // the split with the smallest estimate (wave quantisation is the main loss for small launches).
    const long long base = (long long)a->n_origins * plan->n_chunks;
    const long long n_tiles = (plan->n_groups + kThreads - 1) / kThreads;
    const long long slots = (long long)MDB_BULK_CTAS_PER_SM * sms;
    long long best_tpb = n_tiles;
    double best_cost = 1e300;
    for (long long blocks = 1; blocks <= n_tiles && blocks <= 4096; ++blocks) {
        const long long tpb = (n_tiles + blocks - 1) / blocks;
        if ((n_tiles + tpb - 1) / tpb != blocks) continue;          // same split as a smaller block count
        const long long ctas = base * blocks;
        const long long waves = (ctas + slots - 1) / slots;
        const double cost = (double)waves * ((double)tpb + 1.0);     // ~1 tile of prologue/epilogue per CTA
        if (cost < best_cost * 0.999) {
            best_cost = cost;
            best_tpb = tpb;
        }
    }
    if (a->tuning.atom_blocks > 0) best_tpb = (n_tiles + a->tuning.atom_blocks - 1) / a->tuning.atom_blocks;
    const long long tiles_per_block = best_tpb;
    plan->groups_per_block = (int)(tiles_per_block * kThreads);
    plan->n_ablocks = (int)((n_tiles + tiles_per_block - 1) / tiles_per_block);
    return MDB_OK;
}

int check_args(const mdb_corr_args *a) {
    if (!a->coords || !a->rows) return fail(MDB_ERR_ARGUMENT, "coords and rows are required");
    if (a->coord_type != MDB_F32 && a->coord_type != MDB_F64) return fail(MDB_ERR_ARGUMENT, "bad coord_type");
    const bool any_images = a->images || a->packed || a->images16;
    if (any_images && a->coord_type != MDB_F32)
        return fail(MDB_ERR_ARGUMENT, "image counts go with float32 (wrapped) coordinates");
    if (any_images && !a->box && !(a->box_const[0] > 0.0 && a->box_const[1] > 0.0 && a->box_const[2] > 0.0))
        return fail(MDB_ERR_ARGUMENT, "images need box (per row) or a positive box_const");
    if (a->images && a->img_pitch < 3ll * a->n_atoms) return fail(MDB_ERR_ARGUMENT, "img_pitch < 3*n_atoms");
    if (a->images16 && a->img16_pitch < 3ll * a->n_atoms) return fail(MDB_ERR_ARGUMENT, "img16_pitch < 3*n_atoms");
    if (a->packed && a->packed_pitch < (a->n_atoms + 3) / 4) return fail(MDB_ERR_ARGUMENT, "packed_pitch < ceil(n_atoms/4)");
    if (a->select && (a->sel_pitch < a->n_atoms || a->sel_pitch % 4 != 0 ||
                      reinterpret_cast<uintptr_t>(a->select) % 4 != 0))
        return fail(MDB_ERR_ARGUMENT, "select needs sel_pitch >= n_atoms, a multiple of 4, and a 4-byte aligned base");
    if (a->flags & MDB_HIST) {
        if (a->bin_mode == MDB_BINS_UNIFORM && !(a->dx > 0.0)) return fail(MDB_ERR_ARGUMENT, "dx must be > 0");
        if (a->bin_mode == MDB_BINS_EDGES && !a->edges) return fail(MDB_ERR_ARGUMENT, "edges is required");
        if (a->bin_mode != MDB_BINS_UNIFORM && a->bin_mode != MDB_BINS_EDGES)
            return fail(MDB_ERR_ARGUMENT, "bad bin_mode");
    }
    return MDB_OK;
}

// Kernel parameters of a launch.  `partial` / `hist` are filled in by the caller.
CorrDev device_args(const mdb_corr_args *a, const Plan &plan) {
    CorrDev p;
    memset(&p, 0, sizeof(p));
    const bool hist = (a->flags & MDB_HIST) != 0;
    p.coords = a->coords;
    p.pitch = a->pitch;
    switch (plan.img) {
        case IMG_PACKED: p.images = a->packed; p.img_pitch = a->packed_pitch; break;
        case IMG_BYTES: p.images = a->images; p.img_pitch = a->img_pitch; break;
        case IMG_I16: p.images = a->images16; p.img_pitch = a->img16_pitch; break;
        default: p.images = nullptr; p.img_pitch = 0;
    }
    p.box = plan.img == IMG_NONE ? nullptr : a->box;
    const double scale[3] = {1.0, plan.img == IMG_PACKED ? 4.0 : 1.0, plan.img == IMG_PACKED ? 16.0 : 1.0};
    for (int d = 0; d < 3; ++d) p.box_const[d] = a->box_const[d] / scale[d];
    p.n_atoms = a->n_atoms;
    p.n_groups = plan.n_groups;
    p.groups_per_block = plan.groups_per_block;
    p.S = a->n_origins;
    p.L = a->n_lags;
    p.rows = a->rows;
    p.select = a->select;
    p.sel_pitch = a->sel_pitch;
    p.n_chunks = plan.n_chunks;
    p.n_ablocks = plan.n_ablocks;
    p.n_segments = a->n_segments > 1 ? a->n_segments : 1;
    p.q2 = a->q * a->q;
    // Float32 displacements (process_tile_f32d) on rows with one float32-representable box (what trajectory formats
    // store): the ISF alone, van Hove alone and van Hove + alpha_2 with uniform bins -- unless the caller asks for
    // the float64 chain.
    bool box_f32 = (plan.img == IMG_BYTES || plan.img == IMG_PACKED) && !a->box && plan.bulk;
    for (int d = 0; d < 3; ++d) box_f32 = box_f32 && (double)(float)a->box_const[d] == a->box_const[d];
    const bool hist_flags = (a->flags == MDB_HIST || a->flags == (MDB_SUM_R2 | MDB_SUM_R4 | MDB_HIST)) &&
                            a->bin_mode == MDB_BINS_UNIFORM;
    p.f32d = box_f32 && ((a->flags == MDB_SUM_SINC && !a->tuning.isf_float64) || (hist_flags && !a->tuning.hist_float64));
    p.no_shortcut = (a->tuning.variant & 4) ? 1 : 0;
    p.q2f = (float)p.q2;
    {
        const int dims[6] = {0, 1, 2, 0, 1, 2};
        for (int i = 0; i < 6; ++i) p.neg_box2[i] = -(float)(a->box_const[dims[i]] / scale[dims[i]]);
    }
    p.leg_order = a->legendre_order;
    p.n_bins = hist ? a->n_bins : 0;
    p.bin_mode = a->bin_mode;
    p.dx = a->dx;
    p.inv_dx = hist && a->dx > 0 ? 1.0 / a->dx : 0.0;
    p.inv_dx_f = (float)p.inv_dx;
    // margin of the float32 bin estimate: c1 * t + 1e-6 with c1 = twice the relative error bound of t
    // (3.9e-7 from float64 r^2, 8e-7 from float32 displacements; tests/test_histogram_estimate.py replays both)
    const double c1 = p.f32d ? 1.6e-6 : 8.0e-7, c2 = 1.0e-6;
    p.hist_c1 = (float)(-c1);
    p.hist_lim0 = (float)(0.5 - c2 - 0.5 * c1);
    p.ib_limit = 0x4B400000u + (uint32_t)p.n_bins;
    p.magic23 = 0x4B000000u;
    p.one = 1u;
    p.edges = a->edges;
    return p;
}

int launch_corr(const CorrDev &p, const Plan &plan, const mdb_corr_args *a, cudaStream_t st, LaunchTiming tm) {
    if (!plan.bulk) return corr_launch_ldg(p, plan, a->flags, a->coord_type, st, tm);
    switch (plan.img) {
        case IMG_NONE: return corr_launch_bulk_none(p, plan, a->flags, st, tm);
        case IMG_BYTES: return corr_launch_bulk_bytes(p, plan, a->flags, st, tm);
        default: return corr_launch_bulk_packed(p, plan, a->flags, st, tm);
    }
}

// =================================================================================================
// finish kernel
// =================================================================================================
struct FinishDev {
    const double *partial;         // [n_sums][n_ablocks] (component innermost: see index below)
    int n_ablocks;
    int comps;                     // sums per (origin, lag): 3, or 3 * sub-ensembles
    long long n_sums;              // S*L*comps
    unsigned long long *hist;      // device counts of this shard, replaced by the total over the shards; or null
    long long n_hist;
    double *sums_dev;              // device out
    double *sums_host;             // mapped host out, or null
    unsigned long long *flag_host; // mapped host flag, or null
    unsigned long long seq_host;   // value the flag takes when the results are complete
    unsigned int *ticket;          // two device counters, zero between launches
    const unsigned long long *corr_stamp;   // plans: where the correlation kernel left its start time, or null
    unsigned long long *stamps_host;        // plans: mapped host words [corr start, finish start, finish end] (ns)
    int use_comm;
    unsigned long long seq_comm;
    CommDev comm;
};

constexpr unsigned long long kFlagTimeout = ~0ull;

__device__ __forceinline__ void st_release_sys(unsigned long long *ptr, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(ptr), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *ptr) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(ptr) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long ld_relaxed_sys(const unsigned long long *ptr) {
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(ptr) : "memory");
    return v;
}

// The value of word i of this shard: a sum of the atom-block partials in fixed order, or a histogram count.
__device__ __forceinline__ unsigned long long local_word(const FinishDev &f, long long i) {
    if (i < f.n_sums) {
        const long long sl = i / f.comps;
        const int comp = (int)(i - f.comps * sl);
        double v = 0.0;
        for (int ab = 0; ab < f.n_ablocks; ++ab) v += f.partial[(sl * f.n_ablocks + ab) * f.comps + comp];
        return (unsigned long long)__double_as_longlong(v);
    }
    return f.hist[i - f.n_sums];
}

__device__ __forceinline__ void store_result(const FinishDev &f, long long i, unsigned long long v) {
    if (i < f.n_sums) {
        f.sums_dev[i] = __longlong_as_double((long long)v);
        if (f.sums_host) f.sums_host[i] = __longlong_as_double((long long)v);
    } else {
        f.hist[i - f.n_sums] = v;
    }
}

// Last CTA through a ticket counter.  ONE system-scope fence per CTA: the barrier orders every thread's stores before
// thread 0 at CTA scope, and thread 0's fence makes them visible system-wide before it takes its ticket (fences are
// cumulative).  With a fence in every thread -- 5000 fences behind stores to mapped host memory -- this kernel took
// 14 us (profiles/r2_launches_bench.csv).  The last CTA only RAISES FLAGS afterwards (st.release.sys, which orders its
// own earlier stores); it never reads what the other CTAs wrote, so there is no fence on the acquiring side.
__device__ __forceinline__ bool last_cta(unsigned int *ticket, int *s_flag) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();
        const bool last = atomicAdd(ticket, 1u) == gridDim.x - 1;
        *s_flag = last ? 1 : 0;
    }
    __syncthreads();
    return *s_flag != 0;
}

__device__ __forceinline__ void stamp_end(const FinishDev &f) {
    if (!f.stamps_host) return;
    f.stamps_host[0] = f.corr_stamp ? *f.corr_stamp : 0ull;
    f.stamps_host[2] = global_timer_ns();
}

__global__ void __launch_bounds__(256) finish_kernel(const FinishDev f) {
    __shared__ int s_flag;
    __shared__ int s_timeout;
    const long long n_total = f.n_sums + f.n_hist;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long first = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (f.stamps_host && first == 0) f.stamps_host[1] = global_timer_ns();      // ordered by last_cta's fence
    if (!f.use_comm) {
        for (long long i = first; i < f.n_sums; i += stride) store_result(f, i, local_word(f, i));
        if (last_cta(f.ticket, &s_flag) && threadIdx.x == 0) {
            f.ticket[0] = 0u;
            stamp_end(f);
            if (f.flag_host) st_release_sys(f.flag_host, f.seq_host);
        }
        return;
    }
    const CommDev &c = f.comm;
    const unsigned long long parity = f.seq_comm & 1ull;
    const unsigned long long slot = (parity * c.world + c.rank) * c.capacity;
    // 1. my words into slot `rank` of every rank's inbox (peer stores over NVLink; my own inbox included)
    for (long long i = first; i < n_total; i += stride) {
        const unsigned long long v = local_word(f, i);
        for (int r = 0; r < c.world; ++r) c.inbox[r][slot + i] = v;
    }
    // 2. when every CTA's stores are out: raise my flag in every inbox
    if (last_cta(f.ticket, &s_flag)) {
        if (threadIdx.x < c.world) st_release_sys(c.flags[threadIdx.x] + parity * c.world + c.rank, f.seq_comm);
    }
    // 3. wait for the flags of my own inbox (every CTA: the grid is small enough to be co-resident)
    if (threadIdx.x == 0) s_timeout = 0;
    __syncthreads();
    if (threadIdx.x < c.world) {
        const unsigned long long *flag = c.flags[c.rank] + parity * c.world + threadIdx.x;
        const long long t0 = clock64();
        while (ld_acquire_sys(flag) < f.seq_comm) {
            if (clock64() - t0 > c.timeout_cycles) { s_timeout = 1; break; }
            __nanosleep(20);
        }
    }
    __syncthreads();
    const bool timed_out = s_timeout != 0;
    // 4. add the slots in rank order: every rank gets the same bits
    if (!timed_out) {
        const unsigned long long *mine = c.inbox[c.rank] + parity * c.world * c.capacity;
        for (long long i = first; i < n_total; i += stride) {
            if (i < f.n_sums) {
                double v = 0.0;
                for (int r = 0; r < c.world; ++r) v += __longlong_as_double((long long)ld_relaxed_sys(mine + r * c.capacity + i));
                store_result(f, i, (unsigned long long)__double_as_longlong(v));
            } else {
                unsigned long long v = 0ull;
                for (int r = 0; r < c.world; ++r) v += ld_relaxed_sys(mine + r * c.capacity + i);
                store_result(f, i, v);
            }
        }
    }
    if (timed_out && threadIdx.x == 0) atomicExch(&f.ticket[2], 1u);
    if (last_cta(f.ticket + 1, &s_flag) && threadIdx.x == 0) {
        __threadfence();                 // the other CTAs' time-out marks, set before they took their tickets
        const bool any_timeout = atomicExch(&f.ticket[2], 0u) != 0u;
        f.ticket[0] = 0u;
        f.ticket[1] = 0u;
        stamp_end(f);
        if (f.flag_host) st_release_sys(f.flag_host, any_timeout ? kFlagTimeout : f.seq_host);
    }
}

int finish_grid(long long n_words) {
    int sms = 148, dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const long long want = (n_words + 255) / 256;
    return (int)(want < 1 ? 1 : (want > sms ? sms : want));       // co-resident: the CTAs of a comm launch wait
}

struct Profile {
    bool enabled = false;
    bool valid = false;
    cudaEvent_t start = nullptr, stop = nullptr;
};
Profile &profile() {
    static thread_local Profile pr;
    return pr;
}

}  // namespace
}  // namespace mdb

using namespace mdb;

// =================================================================================================
// one-shot API: caller-owned outputs and workspace
// =================================================================================================
extern "C" int mdb_corr_workspace_bytes(const mdb_corr_args *a, size_t *bytes) {
    Plan plan;
    if (int rc = make_plan(a, &plan)) return rc;
    if (!bytes) return fail(MDB_ERR_ARGUMENT, "bytes is NULL");
    // partial sums of the atom blocks + the ticket counters of the finish kernel
    const int n_seg = a->n_segments > 1 ? a->n_segments : 1;
    *bytes = (size_t)a->n_origins * a->n_lags * plan.n_ablocks * 3 * n_seg * sizeof(double) + 16;
    return MDB_OK;
}

extern "C" int mdb_shifted_corr(const mdb_corr_args *a, void *stream) {
    Plan plan;
    if (int rc = make_plan(a, &plan)) return rc;
    if (int rc = check_args(a)) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (!a->sums) return fail(MDB_ERR_ARGUMENT, "sums is required");
    const bool hist = (a->flags & MDB_HIST) != 0;
    if (hist && !a->hist) return fail(MDB_ERR_ARGUMENT, "hist is required with MDB_HIST");
    const int n_seg = a->n_segments > 1 ? a->n_segments : 1;
    const size_t partial_bytes = (size_t)a->n_origins * a->n_lags * plan.n_ablocks * 3 * n_seg * sizeof(double);
    if (!a->workspace || a->workspace_bytes < partial_bytes + 16)
        return fail(MDB_ERR_WORKSPACE, "workspace of %zu bytes required, %zu given", partial_bytes + 16,
                    a->workspace_bytes);
    CorrDev p = device_args(a, plan);
    p.partial = static_cast<double *>(a->workspace);
    p.hist = a->hist;
    unsigned int *ticket = reinterpret_cast<unsigned int *>(static_cast<char *>(a->workspace) + partial_bytes);
    MDB_CUDA_CHECK(cudaMemsetAsync(ticket, 0, 16, st));
    if (hist)
        MDB_CUDA_CHECK(cudaMemsetAsync(a->hist, 0,
                                       (size_t)a->n_origins * a->n_lags * a->n_bins * sizeof(unsigned long long), st));
    Profile &pr = profile();
    LaunchTiming tm;
    if (pr.enabled) { tm.start = pr.start; tm.stop = pr.stop; }
    if (int rc = launch_corr(p, plan, a, st, tm)) return rc;
    if (pr.enabled) pr.valid = true;
    FinishDev f;
    memset(&f, 0, sizeof(f));
    f.partial = p.partial;
    f.n_ablocks = plan.n_ablocks;
    f.comps = 3 * n_seg;
    f.n_sums = (long long)a->n_origins * a->n_lags * 3 * n_seg;
    f.sums_dev = a->sums;
    f.ticket = ticket;
    finish_kernel<<<finish_grid(f.n_sums), 256, 0, st>>>(f);
    MDB_CUDA_CHECK(cudaGetLastError());
    last_launch_count() = 2;
    return MDB_OK;
}

extern "C" int mdb_corr_last_launches(void) { return last_launch_count(); }

extern "C" int mdb_corr_profile(int enable) {
    Profile &pr = profile();
    if (enable && !pr.start) {
        MDB_CUDA_CHECK(cudaEventCreate(&pr.start));
        MDB_CUDA_CHECK(cudaEventCreate(&pr.stop));
    }
    pr.enabled = enable != 0;
    pr.valid = false;
    return MDB_OK;
}

extern "C" int mdb_corr_last_kernel_ms(float *ms) {
    Profile &pr = profile();
    if (!ms) return fail(MDB_ERR_ARGUMENT, "ms is NULL");
    if (!pr.valid) return fail(MDB_ERR_ARGUMENT, "no profiled launch (call mdb_corr_profile(1) first)");
    MDB_CUDA_CHECK(cudaEventSynchronize(pr.stop));
    MDB_CUDA_CHECK(cudaEventElapsedTime(ms, pr.start, pr.stop));
    return MDB_OK;
}

// =================================================================================================
// plans
// =================================================================================================
struct mdb_corr_plan {
    mdb_corr_args args;
    Plan plan;
    CorrDev dev;
    FinishDev fin;
    mdb_comm *comm = nullptr;
    void *device_block = nullptr;        // [partials][sums][hist][tickets]
    void *host_block = nullptr;          // mapped, page-locked: [flag (64 B)][sums][hist]
    double *sums_host = nullptr;
    unsigned long long *hist_host = nullptr;
    volatile unsigned long long *flag_host = nullptr;
    size_t hist_words = 0;
    unsigned long long seq = 0;          // runs launched
    unsigned long long *comm_seq = nullptr;
    cudaStream_t last_stream = nullptr;
    bool pending = false;
    int wait_mode = 0;                   // mdb_corr_plan_set_wait
    int grid = 1;
    cudaEvent_t events[2 * MDB_PLAN_TIME_SLOTS];
    bool have_events = false;
    int device = 0;
    unsigned long long host_ns[3] = {0, 0, 0};     // steady clock: launch entered, launch returned, result seen
};

static inline unsigned long long steady_ns() {
    return (unsigned long long)std::chrono::duration_cast<std::chrono::nanoseconds>(
               std::chrono::steady_clock::now().time_since_epoch()).count();
}

namespace mdb {
unsigned long long *comm_sequence(mdb_comm *comm);      // comm.cu: the launch counter all plans of a comm share
}

extern "C" int mdb_corr_plan_create(const mdb_corr_args *a, mdb_comm *comm, mdb_corr_plan **out) {
    if (!out) return fail(MDB_ERR_ARGUMENT, "plan is NULL");
    *out = nullptr;
    Plan plan;
    if (int rc = make_plan(a, &plan)) return rc;
    if (int rc = check_args(a)) return rc;
    mdb_corr_plan *pl = new (std::nothrow) mdb_corr_plan();
    if (!pl) return fail(MDB_ERR_CUDA, "out of host memory");
    pl->args = *a;
    pl->plan = plan;
    pl->comm = comm;
    cudaGetDevice(&pl->device);
    const bool hist = (a->flags & MDB_HIST) != 0;
    const int n_seg = a->n_segments > 1 ? a->n_segments : 1;
    const size_t n_sums = (size_t)a->n_origins * a->n_lags * 3 * n_seg;
    pl->hist_words = hist ? (size_t)a->n_origins * a->n_lags * a->n_bins : 0;
    const size_t partial_bytes = n_sums * plan.n_ablocks * sizeof(double);
    const size_t dev_bytes = partial_bytes + n_sums * 8 + pl->hist_words * 8 + 64;
    cudaError_t err = cudaMalloc(&pl->device_block, dev_bytes);
    if (err == cudaSuccess) err = cudaMemset(pl->device_block, 0, dev_bytes);
    if (err == cudaSuccess)
        err = cudaHostAlloc(&pl->host_block, 64 + n_sums * 8 + pl->hist_words * 8, cudaHostAllocMapped | cudaHostAllocPortable);
    if (err != cudaSuccess) {
        const int rc = fail(MDB_ERR_CUDA, "allocating the plan's buffers failed: %s", cudaGetErrorString(err));
        mdb_corr_plan_destroy(pl);
        return rc;
    }
    memset(pl->host_block, 0, 64 + n_sums * 8 + pl->hist_words * 8);
    char *d = static_cast<char *>(pl->device_block);
    pl->dev = device_args(a, plan);
    pl->dev.partial = reinterpret_cast<double *>(d);
    double *sums_dev = reinterpret_cast<double *>(d + partial_bytes);
    pl->dev.hist = hist ? reinterpret_cast<unsigned long long *>(d + partial_bytes + n_sums * 8) : nullptr;
    char *h = static_cast<char *>(pl->host_block);
    pl->flag_host = reinterpret_cast<volatile unsigned long long *>(h);
    pl->sums_host = reinterpret_cast<double *>(h + 64);
    pl->hist_host = hist ? reinterpret_cast<unsigned long long *>(h + 64 + n_sums * 8) : nullptr;

    FinishDev &f = pl->fin;
    memset(&f, 0, sizeof(f));
    f.partial = pl->dev.partial;
    f.n_ablocks = plan.n_ablocks;
    f.comps = 3 * n_seg;
    f.n_sums = (long long)n_sums;
    f.hist = pl->dev.hist;
    f.n_hist = 0;
    f.sums_dev = sums_dev;
    void *mapped = nullptr;
    err = cudaHostGetDevicePointer(&mapped, pl->host_block, 0);
    if (err != cudaSuccess) {
        const int rc = fail(MDB_ERR_CUDA, "cudaHostGetDevicePointer failed: %s", cudaGetErrorString(err));
        mdb_corr_plan_destroy(pl);
        return rc;
    }
    f.flag_host = reinterpret_cast<unsigned long long *>(mapped);
    f.sums_host = reinterpret_cast<double *>(static_cast<char *>(mapped) + 64);
    f.ticket = reinterpret_cast<unsigned int *>(d + partial_bytes + n_sums * 8 + pl->hist_words * 8);
    pl->dev.stamp = reinterpret_cast<unsigned long long *>(d + partial_bytes + n_sums * 8 + pl->hist_words * 8 + 32);
    f.corr_stamp = pl->dev.stamp;
    f.stamps_host = reinterpret_cast<unsigned long long *>(static_cast<char *>(mapped) + 8);
    if (comm) {
        if (int rc = comm_device_view(comm, &f.comm)) {
            mdb_corr_plan_destroy(pl);
            return rc;
        }
        if (n_sums + pl->hist_words > f.comm.capacity) {
            const int rc = fail(MDB_ERR_ARGUMENT, "the communicator's slots hold %llu words, this plan needs %zu",
                                f.comm.capacity, n_sums + pl->hist_words);
            mdb_corr_plan_destroy(pl);
            return rc;
        }
        f.use_comm = 1;
        f.n_hist = (long long)pl->hist_words;
        pl->comm_seq = comm_sequence(comm);
    }
    pl->grid = finish_grid(f.n_sums + f.n_hist);
    *out = pl;
    return MDB_OK;
}

extern "C" int mdb_corr_plan_launch(mdb_corr_plan *pl, void *stream, int32_t time_slot) {
    if (!pl) return fail(MDB_ERR_ARGUMENT, "plan is NULL");
    if (pl->pending) return fail(MDB_ERR_ARGUMENT, "the previous run of this plan has not been waited for");
    if (time_slot >= MDB_PLAN_TIME_SLOTS) return fail(MDB_ERR_ARGUMENT, "time_slot out of range");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    pl->host_ns[0] = steady_ns();
    LaunchTiming tm;
    if (time_slot >= 0) {
        // events are created in groups of 64 slots the first time a slot of the group is used (all 2048
        // cudaEventCreate calls at once cost a millisecond -- inside whatever the caller was timing)
        if (!pl->have_events) {
            for (int i = 0; i < 2 * MDB_PLAN_TIME_SLOTS; ++i) pl->events[i] = nullptr;
            pl->have_events = true;
        }
        if (!pl->events[2 * time_slot]) {
            const int first = time_slot & ~63;
            for (int i = 2 * first; i < 2 * (first + 64); ++i)
                if (!pl->events[i]) MDB_CUDA_CHECK(cudaEventCreate(&pl->events[i]));
        }
        tm.start = pl->events[2 * time_slot];
        tm.stop = pl->events[2 * time_slot + 1];
    }
    if (pl->hist_words) MDB_CUDA_CHECK(cudaMemsetAsync(pl->dev.hist, 0, pl->hist_words * 8, st));
    if (int rc = launch_corr(pl->dev, pl->plan, &pl->args, st, tm)) return rc;
    pl->fin.seq_host = ++pl->seq;
    if (pl->comm_seq) pl->fin.seq_comm = ++*pl->comm_seq;
    finish_kernel<<<pl->grid, 256, 0, st>>>(pl->fin);
    MDB_CUDA_CHECK(cudaGetLastError());
    if (pl->hist_words)
        MDB_CUDA_CHECK(cudaMemcpyAsync(pl->hist_host, pl->dev.hist, pl->hist_words * 8, cudaMemcpyDeviceToHost, st));
    if (pl->wait_mode == 2)
        MDB_CUDA_CHECK(cudaMemcpyAsync(pl->sums_host, pl->fin.sums_dev, (size_t)pl->fin.n_sums * 8,
                                       cudaMemcpyDeviceToHost, st));
    pl->last_stream = st;
    pl->pending = true;
    pl->host_ns[1] = steady_ns();
    return MDB_OK;
}

extern "C" int mdb_corr_plan_wait(mdb_corr_plan *pl) {
    if (!pl) return fail(MDB_ERR_ARGUMENT, "plan is NULL");
    if (!pl->pending) return fail(MDB_ERR_ARGUMENT, "nothing was launched");
    pl->pending = false;
    if (pl->wait_mode != 0) {
        MDB_CUDA_CHECK(cudaStreamSynchronize(pl->last_stream));
        pl->host_ns[2] = steady_ns();
        if (*pl->flag_host == kFlagTimeout) {
            *pl->flag_host = 0;
            return fail(MDB_ERR_TIMEOUT, "a peer GPU did not deliver its partial sums in time");
        }
        return MDB_OK;
    }
    // the finish kernel raises the flag through the mapped block as its last store: poll it (no driver call on the
    // fast path); look at the stream now and then so that a failed launch cannot hang the caller
    const auto t0 = std::chrono::steady_clock::now();
    unsigned long long spins = 0;
    for (;;) {
        const unsigned long long v = *pl->flag_host;
        if (v == pl->seq) break;
        if (v == kFlagTimeout) {
            *pl->flag_host = 0;
            return fail(MDB_ERR_TIMEOUT, "a peer GPU did not deliver its partial sums in time");
        }
#if defined(__x86_64__)
        __builtin_ia32_pause();
#endif
        if ((++spins & 0xffffu) == 0) {
            const cudaError_t q = cudaStreamQuery(pl->last_stream);
            if (q != cudaSuccess && q != cudaErrorNotReady)
                return fail(MDB_ERR_CUDA, "the correlation launch failed: %s", cudaGetErrorString(q));
            if (q == cudaSuccess && *pl->flag_host != pl->seq && *pl->flag_host != kFlagTimeout)
                return fail(MDB_ERR_CUDA, "the stream drained without the finish kernel raising its flag");
            if (std::chrono::steady_clock::now() - t0 > std::chrono::seconds(120))
                return fail(MDB_ERR_TIMEOUT, "no result after 120 s");
        }
    }
    pl->host_ns[2] = steady_ns();
    std::atomic_thread_fence(std::memory_order_acquire);
    if (pl->hist_words) MDB_CUDA_CHECK(cudaStreamSynchronize(pl->last_stream));     // the histogram's D2H copy
    return MDB_OK;
}

extern "C" int mdb_corr_plan_set_wait(mdb_corr_plan *pl, int32_t mode) {
    if (!pl || mode < 0 || mode > 2) return fail(MDB_ERR_ARGUMENT, "mode must be 0, 1 or 2");
    if (pl->pending) return fail(MDB_ERR_ARGUMENT, "a run is pending");
    pl->wait_mode = mode;
    void *mapped = nullptr;
    MDB_CUDA_CHECK(cudaHostGetDevicePointer(&mapped, pl->host_block, 0));
    pl->fin.sums_host = mode == 2 ? nullptr : reinterpret_cast<double *>(static_cast<char *>(mapped) + 64);
    return MDB_OK;
}

extern "C" int mdb_corr_plan_run(mdb_corr_plan *pl, void *stream, int32_t time_slot) {
    if (int rc = mdb_corr_plan_launch(pl, stream, time_slot)) return rc;
    return mdb_corr_plan_wait(pl);
}

extern "C" int mdb_corr_plan_kernel_ms(mdb_corr_plan *pl, int32_t time_slot, float *ms) {
    if (!pl || !ms) return fail(MDB_ERR_ARGUMENT, "NULL argument");
    if (time_slot < 0 || time_slot >= MDB_PLAN_TIME_SLOTS || !pl->have_events || !pl->events[2 * time_slot + 1])
        return fail(MDB_ERR_ARGUMENT, "no timed run in that slot");
    MDB_CUDA_CHECK(cudaEventSynchronize(pl->events[2 * time_slot + 1]));
    MDB_CUDA_CHECK(cudaEventElapsedTime(ms, pl->events[2 * time_slot], pl->events[2 * time_slot + 1]));
    return MDB_OK;
}

extern "C" const double *mdb_corr_plan_sums(const mdb_corr_plan *pl) { return pl ? pl->sums_host : nullptr; }
extern "C" const unsigned long long *mdb_corr_plan_hist(const mdb_corr_plan *pl) { return pl ? pl->hist_host : nullptr; }
extern "C" const double *mdb_corr_plan_sums_device(const mdb_corr_plan *pl) { return pl ? pl->fin.sums_dev : nullptr; }
extern "C" int mdb_corr_plan_launches(const mdb_corr_plan *pl) { return pl ? 2 : 0; }

extern "C" int mdb_corr_plan_timeline(const mdb_corr_plan *pl, unsigned long long *out) {
    if (!pl || !out) return fail(MDB_ERR_ARGUMENT, "NULL argument");
    if (pl->pending || pl->seq == 0) return fail(MDB_ERR_ARGUMENT, "no completed run");
    const volatile unsigned long long *stamps = pl->flag_host + 1;
    for (int i = 0; i < 3; ++i) out[i] = pl->host_ns[i];
    for (int i = 0; i < 3; ++i) out[3 + i] = stamps[i];
    return MDB_OK;
}

static __global__ void clock_kernel(unsigned long long *out) {
    out[1] = global_timer_ns();
    st_release_sys(out, 1ull);
}

extern "C" int mdb_clock_pair(void *stream, unsigned long long *gpu_ns, unsigned long long *host_before_ns,
                              unsigned long long *host_seen_ns) {
    if (!gpu_ns || !host_before_ns || !host_seen_ns) return fail(MDB_ERR_ARGUMENT, "NULL argument");
    unsigned long long *block = nullptr;
    MDB_CUDA_CHECK(cudaHostAlloc(reinterpret_cast<void **>(&block), 64, cudaHostAllocMapped));
    block[0] = block[1] = 0ull;
    void *mapped = nullptr;
    cudaError_t err = cudaHostGetDevicePointer(&mapped, block, 0);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (err == cudaSuccess) err = cudaStreamSynchronize(st);
    if (err == cudaSuccess) {
        *host_before_ns = steady_ns();
        clock_kernel<<<1, 1, 0, st>>>(static_cast<unsigned long long *>(mapped));
        err = cudaGetLastError();
    }
    if (err == cudaSuccess) {
        volatile unsigned long long *flag = block;
        const auto t0 = std::chrono::steady_clock::now();
        while (*flag == 0ull) {
            if (std::chrono::steady_clock::now() - t0 > std::chrono::seconds(10)) { err = cudaErrorTimeout; break; }
        }
        *host_seen_ns = steady_ns();
        std::atomic_thread_fence(std::memory_order_acquire);
        *gpu_ns = block[1];
    }
    cudaFreeHost(block);
    if (err != cudaSuccess) return fail(MDB_ERR_CUDA, "mdb_clock_pair: %s", cudaGetErrorString(err));
    return MDB_OK;
}

extern "C" int mdb_corr_plan_destroy(mdb_corr_plan *pl) {
    if (!pl) return MDB_OK;
    if (pl->pending && pl->last_stream) cudaStreamSynchronize(pl->last_stream);
    if (pl->have_events)
        for (int i = 0; i < 2 * MDB_PLAN_TIME_SLOTS; ++i)
            if (pl->events[i]) cudaEventDestroy(pl->events[i]);
    if (pl->device_block) cudaFree(pl->device_block);
    if (pl->host_block) cudaFreeHost(pl->host_block);
    delete pl;
    return MDB_OK;
}
