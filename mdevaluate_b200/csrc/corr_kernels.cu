// Correlation kernels: per-(origin, lag) sums of r^2, r^4, sinc(q r) and van Hove histograms.
//
// Work decomposition (one CTA = one (origin s, lag chunk c, atom block ab)):
//   * a thread owns groups of 4 consecutive atoms = 12 interleaved xyz values = three 128-bit
//     loads per frame row, fully coalesced (a warp reads 1536 contiguous bytes per row);
//   * the unwrapped origin positions of the group are kept in registers (12 doubles) and reused
//     for the up to kLagChunk lag frames of the chunk; the lag rows are streamed with
//     L1::no_allocate loads, all chunk rows of a group are independent loads in flight;
//   * CTAs of the same (s, ab) and different chunks are adjacent in launch order, so the origin
//     row is fetched from HBM once and re-read from L2 by the other chunks;
//   * per-lag accumulators live in registers for the whole atom loop; there is ONE block
//     reduction per CTA at the end, written as a partial; a second tiny kernel adds the
//     partials of the atom blocks in fixed order (bit-reproducible, no float atomics);
//   * histograms are accumulated in shared memory (u32, one table per lag of the chunk) and
//     flushed with integer atomics.
//
// Arithmetic: positions are widened to float64, unwrapped as fma(-n, box, x) (bit-identical to
// pbc.nojump's float64 frame, pbc.py:235-247), the displacement and r^2 are float64.  sinc is
// evaluated in float32 from float32(q^2 r^2) and summed per group before it joins the float64
// accumulator.  Histogram bins follow utils.histogram exactly (utils.py:328-338): a float32
// estimate decides every sample that is not within rounding distance of a bin edge, the rest
// take the exact float64 path (np.floor_divide semantics or np.histogram edge search).
#include "common.cuh"

namespace mdb {
namespace {

constexpr int kThreads = 256;
constexpr int kLagChunk = 8;
constexpr int kWarps = kThreads / 32;

struct CorrDev {
    const void *coords;
    long long pitch;
    const int8_t *images;
    long long img_pitch;
    const double *box;
    int n_atoms;
    int n_groups;        // ceil(n_atoms / 4)
    int groups_per_block;
    int S, L;
    const int *rows;
    int n_chunks, n_ablocks;
    double q2;
    int n_bins;
    int bin_mode;
    double dx, inv_dx;
    float inv_dx_f;
    const double *edges;
    double *partial;     // [S*L][n_ablocks][3]
    unsigned long long *hist;
};

// ---- loading one group of 4 atoms as 12 unwrapped float64 components --------------------------
template <typename CT, bool IMG, bool VEC, bool STREAM>
struct GroupLoader;

template <bool IMG, bool VEC, bool STREAM>
struct GroupLoader<float, IMG, VEC, STREAM> {
    // raw values kept for the float32-semantics histogram of raw frames
    __device__ static __forceinline__ void load(const CorrDev &p, int row, int g, int nv, const double *bx,
                                                double (&u)[12], float (&raw)[12]) {
        const float *base = static_cast<const float *>(p.coords) + (long long)row * p.pitch + 12ll * g;
        if (VEC && nv == 4) {
            const float4 *b4 = reinterpret_cast<const float4 *>(base);
            float4 a, b, c;
            if (STREAM) { a = ld_stream(b4); b = ld_stream(b4 + 1); c = ld_stream(b4 + 2); }
            else { a = __ldg(b4); b = __ldg(b4 + 1); c = __ldg(b4 + 2); }
            raw[0] = a.x; raw[1] = a.y; raw[2] = a.z; raw[3] = a.w;
            raw[4] = b.x; raw[5] = b.y; raw[6] = b.z; raw[7] = b.w;
            raw[8] = c.x; raw[9] = c.y; raw[10] = c.z; raw[11] = c.w;
        } else {
#pragma unroll
            for (int c = 0; c < 12; ++c) raw[c] = (c < 3 * nv) ? __ldg(base + c) : 0.0f;
        }
        if (IMG) {
            const int8_t *ib = p.images + (long long)row * p.img_pitch + 12ll * g;
            uint32_t w[3];
            if (VEC && nv == 4) {
                const uint32_t *iw = reinterpret_cast<const uint32_t *>(ib);
                if (STREAM) { w[0] = ld_stream(iw); w[1] = ld_stream(iw + 1); w[2] = ld_stream(iw + 2); }
                else { w[0] = __ldg(iw); w[1] = __ldg(iw + 1); w[2] = __ldg(iw + 2); }
            } else {
                w[0] = w[1] = w[2] = 0u;
#pragma unroll
                for (int c = 0; c < 12; ++c)
                    if (c < 3 * nv) w[c >> 2] |= ((uint32_t)(uint8_t)ib[c]) << (8 * (c & 3));
            }
#pragma unroll
            for (int c = 0; c < 12; ++c) {
                const double n = int_to_double(sbyte(w[c >> 2], c & 3));
                u[c] = fma(-n, bx[c % 3], (double)raw[c]);
            }
        } else {
#pragma unroll
            for (int c = 0; c < 12; ++c) u[c] = (double)raw[c];
        }
    }
};

template <bool VEC, bool STREAM>
struct GroupLoader<double, false, VEC, STREAM> {
    __device__ static __forceinline__ void load(const CorrDev &p, int row, int g, int nv, const double *,
                                                double (&u)[12], float (&)[12]) {
        const double *base = static_cast<const double *>(p.coords) + (long long)row * p.pitch + 12ll * g;
        if (VEC && nv == 4) {
            const double2 *b2 = reinterpret_cast<const double2 *>(base);
#pragma unroll
            for (int j = 0; j < 6; ++j) {
                const double2 v = STREAM ? ld_stream(b2 + j) : __ldg(b2 + j);
                u[2 * j] = v.x;
                u[2 * j + 1] = v.y;
            }
        } else {
#pragma unroll
            for (int c = 0; c < 12; ++c) u[c] = (c < 3 * nv) ? __ldg(base + c) : 0.0;
        }
    }
};

// ---- histogram binning ------------------------------------------------------------------------
// np.floor_divide(r, dx) for r >= 0, dx > 0: floor of the exact real quotient (npy_divmod is
// fmod-based, so it never rounds the quotient up across an integer).
__device__ __forceinline__ long long exact_floor_div(double r, double dx, double inv_dx) {
    double k = floor(r * inv_dx);
    if (fma(-k, dx, r) < 0.0) k -= 1.0;              // quotient estimate was rounded up
    else if (fma(-(k + 1.0), dx, r) >= 0.0) k += 1.0;  // or down
    return (long long)k;
}

// np.histogram(r, bins=edges): bin i holds edges[i] <= r < edges[i+1], the last bin is closed.
__device__ __forceinline__ int edge_bin(double r, const double *edges, int n_bins) {
    if (!(r >= edges[0]) || r > edges[n_bins]) return -1;
    if (r == edges[n_bins]) return n_bins - 1;
    int lo = 0, hi = n_bins;  // invariant: edges[lo] <= r < edges[hi]
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (r >= edges[mid]) lo = mid; else hi = mid;
    }
    return lo;
}

template <typename CT, bool IMG, int FLAGS, bool VEC>
__global__ void __launch_bounds__(kThreads, 2) corr_kernel(const CorrDev p) {
    constexpr bool kR2 = (FLAGS & MDB_SUM_R2) != 0;
    constexpr bool kR4 = (FLAGS & MDB_SUM_R4) != 0;
    constexpr bool kSinc = (FLAGS & MDB_SUM_SINC) != 0;
    constexpr bool kHist = (FLAGS & MDB_HIST) != 0;
    // raw float32 frames: the reference computes r in float32 (correlation.py:246-247 on float32
    // arrays), so the histogram replays those float32 operations to land in the same bins.
    constexpr bool kHistF32 = kHist && sizeof(CT) == 4 && !IMG;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ int s_row[kLagChunk + 1];
    __shared__ double s_box[kLagChunk + 1][3];
    __shared__ double s_red[kWarps][kLagChunk * 3];

    const int bid = blockIdx.x;
    const int c = bid % p.n_chunks;
    const int rest = bid / p.n_chunks;
    const int ab = rest % p.n_ablocks;
    const int s = rest / p.n_ablocks;
    const int l0 = (int)((long long)c * p.L / p.n_chunks);
    const int l1 = (int)((long long)(c + 1) * p.L / p.n_chunks);
    const int nl = l1 - l0;
    const int tid = threadIdx.x;

    unsigned int *s_hist = reinterpret_cast<unsigned int *>(smem_raw);
    double *s_edges = reinterpret_cast<double *>(smem_raw + (size_t)kLagChunk * p.n_bins * sizeof(unsigned int));

    if (tid <= kLagChunk) {
        // slot 0: the origin row, slots 1..nl: the lag rows of this chunk
        const int l = (tid == 0) ? 0 : l0 + tid - 1;
        const int row = (tid == 0 || tid - 1 < nl) ? p.rows[(long long)s * p.L + l] : 0;
        s_row[tid] = row;
        if (IMG) {
            s_box[tid][0] = p.box[3ll * row];
            s_box[tid][1] = p.box[3ll * row + 1];
            s_box[tid][2] = p.box[3ll * row + 2];
        }
    }
    if (kHist) {
        for (int i = tid; i < kLagChunk * p.n_bins; i += kThreads) s_hist[i] = 0u;
        if (p.bin_mode == MDB_BINS_EDGES)
            for (int i = tid; i <= p.n_bins; i += kThreads) s_edges[i] = p.edges[i];
    }
    __syncthreads();

    double acc2[kLagChunk], acc4[kLagChunk], accs[kLagChunk];
#pragma unroll
    for (int k = 0; k < kLagChunk; ++k) acc2[k] = acc4[k] = accs[k] = 0.0;

    const int g_begin = ab * p.groups_per_block;
    const int g_end = min(p.n_groups, g_begin + p.groups_per_block);
    const float q2f_unused = 0.f; (void)q2f_unused;

    for (int g = g_begin + tid; g < g_end; g += kThreads) {
        const int nv = min(4, p.n_atoms - 4 * g);
        double u0[12];
        float raw0[12];
        GroupLoader<CT, IMG, VEC, false>::load(p, s_row[0], g, nv, s_box[0], u0, raw0);
#pragma unroll
        for (int k = 0; k < kLagChunk; ++k) {
            if (k < nl) {
                double ut[12];
                float rawt[12];
                GroupLoader<CT, IMG, VEC, true>::load(p, s_row[k + 1], g, nv, s_box[k + 1], ut, rawt);
                float sinc4 = 0.0f;
#pragma unroll
                for (int a = 0; a < 4; ++a) {
                    if (a < nv) {
                        const double ddx = u0[3 * a] - ut[3 * a];
                        const double ddy = u0[3 * a + 1] - ut[3 * a + 1];
                        const double ddz = u0[3 * a + 2] - ut[3 * a + 2];
                        const double r2 = fma(ddz, ddz, fma(ddy, ddy, ddx * ddx));
                        if (kR2) acc2[k] += r2;
                        if (kR4) acc4[k] = fma(r2, r2, acc4[k]);
                        if (kSinc) sinc4 += sinc_of_square((float)(r2 * p.q2));
                        if (kHist) {
                            int bin = -1;
                            if (kHistF32) {
                                // float32 replay: vec = start - end; (vec**2).sum(-1) ** .5, all float32
                                const float vx = __fsub_rn(raw0[3 * a], rawt[3 * a]);
                                const float vy = __fsub_rn(raw0[3 * a + 1], rawt[3 * a + 1]);
                                const float vz = __fsub_rn(raw0[3 * a + 2], rawt[3 * a + 2]);
                                const float q = __fadd_rn(__fadd_rn(__fmul_rn(vx, vx), __fmul_rn(vy, vy)),
                                                          __fmul_rn(vz, vz));
                                const double r = (double)__fsqrt_rn(q);
                                if (p.bin_mode == MDB_BINS_UNIFORM) {
                                    const long long kk = exact_floor_div(r, p.dx, p.inv_dx);
                                    bin = (kk < p.n_bins) ? (int)kk : -1;
                                } else {
                                    bin = edge_bin(r, s_edges, p.n_bins);
                                }
                            } else if (p.bin_mode == MDB_BINS_UNIFORM) {
                                // float32 estimate; exact path only near a bin edge
                                const float t = __fsqrt_rn((float)r2) * p.inv_dx_f;
                                const float fl = floorf(t);
                                const float fr = t - fl;
                                const float margin = fmaf(t, 6.0e-7f, 1.0e-6f);
                                if (fr > margin && fr < 1.0f - margin) {
                                    bin = (t < (float)p.n_bins) ? (int)fl : -1;
                                } else {
                                    const double r2x = __dadd_rn(__dadd_rn(__dmul_rn(ddx, ddx), __dmul_rn(ddy, ddy)),
                                                                 __dmul_rn(ddz, ddz));
                                    const long long kk = exact_floor_div(sqrt(r2x), p.dx, p.inv_dx);
                                    bin = (kk < p.n_bins) ? (int)kk : -1;
                                }
                            } else {
                                const double r2x = __dadd_rn(__dadd_rn(__dmul_rn(ddx, ddx), __dmul_rn(ddy, ddy)),
                                                             __dmul_rn(ddz, ddz));
                                bin = edge_bin(sqrt(r2x), s_edges, p.n_bins);
                            }
                            if (bin >= 0) atomicAdd(&s_hist[k * p.n_bins + bin], 1u);
                        }
                    }
                }
                if (kSinc) accs[k] += (double)sinc4;
            }
        }
    }

    // ---- one block reduction per CTA --------------------------------------------------------
    const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
    for (int k = 0; k < kLagChunk; ++k) {
        if (k < nl) {
            const double v2 = kR2 ? warp_sum(acc2[k]) : 0.0;
            const double v4 = kR4 ? warp_sum(acc4[k]) : 0.0;
            const double vs = kSinc ? warp_sum(accs[k]) : 0.0;
            if (lane == 0) {
                s_red[warp][3 * k] = v2;
                s_red[warp][3 * k + 1] = v4;
                s_red[warp][3 * k + 2] = vs;
            }
        }
    }
    __syncthreads();
    if (tid < 3 * nl) {
        double v = 0.0;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) v += s_red[w][tid];
        const int k = tid / 3, comp = tid - 3 * k;
        p.partial[(((long long)s * p.L + l0 + k) * p.n_ablocks + ab) * 3 + comp] = v;
    }
    if (kHist) {
        for (int i = tid; i < nl * p.n_bins; i += kThreads) {
            const unsigned int cnt = s_hist[i];
            if (cnt) {
                const int k = i / p.n_bins, b = i - k * p.n_bins;
                atomicAdd(&p.hist[((long long)s * p.L + l0 + k) * p.n_bins + b], (unsigned long long)cnt);
            }
        }
    }
}

__global__ void reduce_partials_kernel(const double *__restrict__ partial, double *__restrict__ sums,
                                       long long n_out, int n_ablocks) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // over S*L*3
    if (i >= n_out) return;
    const long long sl = i / 3;
    const int comp = (int)(i - 3 * sl);
    double v = 0.0;
    for (int ab = 0; ab < n_ablocks; ++ab) v += partial[(sl * n_ablocks + ab) * 3 + comp];
    sums[i] = v;
}

struct Plan {
    int n_groups, n_chunks, n_ablocks, groups_per_block;
    size_t smem;
};

int make_plan(const mdb_corr_args *a, Plan *plan) {
    if (!a) return fail(MDB_ERR_ARGUMENT, "args is NULL");
    if (a->n_atoms <= 0 || a->n_origins <= 0 || a->n_lags <= 0)
        return fail(MDB_ERR_ARGUMENT, "n_atoms, n_origins and n_lags must be positive");
    if (!(a->flags & (MDB_SUM_R2 | MDB_SUM_R4 | MDB_SUM_SINC | MDB_HIST)))
        return fail(MDB_ERR_ARGUMENT, "flags selects nothing to compute");
    if (a->pitch < 3ll * a->n_atoms) return fail(MDB_ERR_ARGUMENT, "pitch < 3*n_atoms");
    plan->n_groups = (a->n_atoms + 3) / 4;
    plan->n_chunks = (a->n_lags + kLagChunk - 1) / kLagChunk;
    int sms = 148;
    int dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    // enough CTAs for ~8 per SM, but at least ~4 iterations of work per thread
    const long long base = (long long)a->n_origins * plan->n_chunks;
    long long want = (8ll * sms + base - 1) / base;
    const long long max_blocks = (plan->n_groups + 4ll * kThreads - 1) / (4ll * kThreads);
    if (want > max_blocks) want = max_blocks;
    if (want < 1) want = 1;
    plan->n_ablocks = (int)want;
    plan->groups_per_block = (plan->n_groups + plan->n_ablocks - 1) / plan->n_ablocks;
    plan->smem = 0;
    if (a->flags & MDB_HIST) {
        if (a->n_bins <= 0) return fail(MDB_ERR_ARGUMENT, "n_bins must be positive with MDB_HIST");
        plan->smem = (size_t)kLagChunk * a->n_bins * sizeof(unsigned int);
        if (a->bin_mode == MDB_BINS_EDGES) plan->smem += (size_t)(a->n_bins + 1) * sizeof(double);
        if (plan->smem > 200 * 1024)
            return fail(MDB_ERR_UNSUPPORTED, "n_bins=%d needs %zu bytes of shared memory (limit 200 KiB)",
                        a->n_bins, plan->smem);
    }
    return MDB_OK;
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

template <typename CT, bool IMG, int FLAGS>
int launch_flags(const CorrDev &p, const Plan &plan, bool vec, cudaStream_t st) {
    const long long blocks = (long long)p.S * p.n_chunks * p.n_ablocks;
    if (blocks > 2147483647ll) return fail(MDB_ERR_UNSUPPORTED, "grid too large");
    auto kern = vec ? corr_kernel<CT, IMG, FLAGS, true> : corr_kernel<CT, IMG, FLAGS, false>;
    if (plan.smem > 40 * 1024)
        MDB_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.smem));
    Profile &pr = profile();
    if (pr.enabled) MDB_CUDA_CHECK(cudaEventRecord(pr.start, st));
    kern<<<(unsigned)blocks, kThreads, plan.smem, st>>>(p);
    MDB_CUDA_CHECK(cudaGetLastError());
    if (pr.enabled) {
        MDB_CUDA_CHECK(cudaEventRecord(pr.stop, st));
        pr.valid = true;
    }
    return MDB_OK;
}

template <typename CT, bool IMG>
int launch_typed(const CorrDev &p, const Plan &plan, int flags, bool vec, cudaStream_t st) {
    constexpr int R2 = MDB_SUM_R2, R4 = MDB_SUM_R4, SC = MDB_SUM_SINC, H = MDB_HIST;
    switch (flags) {
        case R2: return launch_flags<CT, IMG, R2>(p, plan, vec, st);
        case SC: return launch_flags<CT, IMG, SC>(p, plan, vec, st);
        case R2 | R4: return launch_flags<CT, IMG, R2 | R4>(p, plan, vec, st);
        case H: return launch_flags<CT, IMG, H>(p, plan, vec, st);
        default:
            // any other mix: the scalar sums together, optionally with the histogram
            if (flags & H) return launch_flags<CT, IMG, R2 | R4 | SC | H>(p, plan, vec, st);
            return launch_flags<CT, IMG, R2 | R4 | SC>(p, plan, vec, st);
    }
}

}  // namespace
}  // namespace mdb

using namespace mdb;

extern "C" int mdb_corr_workspace_bytes(const mdb_corr_args *a, size_t *bytes) {
    Plan plan;
    if (int rc = make_plan(a, &plan)) return rc;
    if (!bytes) return fail(MDB_ERR_ARGUMENT, "bytes is NULL");
    *bytes = (size_t)a->n_origins * a->n_lags * plan.n_ablocks * 3 * sizeof(double);
    return MDB_OK;
}

extern "C" int mdb_shifted_corr(const mdb_corr_args *a, void *stream) {
    Plan plan;
    if (int rc = make_plan(a, &plan)) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (!a->coords || !a->rows || !a->sums) return fail(MDB_ERR_ARGUMENT, "coords, rows and sums are required");
    if (a->coord_type != MDB_F32 && a->coord_type != MDB_F64) return fail(MDB_ERR_ARGUMENT, "bad coord_type");
    if (a->images && a->coord_type != MDB_F32)
        return fail(MDB_ERR_ARGUMENT, "image counts go with float32 (wrapped) coordinates");
    if (a->images && !a->box) return fail(MDB_ERR_ARGUMENT, "box is required with images");
    if (a->images && a->img_pitch < 3ll * a->n_atoms) return fail(MDB_ERR_ARGUMENT, "img_pitch < 3*n_atoms");
    const bool hist = (a->flags & MDB_HIST) != 0;
    if (hist) {
        if (!a->hist) return fail(MDB_ERR_ARGUMENT, "hist is required with MDB_HIST");
        if (a->bin_mode == MDB_BINS_UNIFORM && !(a->dx > 0.0)) return fail(MDB_ERR_ARGUMENT, "dx must be > 0");
        if (a->bin_mode == MDB_BINS_EDGES && !a->edges) return fail(MDB_ERR_ARGUMENT, "edges is required");
        if (a->bin_mode != MDB_BINS_UNIFORM && a->bin_mode != MDB_BINS_EDGES)
            return fail(MDB_ERR_ARGUMENT, "bad bin_mode");
    }
    const size_t need = (size_t)a->n_origins * a->n_lags * plan.n_ablocks * 3 * sizeof(double);
    if (!a->workspace || a->workspace_bytes < need)
        return fail(MDB_ERR_WORKSPACE, "workspace of %zu bytes required, %zu given", need, a->workspace_bytes);

    CorrDev p;
    p.coords = a->coords;
    p.pitch = a->pitch;
    p.images = a->images;
    p.img_pitch = a->img_pitch;
    p.box = a->box;
    p.n_atoms = a->n_atoms;
    p.n_groups = plan.n_groups;
    p.groups_per_block = plan.groups_per_block;
    p.S = a->n_origins;
    p.L = a->n_lags;
    p.rows = a->rows;
    p.n_chunks = plan.n_chunks;
    p.n_ablocks = plan.n_ablocks;
    p.q2 = a->q * a->q;
    p.n_bins = hist ? a->n_bins : 0;
    p.bin_mode = a->bin_mode;
    p.dx = a->dx;
    p.inv_dx = hist && a->dx > 0 ? 1.0 / a->dx : 0.0;
    p.inv_dx_f = (float)p.inv_dx;
    p.edges = a->edges;
    p.partial = static_cast<double *>(a->workspace);
    p.hist = a->hist;

    const size_t esz = a->coord_type == MDB_F32 ? 4 : 8;
    // 128-bit loads need 16-byte aligned rows: base pointer and pitch (and the image rows: 4 bytes)
    bool vec = (reinterpret_cast<uintptr_t>(a->coords) % 16 == 0) && ((a->pitch * esz) % 16 == 0);
    if (a->images)
        vec = vec && (reinterpret_cast<uintptr_t>(a->images) % 4 == 0) && (a->img_pitch % 4 == 0);

    int launches = 0;
    if (hist) {
        MDB_CUDA_CHECK(cudaMemsetAsync(a->hist, 0,
                                       (size_t)a->n_origins * a->n_lags * a->n_bins * sizeof(unsigned long long), st));
    }
    int rc;
    if (a->coord_type == MDB_F64) rc = launch_typed<double, false>(p, plan, a->flags, vec, st);
    else if (a->images) rc = launch_typed<float, true>(p, plan, a->flags, vec, st);
    else rc = launch_typed<float, false>(p, plan, a->flags, vec, st);
    if (rc) return rc;
    ++launches;
    const long long n_out = (long long)a->n_origins * a->n_lags * 3;
    reduce_partials_kernel<<<(unsigned)((n_out + 255) / 256), 256, 0, st>>>(p.partial, a->sums, n_out, plan.n_ablocks);
    MDB_CUDA_CHECK(cudaGetLastError());
    ++launches;
    last_launch_count() = launches;
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
