// nojump kernels: periodic-jump detection + running image counts, and the float64 unwrap.
//
// Producer (reader.py:165-186 + the prefix sum of pbc.py:235-237).  Every (atom, dimension) is an
// independent running sum over time, so the kernel is parallel over the 3N components and
// sequential over frames: a thread owns 4 consecutive floats of the row (one 128-bit load per
// frame, warps read 512 contiguous bytes), keeps `prev` and the running count in registers and
// walks the frames of the segment; the loads of the next frames do not depend on the running
// state, so several rows are in flight per thread.  The jump itself must be bit-exact with
//     ((curr - prev) / box).round().astype(int8)          (reader.py:178)
// in the dtype NumPy uses: float32 subtract, IEEE float32 divide (float64 divide when the box is
// float64), round-half-even.  A division-free test decides the common case: |curr - prev| <
// 0.49 * box implies |quotient| < 0.5 after rounding, i.e. a jump of 0.
#include "common.cuh"

namespace mdb {
namespace {

constexpr int kThreads = 256;
constexpr int kBatch = 8;

struct NojumpDev {
    const float *coords;
    long long pitch;
    int n_frames;
    int n_vals;  // 3 * n_atoms
    float boxf[3];
    double boxd[3];
    int first_segment;
    const float *whole0;
    float *prev;
    int *cum;
    const int *out_row;
    uint8_t *images;   // excess-128
    long long img_pitch;
    int8_t *jumps;
    float *coords_out;
    long long pitch_out;
    int *overflow;
    int16_t *images16;   // plain int16 counts (the fallback beyond +-127), general kernel only
    long long img16_pitch;
};

template <bool BOXF32>
__device__ __forceinline__ int jump_of(float diff, float bf, double bd, float safe) {
    if (fabsf(diff) < safe) return 0;
    if (BOXF32) return (int)(int8_t)(int)rintf(__fdiv_rn(diff, bf));
    return (int)(int8_t)(int)rint(__ddiv_rn((double)diff, bd));
}

// The rare path (some component moved by ~half a box): kept out of line so that the unrolled
// streaming loop stays small enough for the instruction cache.
template <bool BOXF32>
__device__ __noinline__ int4 slow_jumps(float d0, float d1, float d2, float d3, int e0, float bx, float by, float bz,
                                        double dbx, double dby, double dbz) {
    const float bf[3] = {bx, by, bz};
    const double bd[3] = {dbx, dby, dbz};
    const float d[4] = {d0, d1, d2, d3};
    int out[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int dim = (e0 + j) % 3;
        const float safe = 0.49f * fabsf(BOXF32 ? bf[dim] : (float)bd[dim]);
        out[j] = jump_of<BOXF32>(d[j], bf[dim], bd[dim], safe);
    }
    return make_int4(out[0], out[1], out[2], out[3]);
}

__device__ __forceinline__ uint32_t pack4(int a, int b, int c, int d) {
    return ((uint32_t)a & 0xffu) | (((uint32_t)b & 0xffu) << 8) | (((uint32_t)c & 0xffu) << 16) |
           (((uint32_t)d & 0xffu) << 24);
}

// VEC: every thread moves whole float4s (rows are 16-byte aligned, all pitches multiples of 4): the
// last thread may run up to 3 components into the row padding -- they are read and written inside
// the padded rows but never reach the carry buffers or the overflow flag.  !VEC: guarded scalars.
// PLAIN: the dense case (every frame kept in place, counts only) with the per-frame work cut to
// load / 4 subtract+compare / one packed store; otherwise the general case (row map, per-step
// jumps, compaction copy), decided per frame at run time.
template <bool BOXF32, bool VEC, bool PLAIN>
__global__ void __launch_bounds__(kThreads) nojump_scan_kernel(const NojumpDev p) {
    const int i = blockIdx.x * kThreads + threadIdx.x;
    const int e0 = 4 * i;
    if (e0 >= p.n_vals) return;
    const int nv = min(4, p.n_vals - e0);

    float safe[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int d = (e0 + j) % 3;
        safe[j] = 0.49f * fabsf(BOXF32 ? p.boxf[d] : (float)p.boxd[d]);
    }

    float prev[4];
    int cum[4];
    bool overflow = false;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        prev[j] = 0.f;
        cum[j] = 0;
        if (j < nv) {
            if (p.first_segment) {
                // reader.py:166: prev = trajectory[0].whole (frame 0 itself for single-atom residues)
                prev[j] = p.whole0 ? p.whole0[e0 + j] : p.coords[e0 + j];
            } else {
                prev[j] = p.prev[e0 + j];
                cum[j] = p.cum[e0 + j];
            }
        }
    }
    // the packed excess-128 word of the running counts only changes on the rare path
    uint32_t packed = pack4(cum[0] + MDB_IMAGE_BIAS, cum[1] + MDB_IMAGE_BIAS, cum[2] + MDB_IMAGE_BIAS,
                            cum[3] + MDB_IMAGE_BIAS);

    auto load_row = [&](int t, float (&v)[4]) {
        const float *row = p.coords + (long long)t * p.pitch + e0;
        if (VEC) {
            const float4 q = ld_stream(reinterpret_cast<const float4 *>(row));
            v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) v[j] = (j < nv) ? __ldg(row + j) : 0.f;
        }
    };
    auto step = [&](int t, const float (&cur)[4]) {
        float diff[4];
        bool any = false;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            diff[j] = __fsub_rn(cur[j], prev[j]);
            any |= !(fabsf(diff[j]) < safe[j]);
            prev[j] = cur[j];
        }
        int jmp[4] = {0, 0, 0, 0};
        if (any) {   // rare: some component of this float4 moved by about half a box or more
            const int4 jv = slow_jumps<BOXF32>(diff[0], diff[1], diff[2], diff[3], e0, p.boxf[0], p.boxf[1], p.boxf[2],
                                               p.boxd[0], p.boxd[1], p.boxd[2]);
            jmp[0] = jv.x; jmp[1] = jv.y; jmp[2] = jv.z; jmp[3] = jv.w;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (j >= nv) jmp[j] = 0;
                cum[j] += jmp[j];
                overflow |= (p.images != nullptr) && (cum[j] > 127 || cum[j] < -128);
            }
            packed = pack4(cum[0] + MDB_IMAGE_BIAS, cum[1] + MDB_IMAGE_BIAS, cum[2] + MDB_IMAGE_BIAS,
                           cum[3] + MDB_IMAGE_BIAS);
        }
        if (PLAIN) {
            uint8_t *dst = p.images + (long long)t * p.img_pitch + e0;
            if (VEC) *reinterpret_cast<uint32_t *>(dst) = packed;
            else
                for (int j = 0; j < nv; ++j) dst[j] = (uint8_t)(cum[j] + MDB_IMAGE_BIAS);
            return;
        }
        const int orow = p.out_row ? p.out_row[t] : t;
        if (orow < 0) return;
        if (p.images) {
            uint8_t *dst = p.images + (long long)orow * p.img_pitch + e0;
            if (VEC) *reinterpret_cast<uint32_t *>(dst) = packed;
            else
                for (int j = 0; j < nv; ++j) dst[j] = (uint8_t)(cum[j] + MDB_IMAGE_BIAS);
        }
        if (p.images16) {
            int16_t *dst = p.images16 + (long long)orow * p.img16_pitch + e0;
            for (int j = 0; j < nv; ++j) dst[j] = (int16_t)cum[j];
            overflow |= (cum[0] > 32767 || cum[0] < -32768 || cum[1] > 32767 || cum[1] < -32768 || cum[2] > 32767 ||
                         cum[2] < -32768 || cum[3] > 32767 || cum[3] < -32768);
        }
        if (p.jumps) {
            int8_t *dst = p.jumps + (long long)orow * p.img_pitch + e0;
            if (VEC) *reinterpret_cast<uint32_t *>(dst) = pack4(jmp[0], jmp[1], jmp[2], jmp[3]);
            else
                for (int j = 0; j < nv; ++j) dst[j] = (int8_t)jmp[j];
        }
        if (p.coords_out) {
            float *dst = p.coords_out + (long long)orow * p.pitch_out + e0;
            if (VEC) *reinterpret_cast<float4 *>(dst) = make_float4(cur[0], cur[1], cur[2], cur[3]);
            else
                for (int j = 0; j < nv; ++j) dst[j] = cur[j];
        }
    };

    // Software pipeline over whole double-batches: while batch A (kBatch rows, in registers) is
    // processed the loads of batch B are in flight and vice versa; the running state does not feed
    // the addresses, so a thread keeps up to 2*kBatch independent 128-bit loads outstanding.
    float buf_a[kBatch][4], buf_b[kBatch][4];
    const int n_main = p.n_frames / (2 * kBatch) * (2 * kBatch);
    int tb = 0;
    if (n_main > 0) {
#pragma unroll
        for (int u = 0; u < kBatch; ++u) load_row(u, buf_a[u]);
        for (; tb < n_main; tb += 2 * kBatch) {
#pragma unroll
            for (int u = 0; u < kBatch; ++u) load_row(tb + kBatch + u, buf_b[u]);
#pragma unroll
            for (int u = 0; u < kBatch; ++u) step(tb + u, buf_a[u]);
            if (tb + 2 * kBatch < n_main) {
#pragma unroll
                for (int u = 0; u < kBatch; ++u) load_row(tb + 2 * kBatch + u, buf_a[u]);
            }
#pragma unroll
            for (int u = 0; u < kBatch; ++u) step(tb + kBatch + u, buf_b[u]);
        }
    }
    for (int t = n_main; t < p.n_frames; ++t) {   // fewer than 2*kBatch frames left
        float cur[4];
        load_row(t, cur);
        step(t, cur);
    }

    for (int j = 0; j < nv; ++j) {
        p.prev[e0 + j] = prev[j];
        p.cum[e0 + j] = cum[j];
    }
    if (overflow && p.overflow) atomicExch(p.overflow, 1);
}

// The dense case again, fed by the TMA engine: a producer warp streams the CTA's column slab
// (kThreads float4 = 4 KiB per frame row) into a ring of shared-memory stages of kRowsPerStage rows
// with cp.async.bulk + mbarriers; the compute threads take their float4 per row from shared memory.
// Only 3N/4 threads exist (75 k for 10^5 atoms), too few to cover HBM latency with register loads;
// here the ring keeps kStages*kRowsPerStage*4 KiB per CTA in flight.
constexpr int kRowsPerStage = 4;
constexpr int kScanStages = 6;
constexpr int kSlabBytes = kThreads * 16;
constexpr int kScanBulkThreads = kThreads + 32;

template <bool BOXF32>
__global__ void __launch_bounds__(kScanBulkThreads, 2) nojump_scan_bulk_kernel(const NojumpDev p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];   // [kScanStages][kRowsPerStage][kSlabBytes]
    __shared__ __align__(8) uint64_t s_full[kScanStages];
    __shared__ __align__(8) uint64_t s_empty[kScanStages];
    const int tid = threadIdx.x;
    if (tid == kThreads) {
        for (int i = 0; i < kScanStages; ++i) {
            mbar_init(&s_full[i], 1);
            mbar_init(&s_empty[i], kThreads / 32);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const uint32_t full0 = smem_u32(&s_full[0]), empty0 = smem_u32(&s_empty[0]);
    const long long col0 = (long long)blockIdx.x * (kThreads * 4);          // first component of the slab
    const int n_stage_rows = (p.n_frames + kRowsPerStage - 1) / kRowsPerStage;

    if (tid >= kThreads) {
        if (tid == kThreads) {
            const long long left = p.n_vals - col0;
            const uint32_t bytes = (uint32_t)min((long long)kSlabBytes, ((left + 3) / 4) * 16);
            const uint32_t stages0 = smem_u32(smem_raw);
            uint32_t stage = 0, phase = 0;
            for (int sr = 0; sr < n_stage_rows; ++sr) {
                const int t0 = sr * kRowsPerStage;
                const int nr = min(kRowsPerStage, p.n_frames - t0);
                mbar_wait(empty0 + 8 * stage, phase ^ 1);
                const uint32_t bar = full0 + 8 * stage;
                mbar_expect_tx(bar, bytes * nr);
                for (int r = 0; r < nr; ++r)
                    bulk_g2s(stages0 + (stage * kRowsPerStage + r) * kSlabBytes,
                             p.coords + (long long)(t0 + r) * p.pitch + col0, bytes, bar);
                if (++stage == kScanStages) { stage = 0; phase ^= 1; }
            }
        }
        return;
    }

    const int e0 = (int)col0 + 4 * tid;
    const bool live = e0 < p.n_vals;
    const int nv = live ? min(4, p.n_vals - e0) : 0;
    float safe[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int d = (e0 + j) % 3;
        safe[j] = 0.49f * fabsf(BOXF32 ? p.boxf[d] : (float)p.boxd[d]);
    }
    float prev[4];
    int cum[4];
    bool overflow = false;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        prev[j] = 0.f;
        cum[j] = 0;
        if (j < nv) {
            if (p.first_segment) {
                prev[j] = p.whole0 ? p.whole0[e0 + j] : p.coords[e0 + j];
            } else {
                prev[j] = p.prev[e0 + j];
                cum[j] = p.cum[e0 + j];
            }
        }
    }
    uint32_t packed = pack4(cum[0] + MDB_IMAGE_BIAS, cum[1] + MDB_IMAGE_BIAS, cum[2] + MDB_IMAGE_BIAS,
                            cum[3] + MDB_IMAGE_BIAS);
    uint8_t *dst = p.images + e0;
    uint32_t stage = 0, phase = 0;
    for (int sr = 0; sr < n_stage_rows; ++sr) {
        const int t0 = sr * kRowsPerStage;
        mbar_wait(full0 + 8 * stage, phase);
        float4 rows[kRowsPerStage];
#pragma unroll
        for (int r = 0; r < kRowsPerStage; ++r)
            rows[r] = reinterpret_cast<const float4 *>(smem_raw + (size_t)(stage * kRowsPerStage + r) * kSlabBytes)[tid];
        __syncwarp();                                  // every lane has read its rows (lanes diverge in slow_jumps)
        mbar_arrive_if(empty0 + 8 * stage, (tid & 31) == 0);
        if (++stage == kScanStages) { stage = 0; phase ^= 1; }
#pragma unroll
        for (int r = 0; r < kRowsPerStage; ++r) {
            if (t0 + r < p.n_frames && live) {
                const float cur[4] = {rows[r].x, rows[r].y, rows[r].z, rows[r].w};
                float diff[4];
                bool any = false;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    diff[j] = __fsub_rn(cur[j], prev[j]);
                    any |= !(fabsf(diff[j]) < safe[j]);
                    prev[j] = cur[j];
                }
                if (any) {
                    const int4 jv = slow_jumps<BOXF32>(diff[0], diff[1], diff[2], diff[3], e0, p.boxf[0], p.boxf[1],
                                                       p.boxf[2], p.boxd[0], p.boxd[1], p.boxd[2]);
                    const int jmp[4] = {jv.x, jv.y, jv.z, jv.w};
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        if (j < nv) cum[j] += jmp[j];
                        overflow |= (cum[j] > 127 || cum[j] < -128);
                    }
                    packed = pack4(cum[0] + MDB_IMAGE_BIAS, cum[1] + MDB_IMAGE_BIAS, cum[2] + MDB_IMAGE_BIAS,
                                   cum[3] + MDB_IMAGE_BIAS);
                }
                *reinterpret_cast<uint32_t *>(dst + (long long)(t0 + r) * p.img_pitch) = packed;
            }
        }
    }
    for (int j = 0; j < nv; ++j) {
        p.prev[e0 + j] = prev[j];
        p.cum[e0 + j] = cum[j];
    }
    if (overflow && p.overflow) atomicExch(p.overflow, 1);
}

// Image counts from stored jump matrices (the reference's `.nojump.npz` cache, reader.py:189-241): the
// three CSR matrices J_d (frames x atoms) arrive re-sorted by component column (CSC over the 3N interleaved
// components c = 3*atom + d), so a thread owns 4 consecutive components, walks the kept frames in ascending
// order and advances through each component's (frame, jump) list -- the prefix sum
// `m[:step+1].sum(axis=0)` of pbc.py:235-237 -- writing one packed excess-128 word per kept frame
// (a warp writes 128 contiguous bytes of the image row).  Jumps are rare (one entry per wall crossing),
// so the lists are short and the kernel is bound by the image rows it writes.
struct JumpMatDev {
    const long long *indptr;
    const int *frames;
    const int8_t *jumps;
    int n_vals;
    const int *keep;
    int n_keep;
    uint8_t *images;
    long long img_pitch;
    int *overflow;
    int vec;
    int16_t *images16;
    long long img16_pitch;
};

__global__ void __launch_bounds__(kThreads) nojump_images_from_jumps_kernel(const JumpMatDev p) {
    const int c0 = (blockIdx.x * kThreads + threadIdx.x) * 4;
    if (c0 >= p.n_vals) return;
    long long pos[4], end[4];
    int cum[4] = {0, 0, 0, 0};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const bool live = c0 + j < p.n_vals;
        pos[j] = live ? p.indptr[c0 + j] : 0;
        end[j] = live ? p.indptr[c0 + j + 1] : 0;
    }
    bool overflow = false;
    for (int r = 0; r < p.n_keep; ++r) {
        const int frame = p.keep ? p.keep[r] : r;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            while (pos[j] < end[j] && p.frames[pos[j]] <= frame) cum[j] += (int)p.jumps[pos[j]++];
            overflow |= p.images16 ? ((cum[j] < -32768) | (cum[j] > 32767)) : ((cum[j] < -128) | (cum[j] > 127));
        }
        if (p.images16) {
            int16_t *row16 = p.images16 + (long long)r * p.img16_pitch + c0;
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (c0 + j < p.n_vals) row16[j] = (int16_t)cum[j];
            continue;
        }
        uint8_t *row = p.images + (long long)r * p.img_pitch + c0;
        if (p.vec) {
            *reinterpret_cast<uint32_t *>(row) =
                pack4(cum[0] + MDB_IMAGE_BIAS, cum[1] + MDB_IMAGE_BIAS, cum[2] + MDB_IMAGE_BIAS, cum[3] + MDB_IMAGE_BIAS);
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (c0 + j < p.n_vals) row[j] = (uint8_t)(cum[j] + MDB_IMAGE_BIAS);
        }
    }
    if (overflow && p.overflow) atomicExch(p.overflow, 1);
}

__global__ void nojump_apply_kernel(const float *__restrict__ x, const uint8_t *__restrict__ n, double b0, double b1,
                                    double b2, int n_vals, double *__restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_vals) return;
    const int d = i % 3;
    const double b = d == 0 ? b0 : (d == 1 ? b1 : b2);
    // pbc.py:235-247: frame - counts * box, float64 (the product is exact, so fma == mul then sub)
    out[i] = fma(-(double)((int)n[i] - MDB_IMAGE_BIAS), b, (double)x[i]);
}


// MDB_IMAGE_BYTES rows -> MDB_IMAGE_PACKED rows: one thread per group of 4 atoms reads its 12 excess-128 bytes
// (three words) and writes one word, byte a = atom a, bits 2d = n + 1.  Reads 3 B, writes 1 B per atom and frame.
__global__ void __launch_bounds__(256) images_pack_kernel(const uint8_t *__restrict__ images, long long img_pitch,
                                                           long long n_rows, int n_atoms, int n_groups,
                                                           uint32_t *__restrict__ packed, long long packed_pitch,
                                                           int *__restrict__ max_abs, int vec) {
    const long long row = blockIdx.y;
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    int worst = 0;
    if (g < n_groups && row < n_rows) {
        const uint8_t *src = images + row * img_pitch + 12ll * g;
        const int nc = min(12, 3 * n_atoms - 12 * g);
        uint32_t w[3] = {0x80808080u, 0x80808080u, 0x80808080u};
        if (vec && nc == 12) {
            const uint32_t *s4 = reinterpret_cast<const uint32_t *>(src);
            w[0] = __ldg(s4); w[1] = __ldg(s4 + 1); w[2] = __ldg(s4 + 2);
        } else {
            for (int c = 0; c < nc; ++c) w[c >> 2] = (w[c >> 2] & ~(0xffu << (8 * (c & 3)))) | ((uint32_t)src[c] << (8 * (c & 3)));
        }
        uint32_t out = 0u;
#pragma unroll
        for (int c = 0; c < 12; ++c) {
            const int n = (int)((w[c >> 2] >> (8 * (c & 3))) & 0xffu) - MDB_IMAGE_BIAS;
            worst = max(worst, abs(n));
            out |= (uint32_t)(min(max(n, -1), 1) + 1) << (8 * (c / 3) + 2 * (c % 3));
        }
        packed[row * packed_pitch + g] = out;
    }
    worst = __reduce_max_sync(0xffffffffu, worst);
    if ((threadIdx.x & 31) == 0 && worst > 1) atomicMax(max_abs, worst);
}

__global__ void nojump_apply_format_kernel(const float *__restrict__ x, const void *__restrict__ images, int format,
                                           double b0, double b1, double b2, int n_vals, double *__restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_vals) return;
    const int d = i % 3;
    const double b = d == 0 ? b0 : (d == 1 ? b1 : b2);
    int n;
    if (format == MDB_IMAGE_BYTES) n = (int)static_cast<const uint8_t *>(images)[i] - MDB_IMAGE_BIAS;
    else if (format == MDB_IMAGE_I16) n = (int)static_cast<const int16_t *>(images)[i];
    else n = (int)((static_cast<const uint32_t *>(images)[i / 12] >> (8 * ((i % 12) / 3) + 2 * d)) & 3u) - 1;
    out[i] = fma(-(double)n, b, (double)x[i]);     // pbc.py:235-247
}

}  // namespace
}  // namespace mdb

using namespace mdb;

extern "C" int mdb_nojump_scan(const mdb_nojump_args *a, void *stream) {
    if (!a) return fail(MDB_ERR_ARGUMENT, "args is NULL");
    if (a->n_frames < 0 || a->n_atoms <= 0) return fail(MDB_ERR_ARGUMENT, "bad n_frames / n_atoms");
    if (!a->coords || !a->prev || !a->cum) return fail(MDB_ERR_ARGUMENT, "coords, prev and cum are required");
    if (a->pitch < 3ll * a->n_atoms) return fail(MDB_ERR_ARGUMENT, "pitch < 3*n_atoms");
    if ((a->images || a->jumps) && a->img_pitch < 3ll * a->n_atoms)
        return fail(MDB_ERR_ARGUMENT, "img_pitch < 3*n_atoms");
    if (a->coords_out && a->pitch_out < 3ll * a->n_atoms) return fail(MDB_ERR_ARGUMENT, "pitch_out < 3*n_atoms");
    for (int d = 0; d < 3; ++d)
        if (!(a->box0[d] > 0.0)) return fail(MDB_ERR_ARGUMENT, "box0 must be positive");
    if (a->n_frames == 0) return MDB_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);

    NojumpDev p;
    p.coords = a->coords;
    p.pitch = a->pitch;
    p.n_frames = a->n_frames;
    p.n_vals = 3 * a->n_atoms;
    for (int d = 0; d < 3; ++d) {
        p.boxd[d] = a->box0[d];
        p.boxf[d] = (float)a->box0[d];
    }
    p.first_segment = a->first_segment;
    p.whole0 = a->whole0;
    p.prev = a->prev;
    p.cum = a->cum;
    p.out_row = a->out_row;
    p.images = a->images;
    p.img_pitch = a->img_pitch;
    p.jumps = a->jumps;
    p.coords_out = a->coords_out;
    p.pitch_out = a->pitch_out;
    p.overflow = a->overflow;
    p.images16 = a->images16;
    p.img16_pitch = a->img16_pitch;
    if (a->images16 && a->img16_pitch < 3ll * a->n_atoms) return fail(MDB_ERR_ARGUMENT, "img16_pitch < 3*n_atoms");

    bool vec = (reinterpret_cast<uintptr_t>(a->coords) % 16 == 0) && (a->pitch % 4 == 0);
    if (a->images) vec = vec && (reinterpret_cast<uintptr_t>(a->images) % 4 == 0) && (a->img_pitch % 4 == 0);
    if (a->jumps) vec = vec && (reinterpret_cast<uintptr_t>(a->jumps) % 4 == 0) && (a->img_pitch % 4 == 0);
    if (a->coords_out) vec = vec && (reinterpret_cast<uintptr_t>(a->coords_out) % 16 == 0) && (a->pitch_out % 4 == 0);

    const int n_threads = (p.n_vals + 3) / 4;
    const unsigned blocks = (unsigned)((n_threads + kThreads - 1) / kThreads);
    const bool plain = vec && a->images && !a->out_row && !a->jumps && !a->coords_out && !a->images16;
    // the bulk-copy variant needs enough frames to fill its ring and is the default for the dense case
    const bool bulk = plain && a->n_frames >= 64 && !a->tuning.scan_direct;
    if (bulk) {
        const unsigned bblocks = (unsigned)((p.n_vals + kThreads * 4 - 1) / (kThreads * 4));
        const size_t smem = (size_t)kScanStages * kRowsPerStage * kSlabBytes;
        if (a->box_is_f32) {
            MDB_CUDA_CHECK(cudaFuncSetAttribute(nojump_scan_bulk_kernel<true>,
                                                cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            nojump_scan_bulk_kernel<true><<<bblocks, kScanBulkThreads, smem, st>>>(p);
        } else {
            MDB_CUDA_CHECK(cudaFuncSetAttribute(nojump_scan_bulk_kernel<false>,
                                                cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            nojump_scan_bulk_kernel<false><<<bblocks, kScanBulkThreads, smem, st>>>(p);
        }
        MDB_CUDA_CHECK(cudaGetLastError());
        return MDB_OK;
    }
    if (a->box_is_f32) {
        if (plain) nojump_scan_kernel<true, true, true><<<blocks, kThreads, 0, st>>>(p);
        else if (vec) nojump_scan_kernel<true, true, false><<<blocks, kThreads, 0, st>>>(p);
        else nojump_scan_kernel<true, false, false><<<blocks, kThreads, 0, st>>>(p);
    } else {
        if (plain) nojump_scan_kernel<false, true, true><<<blocks, kThreads, 0, st>>>(p);
        else if (vec) nojump_scan_kernel<false, true, false><<<blocks, kThreads, 0, st>>>(p);
        else nojump_scan_kernel<false, false, false><<<blocks, kThreads, 0, st>>>(p);
    }
    MDB_CUDA_CHECK(cudaGetLastError());
    return MDB_OK;
}

extern "C" int mdb_nojump_apply(const float *coords_row, const uint8_t *images_row, const double box[3],
                                int32_t n_atoms, double *out, void *stream) {
    if (!coords_row || !images_row || !box || !out || n_atoms <= 0)
        return fail(MDB_ERR_ARGUMENT, "mdb_nojump_apply: NULL pointer or n_atoms <= 0");
    const int n_vals = 3 * n_atoms;
    nojump_apply_kernel<<<(n_vals + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        coords_row, images_row, box[0], box[1], box[2], n_vals, out);
    MDB_CUDA_CHECK(cudaGetLastError());
    return MDB_OK;
}

extern "C" int mdb_nojump_images_from_jumps(const mdb_jumpmat_args *a, void *stream) {
    if (!a) return fail(MDB_ERR_ARGUMENT, "args is NULL");
    if (a->n_atoms <= 0 || a->n_keep < 0) return fail(MDB_ERR_ARGUMENT, "bad n_atoms / n_keep");
    if (!a->indptr || (!a->images && !a->images16)) return fail(MDB_ERR_ARGUMENT, "indptr and images (or images16) are required");
    if (a->nnz < 0 || (a->nnz > 0 && (!a->frames || !a->jumps)))
        return fail(MDB_ERR_ARGUMENT, "frames and jumps are required when nnz > 0");
    if (a->images && a->img_pitch < 3ll * a->n_atoms) return fail(MDB_ERR_ARGUMENT, "img_pitch < 3*n_atoms");
    if (a->images16 && a->img16_pitch < 3ll * a->n_atoms) return fail(MDB_ERR_ARGUMENT, "img16_pitch < 3*n_atoms");
    if (a->n_keep == 0) return MDB_OK;
    JumpMatDev p;
    p.images16 = a->images ? nullptr : a->images16;
    p.img16_pitch = a->img16_pitch;
    p.indptr = reinterpret_cast<const long long *>(a->indptr);
    p.frames = a->frames;
    p.jumps = a->jumps;
    p.n_vals = 3 * a->n_atoms;
    p.keep = a->keep;
    p.n_keep = a->n_keep;
    p.images = a->images;
    p.img_pitch = a->img_pitch;
    p.overflow = a->overflow;
    // whole words only when every row is word-aligned and the padding can take the last partial word
    p.vec = (reinterpret_cast<uintptr_t>(a->images) % 4 == 0) && (a->img_pitch % 4 == 0) &&
            (a->img_pitch >= (p.n_vals + 3) / 4 * 4);
    const int n_threads = (p.n_vals + 3) / 4;
    nojump_images_from_jumps_kernel<<<(n_threads + kThreads - 1) / kThreads, kThreads, 0,
                                      static_cast<cudaStream_t>(stream)>>>(p);
    MDB_CUDA_CHECK(cudaGetLastError());
    return MDB_OK;
}

extern "C" int mdb_images_pack(const uint8_t *images, int64_t img_pitch, int64_t n_rows, int32_t n_atoms,
                               uint32_t *packed, int64_t packed_pitch, int32_t *max_abs, void *stream) {
    if (!images || !packed || !max_abs) return fail(MDB_ERR_ARGUMENT, "mdb_images_pack: NULL pointer");
    if (n_atoms <= 0 || n_rows < 0) return fail(MDB_ERR_ARGUMENT, "mdb_images_pack: bad n_atoms / n_rows");
    const int n_groups = (n_atoms + 3) / 4;
    if (img_pitch < 3ll * n_atoms || packed_pitch < n_groups)
        return fail(MDB_ERR_ARGUMENT, "mdb_images_pack: pitch too small");
    if (n_rows == 0) return MDB_OK;
    const int vec = (reinterpret_cast<uintptr_t>(images) % 4 == 0) && (img_pitch % 4 == 0);
    for (int64_t r0 = 0; r0 < n_rows; r0 += 65535) {          // gridDim.y limit
        const int64_t nr = n_rows - r0 < 65535 ? n_rows - r0 : 65535;
        const dim3 grid((unsigned)((n_groups + 255) / 256), (unsigned)nr);
        images_pack_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
            images + r0 * img_pitch, img_pitch, nr, n_atoms, n_groups, packed + r0 * packed_pitch, packed_pitch, max_abs, vec);
        MDB_CUDA_CHECK(cudaGetLastError());
    }
    return MDB_OK;
}

extern "C" int mdb_nojump_apply_format(const float *coords_row, const void *images_row, int32_t image_format,
                                       const double box[3], int32_t n_atoms, double *out, void *stream) {
    if (!coords_row || !images_row || !box || !out || n_atoms <= 0)
        return fail(MDB_ERR_ARGUMENT, "mdb_nojump_apply_format: NULL pointer or n_atoms <= 0");
    if (image_format != MDB_IMAGE_BYTES && image_format != MDB_IMAGE_PACKED && image_format != MDB_IMAGE_I16)
        return fail(MDB_ERR_ARGUMENT, "mdb_nojump_apply_format: unknown image format %d", image_format);
    const int n_vals = 3 * n_atoms;
    nojump_apply_format_kernel<<<(n_vals + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        coords_row, images_row, image_format, box[0], box[1], box[2], n_vals, out);
    MDB_CUDA_CHECK(cudaGetLastError());
    return MDB_OK;
}
