// Correlation kernels: per-(origin, lag) sums of r^2, r^4, sinc(q r) and van Hove histograms.
//
// Work decomposition (one CTA = one (origin s, lag chunk c, atom block ab)):
//   * a thread owns groups of 4 consecutive atoms = 12 interleaved xyz values = 48 bytes per frame
//     row (three 128-bit accesses, a warp covers 1536 contiguous bytes);
//   * the origin positions of the group are kept in registers and reused for the up to kLagChunk lag
//     frames of the chunk;
//   * CTAs of the same (s, ab) and different chunks are adjacent in launch order, so the origin
//     row is fetched from HBM once and re-read from L2 by the other chunks;
//   * per-lag accumulators live in shared memory (one private column per thread) for the whole atom loop;
//     there is ONE block reduction per CTA at the end, written as a partial; the finish kernel
//     (corr_host.cu) adds the partials of the atom blocks in fixed order (bit-reproducible, no float atomics);
//   * histograms are accumulated in shared memory (u32, one table per lag of the chunk) with
//     red.shared (measured on B200: one warp-wide update per clock and SM, scripts/ubench_atoms.cu) and
//     flushed with integer atomics.
//
// Two data paths feed the same arithmetic:
//   corr_bulk_kernel  (the product path for float32 rows with 16-byte aligned pitch): a producer
//       warp streams row tiles (1024 atoms = 12 KiB coordinates + 1 KiB packed / 3 KiB byte image counts)
//       from HBM into a ring of shared-memory stages with cp.async.bulk (the TMA engine, SASS UBLKCP),
//       completion on mbarriers; 8 consumer warps read their bytes per stage with conflict-free LDS and
//       release the stage as soon as the values are in registers.
//   corr_ldg_kernel   (fallback: float64 rows, unaligned pitch, int16 counts): the same loop with direct
//       128-bit / scalar global loads.
//
// Arithmetic, float64 chain: positions are widened to float64, unwrapped as fma(-n, box, x) (bit-identical to
// pbc.nojump's float64 frame, pbc.py:235-247), the displacement and r^2 are float64; sinc is evaluated in
// float32 from float32(q^2 r^2).  Float32 displacements (process_tile_f32d; ISF alone, van Hove alone and
// van Hove + alpha_2 on rows with one constant float32 box): d = x0 - xt - k L with k = n0 - nt, where the
// operand on the far side of the periodic boundary is moved first, which keeps every component good to 2^-23.
// Histogram bins follow utils.histogram exactly (utils.py:328-338): a branch-free float32 estimate decides
// every sample that is not within rounding distance of a bin edge, the rest are recomputed with the
// reference's float64 arithmetic (np.floor_divide semantics or np.histogram edge search).
//
// This file is compiled once per MDB_CORR_PART (see corr_core.cuh).
#include "corr_core.cuh"

#ifndef MDB_CORR_PART
#error "compile with -DMDB_CORR_PART=0..3"
#endif

namespace mdb {
namespace {

// ---- image counts of one group of 4 atoms (12 components) -----------------------------------------------
// raw words: IMG_BYTES 3 words (byte c), IMG_PACKED 1 word (byte a, bits 2d), IMG_I16 6 words (int16 c)
template <int IMG>
__device__ __forceinline__ int image_count_int(const uint32_t (&w)[6], int c) {
    if (IMG == IMG_BYTES) return (int)((w[c >> 2] >> (8 * (c & 3))) & 0xffu) - MDB_IMAGE_BIAS;
    if (IMG == IMG_PACKED) return (int)((w[0] >> (8 * (c / 3) + 2 * (c % 3))) & 3u) - 1;
    if (IMG == IMG_I16) return (int)(short)(w[c >> 1] >> (16 * (c & 1)));
    return 0;
}

// ---- unwrap: 12 raw float32 components (+ image counts) -> 12 float64 ------------------------------------
// bx: the box of the row; for IMG_PACKED it is already divided by 4^d (the packed fields come scaled).
template <int IMG>
__device__ __forceinline__ void unwrap12(const float (&raw)[12], const uint32_t (&w)[6], const double *bx,
                                         double (&u)[12]) {
    if (IMG == IMG_NONE) {
#pragma unroll
        for (int c = 0; c < 12; ++c) u[c] = (double)raw[c];
        return;
    }
    const double b[3] = {bx[0], bx[1], bx[2]};
    if (IMG == IMG_PACKED) {
        const uint32_t m[3] = {packed_dim(w[0], 0), packed_dim(w[0], 1), packed_dim(w[0], 2)};
#pragma unroll
        for (int c = 0; c < 12; ++c) u[c] = fma(-packed_count(m[c % 3], c / 3, c % 3), b[c % 3], (double)raw[c]);
    } else if (IMG == IMG_BYTES) {
#pragma unroll
        for (int c = 0; c < 12; ++c) u[c] = fma(-image_count(w[c >> 2], c & 3), b[c % 3], (double)raw[c]);
    } else {
#pragma unroll
        for (int c = 0; c < 12; ++c) u[c] = fma(-(double)image_count_int<IMG_I16>(w, c), b[c % 3], (double)raw[c]);
    }
}

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

__device__ __forceinline__ double r2_unfused(double ddx, double ddy, double ddz) {
    // (vec ** 2).sum(axis=-1) as NumPy evaluates it: three rounded squares, summed left to right
    return __dadd_rn(__dadd_rn(__dmul_rn(ddx, ddx), __dmul_rn(ddy, ddy)), __dmul_rn(ddz, ddz));
}

// Exact bin of one float64 displacement (the rare / general path, kept out of line so that the streaming
// loop stays small): float64 sqrt of the unfused r^2, then utils.histogram's rule; counts it.
__device__ __noinline__ void exact_count(double ddx, double ddy, double ddz, int bin_mode, double dx, double inv_dx,
                                         int n_bins, const double *s_edges, unsigned int *s_hist_k) {
    const double r = sqrt(r2_unfused(ddx, ddy, ddz));
    int bin;
    if (bin_mode == MDB_BINS_UNIFORM) {
        const long long kk = exact_floor_div(r, dx, inv_dx);
        bin = (kk < n_bins) ? (int)kk : -1;
    } else {
        bin = edge_bin(r, s_edges, n_bins);
    }
    if (bin >= 0) atomicAdd(&s_hist_k[bin], 1u);
}

// The same from the raw values of one atom (float32 position + integer image counts of the two frames, one
// constant box): the reference's float64 arithmetic (pbc.py:235-247, correlation.py:246-247, utils.py:334),
// for the samples the float32-displacement path cannot decide.  Uniform bins only.
__device__ __noinline__ void exact_count_raw(double bx, double by, double bz, double dx, double inv_dx, int n_bins,
                                             float x0, float y0, float z0, int n0x, int n0y, int n0z, float xt,
                                             float yt, float zt, int ntx, int nty, int ntz, unsigned int *s_hist_k) {
    const double ddx = fma(-(double)n0x, bx, (double)x0) - fma(-(double)ntx, bx, (double)xt);
    const double ddy = fma(-(double)n0y, by, (double)y0) - fma(-(double)nty, by, (double)yt);
    const double ddz = fma(-(double)n0z, bz, (double)z0) - fma(-(double)ntz, bz, (double)zt);
    const long long kk = exact_floor_div(sqrt(r2_unfused(ddx, ddy, ddz)), dx, inv_dx);
    if (kk < n_bins) atomicAdd(&s_hist_k[(int)kk], 1u);
}

// Branch-free bin estimate of 4 samples from float32 r^2 (uniform bins).  th = r/dx - 1/2, so that the
// nearest integer of th (the 1.5*2^23 trick: no conversion instructions) is floor(r/dx) for every sample that is
// not next to a bin edge, and dd = th - rint(th) = frac(r/dx) - 1/2.  A sample is decided here when
// |dd| < 1/2 - margin(t), margin = twice the float32 error bound of t (p.hist_c1, p.hist_lim0); the others
// come back in the returned mask and are recomputed exactly.  r = 0 gives NaN (0 * inf): not decided.  The bin
// of a decided sample is (bits of tb) - 0x4B400000 >= 0; the subtraction is folded into the table address
// (`table` = shared address of the lag's table - 4 * 0x4B400000) and t >= 2^22 (dropped like every r >= n_bins dx)
// fails the range test.  The update itself is unconditional (see red_shared_add).
template <bool MASKED>
__device__ __forceinline__ uint32_t hist_fast4(const CorrDev &p, const float (&r2)[4], uint32_t vmask, uint32_t table,
                                               uint32_t trash) {
    const f32x2 inv_dx = bcast2(p.inv_dx_f), mhalf = bcast2(-0.5f), big = bcast2(12582912.0f),
                nbig = bcast2(-12582912.0f), c1 = bcast2(p.hist_c1), lim0 = bcast2(p.hist_lim0);
    const uint32_t one = p.one;
    float tb[4], dd[4], lim[4];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const f32x2 r = mul2(pack2(r2[2 * h], r2[2 * h + 1]), pack2(rsqrt_approx(r2[2 * h]), rsqrt_approx(r2[2 * h + 1])));
        const f32x2 th = fma2(r, inv_dx, mhalf);
        const f32x2 b = add2(th, big);
        unpack2(b, tb[2 * h], tb[2 * h + 1]);
        unpack2(sub2(th, add2(b, nbig)), dd[2 * h], dd[2 * h + 1]);
        unpack2(fma2(th, c1, lim0), lim[2 * h], lim[2 * h + 1]);
    }
    uint32_t slow = 0u;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        const bool live = !MASKED || ((vmask >> a) & 1u);
        const bool fast = fabsf(dd[a]) < lim[a];
        const uint32_t ib = __float_as_uint(tb[a]);
        // samples that are undecided, out of range or masked go to the scratch word `trash`
#ifdef MDB_DBG_FAST_ATOMIC
        if (live && fast && ib < p.ib_limit) {
            unsigned int *ptr;
            asm("cvta.shared.u64 %0, %1;" : "=l"(ptr) : "l"((unsigned long long)(table + 4u * ib)));
            atomicAdd(ptr, 1u);
        }
#else
        red_shared_add((live && fast && ib < p.ib_limit) ? table + 4u * ib : trash, one);
#endif
        slow |= (live && !fast) ? (1u << a) : 0u;
    }
    return slow;
}

// Accumulators of one thread for the lags of its chunk.
template <int FLAGS>
struct Accum {
    static constexpr bool kR2 = (FLAGS & (MDB_SUM_R2 | MDB_SUM_LEGENDRE)) != 0;   // slot 0
    static constexpr bool kR4 = (FLAGS & MDB_SUM_R4) != 0;
    static constexpr bool kSinc = (FLAGS & MDB_SUM_SINC) != 0;
    double r2[kR2 ? kLagChunk : 1];
    double r4[kR4 ? kLagChunk : 1];
    double sc[kSinc ? kLagChunk : 1];
    __device__ __forceinline__ void clear() {
#pragma unroll
        for (int k = 0; k < kLagChunk; ++k) {
            if constexpr (kR2) r2[k] = 0.0;
            if constexpr (kR4) r4[k] = 0.0;
            if constexpr (kSinc) sc[k] = 0.0;
        }
    }
};

// One group of 4 atoms against one lag frame in float64: partial sums of r^2, r^4 (float64) and sinc (float32)
// over the group, histogram counts straight into shared memory.  MASKED is instantiated for the
// last tile of a shard and for sub-ensemble launches: bit a of `vmask` says whether atom a of the
// group takes part (it exists, and the origin's selector picked it); the others are computed and
// discarded with selects (their inputs may be row padding).
template <int FLAGS, bool HIST_F32, bool MASKED>
__device__ __forceinline__ void eval_group(const CorrDev &p, const double (&u0)[12], const float (&raw0)[12],
                                           const double (&ut)[12], const float (&rawt)[12], uint32_t vmask,
                                           double &sum2, double &sum4, float &sums, unsigned int *s_hist_k,
                                           const double *s_edges, uint32_t trash) {
    constexpr bool kR2 = (FLAGS & MDB_SUM_R2) != 0;
    constexpr bool kR4 = (FLAGS & MDB_SUM_R4) != 0;
    constexpr bool kSinc = (FLAGS & MDB_SUM_SINC) != 0;
    constexpr bool kHist = (FLAGS & MDB_HIST) != 0;
    sum2 = 0.0;
    sum4 = 0.0;
    sums = 0.0f;
    float yq[4];
    float r2f[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        const double ddx = u0[3 * a] - ut[3 * a];
        const double ddy = u0[3 * a + 1] - ut[3 * a + 1];
        const double ddz = u0[3 * a + 2] - ut[3 * a + 2];
        const double r2 = fma(ddz, ddz, fma(ddy, ddy, ddx * ddx));
        const bool on = !MASKED || ((vmask >> a) & 1u);
        if constexpr ((FLAGS & MDB_SUM_LEGENDRE) != 0) {
            // rows are vectors: P_l of the scalar product (Bonnet recurrence), summed into slot 0
            const double x = fma(u0[3 * a + 2], ut[3 * a + 2], fma(u0[3 * a + 1], ut[3 * a + 1], u0[3 * a] * ut[3 * a]));
            double pm = 1.0, pl = (p.leg_order == 0) ? 1.0 : x;
            for (int k = 1; k < p.leg_order; ++k) {
                const double pn = (fma((double)(2 * k + 1) * x, pl, -(double)k * pm)) / (double)(k + 1);
                pm = pl;
                pl = pn;
            }
            sum2 += on ? pl : 0.0;
        }
        if constexpr (kR2) sum2 += on ? r2 : 0.0;
        if constexpr (kR4) sum4 = MASKED ? sum4 + (on ? r2 * r2 : 0.0) : fma(r2, r2, sum4);
        if constexpr (kSinc) yq[a] = (float)(r2 * p.q2);
        if constexpr (kHist) r2f[a] = (float)r2;
    }
    if constexpr (kSinc) {
        float v0, v1, v2, v3;
        sinc_of_square_x2(yq[0], yq[1], v0, v1);
        sinc_of_square_x2(yq[2], yq[3], v2, v3);
        if constexpr (MASKED) {
            v0 = (vmask & 1u) ? v0 : 0.0f;
            v1 = (vmask & 2u) ? v1 : 0.0f;
            v2 = (vmask & 4u) ? v2 : 0.0f;
            v3 = (vmask & 8u) ? v3 : 0.0f;
        }
        sums = (v0 + v1) + (v2 + v3);
    }
    if constexpr (kHist) {
        if (HIST_F32) {
            // raw float32 frames: the reference computes r in float32 (correlation.py:246-247 on float32
            // arrays); replay vec = start - end; (vec**2).sum(-1) ** .5 in float32, then the exact rule
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                if (!MASKED || ((vmask >> a) & 1u)) {
                    const float vx = __fsub_rn(raw0[3 * a], rawt[3 * a]);
                    const float vy = __fsub_rn(raw0[3 * a + 1], rawt[3 * a + 1]);
                    const float vz = __fsub_rn(raw0[3 * a + 2], rawt[3 * a + 2]);
                    const float q = __fadd_rn(__fadd_rn(__fmul_rn(vx, vx), __fmul_rn(vy, vy)), __fmul_rn(vz, vz));
                    const double r = (double)__fsqrt_rn(q);
                    int bin;
                    if (p.bin_mode == MDB_BINS_UNIFORM) {
                        const long long kk = exact_floor_div(r, p.dx, p.inv_dx);
                        bin = (kk < p.n_bins) ? (int)kk : -1;
                    } else {
                        bin = edge_bin(r, s_edges, p.n_bins);
                    }
                    if (bin >= 0) atomicAdd(&s_hist_k[bin], 1u);
                }
            }
        } else {
            // uniform bins: the float32 estimate decides what is clear of the bin edges; everything else (and
            // every sample of non-uniform bins) is counted by exact_count()
            uint32_t slow = MASKED ? (vmask & 0xfu) : 0xfu;
            if (p.bin_mode == MDB_BINS_UNIFORM)
                slow = hist_fast4<MASKED>(p, r2f, vmask, smem_u32(s_hist_k) - 4u * 0x4B400000u, trash);
            if (slow) {
#pragma unroll
                for (int a = 0; a < 4; ++a)
                    if (slow & (1u << a))
                        exact_count(u0[3 * a] - ut[3 * a], u0[3 * a + 1] - ut[3 * a + 1], u0[3 * a + 2] - ut[3 * a + 2],
                                    p.bin_mode, p.dx, p.inv_dx, p.n_bins, s_edges, s_hist_k);
            }
        }
    }
}

// Which atoms of group g take part: they exist (tail of the shard) and, for sub-ensemble launches, the selector of
// origin s picked them (one byte per atom, four atoms = one 32-bit word).  Bits 0..3: atom a takes part; bits
// 4+6a .. 9+6a: the byte itself (1..63) -- the sub-ensemble the atom belongs to when ONE pass serves several of them
// (p.n_segments > 1, multi_subensemble_correlation); with a plain mask the byte is 1.
__device__ __forceinline__ uint32_t group_mask(const CorrDev &p, int s, int g) {
    const int nv = max(0, min(4, p.n_atoms - 4 * g));
    uint32_t m = (1u << nv) - 1u;
    if (p.select && nv > 0) {
        const uint32_t w = __ldg(reinterpret_cast<const uint32_t *>(p.select + (long long)s * p.sel_pitch) + g);
        m &= ((w & 0xffu) ? 1u : 0u) | ((w & 0xff00u) ? 2u : 0u) | ((w & 0xff0000u) ? 4u : 0u) |
             ((w & 0xff000000u) ? 8u : 0u);
        m |= ((w & 0x3fu) << 4) | (((w >> 8) & 0x3fu) << 10) | (((w >> 16) & 0x3fu) << 16) | (((w >> 24) & 0x3fu) << 22);
    }
    return m;
}

// Scratch words for the histogram updates that must not count (hist_fast4): one word per lane in a 128-byte line of
// its own at the end of the dynamic shared memory (kHistTailBytes, reserved by make_plan) -- no two lanes of a warp
// share a word, and the line holds nothing else.  (A single scratch word placed next to the mbarriers of the ring
// made samples read overwritten stages -- found with scripts/debug_hist.py; the barriers share no line with anything
// the kernels update with shared-memory atomics any more.)
__device__ __forceinline__ uint32_t trash_slot(const CorrDev &p, const double *s_edges, int tid) {
    const uint32_t tail = smem_u32(s_edges) + (p.bin_mode == MDB_BINS_EDGES ? 8u * (uint32_t)(p.n_bins + 1) : 0u);
    return ((tail + 127u) & ~127u) + 4u * (uint32_t)(tid & 31);
}

// CTA prologue shared by both kernels: which (origin, chunk, atom block) am I, rows and boxes.
struct CtaWork {
    int s, ab, l0, nl;
};

__device__ __forceinline__ CtaWork cta_work(const CorrDev &p) {
    CtaWork w;
    const int bid = blockIdx.x;
    const int c = bid % p.n_chunks;
    const int rest = bid / p.n_chunks;
    w.ab = rest % p.n_ablocks;
    w.s = rest / p.n_ablocks;
    w.l0 = (int)((long long)c * p.L / p.n_chunks);
    w.nl = (int)((long long)(c + 1) * p.L / p.n_chunks) - w.l0;
    return w;
}

template <int IMG>
__device__ __forceinline__ void load_rows(const CorrDev &p, const CtaWork &w, int tid, int *s_row,
                                          double (*s_box)[3]) {
    if (tid <= kLagChunk) {
        // slot 0: the origin row, slots 1..nl: the lag rows of this chunk
        const int l = (tid == 0) ? 0 : w.l0 + tid - 1;
        const int row = (tid == 0 || tid - 1 < w.nl) ? p.rows[(long long)w.s * p.L + l] : 0;
        s_row[tid] = row;
        if (IMG != IMG_NONE) {
            // packed counts come scaled by 4^d: the box is divided accordingly (exact)
            const double sy = IMG == IMG_PACKED ? 0.25 : 1.0, sz = IMG == IMG_PACKED ? 0.0625 : 1.0;
            s_box[tid][0] = p.box ? p.box[3ll * row] : p.box_const[0];
            s_box[tid][1] = p.box ? p.box[3ll * row + 1] * sy : p.box_const[1];
            s_box[tid][2] = p.box ? p.box[3ll * row + 2] * sz : p.box_const[2];
        }
    }
}

// CTA epilogue: one block reduction, partial sums out, histogram flush.  Called by the kThreads
// compute threads only; `sync` is the barrier that joins exactly those threads.  `seg` / `n_seg`: the sub-ensemble
// these accumulators belong to (0 / 1 for ordinary launches).
template <int FLAGS, typename Sync>
__device__ __forceinline__ void cta_epilogue(const CorrDev &p, const CtaWork &w, int tid, Accum<FLAGS> &acc,
                                             double (*s_red)[kLagChunk * 3], const unsigned int *s_hist, Sync sync,
                                             int seg = 0, int n_seg = 1) {
    constexpr bool kR2 = (FLAGS & (MDB_SUM_R2 | MDB_SUM_LEGENDRE)) != 0;
    constexpr bool kR4 = (FLAGS & MDB_SUM_R4) != 0;
    constexpr bool kSinc = (FLAGS & MDB_SUM_SINC) != 0;
    constexpr bool kHist = (FLAGS & MDB_HIST) != 0;
    const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
    for (int k = 0; k < kLagChunk; ++k) {
        if (k < w.nl) {
            double v2 = 0.0, v4 = 0.0, vs = 0.0;
            if constexpr (kR2) v2 = warp_sum(acc.r2[k]);
            if constexpr (kR4) v4 = warp_sum(acc.r4[k]);
            if constexpr (kSinc) vs = warp_sum(acc.sc[k]);
            if (lane == 0) {
                s_red[warp][3 * k] = v2;
                s_red[warp][3 * k + 1] = v4;
                s_red[warp][3 * k + 2] = vs;
            }
        }
    }
    sync();
    if (tid < 3 * w.nl) {
        double v = 0.0;
#pragma unroll
        for (int wi = 0; wi < kWarps; ++wi) v += s_red[wi][tid];
        const int k = tid / 3, comp = tid - 3 * k;
        p.partial[((((long long)w.s * p.L + w.l0 + k) * p.n_ablocks + w.ab) * n_seg + seg) * 3 + comp] = v;
    }
    if (kHist && seg == 0) {
        // the chunk's tables [lag][bin] are one contiguous piece of the [origin][lag][bin] result
        unsigned long long *out = p.hist + ((long long)w.s * p.L + w.l0) * p.n_bins;
        for (int i = tid; i < w.nl * p.n_bins; i += kThreads) {
            const unsigned int cnt = s_hist[i];
            if (cnt) atomicAdd(&out[i], (unsigned long long)cnt);
        }
    }
    if (n_seg > 1) sync();                 // s_red is reused by the next sub-ensemble
}

#if MDB_CORR_PART == 0
// =================================================================================================
// fallback path: direct global loads (float64 rows, unaligned pitch, int16 image counts)
// =================================================================================================
template <typename CT, int IMG, bool VEC, bool STREAM>
struct GroupLoader;

template <int IMG, bool VEC, bool STREAM>
struct GroupLoader<float, IMG, VEC, STREAM> {
    __device__ static __forceinline__ void load(const CorrDev &p, int row, int g, int nv, const double *bx,
                                                double (&u)[12], float (&raw)[12]) {
        const float *base = static_cast<const float *>(p.coords) + (long long)row * p.pitch + 12ll * g;
        uint32_t w[6] = {0u, 0u, 0u, 0u, 0u, 0u};
        if (IMG == IMG_BYTES) w[0] = w[1] = w[2] = 0x80808080u;   // n = 0 in excess-128
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
        if (IMG == IMG_BYTES) {
            const uint8_t *ib = static_cast<const uint8_t *>(p.images) + (long long)row * p.img_pitch + 12ll * g;
            if (VEC && nv == 4) {
                const uint32_t *iw = reinterpret_cast<const uint32_t *>(ib);
                if (STREAM) { w[0] = ld_stream(iw); w[1] = ld_stream(iw + 1); w[2] = ld_stream(iw + 2); }
                else { w[0] = __ldg(iw); w[1] = __ldg(iw + 1); w[2] = __ldg(iw + 2); }
            } else {
#pragma unroll
                for (int c = 0; c < 12; ++c)
                    if (c < 3 * nv)
                        w[c >> 2] = (w[c >> 2] & ~(0xffu << (8 * (c & 3)))) | ((uint32_t)ib[c] << (8 * (c & 3)));
            }
        }
        if (IMG == IMG_I16) {
            const uint16_t *ib = reinterpret_cast<const uint16_t *>(p.images) + (long long)row * p.img_pitch + 12ll * g;
#pragma unroll
            for (int c = 0; c < 12; ++c)
                if (c < 3 * nv) w[c >> 1] |= (uint32_t)__ldg(ib + c) << (16 * (c & 1));
        }
        unwrap12<IMG>(raw, w, bx, u);
    }
};

template <bool VEC, bool STREAM>
struct GroupLoader<double, IMG_NONE, VEC, STREAM> {
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

template <typename CT, int IMG, int FLAGS, bool VEC>
__global__ void __launch_bounds__(kThreads, 2) corr_ldg_kernel(const CorrDev p) {
    constexpr bool kHist = (FLAGS & MDB_HIST) != 0;
    constexpr bool kHistF32 = kHist && sizeof(CT) == 4 && IMG == IMG_NONE;

    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ int s_row[kLagChunk + 1];
    __shared__ double s_box[kLagChunk + 1][3];
    __shared__ double s_red[kWarps][kLagChunk * 3];

    const CtaWork w = cta_work(p);
    const int tid = threadIdx.x;
    if (p.stamp != nullptr && (blockIdx.x | blockIdx.y | tid) == 0) *p.stamp = global_timer_ns();
    unsigned int *s_hist = reinterpret_cast<unsigned int *>(smem_raw);
    double *s_edges = reinterpret_cast<double *>(smem_raw + (size_t)kLagChunk * p.n_bins * sizeof(unsigned int));
    const uint32_t trash = trash_slot(p, s_edges, tid);

    load_rows<IMG>(p, w, tid, s_row, s_box);
    if (kHist) {
        for (int i = tid; i < kLagChunk * p.n_bins; i += kThreads) s_hist[i] = 0u;
        if (p.bin_mode == MDB_BINS_EDGES)
            for (int i = tid; i <= p.n_bins; i += kThreads) s_edges[i] = p.edges[i];
    }
    __syncthreads();

    Accum<FLAGS> acc;
    acc.clear();
    const int g_begin = w.ab * p.groups_per_block;
    const int g_end = min(p.n_groups, g_begin + p.groups_per_block);
    for (int g = g_begin + tid; g < g_end; g += kThreads) {
        const int nv = min(4, p.n_atoms - 4 * g);
        const uint32_t vmask = group_mask(p, w.s, g);
        double u0[12];
        float raw0[12];
        GroupLoader<CT, IMG, VEC, false>::load(p, s_row[0], g, nv, s_box[0], u0, raw0);
#pragma unroll
        for (int k = 0; k < kLagChunk; ++k) {
            if (k < w.nl) {
                double ut[12];
                float rawt[12];
                GroupLoader<CT, IMG, VEC, true>::load(p, s_row[k + 1], g, nv, s_box[k + 1], ut, rawt);
                double s2, s4;
                float ss;
                eval_group<FLAGS, kHistF32, true>(p, u0, raw0, ut, rawt, vmask, s2, s4, ss, s_hist + k * p.n_bins,
                                                  s_edges, trash);
                if constexpr (Accum<FLAGS>::kR2) acc.r2[k] += s2;
                if constexpr (Accum<FLAGS>::kR4) acc.r4[k] += s4;
                if constexpr (Accum<FLAGS>::kSinc) acc.sc[k] += (double)ss;
            }
        }
    }
    cta_epilogue<FLAGS>(p, w, tid, acc, s_red, s_hist, [] { __syncthreads(); });
}
#endif  // MDB_CORR_PART == 0

#if MDB_CORR_PART != 0
// =================================================================================================
// product path: cp.async.bulk (TMA) -> shared-memory ring -> compute warps
// =================================================================================================
__device__ __forceinline__ void compute_barrier() {  // named barrier 1: the kThreads compute threads
    asm volatile("bar.sync 1, %0;" ::"n"(kThreads) : "memory");
}

// Ring position of a compute thread / the producer.
struct Ring {
    uint32_t stage, phase, n_stages;
    __device__ __forceinline__ void advance() {
        if (++stage == n_stages) { stage = 0; phase ^= 1; }
    }
};

// Take the calling thread's 4 atoms out of the current stage and release the stage.
template <int IMG>
__device__ __forceinline__ void fetch_raw(Ring &ring, uint32_t full0, uint32_t empty0, const unsigned char *stages,
                                          int stage_bytes, int tid, float (&raw)[12], uint32_t (&wds)[6]) {
    mbar_wait(full0 + 8 * ring.stage, ring.phase);
    const unsigned char *src = stages + (size_t)ring.stage * stage_bytes;
    const float4 *c4 = reinterpret_cast<const float4 *>(src) + 3 * tid;
    const float4 a = c4[0], b = c4[1], c = c4[2];
    wds[0] = wds[1] = wds[2] = wds[3] = wds[4] = wds[5] = 0u;
    if (IMG == IMG_BYTES) {
        const uint32_t *iw = reinterpret_cast<const uint32_t *>(src + kCoordTileBytes) + 3 * tid;
        wds[0] = iw[0]; wds[1] = iw[1]; wds[2] = iw[2];
    }
    if (IMG == IMG_PACKED) wds[0] = reinterpret_cast<const uint32_t *>(src + kCoordTileBytes)[tid];
    raw[0] = a.x; raw[1] = a.y; raw[2] = a.z; raw[3] = a.w;
    raw[4] = b.x; raw[5] = b.y; raw[6] = b.z; raw[7] = b.w;
    raw[8] = c.x; raw[9] = c.y; raw[10] = c.z; raw[11] = c.w;
    // the values are in registers: hand the stage back.  One arrive per warp: __syncwarp() makes sure
    // every lane has issued its loads of this stage (lanes may have diverged in the rare exact-bin
    // path of the previous group) before lane 0 releases it.
    __syncwarp();
    mbar_arrive_if(empty0 + 8 * ring.stage, (tid & 31) == 0);
    ring.advance();
}

// ... and unwrap to float64.
template <int IMG, bool BOXC>
__device__ __forceinline__ void fetch_group(const CorrDev &p, Ring &ring, uint32_t full0, uint32_t empty0,
                                            const unsigned char *stages, int stage_bytes, int tid,
                                            const double *bx, double (&u)[12], float (&raw)[12]) {
    uint32_t wds[6];
    fetch_raw<IMG>(ring, full0, empty0, stages, stage_bytes, tid, raw, wds);
    // a box that is the same for every row comes from the kernel parameters (constant bank operands)
    unwrap12<IMG>(raw, wds, BOXC ? p.box_const : bx, u);
}

// Float32 displacements on rows with one constant float32 box, without conversions and without the float64
// pipe.  d = (x0 - n0 L) - (xt - nt L) = x0 - xt - k L with k = n0 - nt.  k = 0 (the common case): fl(x0 - xt) is
// the exact difference rounded once (relative 2^-24).  k != 0: the operand on the far side of the boundary is moved
// first -- k > 0: (x0 - kL) - xt, k < 0: x0 - (xt + kL) -- one FMA each, so the moved operand is the exact value
// rounded once and, being within |d| of the boundary (Sterbenz: exact when it lies in [L/2, L]), carries an error
// of at most ulp(d)/2; moving the OTHER operand would cost ulp(L)/2.  Every component of d is thus good to ~2^-23
// relative -- the accuracy of the float32 sinc (ISF), far inside the margin of the bin estimate (van Hove; the
// undecided samples are recomputed in float64 from the raw values, so the counts stay exact), and 1e-7 relative
// per atom on r^2, r^4 for alpha_2 when it rides along with the histogram (BASELINE config 3; alone, MSD and
// alpha_2 keep the float64 chain).  Image counts become floats by one byte-permute (image_float); all adds / fmas
// are packed pairs (components 2j, 2j+1 of the 12).  Packed counts: the three masked words (one per dimension, see
// packed_dim) of the lag row are subtracted from the biased ones of the origin as integers, the four bytes of each
// difference are 64 + 4^d k.
template <int IMG, int FLAGS, bool MASKED>
__device__ __forceinline__ void process_tile_f32d(const CorrDev &p, Ring &ring, uint32_t full0, uint32_t empty0,
                                                  const unsigned char *stages, int stage_bytes, int tid,
                                                  uint32_t vmask, int nl, double *s_acc, unsigned int *s_hist,
                                                  uint32_t trash) {
    constexpr bool kR2 = (FLAGS & MDB_SUM_R2) != 0;          // with R4: alpha_2 next to the histogram
    constexpr bool kSinc = (FLAGS & MDB_SUM_SINC) != 0;
    constexpr bool kHist = (FLAGS & MDB_HIST) != 0;
    constexpr int NV = (kR2 ? 2 : 0) + (kSinc ? 1 : 0);
    static_assert(IMG == IMG_BYTES || IMG == IMG_PACKED, "float32 displacements need image counts");
    const f32x2 *neg_box = reinterpret_cast<const f32x2 *>(p.neg_box2);      // dims of pair j: j % 3
    const f32x2 q2 = bcast2(p.q2f);
    const uint32_t magic = p.magic23;
    float raw0[12];
    uint32_t w0[6];
    fetch_raw<IMG>(ring, full0, empty0, stages, stage_bytes, tid, raw0, w0);
    f32x2 x0[6], f0[6];
    uint32_t o[3];
#pragma unroll
    for (int j = 0; j < 6; ++j) {
        x0[j] = pack2(raw0[2 * j], raw0[2 * j + 1]);
        if (IMG == IMG_BYTES)
            f0[j] = pack2(image_float(w0[(2 * j) >> 2], (2 * j) & 3, magic),
                          image_float(w0[(2 * j + 1) >> 2], (2 * j + 1) & 3, magic));
    }
    if (IMG == IMG_PACKED) {
#pragma unroll
        for (int d = 0; d < 3; ++d) o[d] = packed_dim(w0[0], d) + 0x40404040u;
    }
    const uint32_t table0 = smem_u32(s_hist) - 4u * 0x4B400000u;
#pragma unroll 1
    for (int k = 0; k < nl; ++k) {
        float rawt[12];
        uint32_t wt[6];
        fetch_raw<IMG>(ring, full0, empty0, stages, stage_bytes, tid, rawt, wt);
        float dv[12];
        // Between an origin and a lag frame few atoms change their image: when NO atom of the warp did (the image
        // words of the two rows are equal in every lane) the displacement is x0 - xt -- exactly what the general
        // formula below gives for k = 0 (fma(+0, -L, x) = x), so the results are bit-identical; it saves 48 of the
        // 155 (ISF) instructions per 4 evaluations.  Warp-uniform, so there is no divergence.
        bool same = wt[0] == w0[0];
        if (IMG == IMG_BYTES) same = same && wt[1] == w0[1] && wt[2] == w0[2];
        if (__all_sync(0xffffffffu, same && !p.no_shortcut)) {
#pragma unroll
            for (int j = 0; j < 6; ++j)
                unpack2(sub2(x0[j], pack2(rawt[2 * j], rawt[2 * j + 1])), dv[2 * j], dv[2 * j + 1]);
        } else {
        uint32_t kd[3];
        if (IMG == IMG_PACKED) {
#pragma unroll
            for (int d = 0; d < 3; ++d) kd[d] = o[d] - packed_dim(wt[0], d);     // bytes: 64 + 4^d (n0 - nt)
        }
#pragma unroll
        for (int j = 0; j < 6; ++j) {
            f32x2 dn;                                                            // (scaled) n0 - nt, exact
            if (IMG == IMG_BYTES) {
                const f32x2 ft = pack2(image_float(wt[(2 * j) >> 2], (2 * j) & 3, magic),
                                       image_float(wt[(2 * j + 1) >> 2], (2 * j + 1) & 3, magic));
                dn = sub2(f0[j], ft);
            } else {
                const int c0 = 2 * j, c1 = 2 * j + 1;
                dn = add2(pack2(image_float(kd[c0 % 3], c0 / 3, magic), image_float(kd[c1 % 3], c1 / 3, magic)),
                          bcast2(-8388672.0f));
            }
            float dn0, dn1;
            unpack2(dn, dn0, dn1);
            const f32x2 up = pack2(fmaxf(dn0, 0.0f), fmaxf(dn1, 0.0f));     // images the origin is moved by
            const f32x2 dw = sub2(up, dn);                                   // images the lag frame is moved by
            const f32x2 om = fma2(up, neg_box[j % 3], x0[j]);
            const f32x2 tm = fma2(dw, neg_box[j % 3], pack2(rawt[2 * j], rawt[2 * j + 1]));
            unpack2(sub2(om, tm), dv[2 * j], dv[2 * j + 1]);
        }
        }
        float r2[4];
#pragma unroll
        for (int a = 0; a < 4; ++a)
            r2[a] = fmaf(dv[3 * a + 2], dv[3 * a + 2], fmaf(dv[3 * a + 1], dv[3 * a + 1], dv[3 * a] * dv[3 * a]));
        double *col = s_acc + (size_t)k * NV * kThreads + tid;
        if constexpr (kR2) {
            float m0 = r2[0], m1 = r2[1], m2 = r2[2], m3 = r2[3];
            if (MASKED) {
                m0 = (vmask & 1u) ? m0 : 0.0f;
                m1 = (vmask & 2u) ? m1 : 0.0f;
                m2 = (vmask & 4u) ? m2 : 0.0f;
                m3 = (vmask & 8u) ? m3 : 0.0f;
            }
            col[0] += (double)((m0 + m1) + (m2 + m3));
            col[kThreads] += (double)fmaf(m0, m0, fmaf(m1, m1, fmaf(m2, m2, m3 * m3)));
        }
        if constexpr (kSinc) {
            float y0, y1, y2, y3, v0, v1, v2, v3;
            unpack2(mul2(pack2(r2[0], r2[1]), q2), y0, y1);
            unpack2(mul2(pack2(r2[2], r2[3]), q2), y2, y3);
            sinc_of_square_x2(y0, y1, v0, v1);
            sinc_of_square_x2(y2, y3, v2, v3);
            if (MASKED) {
                v0 = (vmask & 1u) ? v0 : 0.0f;
                v1 = (vmask & 2u) ? v1 : 0.0f;
                v2 = (vmask & 4u) ? v2 : 0.0f;
                v3 = (vmask & 8u) ? v3 : 0.0f;
            }
            if (MASKED && FLAGS == MDB_SUM_SINC && p.n_segments > 1) {
                // several sub-ensembles in one pass: atom a adds to the accumulator of its sub-ensemble (the byte of
                // the selector, carried in vmask); column layout [lag][sub-ensemble][thread]
                const uint32_t s0 = (vmask >> 4) & 63u, s1 = (vmask >> 10) & 63u, s2 = (vmask >> 16) & 63u,
                               s3 = (vmask >> 22) & 63u;
                double *seg_col = s_acc + (size_t)k * p.n_segments * kThreads + tid;
                for (uint32_t j = 1; j <= (uint32_t)p.n_segments; ++j) {
                    const float part = ((s0 == j ? v0 : 0.0f) + (s1 == j ? v1 : 0.0f)) +
                                       ((s2 == j ? v2 : 0.0f) + (s3 == j ? v3 : 0.0f));
                    seg_col[(size_t)(j - 1) * kThreads] += (double)part;
                }
            } else {
                col[(kR2 ? 2 : 0) * kThreads] += (double)((v0 + v1) + (v2 + v3));
            }
        }
        if constexpr (kHist) {
            const uint32_t slow = hist_fast4<MASKED>(p, r2, vmask, table0 + 4u * (uint32_t)(k * p.n_bins), trash);
            if (slow) {
#pragma unroll
                for (int a = 0; a < 4; ++a)
                    if (slow & (1u << a))
                        exact_count_raw(p.box_const[0], p.box_const[1] * (IMG == IMG_PACKED ? 4.0 : 1.0),
                                        p.box_const[2] * (IMG == IMG_PACKED ? 16.0 : 1.0), p.dx, p.inv_dx, p.n_bins,
                                        raw0[3 * a], raw0[3 * a + 1], raw0[3 * a + 2], image_count_int<IMG>(w0, 3 * a),
                                        image_count_int<IMG>(w0, 3 * a + 1), image_count_int<IMG>(w0, 3 * a + 2),
                                        rawt[3 * a], rawt[3 * a + 1], rawt[3 * a + 2], image_count_int<IMG>(wt, 3 * a),
                                        image_count_int<IMG>(wt, 3 * a + 1), image_count_int<IMG>(wt, 3 * a + 2),
                                        s_hist + k * p.n_bins);
            }
        }
    }
}

// Does this instantiation have a float32-displacement variant?  (selected per launch by p.f32d)
template <int IMG, int FLAGS, bool BOXC>
struct HasF32d {
    static constexpr bool value = (IMG == IMG_BYTES || IMG == IMG_PACKED) && BOXC &&
                                  (FLAGS == MDB_SUM_SINC || FLAGS == MDB_HIST ||
                                   FLAGS == (MDB_SUM_R2 | MDB_SUM_R4 | MDB_HIST));
};

// One tile of 1024 atoms: origin group into registers, then the lag rows of the chunk one by one.
// The loop over lags is NOT unrolled (the body is 4 evaluations, ~1 KiB of code, and stays in the
// instruction cache); per-lag accumulators live in shared memory, one private column per thread.
template <int IMG, int FLAGS, bool MASKED, bool BOXC>
__device__ __forceinline__ void process_tile(const CorrDev &p, Ring &ring, uint32_t full0, uint32_t empty0,
                                             const unsigned char *stages, int stage_bytes, int tid, uint32_t vmask,
                                             int nl,
                                             const double (*s_box)[3], double *s_acc, unsigned int *s_hist,
                                             const double *s_edges, uint32_t trash) {
    constexpr bool kR2 = (FLAGS & (MDB_SUM_R2 | MDB_SUM_LEGENDRE)) != 0;
    constexpr bool kR4 = (FLAGS & MDB_SUM_R4) != 0;
    constexpr bool kSinc = (FLAGS & MDB_SUM_SINC) != 0;
    constexpr bool kHist = (FLAGS & MDB_HIST) != 0;
    constexpr bool kHistF32 = kHist && IMG == IMG_NONE;
    constexpr int NV = (kR2 ? 1 : 0) + (kR4 ? 1 : 0) + (kSinc ? 1 : 0);
    if constexpr (HasF32d<IMG, FLAGS, BOXC>::value) {
        if (p.f32d) {          // uniform over the launch
            process_tile_f32d<IMG, FLAGS, MASKED>(p, ring, full0, empty0, stages, stage_bytes, tid, vmask, nl, s_acc,
                                                  s_hist, trash);
            return;
        }
    }
    double u0[12];
    float raw0[12];
    fetch_group<IMG, BOXC>(p, ring, full0, empty0, stages, stage_bytes, tid, s_box[0], u0, raw0);
#pragma unroll 1
    for (int k = 0; k < nl; ++k) {
        double ut[12];
        float rawt[12];
        fetch_group<IMG, BOXC>(p, ring, full0, empty0, stages, stage_bytes, tid, s_box[k + 1], ut, rawt);
        double s2, s4;
        float ss;
        eval_group<FLAGS, kHistF32, MASKED>(p, u0, raw0, ut, rawt, vmask, s2, s4, ss, s_hist + k * p.n_bins, s_edges,
                                            trash);
        double *col = s_acc + (size_t)k * NV * kThreads + tid;
        int v = 0;
        if constexpr (kR2) { col[v * kThreads] += s2; ++v; }
        if constexpr (kR4) { col[v * kThreads] += s4; ++v; }
        if constexpr (kSinc) { col[v * kThreads] += (double)ss; ++v; }
    }
}

template <int IMG, int FLAGS, bool BOXC>
__global__ void __launch_bounds__(kBulkThreads, MDB_BULK_CTAS_PER_SM) corr_bulk_kernel(const CorrDev p, const int n_stages) {
    constexpr bool kR2 = (FLAGS & (MDB_SUM_R2 | MDB_SUM_LEGENDRE)) != 0;
    constexpr bool kR4 = (FLAGS & MDB_SUM_R4) != 0;
    constexpr bool kSinc = (FLAGS & MDB_SUM_SINC) != 0;
    constexpr bool kHist = (FLAGS & MDB_HIST) != 0;
    constexpr int NV = (kR2 ? 1 : 0) + (kR4 ? 1 : 0) + (kSinc ? 1 : 0);
    constexpr int kImgBytes = img_tile_bytes(IMG);
    constexpr int kStageBytes = kCoordTileBytes + kImgBytes;
    constexpr int kImgElem = IMG == IMG_PACKED ? 4 : 1;          // bytes per element of an image row

    // dynamic shared memory: [n_stages stages][accumulators NV*kLagChunk*kThreads doubles][hist][edges]
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ int s_row[kLagChunk + 1];
    __shared__ double s_box[kLagChunk + 1][3];
    __shared__ double s_red[kWarps][kLagChunk * 3];
    __shared__ __align__(8) uint64_t s_full[kMaxStages];
    __shared__ __align__(8) uint64_t s_empty[kMaxStages];

    const CtaWork w = cta_work(p);
    const int tid = threadIdx.x;
    if (p.stamp != nullptr && (blockIdx.x | blockIdx.y | tid) == 0) *p.stamp = global_timer_ns();
    const int n_seg = (FLAGS == MDB_SUM_SINC && p.n_segments > 1) ? p.n_segments : 1;     // sub-ensembles per pass
    double *s_acc = reinterpret_cast<double *>(smem_raw + (size_t)n_stages * kStageBytes);
    unsigned int *s_hist = reinterpret_cast<unsigned int *>(s_acc + (size_t)NV * kLagChunk * n_seg * kThreads);
    double *s_edges = reinterpret_cast<double *>(s_hist + (size_t)kLagChunk * p.n_bins);
    const uint32_t trash = trash_slot(p, s_edges, tid);

    load_rows<IMG>(p, w, tid, s_row, s_box);
    if (tid == kThreads) {
        for (int i = 0; i < n_stages; ++i) {
            mbar_init(&s_full[i], 1);          // one arrive.expect_tx by the producer
            mbar_init(&s_empty[i], kWarps);    // one arrive per compute warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < kThreads)
        for (int i = 0; i < NV * kLagChunk * n_seg; ++i) s_acc[(size_t)i * kThreads + tid] = 0.0;
    if (kHist) {
        for (int i = tid; i < kLagChunk * p.n_bins; i += kBulkThreads) s_hist[i] = 0u;
        if (p.bin_mode == MDB_BINS_EDGES)
            for (int i = tid; i <= p.n_bins; i += kBulkThreads) s_edges[i] = p.edges[i];
    }
    __syncthreads();

    // atom range of this CTA in tiles of kTileAtoms (groups_per_block is a multiple of kThreads)
    const int g_begin = w.ab * p.groups_per_block;
    const int g_end = min(p.n_groups, g_begin + p.groups_per_block);
    const int n_tiles = (g_end - g_begin + kThreads - 1) / kThreads;
    const int n_units = w.nl + 1;   // origin + lag rows per tile
    const uint32_t full0 = smem_u32(&s_full[0]), empty0 = smem_u32(&s_empty[0]);
    Ring ring{0u, 0u, (uint32_t)n_stages};

    if (tid >= kThreads) {
        // ===== producer warp: one elected lane feeds the ring =====
        if (tid == kThreads) {
            const long long row_floats = p.pitch;
            const uint32_t stages0 = smem_u32(smem_raw);
            for (int t = 0; t < n_tiles; ++t) {
                const long long first_group = g_begin + (long long)t * kThreads;
                const long long first_float = 12ll * first_group;
                // copy what the shard's atoms need, rounded up to the 16-byte granule of the bulk copy (rows
                // are readable that far: pitch % 4 == 0 floats, image pitch % 16 == 0 bytes)
                const long long left = 3ll * p.n_atoms - first_float;          // components still to come
                const uint32_t cbytes = (uint32_t)min((long long)kCoordTileBytes, ((left + 3) / 4) * 16);
                uint32_t ibytes = 0u;
                if (IMG == IMG_BYTES) ibytes = (uint32_t)min((long long)kImgBytes, ((left + 15) / 16) * 16);
                if (IMG == IMG_PACKED) ibytes = (uint32_t)min((long long)kImgBytes, ((left + 47) / 48) * 16);
                const long long img_first = IMG == IMG_PACKED ? first_group : first_float;   // elements into the row
                for (int u = 0; u < n_units; ++u) {
                    const int row = s_row[u];
                    mbar_wait(empty0 + 8 * ring.stage, ring.phase ^ 1);
                    const uint32_t dst = stages0 + ring.stage * kStageBytes;
                    const uint32_t bar = full0 + 8 * ring.stage;
                    mbar_expect_tx(bar, cbytes + ibytes);
                    bulk_g2s(dst, static_cast<const float *>(p.coords) + (long long)row * row_floats + first_float,
                             cbytes, bar);
                    if (IMG != IMG_NONE)
                        bulk_g2s(dst + kCoordTileBytes,
                                 static_cast<const unsigned char *>(p.images) +
                                     ((long long)row * p.img_pitch + img_first) * kImgElem,
                                 ibytes, bar);
                    ring.advance();
                }
            }
        }
        return;
    }

    // ===== compute warps =====
    for (int t = 0; t < n_tiles; ++t) {
        const int g_last = g_begin + t * kThreads + kThreads - 1;          // last group of this tile
        if (4 * g_last + 3 < p.n_atoms && p.select == nullptr) {
            process_tile<IMG, FLAGS, false, BOXC>(p, ring, full0, empty0, smem_raw, kStageBytes, tid, 0xfu, w.nl, s_box,
                                                  s_acc, s_hist, s_edges, trash);
        } else {
            const uint32_t vmask = group_mask(p, w.s, g_begin + t * kThreads + tid);
            process_tile<IMG, FLAGS, true, BOXC>(p, ring, full0, empty0, smem_raw, kStageBytes, tid, vmask, w.nl, s_box,
                                                 s_acc, s_hist, s_edges, trash);
        }
    }

    for (int seg = 0; seg < n_seg; ++seg) {
        Accum<FLAGS> acc;
#pragma unroll
        for (int k = 0; k < kLagChunk; ++k) {
            const double *col = s_acc + ((size_t)k * NV * n_seg + seg) * kThreads + tid;
            int v = 0;
            if constexpr (kR2) { acc.r2[k] = col[v * kThreads]; ++v; }
            if constexpr (kR4) { acc.r4[k] = col[v * kThreads]; ++v; }
            if constexpr (kSinc) { acc.sc[k] = col[v * kThreads]; ++v; }
        }
        cta_epilogue<FLAGS>(p, w, tid, acc, s_red, s_hist, [] { compute_barrier(); }, seg, n_seg);
    }
}
#endif  // MDB_CORR_PART != 0

// =================================================================================================
// launchers of this translation unit
// =================================================================================================
template <typename Kernel, typename... Extra>
int launch_kernel(Kernel kern, const CorrDev &p, const Plan &plan, int threads, cudaStream_t st, LaunchTiming tm,
                  Extra... extra) {
    const long long blocks = (long long)p.S * p.n_chunks * p.n_ablocks;
    if (blocks > 2147483647ll) return fail(MDB_ERR_UNSUPPORTED, "grid too large");
    if (plan.smem > 40 * 1024)
        MDB_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.smem));
    if (tm.start) MDB_CUDA_CHECK(cudaEventRecord(tm.start, st));
    kern<<<(unsigned)blocks, threads, plan.smem, st>>>(p, extra...);
    MDB_CUDA_CHECK(cudaGetLastError());
    if (tm.stop) MDB_CUDA_CHECK(cudaEventRecord(tm.stop, st));
    return MDB_OK;
}

#if MDB_CORR_PART == 0
template <typename CT, int IMG, int FLAGS>
int launch_ldg_flags(const CorrDev &p, const Plan &plan, cudaStream_t st, LaunchTiming tm) {
    if (plan.vec) return launch_kernel(corr_ldg_kernel<CT, IMG, FLAGS, true>, p, plan, kThreads, st, tm);
    return launch_kernel(corr_ldg_kernel<CT, IMG, FLAGS, false>, p, plan, kThreads, st, tm);
}

template <typename CT, int IMG>
int launch_ldg_typed(const CorrDev &p, const Plan &plan, int flags, cudaStream_t st, LaunchTiming tm) {
    constexpr int R2 = MDB_SUM_R2, R4 = MDB_SUM_R4, SC = MDB_SUM_SINC, H = MDB_HIST;
    switch (kernel_flags(flags)) {
        case R2: return launch_ldg_flags<CT, IMG, R2>(p, plan, st, tm);
        case SC: return launch_ldg_flags<CT, IMG, SC>(p, plan, st, tm);
        case R2 | R4: return launch_ldg_flags<CT, IMG, R2 | R4>(p, plan, st, tm);
        case H: return launch_ldg_flags<CT, IMG, H>(p, plan, st, tm);
        case R2 | R4 | H: return launch_ldg_flags<CT, IMG, R2 | R4 | H>(p, plan, st, tm);
        case MDB_SUM_LEGENDRE:
            if constexpr (IMG == IMG_NONE) return launch_ldg_flags<CT, IMG_NONE, MDB_SUM_LEGENDRE>(p, plan, st, tm);
            return fail(MDB_ERR_ARGUMENT, "MDB_SUM_LEGENDRE needs raw rows");
        case R2 | R4 | SC | H: return launch_ldg_flags<CT, IMG, R2 | R4 | SC | H>(p, plan, st, tm);
        default: return launch_ldg_flags<CT, IMG, R2 | R4 | SC>(p, plan, st, tm);
    }
}
#else
template <int IMG, int FLAGS>
int launch_bulk_flags(const CorrDev &p, const Plan &plan, cudaStream_t st, LaunchTiming tm) {
    if (IMG != IMG_NONE && p.box == nullptr)
        return launch_kernel(corr_bulk_kernel<IMG, FLAGS, IMG != IMG_NONE>, p, plan, kBulkThreads, st, tm, plan.stages);
    return launch_kernel(corr_bulk_kernel<IMG, FLAGS, false>, p, plan, kBulkThreads, st, tm, plan.stages);
}

template <int IMG>
int launch_bulk(const CorrDev &p, const Plan &plan, int flags, cudaStream_t st, LaunchTiming tm) {
    constexpr int R2 = MDB_SUM_R2, R4 = MDB_SUM_R4, SC = MDB_SUM_SINC, H = MDB_HIST;
    switch (kernel_flags(flags)) {
        case R2: return launch_bulk_flags<IMG, R2>(p, plan, st, tm);
        case SC: return launch_bulk_flags<IMG, SC>(p, plan, st, tm);
        case R2 | R4: return launch_bulk_flags<IMG, R2 | R4>(p, plan, st, tm);
        case H: return launch_bulk_flags<IMG, H>(p, plan, st, tm);
        case R2 | R4 | H: return launch_bulk_flags<IMG, R2 | R4 | H>(p, plan, st, tm);
        case MDB_SUM_LEGENDRE:
            if constexpr (IMG == IMG_NONE) return launch_bulk_flags<IMG_NONE, MDB_SUM_LEGENDRE>(p, plan, st, tm);
            return fail(MDB_ERR_ARGUMENT, "MDB_SUM_LEGENDRE needs raw rows");
        case R2 | R4 | SC | H: return launch_bulk_flags<IMG, R2 | R4 | SC | H>(p, plan, st, tm);
        default: return launch_bulk_flags<IMG, R2 | R4 | SC>(p, plan, st, tm);
    }
}
#endif

}  // namespace

#if MDB_CORR_PART == 0
int corr_launch_ldg(const CorrDev &p, const Plan &plan, int flags, int coord_type, cudaStream_t st, LaunchTiming tm) {
    if (coord_type == MDB_F64) return launch_ldg_typed<double, IMG_NONE>(p, plan, flags, st, tm);
    switch (plan.img) {
        case IMG_NONE: return launch_ldg_typed<float, IMG_NONE>(p, plan, flags, st, tm);
        case IMG_BYTES: return launch_ldg_typed<float, IMG_BYTES>(p, plan, flags, st, tm);
        case IMG_I16: return launch_ldg_typed<float, IMG_I16>(p, plan, flags, st, tm);
        default: return fail(MDB_ERR_UNSUPPORTED, "the direct-load kernel reads byte or int16 image counts");
    }
}
#elif MDB_CORR_PART == 1
int corr_launch_bulk_none(const CorrDev &p, const Plan &plan, int flags, cudaStream_t st, LaunchTiming tm) {
    return launch_bulk<IMG_NONE>(p, plan, flags, st, tm);
}
#elif MDB_CORR_PART == 2
int corr_launch_bulk_bytes(const CorrDev &p, const Plan &plan, int flags, cudaStream_t st, LaunchTiming tm) {
    return launch_bulk<IMG_BYTES>(p, plan, flags, st, tm);
}
#elif MDB_CORR_PART == 3
int corr_launch_bulk_packed(const CorrDev &p, const Plan &plan, int flags, cudaStream_t st, LaunchTiming tm) {
    return launch_bulk<IMG_PACKED>(p, plan, flags, st, tm);
}
#endif

}  // namespace mdb
