// Dense schedules (every origin x a contiguous range of lags, averaged over origins).
//
// The streaming kernel of corr_kernels.cu reads one lag row per (origin, lag) pair: 15 bytes per
// evaluation, HBM-bound.  With linear lags neighbouring origins need almost the same frames, so
// here a CTA works on a BLOCK of kS origins x kL lags: it brings the kS origin rows and the
// (kS-1)*step + kL partner rows of one 64-atom tile into shared memory -- unwrapped to float64
// once per row (structure of arrays, conflict-free) -- and evaluates the kS*kL pairs from there.
// Each row is reused by up to kL pairs: ~3 bytes of HBM traffic per evaluation, and no per-pair
// widening / unwrapping, which leaves ~30 instructions per evaluation (bound by instruction issue).
//
// One CTA owns one lag block for a slice of the atom tiles and walks over ALL origin blocks, so the
// per-lag sums over origins and atoms live in registers until the very end: thread (atom i, slice j)
// handles the lags l = j, j+4, ..., j+28 of the block (kL = 32).  Partials per atom split are added by a small
// second kernel in fixed order (bit-reproducible).
#include <stdlib.h>

#include <type_traits>

#include "common.cuh"

namespace mdb {
namespace {

#ifndef MDB_DENSE_ATOMS
#define MDB_DENSE_ATOMS 64
#endif
#ifndef MDB_DENSE_CTAS
#define MDB_DENSE_CTAS 2
#endif
constexpr int kTileAtoms = MDB_DENSE_ATOMS;           // atoms per tile (a multiple of 32)
constexpr int kSlices = 256 / kTileAtoms;             // lag slices: thread = (atom, slice)
constexpr int kWarpsPerSlice = kTileAtoms / 32;
constexpr int kDenseThreads = kTileAtoms * kSlices;   // 256
#ifndef MDB_DENSE_KL
#define MDB_DENSE_KL 32
#endif
constexpr int kL = MDB_DENSE_KL;                      // lags per block
constexpr int kSlots = kL / kSlices;                  // lag slots per thread
constexpr int kMaxS = 16;                             // origins per block (fewer for origin_step > 1)
constexpr int kMaxPartnerRows = 32 + kL;
constexpr int kQuadsPerRow = 3 * kTileAtoms / 4;     // 48 float4 per row of the tile
constexpr int kStageBatch = 5;                       // independent 128-bit loads per thread in flight while staging (12 was slower: 128 registers)

struct DenseDev {
    const float *coords;
    long long pitch;
    const uint8_t *images;
    long long img_pitch;
    const double *box;
    double box_const[3];
    const int *frame_rows;
    int n_atoms;
    int o0, ostep, S;
    int i0, L;
    double q2;
    int ks;                 // origins per block
    int n_sblocks, n_lblocks, n_splits, tiles_per_split, n_tiles;
    double *partial;        // [n_splits][L][2]
    int stage_f32;          // v2: the box is a float32 number in every dimension -> offsets without the float64 pipe
    float neg_box_f[3];     // -box as float32 (exact when stage_f32)
    uint32_t magic23;       // 0x4B000000 as a register operand (image_float, common.cuh)
};

__device__ __forceinline__ int row_of(const DenseDev &p, int frame) {
    return p.frame_rows ? p.frame_rows[frame] : frame;
}

// One quad (4 consecutive interleaved values = one 128-bit load + one 32-bit image word) of a frame
// row of the tile.  Rows are 16-byte aligned (checked on the host); quads that start beyond the
// shard's last component read nothing.
struct Quad {
    float4 x;
    uint32_t n;
};

template <bool IMG>
__device__ __forceinline__ Quad load_quad(const DenseDev &p, int frame, int atom0, int qd) {
    Quad q;
    q.x = make_float4(0.f, 0.f, 0.f, 0.f);
    q.n = 0x80808080u;
    const long long e = 3ll * atom0 + 4 * qd;
    if (e < 3ll * p.n_atoms) {
        const int row = row_of(p, frame);
        q.x = __ldg(reinterpret_cast<const float4 *>(p.coords + (long long)row * p.pitch + e));
        if (IMG) q.n = __ldg(reinterpret_cast<const uint32_t *>(p.images + (long long)row * p.img_pitch + e));
    }
    return q;
}

// Unwrap the quad (pbc.py:235-247, the same expression as the streaming kernel) into the row's
// structure-of-arrays float64 tile [3][64].
template <bool IMG>
__device__ __forceinline__ void store_quad(const DenseDev &p, const Quad &q, int frame, int qd, double *dst) {
    double bx = 0.0, by = 0.0, bz = 0.0;
    if (IMG) {
        if (p.box) {
            const int row = row_of(p, frame);
            bx = __ldg(p.box + 3ll * row); by = __ldg(p.box + 3ll * row + 1); bz = __ldg(p.box + 3ll * row + 2);
        } else {
            bx = p.box_const[0]; by = p.box_const[1]; bz = p.box_const[2];
        }
    }
    const float xs[4] = {q.x.x, q.x.y, q.x.z, q.x.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int v = 4 * qd + k;
        const int atom = v / 3, c = v - 3 * atom;
        double value = (double)xs[k];
        if (IMG) value = fma(-image_count(q.n, k), c == 0 ? bx : (c == 1 ? by : bz), value);
        dst[c * kTileAtoms + atom] = value;
    }
}

// Candidate staging (MDB_DENSE_STAGE=warp, NOT the default: written at the end of round 1 without GPU time left
// to measure it).  The loop above spends ~170 instructions per quad on index arithmetic -- a division by 48
// per quad, v / 3 per value, bounds and row-map tests -- which is 58 % (MSD) / 39 % (ISF) of all instructions the
// kernel executes (scripts/sass_budget.py).  Here a WARP owns rows r = warp, warp + 8, ...; lane l always takes
// quad l and, for l < kQuadsPerRow - 32, quad l + 32 of a row, so the position of each of its values in the
// [3][tile] row (component, atom) and the box length it needs are per-lane constants computed once per kernel;
// kRowBatch rows are loaded before any is unwrapped (memory parallelism as before).
constexpr int kRowBatch = 4;

struct LaneMap {
    double *dst[2][4];    // where value k of the lane's quad 0 / 1 goes in row 0 of the structure-of-arrays tile
    double box[2][4];     // box length of that value's component (constant box only)
};

__device__ __forceinline__ LaneMap make_lane_map(const DenseDev &p, int lane, double *s_rows) {
    LaneMap m;
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int v = 4 * (lane + 32 * h) + k;
            const int atom = v / 3, c = v - 3 * atom;
            m.dst[h][k] = s_rows + c * kTileAtoms + atom;
            m.box[h][k] = c == 0 ? p.box_const[0] : (c == 1 ? p.box_const[1] : p.box_const[2]);
        }
    return m;
}

// FULL: every quad of the tile lies inside the shard (all tiles but the last): no bounds tests at all.
// Rows past the end of the block are clamped to its last row: a few warps then stage that row again with the
// same values, which is cheaper than predicating every load and store of the batch.
template <bool IMG, bool FULL>
__device__ __forceinline__ void stage_rows_warp(const DenseDev &p, const LaneMap &map, int lane, int warp, int atom0,
                                                int ns, int n_rows, int f_origin0, int f_partner0) {
    static_assert(kQuadsPerRow > 32 && kQuadsPerRow <= 64, "two quads per lane cover a row");
    constexpr int row_doubles = 3 * kTileAtoms;
    const bool second = lane + 32 < kQuadsPerRow;
    const long long e0 = 3ll * atom0 + 4 * lane, e_end = 3ll * p.n_atoms;
    const bool live0 = FULL || e0 < e_end, live1 = second && (FULL || e0 + 128 < e_end);
    const float *src0 = p.coords + e0;
    const uint8_t *img0 = IMG ? p.images + e0 : nullptr;
    for (int rb = warp; rb < n_rows; rb += 8 * kRowBatch) {
        float4 x[kRowBatch][2];
        uint32_t n[kRowBatch][2];
#pragma unroll
        for (int u = 0; u < kRowBatch; ++u) {
            const int r = min(rb + 8 * u, n_rows - 1);
            const int frame = (r < ns) ? f_origin0 + r * p.ostep : f_partner0 + (r - ns);
            const long long row = row_of(p, frame);
            x[u][0] = x[u][1] = make_float4(0.f, 0.f, 0.f, 0.f);
            n[u][0] = n[u][1] = 0x80808080u;
            if (live0) {
                x[u][0] = __ldg(reinterpret_cast<const float4 *>(src0 + row * p.pitch));
                if (IMG) n[u][0] = __ldg(reinterpret_cast<const uint32_t *>(img0 + row * p.img_pitch));
            }
            if (live1) {
                x[u][1] = __ldg(reinterpret_cast<const float4 *>(src0 + row * p.pitch + 128));
                if (IMG) n[u][1] = __ldg(reinterpret_cast<const uint32_t *>(img0 + row * p.img_pitch + 128));
            }
        }
#pragma unroll
        for (int u = 0; u < kRowBatch; ++u) {
            const int r = min(rb + 8 * u, n_rows - 1);
            const int shift = r * row_doubles;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                if (h == 0 || second) {
                    const float xs[4] = {x[u][h].x, x[u][h].y, x[u][h].z, x[u][h].w};
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        double value = (double)xs[k];
                        if (IMG) value = fma(-image_count(n[u][h], k), map.box[h][k], value);
                        map.dst[h][k][shift] = value;
                    }
                }
            }
        }
    }
}

template <bool IMG, int FLAGS, bool STAGE_WARP>
__global__ void __launch_bounds__(kDenseThreads, MDB_DENSE_CTAS) dense_corr_kernel(const DenseDev p) {
    constexpr bool kR2 = (FLAGS & MDB_SUM_R2) != 0;
    constexpr bool kSinc = (FLAGS & MDB_SUM_SINC) != 0;
    extern __shared__ __align__(16) double s_rows[];     // [ks + n_partner][3][64]
    __shared__ double s_red[kDenseThreads / 32][kSlots][2];

    const int lb = blockIdx.x % p.n_lblocks;
    const int split = blockIdx.x / p.n_lblocks;
    const int lam = lb * kL;                              // first lag of the block
    const int nl = min(kL, p.L - lam);
    const int tid = threadIdx.x;
    const int i = tid % kTileAtoms, j = tid / kTileAtoms;
    const int row_doubles = 3 * kTileAtoms;

    double acc2[kSlots], accs[kSlots];
#pragma unroll
    for (int m = 0; m < kSlots; ++m) acc2[m] = accs[m] = 0.0;
    LaneMap lane_map;
    if constexpr (STAGE_WARP) lane_map = make_lane_map(p, tid & 31, s_rows);

    const int t_begin = split * p.tiles_per_split;
    const int t_end = min(p.n_tiles, t_begin + p.tiles_per_split);
    for (int tile = t_begin; tile < t_end; ++tile) {
        const int atom0 = tile * kTileAtoms;
        const bool live = atom0 + i < p.n_atoms;
        for (int sb = 0; sb < p.n_sblocks; ++sb) {
            const int s0 = sb * p.ks;
            const int ns = min(p.ks, p.S - s0);
            const int f_origin0 = p.o0 + s0 * p.ostep;
            const int f_partner0 = f_origin0 + p.i0 + lam;
            const int n_partner = (ns - 1) * p.ostep + nl;
            __syncthreads();                              // the previous block is done with the rows
            const int n_rows = ns + n_partner;
            if constexpr (STAGE_WARP) {
                if (atom0 + kTileAtoms <= p.n_atoms)
                    stage_rows_warp<IMG, true>(p, lane_map, tid & 31, tid >> 5, atom0, ns, n_rows, f_origin0, f_partner0);
                else
                    stage_rows_warp<IMG, false>(p, lane_map, tid & 31, tid >> 5, atom0, ns, n_rows, f_origin0, f_partner0);
            } else {
            // stage the rows: kStageBatch independent 128-bit loads per thread in flight, then unwrap + store
            const int n_quads = n_rows * kQuadsPerRow;
            for (int w0 = tid; w0 < n_quads; w0 += kDenseThreads * kStageBatch) {
                Quad q[kStageBatch];
#pragma unroll
                for (int u = 0; u < kStageBatch; ++u) {
                    const int w = w0 + u * kDenseThreads;
                    if (w < n_quads) {
                        const int r = w / kQuadsPerRow, qd = w - kQuadsPerRow * r;
                        const int frame = (r < ns) ? f_origin0 + r * p.ostep : f_partner0 + (r - ns);
                        q[u] = load_quad<IMG>(p, frame, atom0, qd);
                    }
                }
#pragma unroll
                for (int u = 0; u < kStageBatch; ++u) {
                    const int w = w0 + u * kDenseThreads;
                    if (w < n_quads) {
                        const int r = w / kQuadsPerRow, qd = w - kQuadsPerRow * r;
                        const int frame = (r < ns) ? f_origin0 + r * p.ostep : f_partner0 + (r - ns);
                        store_quad<IMG>(p, q[u], frame, qd, s_rows + (size_t)r * row_doubles);
                    }
                }
            }
            }
            __syncthreads();
            if (live) {
                // Branch-free inner loop: the 4 lag slots of a thread are always evaluated.  In the last
                // lag block some slots lie beyond n_lags: they read rows that are allocated (stale data)
                // and their sums are simply never written out.
                float fs[kSlots];                          // float32 sinc sums over the origins of this block
#pragma unroll
                for (int m = 0; m < kSlots; ++m) fs[m] = 0.f;
                const double *rp0 = s_rows + (size_t)(ns + j) * row_doubles + i;
                for (int a = 0; a < ns; ++a) {
                    const double *ro = s_rows + (size_t)a * row_doubles + i;
                    const double ox = ro[0], oy = ro[kTileAtoms], oz = ro[2 * kTileAtoms];
                    const double *rp = rp0 + (size_t)a * p.ostep * row_doubles;
                    float yq[kSlots];
#pragma unroll
                    for (int m = 0; m < kSlots; ++m) {
                        const double *rq = rp + (size_t)(kSlices * m) * row_doubles;
                        const double dx = ox - rq[0], dy = oy - rq[kTileAtoms], dz = oz - rq[2 * kTileAtoms];
                        const double r2 = fma(dz, dz, fma(dy, dy, dx * dx));
                        if constexpr (kR2) acc2[m] += r2;
                        if constexpr (kSinc) yq[m] = (float)(r2 * p.q2);
                    }
                    if constexpr (kSinc) {
#pragma unroll
                        for (int m = 0; m < kSlots; m += 2) {
                            float v0, v1;
                            sinc_of_square_x2(yq[m], yq[m + 1], v0, v1);
                            fs[m] += v0;
                            fs[m + 1] += v1;
                        }
                    }
                }
                if constexpr (kSinc) {
#pragma unroll
                    for (int m = 0; m < kSlots; ++m) accs[m] += (double)fs[m];
                }
            }
        }
    }

    // ---- reduce over the 64 atoms of every (slice j, slot m): two warps per slice ----
    const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
    for (int m = 0; m < kSlots; ++m) {
        const double v2 = warp_sum(acc2[m]);
        const double vs = warp_sum(accs[m]);
        if (lane == 0) {
            s_red[warp][m][0] = v2;
            s_red[warp][m][1] = vs;
        }
    }
    __syncthreads();
    if (tid < kL * 2) {
        const int l = tid / 2, comp = tid & 1;          // lag of the block, component
        const int jj = l % kSlices, m = l / kSlices;    // which slice / slot holds it
        if (l < nl) {
            double v = 0.0;
#pragma unroll
            for (int wv = 0; wv < kWarpsPerSlice; ++wv) v += s_red[kWarpsPerSlice * jj + wv][m][comp];
            p.partial[((long long)split * p.L + lam + l) * 2 + comp] = v;
        }
    }
}

// =================================================================================================
// Version 2 (round 2): the same blocking with the rows held in shared memory as FLOAT32 OFFSETS from a per-atom
// reference -- the unwrapped position at the first origin frame of the block, taken in float64 -- and packed
// float32 arithmetic on PAIRS OF ATOMS.
//
// Why: the float64 tile above is bound by shared-memory bandwidth and the half-rate FP64 pipe (three LDS.64 and seven
// FP64 operations per evaluation: 8.0e11 evals/s for the MSD).  An offset from a reference that is at most a few
// hundred frames away from the origin rows is small where the displacement is small and of the size of the
// displacement where it is large, so rounding it to float32 costs 2^-24 of max(|offset|, |d|): every displacement is
// good to ~1e-7 relative (the float64 chain: 1e-16), inside the north star's 1e-5 for the averaged functions and far
// below the float32 sinc.  In exchange a thread (atom pair, lag slice) needs three LDS.64 per TWO evaluations and the
// whole displacement / r^2 / accumulate chain is seven packed instructions per two evaluations; sinc_of_square_x2
// takes its pair as it comes.  `tuning.dense_float64` selects the float64 tile (and per-frame boxes always do).
//
// Thread layout: lane = atom pair (atoms 2*lane, 2*lane+1 of the 64-atom tile), warp = lag slice j: lags 4j .. 4j+3
// of the 32-lag block; every warp owns its lags, so there is no cross-warp reduction.  Rows: [row][3][64] float32,
// 768 bytes each; 32 origins x 32 lags need 95 rows = 71 KiB: three CTAs per SM.
// =================================================================================================
constexpr int kV2Lags = 32;                 // lags per block
constexpr int kV2Slots = 4;                 // consecutive lags per warp
constexpr int kV2MaxOrigins = 32;
constexpr int kV2MaxRows = kV2MaxOrigins + (kV2MaxOrigins - 1) + kV2Lags;     // 95 at origin step 1
constexpr int kV2RowFloats = 3 * kTileAtoms;                                   // 192
static_assert(kTileAtoms == 64 && kDenseThreads == 256, "dense v2 is laid out for 64-atom tiles and 8 warps");

constexpr int kV2RowBatch = 2;              // row pairs in flight per warp while staging

template <bool IMG>
__device__ __forceinline__ double unwrapped(float x, uint32_t n, int k, double box) {
    double value = (double)x;
    if (IMG) value = fma(-image_count(n, k), box, value);      // pbc.py:235-247
    return value;
}

// Staging: a HALF-WARP per row.  Lane t (0..15) of a half takes atoms 4t .. 4t+3 of the tile: 12 consecutive values
// = three 128-bit loads (+ three image words), whose component pattern x y z x y z ... is the same for every lane,
// so the box lengths are uniform operands; every value is unwrapped in float64, the lane's reference (the same 12
// values of the block's first origin row) is subtracted, and the four float32 offsets of each component go out as
// one 128-bit store into the [component][atom] row.  A warp stages rows 2w, 2w+1 (halves), then 2w+16, ...
template <bool IMG, bool FULL>
__device__ __forceinline__ void stage_rows_v2(const DenseDev &p, int lane, int warp, int atom0, int ns, int n_rows,
                                              int f_origin0, int f_partner0, float *s_rows) {
    const int t = lane & 15, half = lane >> 4;
    const long long e0 = 3ll * atom0 + 12 * t, e_end = 3ll * p.n_atoms;
    const float *src0 = p.coords + e0;
    const uint8_t *img0 = IMG ? p.images + e0 : nullptr;
    const double bx = p.box_const[0], by = p.box_const[1], bz = p.box_const[2];
    auto load = [&](long long row, float (&x)[12], uint32_t (&n)[3]) {
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            n[q] = 0x80808080u;
            if (FULL || e0 + 4 * q < e_end) {          // rows are padded: a quad that starts inside may run over
                v = __ldg(reinterpret_cast<const float4 *>(src0 + row * p.pitch) + q);
                if (IMG) n[q] = __ldg(reinterpret_cast<const uint32_t *>(img0 + row * p.img_pitch) + q);
            }
            x[4 * q] = v.x; x[4 * q + 1] = v.y; x[4 * q + 2] = v.z; x[4 * q + 3] = v.w;
        }
    };
    double ref[12];
    {
        float x[12];
        uint32_t n[3];
        load(row_of(p, f_origin0), x, n);
#pragma unroll
        for (int k = 0; k < 12; ++k)
            ref[k] = unwrapped<IMG>(x[k], n[k >> 2], k & 3, k % 3 == 0 ? bx : (k % 3 == 1 ? by : bz));
    }
    for (int rb = 2 * warp + half; rb < n_rows; rb += 16 * kV2RowBatch) {
        float x[kV2RowBatch][12];
        uint32_t n[kV2RowBatch][3];
#pragma unroll
        for (int u = 0; u < kV2RowBatch; ++u) {
            const int r = min(rb + 16 * u, n_rows - 1);
            const int frame = (r < ns) ? f_origin0 + r * p.ostep : f_partner0 + (r - ns);
            load(row_of(p, frame), x[u], n[u]);
        }
#pragma unroll
        for (int u = 0; u < kV2RowBatch; ++u) {
            const int r = min(rb + 16 * u, n_rows - 1);
            float w[12];
#pragma unroll
            for (int k = 0; k < 12; ++k)
                w[k] = (float)(unwrapped<IMG>(x[u][k], n[u][k >> 2], k & 3, k % 3 == 0 ? bx : (k % 3 == 1 ? by : bz)) -
                               ref[k]);
            float4 *dst = reinterpret_cast<float4 *>(s_rows + (size_t)r * kV2RowFloats) + t;      // atoms 4t..4t+3
            dst[0] = make_float4(w[0], w[3], w[6], w[9]);                        // x of the four atoms
            dst[kTileAtoms / 4] = make_float4(w[1], w[4], w[7], w[10]);          // y
            dst[2 * (kTileAtoms / 4)] = make_float4(w[2], w[5], w[8], w[11]);    // z
        }
    }
}

// The same staging WITHOUT float64 and without conversions, for a box whose lengths are float32 numbers (what a
// reader of float32 trajectories delivers).  offset = (x - n L) - (xr - nr L) = x - xr - k L with k = n - nr:
// k = 0 (nearly every value): fl(x - xr) is the exact difference rounded once -- bit-identical to the float64 chain
// followed by its float32 rounding.  k != 0: the operand on the far side of the boundary is moved first (k > 0:
// (x - kL) - xr, k < 0: x - (xr + kL)), one FMA each, exact by Sterbenz when the moved operand lies in [L/2, L] and
// good to ulp(offset) otherwise -- the construction of process_tile_f32d in corr_kernels.cu.  Image counts become
// floats by one byte-permute.  About 6 instructions per value instead of 11, none on the XU or FP64 pipes (which the
// ISF's rsqrt needs): profiles/r2_ncu_dense_v2.txt had the staging at 36 % of all instructions.
template <bool IMG, bool FULL>
__device__ __forceinline__ void stage_rows_v2_f32(const DenseDev &p, int lane, int warp, int atom0, int ns, int n_rows,
                                                  int f_origin0, int f_partner0, float *s_rows) {
    const int t = lane & 15, half = lane >> 4;
    const long long e0 = 3ll * atom0 + 12 * t, e_end = 3ll * p.n_atoms;
    const float *src0 = p.coords + e0;
    const uint8_t *img0 = IMG ? p.images + e0 : nullptr;
    const uint32_t magic = p.magic23;
    auto load = [&](long long row, float (&x)[12], uint32_t (&n)[3]) {
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            n[q] = 0x80808080u;
            if (FULL || e0 + 4 * q < e_end) {
                v = __ldg(reinterpret_cast<const float4 *>(src0 + row * p.pitch) + q);
                if (IMG) n[q] = __ldg(reinterpret_cast<const uint32_t *>(img0 + row * p.img_pitch) + q);
            }
            x[4 * q] = v.x; x[4 * q + 1] = v.y; x[4 * q + 2] = v.z; x[4 * q + 3] = v.w;
        }
    };
    float xr[12];
    uint32_t nr[3];
    load(row_of(p, f_origin0), xr, nr);
    f32x2 fr[6];                                               // the reference's image counts as floats 2^23 + byte
#pragma unroll
    for (int k = 0; k < 12; k += 2)
        fr[k / 2] = pack2(image_float(nr[k >> 2], k & 3, magic), image_float(nr[(k + 1) >> 2], (k + 1) & 3, magic));
    for (int rb = 2 * warp + half; rb < n_rows; rb += 16 * kV2RowBatch) {
        float x[kV2RowBatch][12];
        uint32_t n[kV2RowBatch][3];
#pragma unroll
        for (int u = 0; u < kV2RowBatch; ++u) {
            const int r = min(rb + 16 * u, n_rows - 1);
            const int frame = (r < ns) ? f_origin0 + r * p.ostep : f_partner0 + (r - ns);
            load(row_of(p, frame), x[u], n[u]);
        }
#pragma unroll
        for (int u = 0; u < kV2RowBatch; ++u) {
            const int r = min(rb + 16 * u, n_rows - 1);
            float w[12];
#pragma unroll
            for (int k = 0; k < 12; k += 2) {                  // packed pairs: components (k, k + 1) % 3
                const f32x2 xv = pack2(x[u][k], x[u][k + 1]), rv = pack2(xr[k], xr[k + 1]);
                f32x2 off;
                if (IMG) {
                    const f32x2 nb = pack2(p.neg_box_f[k % 3], p.neg_box_f[(k + 1) % 3]);
                    const f32x2 dn = sub2(pack2(image_float(n[u][k >> 2], k & 3, magic),
                                                image_float(n[u][(k + 1) >> 2], (k + 1) & 3, magic)), fr[k / 2]);
                    float d0, d1;
                    unpack2(dn, d0, d1);
                    const f32x2 up = pack2(fmaxf(d0, 0.0f), fmaxf(d1, 0.0f));   // images this row's value is moved by
                    const f32x2 dw = sub2(up, dn);                              // images the reference is moved by
                    off = sub2(fma2(up, nb, xv), fma2(dw, nb, rv));
                } else {
                    off = sub2(xv, rv);
                }
                unpack2(off, w[k], w[k + 1]);
            }
            float4 *dst = reinterpret_cast<float4 *>(s_rows + (size_t)r * kV2RowFloats) + t;      // atoms 4t..4t+3
            dst[0] = make_float4(w[0], w[3], w[6], w[9]);                        // x of the four atoms
            dst[kTileAtoms / 4] = make_float4(w[1], w[4], w[7], w[10]);          // y
            dst[2 * (kTileAtoms / 4)] = make_float4(w[2], w[5], w[8], w[11]);    // z
        }
    }
}

__device__ __forceinline__ f32x2 lds_pair(const float *ptr) {
    return *reinterpret_cast<const f32x2 *>(ptr);
}

// STEP1 (every frame an origin, BASELINE config 5): origin a and lag slot m meet partner row a + m of the warp's
// slice, so consecutive origins share three of their four partner rows.  The four rows live in registers as a
// rotating window: per origin ONE new partner row and the origin row are read from shared memory -- 6 LDS.64 per 8
// evaluations instead of 15.  That matters because the general loop is bound by shared-memory bandwidth, not by
// arithmetic: 15 x 256 B per origin and warp against 28 packed FP32 instructions (profiles/r2_ncu_dense_v2.txt).
template <bool IMG, int FLAGS, bool STEP1>
__global__ void __launch_bounds__(kDenseThreads, 3) dense_corr_kernel_v2(const DenseDev p) {
    constexpr bool kR2 = (FLAGS & MDB_SUM_R2) != 0;
    constexpr bool kSinc = (FLAGS & MDB_SUM_SINC) != 0;
    extern __shared__ __align__(16) float s_rows_f[];      // [ks + n_partner][3][64]

    const int lb = blockIdx.x % p.n_lblocks;
    const int split = blockIdx.x / p.n_lblocks;
    const int lam = lb * kV2Lags;                          // first lag of the block
    const int nl = min(kV2Lags, p.L - lam);
    const int tid = threadIdx.x;
    const int lane = tid & 31, j = tid >> 5;               // atom pair, lag slice
    const f32x2 q2 = bcast2((float)p.q2);

    double acc2[kV2Slots], accs[kV2Slots];
#pragma unroll
    for (int m = 0; m < kV2Slots; ++m) acc2[m] = accs[m] = 0.0;

    const int t_begin = split * p.tiles_per_split;
    const int t_end = min(p.n_tiles, t_begin + p.tiles_per_split);
    for (int tile = t_begin; tile < t_end; ++tile) {
        const int atom0 = tile * kTileAtoms;
        const bool live_lo = atom0 + 2 * lane < p.n_atoms, live_hi = atom0 + 2 * lane + 1 < p.n_atoms;
        for (int sb = 0; sb < p.n_sblocks; ++sb) {
            const int s0 = sb * p.ks;
            const int ns = min(p.ks, p.S - s0);
            const int f_origin0 = p.o0 + s0 * p.ostep;
            const int f_partner0 = f_origin0 + p.i0 + lam;
            const int n_rows = ns + (ns - 1) * p.ostep + nl;
            __syncthreads();                               // the previous block is done with the rows
            if (p.stage_f32) {
                if (atom0 + kTileAtoms <= p.n_atoms)
                    stage_rows_v2_f32<IMG, true>(p, lane, j, atom0, ns, n_rows, f_origin0, f_partner0, s_rows_f);
                else
                    stage_rows_v2_f32<IMG, false>(p, lane, j, atom0, ns, n_rows, f_origin0, f_partner0, s_rows_f);
            } else if (atom0 + kTileAtoms <= p.n_atoms)
                stage_rows_v2<IMG, true>(p, lane, j, atom0, ns, n_rows, f_origin0, f_partner0, s_rows_f);
            else
                stage_rows_v2<IMG, false>(p, lane, j, atom0, ns, n_rows, f_origin0, f_partner0, s_rows_f);
            __syncthreads();
            // The 4 lag slots of a warp are always evaluated; in the last lag block some lie beyond n_lags: they read
            // rows that are allocated (stale data) and their sums are never written out.
            f32x2 part2[kV2Slots], parts[kV2Slots];
#pragma unroll
            for (int m = 0; m < kV2Slots; ++m) part2[m] = parts[m] = pack2(0.f, 0.f);
            const float *ro = s_rows_f + 2 * lane;
            const float *rp = s_rows_f + (size_t)(ns + kV2Slots * j) * kV2RowFloats + 2 * lane;
            const int partner_stride = p.ostep * kV2RowFloats;
            auto evaluate = [&](const f32x2 &ox, const f32x2 &oy, const f32x2 &oz, const f32x2 &px, const f32x2 &py,
                                const f32x2 &pz, int m) {
                const f32x2 dx = sub2(ox, px), dy = sub2(oy, py), dz = sub2(oz, pz);
                const f32x2 r2 = fma2(dz, dz, fma2(dy, dy, mul2(dx, dx)));
                if constexpr (kR2) part2[m] = add2(part2[m], r2);
                if constexpr (kSinc) {
                    float y0, y1, v0, v1;
                    unpack2(mul2(r2, q2), y0, y1);
                    sinc_of_square_x2(y0, y1, v0, v1);
                    parts[m] = add2(parts[m], pack2(v0, v1));
                }
            };
            if constexpr (STEP1) {
                static_assert(kV2Slots == 4, "the rotating window holds four partner rows");
                f32x2 wx[4], wy[4], wz[4];
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    wx[i] = lds_pair(rp + i * kV2RowFloats);
                    wy[i] = lds_pair(rp + i * kV2RowFloats + kTileAtoms);
                    wz[i] = lds_pair(rp + i * kV2RowFloats + 2 * kTileAtoms);
                }
                rp += 3 * kV2RowFloats;                    // the next row the window takes in
                for (int a = 0; a < ns; a += 4) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        if (a + k < ns) {
                            wx[(k + 3) & 3] = lds_pair(rp);
                            wy[(k + 3) & 3] = lds_pair(rp + kTileAtoms);
                            wz[(k + 3) & 3] = lds_pair(rp + 2 * kTileAtoms);
                            const f32x2 ox = lds_pair(ro), oy = lds_pair(ro + kTileAtoms),
                                        oz = lds_pair(ro + 2 * kTileAtoms);
#pragma unroll
                            for (int m = 0; m < kV2Slots; ++m)
                                evaluate(ox, oy, oz, wx[(k + m) & 3], wy[(k + m) & 3], wz[(k + m) & 3], m);
                            ro += kV2RowFloats;
                            rp += kV2RowFloats;
                        }
                    }
                }
            } else {
#pragma unroll 2
                for (int a = 0; a < ns; ++a) {
                    const f32x2 ox = lds_pair(ro), oy = lds_pair(ro + kTileAtoms), oz = lds_pair(ro + 2 * kTileAtoms);
#pragma unroll
                    for (int m = 0; m < kV2Slots; ++m) {
                        const float *rq = rp + m * kV2RowFloats;
                        evaluate(ox, oy, oz, lds_pair(rq), lds_pair(rq + kTileAtoms), lds_pair(rq + 2 * kTileAtoms), m);
                    }
                    ro += kV2RowFloats;
                    rp += partner_stride;
                }
            }
#pragma unroll
            for (int m = 0; m < kV2Slots; ++m) {
                float lo, hi;
                if constexpr (kR2) {
                    unpack2(part2[m], lo, hi);
                    acc2[m] += (double)((live_lo ? lo : 0.f) + (live_hi ? hi : 0.f));
                }
                if constexpr (kSinc) {
                    unpack2(parts[m], lo, hi);
                    accs[m] += (double)((live_lo ? lo : 0.f) + (live_hi ? hi : 0.f));
                }
            }
        }
    }
#pragma unroll
    for (int m = 0; m < kV2Slots; ++m) {
        const double v2 = warp_sum(acc2[m]);
        const double vs = warp_sum(accs[m]);
        const int l = kV2Slots * j + m;
        if (lane == 0 && l < nl) {
            p.partial[((long long)split * p.L + lam + l) * 2] = v2;
            p.partial[((long long)split * p.L + lam + l) * 2 + 1] = vs;
        }
    }
}

__global__ void dense_reduce_kernel(const double *__restrict__ partial, double *__restrict__ sums, int n_out,
                                    int n_splits) {
    const int o = blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= n_out) return;
    double v = 0.0;
    for (int s = 0; s < n_splits; ++s) v += partial[(long long)s * n_out + o];
    sums[o] = v;
}

struct DensePlan {
    int ks, n_sblocks, n_lblocks, n_tiles, n_splits, tiles_per_split;
    size_t smem;
    bool v2;              // float32 offsets in shared memory (dense_corr_kernel_v2)
};

int make_dense_plan(const mdb_dense_args *a, DensePlan *pl) {
    if (!a) return fail(MDB_ERR_ARGUMENT, "args is NULL");
    if (a->n_atoms <= 0 || a->n_origins <= 0 || a->n_lags <= 0 || a->origin_step <= 0 || a->first_origin < 0 ||
        a->first_lag < 0)
        return fail(MDB_ERR_ARGUMENT, "dense: sizes must be positive, origin_step >= 1, offsets >= 0");
    if (!(a->flags & (MDB_SUM_R2 | MDB_SUM_SINC)) || (a->flags & ~(MDB_SUM_R2 | MDB_SUM_SINC)))
        return fail(MDB_ERR_UNSUPPORTED, "dense: flags must be MDB_SUM_R2 and/or MDB_SUM_SINC");
    if (a->pitch < 3ll * a->n_atoms) return fail(MDB_ERR_ARGUMENT, "pitch < 3*n_atoms");
    if (reinterpret_cast<uintptr_t>(a->coords) % 16 || a->pitch % 4 ||
        (a->images && (reinterpret_cast<uintptr_t>(a->images) % 4 || a->img_pitch % 4)))
        return fail(MDB_ERR_UNSUPPORTED, "dense: rows must be 16-byte aligned (pitch % 4 == 0, image pitch % 4 == 0)");
    // version 2 wherever it applies: a constant box (per-frame boxes keep the float64 tile), unless the caller asks
    pl->v2 = !a->tuning.dense_float64 && !(a->images && a->box);
    int ks;
    if (pl->v2) {
        ks = (kV2MaxRows - kV2Lags + a->origin_step) / (1 + a->origin_step);     // ks + (ks-1)*step + 32 <= 95
        if (ks > kV2MaxOrigins) ks = kV2MaxOrigins;
    } else {
        ks = (kMaxPartnerRows - kL) / a->origin_step + 1;
        if (ks > kMaxS) ks = kMaxS;
    }
    if (ks < 2) return fail(MDB_ERR_UNSUPPORTED, "dense: origin_step too large for row reuse (use mdb_shifted_corr)");
    if (ks > a->n_origins) ks = a->n_origins;
    pl->ks = ks;
    pl->n_sblocks = (a->n_origins + ks - 1) / ks;
    pl->n_lblocks = (a->n_lags + kL - 1) / kL;
    static_assert(kL == kV2Lags, "both versions cut the lags into blocks of 32");
    pl->n_tiles = (a->n_atoms + kTileAtoms - 1) / kTileAtoms;
    int sms = 148, dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    long long want = (8ll * sms + pl->n_lblocks - 1) / pl->n_lblocks;     // ~8 CTAs per SM in total
    if (want > pl->n_tiles) want = pl->n_tiles;
    if (want < 1) want = 1;
    pl->tiles_per_split = (int)((pl->n_tiles + want - 1) / want);
    pl->n_splits = (pl->n_tiles + pl->tiles_per_split - 1) / pl->tiles_per_split;
    const int max_partner = (ks - 1) * a->origin_step + kL;       // also what the padded lag slots may touch
    pl->smem = (size_t)(ks + max_partner) * 3 * kTileAtoms * (pl->v2 ? sizeof(float) : sizeof(double));
    return MDB_OK;
}

}  // namespace
}  // namespace mdb

using namespace mdb;

extern "C" int mdb_dense_workspace_bytes(const mdb_dense_args *a, size_t *bytes) {
    DensePlan pl;
    if (int rc = make_dense_plan(a, &pl)) return rc;
    if (!bytes) return fail(MDB_ERR_ARGUMENT, "bytes is NULL");
    *bytes = (size_t)pl.n_splits * a->n_lags * 2 * sizeof(double);
    return MDB_OK;
}

extern "C" int mdb_dense_corr(const mdb_dense_args *a, void *stream) {
    DensePlan pl;
    if (int rc = make_dense_plan(a, &pl)) return rc;
    if (!a->coords || !a->sums) return fail(MDB_ERR_ARGUMENT, "dense: coords and sums are required");
    if (a->images && !a->box && !(a->box_const[0] > 0.0 && a->box_const[1] > 0.0 && a->box_const[2] > 0.0))
        return fail(MDB_ERR_ARGUMENT, "dense: images need box (per row) or a positive box_const");
    if (a->images && a->img_pitch < 3ll * a->n_atoms) return fail(MDB_ERR_ARGUMENT, "img_pitch < 3*n_atoms");
    const size_t need = (size_t)pl.n_splits * a->n_lags * 2 * sizeof(double);
    if (!a->workspace || a->workspace_bytes < need)
        return fail(MDB_ERR_WORKSPACE, "workspace of %zu bytes required, %zu given", need, a->workspace_bytes);
    cudaStream_t st = static_cast<cudaStream_t>(stream);

    DenseDev p;
    p.coords = a->coords;
    p.pitch = a->pitch;
    p.images = a->images;
    p.img_pitch = a->img_pitch;
    p.box = a->box;
    for (int d = 0; d < 3; ++d) p.box_const[d] = a->box_const[d];
    p.frame_rows = a->frame_rows;
    p.n_atoms = a->n_atoms;
    p.o0 = a->first_origin;
    p.ostep = a->origin_step;
    p.S = a->n_origins;
    p.i0 = a->first_lag;
    p.L = a->n_lags;
    p.q2 = a->q * a->q;
    p.ks = pl.ks;
    p.n_sblocks = pl.n_sblocks;
    p.n_lblocks = pl.n_lblocks;
    p.n_splits = pl.n_splits;
    p.tiles_per_split = pl.tiles_per_split;
    p.n_tiles = pl.n_tiles;
    p.partial = static_cast<double *>(a->workspace);
    p.stage_f32 = (a->tuning.variant & 2) ? 0 : 1;
    for (int d = 0; d < 3; ++d) {
        p.neg_box_f[d] = -(float)a->box_const[d];
        if ((double)(float)a->box_const[d] != a->box_const[d]) p.stage_f32 = 0;
    }
    p.magic23 = 0x4B000000u;

    const unsigned blocks = (unsigned)((long long)pl.n_lblocks * pl.n_splits);
    auto launch = [&](auto kern) -> int {
        if (pl.smem > 40 * 1024)
            MDB_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem));
        kern<<<blocks, kDenseThreads, pl.smem, st>>>(p);
        MDB_CUDA_CHECK(cudaGetLastError());
        return MDB_OK;
    };
    int rc;
    const bool img = a->images != nullptr;
    if (pl.v2) {
        auto pick2 = [&](auto flags_c) -> int {
            constexpr int F = decltype(flags_c)::value;
            if (a->origin_step == 1 && !(a->tuning.variant & 1))
                return img ? launch(dense_corr_kernel_v2<true, F, true>) : launch(dense_corr_kernel_v2<false, F, true>);
            return img ? launch(dense_corr_kernel_v2<true, F, false>) : launch(dense_corr_kernel_v2<false, F, false>);
        };
        switch (a->flags) {
            case MDB_SUM_R2: rc = pick2(std::integral_constant<int, MDB_SUM_R2>{}); break;
            case MDB_SUM_SINC: rc = pick2(std::integral_constant<int, MDB_SUM_SINC>{}); break;
            default: rc = pick2(std::integral_constant<int, MDB_SUM_R2 | MDB_SUM_SINC>{});
        }
        if (rc) return rc;
        const int n_out2 = a->n_lags * 2;
        dense_reduce_kernel<<<(n_out2 + 255) / 256, 256, 0, st>>>(p.partial, a->sums, n_out2, pl.n_splits);
        MDB_CUDA_CHECK(cudaGetLastError());
        last_launch_count() = 2;
        return MDB_OK;
    }
    // warp-per-row staging (measured: MSD 7.0e11 -> 8.0e11, ISF 5.2e11 -> 6.4e11 evals/s, profiles/r2_call1.log) is
    // the default where it applies (constant box); tuning.dense_stage_cta = 1 selects the CTA-wide staging loop
    const bool warp_stage = !a->tuning.dense_stage_cta && kDenseThreads == 256 && !(img && a->box);
    auto pick = [&](auto flags_c) -> int {
        constexpr int F = decltype(flags_c)::value;
        if (warp_stage) return img ? launch(dense_corr_kernel<true, F, true>) : launch(dense_corr_kernel<false, F, true>);
        return img ? launch(dense_corr_kernel<true, F, false>) : launch(dense_corr_kernel<false, F, false>);
    };
    switch (a->flags) {
        case MDB_SUM_R2:
            rc = pick(std::integral_constant<int, MDB_SUM_R2>{});
            break;
        case MDB_SUM_SINC:
            rc = pick(std::integral_constant<int, MDB_SUM_SINC>{});
            break;
        default:
            rc = pick(std::integral_constant<int, MDB_SUM_R2 | MDB_SUM_SINC>{});
    }
    if (rc) return rc;
    const int n_out = a->n_lags * 2;
    dense_reduce_kernel<<<(n_out + 255) / 256, 256, 0, st>>>(p.partial, a->sums, n_out, pl.n_splits);
    MDB_CUDA_CHECK(cudaGetLastError());
    last_launch_count() = 2;
    return MDB_OK;
}
