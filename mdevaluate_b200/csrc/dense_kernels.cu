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
// handles the lags l = j, j+4, j+8, j+12 of the block.  Partials per atom split are added by a small
// second kernel in fixed order (bit-reproducible).
#include "common.cuh"

namespace mdb {
namespace {

constexpr int kTileAtoms = 64;
constexpr int kSlices = 4;
constexpr int kDenseThreads = kTileAtoms * kSlices;   // 256
constexpr int kL = 16;                                // lags per block (4 per thread)
constexpr int kMaxS = 16;                             // origins per block (fewer for origin_step > 1)
constexpr int kMaxPartnerRows = 48;

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
};

__device__ __forceinline__ int row_of(const DenseDev &p, int frame) {
    return p.frame_rows ? p.frame_rows[frame] : frame;
}

// Load one frame row of the tile (64 atoms = 192 interleaved values), unwrap, store as SoA doubles.
template <bool IMG>
__device__ __forceinline__ void stage_row(const DenseDev &p, int frame, int atom0, double *dst /*[3][64]*/, int v) {
    // v in [0, 192): interleaved component index inside the tile
    const int atom = v / 3, c = v - 3 * atom;
    double value = 0.0;
    if (atom0 + atom < p.n_atoms) {
        const int row = row_of(p, frame);
        const long long e = 3ll * atom0 + v;
        const float x = __ldg(p.coords + (long long)row * p.pitch + e);
        value = (double)x;
        if (IMG) {
            const int n = (int)__ldg(p.images + (long long)row * p.img_pitch + e) - MDB_IMAGE_BIAS;
            const double b = p.box ? __ldg(p.box + 3ll * row + c) : p.box_const[c];
            value = fma(-(double)n, b, value);     // pbc.py:235-247, same expression as the streaming kernel
        }
    }
    dst[c * kTileAtoms + atom] = value;
}

template <bool IMG, int FLAGS>
__global__ void __launch_bounds__(kDenseThreads, 2) dense_corr_kernel(const DenseDev p) {
    constexpr bool kR2 = (FLAGS & MDB_SUM_R2) != 0;
    constexpr bool kSinc = (FLAGS & MDB_SUM_SINC) != 0;
    extern __shared__ __align__(16) double s_rows[];     // [ks + n_partner][3][64]
    __shared__ double s_red[kDenseThreads / 32][4][2];

    const int lb = blockIdx.x % p.n_lblocks;
    const int split = blockIdx.x / p.n_lblocks;
    const int lam = lb * kL;                              // first lag of the block
    const int nl = min(kL, p.L - lam);
    const int tid = threadIdx.x;
    const int i = tid % kTileAtoms, j = tid / kTileAtoms;
    const int row_doubles = 3 * kTileAtoms;

    double acc2[4] = {0.0, 0.0, 0.0, 0.0}, accs[4] = {0.0, 0.0, 0.0, 0.0};

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
            for (int w = tid; w < n_rows * 192; w += kDenseThreads) {
                const int r = w / 192, v = w - 192 * r;
                const int frame = (r < ns) ? f_origin0 + r * p.ostep : f_partner0 + (r - ns);
                stage_row<IMG>(p, frame, atom0, s_rows + (size_t)r * row_doubles, v);
            }
            __syncthreads();
            if (live) {
                for (int a = 0; a < ns; ++a) {
                    const double *ro = s_rows + (size_t)a * row_doubles + i;
                    const double ox = ro[0], oy = ro[kTileAtoms], oz = ro[2 * kTileAtoms];
                    float yq[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                    for (int m = 0; m < 4; ++m) {
                        const int l = j + kSlices * m;
                        if (l < nl) {
                            const double *rp = s_rows + (size_t)(ns + a * p.ostep + l) * row_doubles + i;
                            const double dx = ox - rp[0], dy = oy - rp[kTileAtoms], dz = oz - rp[2 * kTileAtoms];
                            const double r2 = fma(dz, dz, fma(dy, dy, dx * dx));
                            if constexpr (kR2) acc2[m] += r2;
                            if constexpr (kSinc) yq[m] = (float)(r2 * p.q2);
                        }
                    }
                    if constexpr (kSinc) {
                        float v0, v1, v2, v3;
                        sinc_of_square_x2(yq[0], yq[1], v0, v1);
                        sinc_of_square_x2(yq[2], yq[3], v2, v3);
                        if (j < nl) accs[0] += (double)v0;
                        if (j + kSlices < nl) accs[1] += (double)v1;
                        if (j + 2 * kSlices < nl) accs[2] += (double)v2;
                        if (j + 3 * kSlices < nl) accs[3] += (double)v3;
                    }
                }
            }
        }
    }

    // ---- reduce over the 64 atoms of every (slice j, slot m): two warps per slice ----
    const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
    for (int m = 0; m < 4; ++m) {
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
            const double v = s_red[2 * jj][m][comp] + s_red[2 * jj + 1][m][comp];
            p.partial[((long long)split * p.L + lam + l) * 2 + comp] = v;
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
};

int make_dense_plan(const mdb_dense_args *a, DensePlan *pl) {
    if (!a) return fail(MDB_ERR_ARGUMENT, "args is NULL");
    if (a->n_atoms <= 0 || a->n_origins <= 0 || a->n_lags <= 0 || a->origin_step <= 0 || a->first_origin < 0 ||
        a->first_lag < 0)
        return fail(MDB_ERR_ARGUMENT, "dense: sizes must be positive, origin_step >= 1, offsets >= 0");
    if (!(a->flags & (MDB_SUM_R2 | MDB_SUM_SINC)) || (a->flags & ~(MDB_SUM_R2 | MDB_SUM_SINC)))
        return fail(MDB_ERR_UNSUPPORTED, "dense: flags must be MDB_SUM_R2 and/or MDB_SUM_SINC");
    if (a->pitch < 3ll * a->n_atoms) return fail(MDB_ERR_ARGUMENT, "pitch < 3*n_atoms");
    int ks = (kMaxPartnerRows - kL) / a->origin_step + 1;
    if (ks > kMaxS) ks = kMaxS;
    if (ks < 2) return fail(MDB_ERR_UNSUPPORTED, "dense: origin_step too large for row reuse (use mdb_shifted_corr)");
    if (ks > a->n_origins) ks = a->n_origins;
    pl->ks = ks;
    pl->n_sblocks = (a->n_origins + ks - 1) / ks;
    pl->n_lblocks = (a->n_lags + kL - 1) / kL;
    pl->n_tiles = (a->n_atoms + kTileAtoms - 1) / kTileAtoms;
    int sms = 148, dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    long long want = (8ll * sms + pl->n_lblocks - 1) / pl->n_lblocks;     // ~8 CTAs per SM in total
    if (want > pl->n_tiles) want = pl->n_tiles;
    if (want < 1) want = 1;
    pl->tiles_per_split = (int)((pl->n_tiles + want - 1) / want);
    pl->n_splits = (pl->n_tiles + pl->tiles_per_split - 1) / pl->tiles_per_split;
    const int max_partner = (ks - 1) * a->origin_step + kL;
    pl->smem = (size_t)(ks + max_partner) * 3 * kTileAtoms * sizeof(double);
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
    switch (a->flags) {
        case MDB_SUM_R2:
            rc = img ? launch(dense_corr_kernel<true, MDB_SUM_R2>) : launch(dense_corr_kernel<false, MDB_SUM_R2>);
            break;
        case MDB_SUM_SINC:
            rc = img ? launch(dense_corr_kernel<true, MDB_SUM_SINC>) : launch(dense_corr_kernel<false, MDB_SUM_SINC>);
            break;
        default:
            rc = img ? launch(dense_corr_kernel<true, MDB_SUM_R2 | MDB_SUM_SINC>)
                     : launch(dense_corr_kernel<false, MDB_SUM_R2 | MDB_SUM_SINC>);
    }
    if (rc) return rc;
    const int n_out = a->n_lags * 2;
    dense_reduce_kernel<<<(n_out + 255) / 256, 256, 0, st>>>(p.partial, a->sums, n_out, pl.n_splits);
    MDB_CUDA_CHECK(cudaGetLastError());
    last_launch_count() = 2;
    return MDB_OK;
}
