// Overlap with a reference configuration (reference: correlation.overlap, correlation.py:288-312, on a
// coordinates.CoordinatesKDTree, coordinates.py:415-455):
//
//   overlap(onset, frame) = #{ atoms i of `frame` : min_j |frame_i - tree_j| <= radius } / tree.n
//
// where the tree is scipy's cKDTree of the (selected) atoms of the ORIGIN frame, optionally periodic (`boxsize`).
// scipy is not part of the reference tree; what it computes for k = 1, p = 2 is restated in oracle/md_oracle.py and
// checked there against scipy itself: float64 throughout, squared distance s = ((dx*dx) + dy*dy) + dz*dz without fused
// multiply-adds, periodic trees wrap the query point (x - floor(x / L) * L) and every difference (d < -L/2: d + L,
// d > L/2: d - L), the distance is sqrt(min s).  sqrt is monotone, so "nearest distance <= radius" is "some s <= T"
// with T the largest double whose square root does not exceed the radius -- found on the host, and the kernels never
// take a root.
//
// One launch sequence serves all S origins and S*L (origin, lag) pairs of a schedule:
//   bounds  -> per origin: bounding box of the tree's atoms, their number           (atomics on order-preserving keys)
//   grid    -> per origin: a uniform cell grid, cell edge >= radius, about one atom per cell
//   count / scan (cub) / scatter -> the tree's atoms sorted by cell, float64 x y z   (counting sort, all origins at once)
//   query   -> one thread per atom of the lag frame walks the 3 x 3 z-runs of cells around its own cell and stops at
//              the first hit; block reduction, one atomic per block into counts[s*L + l]
// The sorted trees of a schedule stay in L2 (1M atoms: 32 MB); the query frames stream from HBM once.
#include <cub/device/device_scan.cuh>

#include <cmath>
#include <cstring>

#include "common.cuh"

namespace mdb {
namespace {

constexpr int kOvThreads = 256;

struct FrameRows {
    const void *coords;
    long long pitch;
    int coord_type;          // MDB_F32 / MDB_F64
    const void *images;      // or nullptr (raw frames)
    int image_format;
    long long img_pitch;
    const double *box;       // per storage row (nojump), or nullptr: box_const
    double box_const[3];
};

struct Grid {
    double lo[3];
    double inv[3];           // cells per length
    int n[3];
    int total;
};

struct OvDev {
    FrameRows ref, qry;
    const int *ref_rows;     // [S]
    const uint8_t *ref_select;
    long long ref_sel_pitch;
    int n_ref;
    const int *rows;         // [S*L]
    int n_atoms;
    int S, L;
    double threshold;        // largest s with sqrt(s) <= radius, or < 0: nothing matches
    double reach;            // radius plus a rounding margin: cell edges are at least this long
    int periodic;
    double boxsize[3], half_box[3];
    int max_cells;           // per origin
    int tree_order;          // query the atoms in the order of the origin's sorted tree (see overlap_query_kernel)
    // workspace
    unsigned long long *bounds_key;   // [S][6] order-preserving keys of min xyz / max xyz
    int *tree_n;                      // [S]
    Grid *grids;                      // [S]
    int *cell_count;                  // [S*max_cells + 1]
    int *cell_start;                  // [S*max_cells + 1] exclusive scan of the counts
    int *atom_cell;                   // [S][n_ref] cell of every tree atom, -1 = not selected
    double4 *sorted;                  // [sum of tree sizes]
    int *bad;                         // periodic trees: tree atoms outside [0, boxsize)
    unsigned long long *counts;       // [S*L] out
    double *bounds_out;               // [S][6] out, or nullptr
};

__device__ __forceinline__ int image_count(const FrameRows &f, long long row, int atom, int d) {
    const long long c = 3ll * atom + d;
    if (f.image_format == MDB_IMAGE_BYTES)
        return (int)static_cast<const uint8_t *>(f.images)[row * f.img_pitch + c] - MDB_IMAGE_BIAS;
    if (f.image_format == MDB_IMAGE_I16) return (int)static_cast<const int16_t *>(f.images)[row * f.img_pitch + c];
    return (int)((static_cast<const uint32_t *>(f.images)[row * f.img_pitch + atom / 4] >> (8 * (atom & 3) + 2 * d)) & 3u) - 1;
}

// One atom of one stored row as the float64 values cKDTree sees: raw frames widen exactly, nojump frames are
// x - n*box as pbc.nojump builds them (pbc.py:235-247; one fused multiply-add, as in the correlation kernels).
__device__ __forceinline__ void position(const FrameRows &f, long long row, int atom, double out[3]) {
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        double x;
        if (f.coord_type == MDB_F32) x = (double)static_cast<const float *>(f.coords)[row * f.pitch + 3ll * atom + d];
        else x = static_cast<const double *>(f.coords)[row * f.pitch + 3ll * atom + d];
        if (f.images != nullptr) {
            const double b = f.box ? f.box[3 * row + d] : f.box_const[d];
            x = fma(-(double)image_count(f, row, atom, d), b, x);
        }
        out[d] = x;
    }
}

// doubles <-> unsigned keys with the same order
__device__ __forceinline__ unsigned long long key_of(double v) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(v);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double value_of(unsigned long long k) {
    const unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
    return __longlong_as_double((long long)b);
}

__global__ void __launch_bounds__(kOvThreads) overlap_init_kernel(const OvDev p) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < p.S * 6) p.bounds_key[i] = (i % 6) < 3 ? ~0ull : 0ull;
    if (i < p.S) p.tree_n[i] = 0;
    if (i == 0) *p.bad = 0;
}

__global__ void __launch_bounds__(kOvThreads) overlap_bounds_kernel(const OvDev p) {
    const int s = blockIdx.y;
    const long long row = p.ref_rows[s];
    double lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    int n = 0, bad = 0;
    for (int a = blockIdx.x * blockDim.x + threadIdx.x; a < p.n_ref; a += gridDim.x * blockDim.x) {
        if (p.ref_select && !p.ref_select[(long long)s * p.ref_sel_pitch + a]) continue;
        double x[3];
        position(p.ref, row, a, x);
        ++n;
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            lo[d] = fmin(lo[d], x[d]);
            hi[d] = fmax(hi[d], x[d]);
            // cKDTree: "Some input data are greater than the size of the periodic box" / "Negative input data ..."
            if (p.periodic && p.boxsize[d] > 0.0 && !(x[d] >= 0.0 && x[d] < p.boxsize[d])) bad = 1;
        }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            lo[d] = fmin(lo[d], __shfl_xor_sync(0xffffffffu, lo[d], off));
            hi[d] = fmax(hi[d], __shfl_xor_sync(0xffffffffu, hi[d], off));
        }
        n += __shfl_xor_sync(0xffffffffu, n, off);
        bad |= __shfl_xor_sync(0xffffffffu, bad, off);
    }
    if ((threadIdx.x & 31) == 0 && n > 0) {
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            atomicMin(&p.bounds_key[s * 6 + d], key_of(lo[d]));
            atomicMax(&p.bounds_key[s * 6 + 3 + d], key_of(hi[d]));
        }
        atomicAdd(&p.tree_n[s], n);
        if (bad) atomicExch(p.bad, 1);
    }
}

// The cell grid of one tree: edge >= reach (so a hit is never further than one cell away), about 0.7 mean atom
// spacings when the radius is smaller than that, and never more than max_cells cells.
__global__ void __launch_bounds__(128) overlap_grid_kernel(const OvDev p) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= p.S) return;
    Grid g;
    const int n_tree = p.tree_n[s];
    double ext[3];
    for (int d = 0; d < 3; ++d) {
        const double lo = n_tree > 0 ? value_of(p.bounds_key[s * 6 + d]) : 0.0;
        const double hi = n_tree > 0 ? value_of(p.bounds_key[s * 6 + 3 + d]) : 0.0;
        if (p.bounds_out) {
            p.bounds_out[s * 6 + d] = lo;
            p.bounds_out[s * 6 + 3 + d] = hi;
        }
        const bool per = p.periodic && p.boxsize[d] > 0.0;
        g.lo[d] = per ? 0.0 : lo;
        ext[d] = per ? p.boxsize[d] : hi - lo;
        if (!(ext[d] >= 0.0) || !isfinite(ext[d])) ext[d] = 0.0;
    }
    double volume = 1.0;
    for (int d = 0; d < 3; ++d) volume *= ext[d] > 0.0 ? ext[d] : 1.0;
    double edge = 0.7 * cbrt(volume / (double)(n_tree > 0 ? n_tree : 1));
    if (!(edge >= p.reach)) edge = p.reach;
    if (!(edge > 0.0) || !isfinite(edge)) edge = 1.0;
    for (int iter = 0; iter < 200; ++iter) {
        double total = 1.0;
        for (int d = 0; d < 3; ++d) {
            const bool per = p.periodic && p.boxsize[d] > 0.0;
            double cells = floor(ext[d] / edge);
            if (!per) cells += 1.0;                      // cells of exactly `edge`, the last one holds the maximum
            if (!(cells >= 1.0)) cells = 1.0;
            if (cells > 1.0e9) cells = 1.0e9;
            g.n[d] = (int)cells;
            g.inv[d] = per ? cells / ext[d] : 1.0 / edge;
            total *= cells;
        }
        if (total <= (double)p.max_cells) {
            g.total = g.n[0] * g.n[1] * g.n[2];
            break;
        }
        edge *= 1.26;
        g.total = 0;
    }
    if (g.total <= 0) {                                  // never reached: 1.26^200 exceeds every finite extent
        g.n[0] = g.n[1] = g.n[2] = 1;
        g.total = 1;
    }
    p.grids[s] = g;
}

__device__ __forceinline__ int cell_coord(const Grid &g, int d, double x) {
    const double c = floor((x - g.lo[d]) * g.inv[d]);
    return c < 0.0 ? 0 : (c >= (double)g.n[d] ? g.n[d] - 1 : (int)c);      // a NaN lands in some cell and never matches
}

__global__ void __launch_bounds__(kOvThreads) overlap_count_kernel(const OvDev p) {
    const int s = blockIdx.y;
    const long long row = p.ref_rows[s];
    const Grid g = p.grids[s];
    for (int a = blockIdx.x * blockDim.x + threadIdx.x; a < p.n_ref; a += gridDim.x * blockDim.x) {
        int cell = -1;
        if (!p.ref_select || p.ref_select[(long long)s * p.ref_sel_pitch + a]) {
            double x[3];
            position(p.ref, row, a, x);
            cell = (cell_coord(g, 0, x[0]) * g.n[1] + cell_coord(g, 1, x[1])) * g.n[2] + cell_coord(g, 2, x[2]);
            atomicAdd(&p.cell_count[(long long)s * p.max_cells + cell], 1);
        }
        p.atom_cell[(long long)s * p.n_ref + a] = cell;
    }
}

__global__ void __launch_bounds__(kOvThreads) overlap_scatter_kernel(const OvDev p) {
    const int s = blockIdx.y;
    const long long row = p.ref_rows[s];
    for (int a = blockIdx.x * blockDim.x + threadIdx.x; a < p.n_ref; a += gridDim.x * blockDim.x) {
        const int cell = p.atom_cell[(long long)s * p.n_ref + a];
        if (cell < 0) continue;
        double x[3];
        position(p.ref, row, a, x);
        const long long c = (long long)s * p.max_cells + cell;
        // cell_count was zeroed after the scan and now counts the atoms already placed
        const int at = p.cell_start[c] + atomicAdd(&p.cell_count[c], 1);
        p.sorted[at] = make_double4(x[0], x[1], x[2], (double)a);
    }
}

// wrap of a periodic tree's query point (scipy: BoxDist1D::wrap_position)
__device__ __forceinline__ double wrap_position(double x, double box) {
    if (!(box > 0.0)) return x;
    const double r = floor(__ddiv_rn(x, box));
    double x1 = __dsub_rn(x, __dmul_rn(r, box));
    while (x1 >= box) x1 = __dsub_rn(x1, box);
    while (x1 < 0.0) x1 = __dadd_rn(x1, box);
    return x1;
}

template <bool PERIODIC>
__device__ __forceinline__ bool any_within(const OvDev &p, const double4 *__restrict__ tree, int begin, int end,
                                           const double x[3]) {
    for (int j = begin; j < end; ++j) {
        const double2 yxy = __ldg(reinterpret_cast<const double2 *>(&tree[j]));
        const double yz = __ldg(&tree[j].z);
        double dx = __dsub_rn(x[0], yxy.x), dy = __dsub_rn(x[1], yxy.y), dz = __dsub_rn(x[2], yz);
        if (PERIODIC) {                                   // scipy: BoxDist1D::wrap_distance
            if (p.boxsize[0] > 0.0) dx = dx < -p.half_box[0] ? __dadd_rn(dx, p.boxsize[0]) : (dx > p.half_box[0] ? __dsub_rn(dx, p.boxsize[0]) : dx);
            if (p.boxsize[1] > 0.0) dy = dy < -p.half_box[1] ? __dadd_rn(dy, p.boxsize[1]) : (dy > p.half_box[1] ? __dsub_rn(dy, p.boxsize[1]) : dy);
            if (p.boxsize[2] > 0.0) dz = dz < -p.half_box[2] ? __dadd_rn(dz, p.boxsize[2]) : (dz > p.half_box[2] ? __dsub_rn(dz, p.boxsize[2]) : dz);
        }
        const double s2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
        if (s2 <= p.threshold) return true;
    }
    return false;
}

// the up-to-three neighbour coordinates of cell c along one axis: clamped (open) or wrapped without repeats (periodic)
__device__ __forceinline__ int neighbours(int c, int n, bool periodic, int out[3]) {
    int k = 0;
    if (!periodic) {
        for (int o = -1; o <= 1; ++o)
            if (c + o >= 0 && c + o < n) out[k++] = c + o;
        return k;
    }
    if (n >= 3) {
        out[0] = c == 0 ? n - 1 : c - 1;
        out[1] = c;
        out[2] = c == n - 1 ? 0 : c + 1;
        return 3;
    }
    for (int i = 0; i < n; ++i) out[k++] = i;
    return k;
}

// tree_order: when the tree holds every atom of the correlated trajectory (no selection, same atom count), thread k of
// a pair takes the atom that sits at position k of the origin's SORTED tree instead of atom k of the file.  An atom
// rarely moves further than a cell within the window, so the lanes of a warp -- neighbours in the tree's cell order
// -- walk the same few cells and their candidate loads coalesce; the price is a gather of the query positions
// (12 B at a random place of a row that stays in L2 while its pair is evaluated).
constexpr int kOvChunk = 64;      // warp passes (32 atoms each) of one (origin, lag) pair that a warp takes in a row

template <bool PERIODIC>
__global__ void __launch_bounds__(kOvThreads) overlap_query_kernel(const OvDev p, const long long n_chunks,
                                                                   const int chunks_per_pair) {
    // A WARP takes chunks of 2048 atoms of one pair, round robin over all warps of the grid (pairs at long lags cost
    // more than pairs at short ones: fine-grained turns keep the warps level); hits are counted in a register and
    // go out as one atomic per chunk -- no block-wide barrier anywhere.
    const int lane = threadIdx.x & 31;
    const long long warp_id = (long long)blockIdx.x * (kOvThreads / 32) + (threadIdx.x >> 5);
    const long long n_warps = (long long)gridDim.x * (kOvThreads / 32);
    for (long long c = warp_id; c < n_chunks; c += n_warps) {
        const long long pair = c / chunks_per_pair;
        const int first = (int)(c - pair * chunks_per_pair) * (kOvChunk * 32);
        const int s = (int)(pair / p.L);
        const Grid g = p.grids[s];
        const long long row = p.rows[pair];
        const int *start = p.cell_start + (long long)s * p.max_cells;
        const bool any_tree = p.tree_n[s] > 0 && p.threshold >= 0.0;
        const int seg0 = p.tree_order ? start[0] : 0;
        int hits = 0;
        for (int a0 = first; a0 < min(p.n_atoms, first + kOvChunk * 32); a0 += 32) {
            int a = a0 + lane;
            bool hit = false;
            if (a < p.n_atoms && any_tree) {
                if (p.tree_order) a = (int)__ldg(&p.sorted[seg0 + a].w);
                double x[3];
                position(p.qry, row, a, x);
                bool reachable = true;
#pragma unroll
                for (int d = 0; d < 3; ++d) {
                    const bool per = PERIODIC && p.boxsize[d] > 0.0;
                    if (per) x[d] = wrap_position(x[d], p.boxsize[d]);
                    else {
                        const double top = g.lo[d] + (double)g.n[d] / g.inv[d];     // >= the largest tree coordinate
                        if (!(x[d] >= g.lo[d] - p.reach && x[d] <= top + p.reach)) reachable = false;
                    }
                }
                if (reachable) {
                    int cx[3], cy[3], cz[3];
                    const int nx = neighbours(cell_coord(g, 0, x[0]), g.n[0], PERIODIC && p.boxsize[0] > 0.0, cx);
                    const int ny = neighbours(cell_coord(g, 1, x[1]), g.n[1], PERIODIC && p.boxsize[1] > 0.0, cy);
                    const int nz = neighbours(cell_coord(g, 2, x[2]), g.n[2], PERIODIC && p.boxsize[2] > 0.0, cz);
                    // consecutive z cells are consecutive in the sorted array: one run, two where the wrap splits it
                    int z_lo[2], z_hi[2], n_runs = 1;
                    z_lo[0] = cz[0];
                    z_hi[0] = cz[0];
                    z_lo[1] = z_hi[1] = 0;
                    for (int i = 1; i < nz; ++i) {
                        if (cz[i] == z_hi[n_runs - 1] + 1) z_hi[n_runs - 1] = cz[i];
                        else if (n_runs < 2) { z_lo[1] = z_hi[1] = cz[i]; n_runs = 2; }
                        else z_hi[1] = cz[i];      // not reached: three cells make at most two runs
                    }
                    for (int ix = 0; ix < nx && !hit; ++ix) {
                        // the bounds of all runs of this x layer first (independent loads), then the scans
                        int begin[3][2], end[3][2];
#pragma unroll
                        for (int iy = 0; iy < 3; ++iy)
#pragma unroll
                            for (int r = 0; r < 2; ++r) {
                                begin[iy][r] = end[iy][r] = 0;
                                if (iy < ny && r < n_runs) {
                                    const int base = (cx[ix] * g.n[1] + cy[iy]) * g.n[2];
                                    begin[iy][r] = start[base + z_lo[r]];
                                    end[iy][r] = start[base + z_hi[r] + 1];
                                }
                            }
#pragma unroll
                        for (int iy = 0; iy < 3; ++iy)
#pragma unroll
                            for (int r = 0; r < 2; ++r)
                                if (!hit && begin[iy][r] < end[iy][r])
                                    hit = any_within<PERIODIC>(p, p.sorted, begin[iy][r], end[iy][r], x);
                    }
                }
            }
            hits += __popc(__ballot_sync(0xffffffffu, hit));
        }
        if (lane == 0 && hits) atomicAdd(&p.counts[pair], (unsigned long long)hits);
    }
}

size_t align_up(size_t v) { return (v + 255) & ~(size_t)255; }

struct Layout {
    size_t bounds, tree_n, grids, cell_count, cell_start, atom_cell, sorted, bad, scan_temp, scan_bytes, total;
    int max_cells;
};

int layout_of(const mdb_overlap_args *a, Layout *lay) {
    if (!a) return fail(MDB_ERR_ARGUMENT, "args is NULL");
    if (a->n_origins <= 0 || a->n_lags <= 0 || a->n_ref < 0 || a->n_atoms < 0)
        return fail(MDB_ERR_ARGUMENT, "sizes must be positive");
    const long long S = a->n_origins;
    const long long cells = 4ll * a->n_ref + 64;
    if (cells * S + 1 > 0x7fffffffll || (long long)a->n_ref * S > 0x7fffffffll)
        return fail(MDB_ERR_UNSUPPORTED, "too many tree atoms for one call (%lld origins x %d atoms): split the origins",
                    S, a->n_ref);
    lay->max_cells = (int)cells;
    size_t off = 0;
    lay->bounds = off;      off = align_up(off + (size_t)S * 6 * 8);
    lay->tree_n = off;      off = align_up(off + (size_t)S * 4);
    lay->grids = off;       off = align_up(off + (size_t)S * sizeof(Grid));
    lay->cell_count = off;  off = align_up(off + ((size_t)S * cells + 1) * 4);
    lay->cell_start = off;  off = align_up(off + ((size_t)S * cells + 1) * 4);
    lay->atom_cell = off;   off = align_up(off + (size_t)S * a->n_ref * 4);
    lay->sorted = off;      off = align_up(off + (size_t)S * a->n_ref * sizeof(double4) + 32);
    lay->bad = off;         off = align_up(off + 4);
    size_t temp = 0;
    if (cub::DeviceScan::ExclusiveSum(nullptr, temp, (const int *)nullptr, (int *)nullptr, (int)(S * cells + 1)) != cudaSuccess)
        return fail(MDB_ERR_CUDA, "cub::DeviceScan size query failed");
    lay->scan_temp = off;
    lay->scan_bytes = temp;
    off = align_up(off + temp);
    lay->total = off;
    return MDB_OK;
}

int frame_rows(const void *coords, long long pitch, int type, const void *images, int format, long long img_pitch,
               const double *box, const double *box_const, int n_atoms, FrameRows *out, const char *what) {
    if (!coords) return fail(MDB_ERR_ARGUMENT, "%s coordinates are NULL", what);
    if (type != MDB_F32 && type != MDB_F64) return fail(MDB_ERR_ARGUMENT, "%s coordinate type", what);
    if (pitch < 3ll * n_atoms) return fail(MDB_ERR_ARGUMENT, "%s pitch is shorter than a frame", what);
    if (images && format != MDB_IMAGE_BYTES && format != MDB_IMAGE_PACKED && format != MDB_IMAGE_I16)
        return fail(MDB_ERR_ARGUMENT, "%s image format", what);
    out->coords = coords;
    out->pitch = pitch;
    out->coord_type = type;
    out->images = images;
    out->image_format = format;
    out->img_pitch = img_pitch;
    out->box = box;
    for (int d = 0; d < 3; ++d) out->box_const[d] = box_const[d];
    return MDB_OK;
}

}  // namespace
}  // namespace mdb

using namespace mdb;

extern "C" int mdb_overlap_workspace_bytes(const mdb_overlap_args *a, size_t *bytes) {
    Layout lay;
    if (int rc = layout_of(a, &lay)) return rc;
    if (!bytes) return fail(MDB_ERR_ARGUMENT, "bytes is NULL");
    *bytes = lay.total;
    return MDB_OK;
}

extern "C" int mdb_overlap(const mdb_overlap_args *a, void *stream) {
    Layout lay;
    if (int rc = layout_of(a, &lay)) return rc;
    if (!a->ref_rows || !a->rows || !a->counts) return fail(MDB_ERR_ARGUMENT, "ref_rows, rows and counts are required");
    if (!a->workspace || a->workspace_bytes < lay.total)
        return fail(MDB_ERR_WORKSPACE, "workspace of %zu bytes required, %zu given", lay.total, a->workspace_bytes);
    OvDev p;
    memset(&p, 0, sizeof(p));
    if (int rc = frame_rows(a->ref_coords, a->ref_pitch, a->ref_type, a->ref_images, a->ref_image_format,
                            a->ref_img_pitch, a->ref_box, a->ref_box_const, a->n_ref, &p.ref, "tree"))
        return rc;
    if (int rc = frame_rows(a->coords, a->pitch, a->coord_type, a->images, a->image_format, a->img_pitch, a->box,
                            a->box_const, a->n_atoms, &p.qry, "frame"))
        return rc;
    p.ref_rows = a->ref_rows;
    p.ref_select = a->ref_select;
    p.ref_sel_pitch = a->ref_sel_pitch;
    if (a->ref_select && a->ref_sel_pitch < a->n_ref) return fail(MDB_ERR_ARGUMENT, "ref_sel_pitch is shorter than a row");
    p.n_ref = a->n_ref;
    p.rows = a->rows;
    p.n_atoms = a->n_atoms;
    p.S = a->n_origins;
    p.L = a->n_lags;
    // "sqrt(s) <= radius" as "s <= T": T = the largest double whose (correctly rounded) root does not exceed the radius
    const double radius = a->radius;
    double T = -1.0;
    if (radius >= 0.0) {
        if (std::isinf(radius)) T = INFINITY;
        else {
            T = radius * radius;
            if (std::isinf(T)) T = 1.7976931348623157e308;
            while (std::sqrt(T) > radius) T = std::nextafter(T, 0.0);
            for (;;) {
                const double up = std::nextafter(T, INFINITY);
                if (std::isinf(up) || std::sqrt(up) > radius) break;
                T = up;
            }
        }
    }
    p.threshold = T;
    p.reach = radius >= 0.0 && std::isfinite(radius) ? radius * (1.0 + 1e-6) + 1e-300 : (radius > 0.0 ? 1.0e300 : 1e-300);
    for (int d = 0; d < 3; ++d) {
        const double b = a->boxsize[d];
        if (b < 0.0 || std::isnan(b)) return fail(MDB_ERR_ARGUMENT, "boxsize must not be negative");
        p.boxsize[d] = b;
        p.half_box[d] = 0.5 * b;
        if (b > 0.0) p.periodic = 1;
    }
    p.max_cells = lay.max_cells;
    p.tree_order = (a->n_ref == a->n_atoms && !a->ref_select && !a->file_order) ? 1 : 0;
    char *w = static_cast<char *>(a->workspace);
    p.bounds_key = reinterpret_cast<unsigned long long *>(w + lay.bounds);
    p.tree_n = reinterpret_cast<int *>(w + lay.tree_n);
    p.grids = reinterpret_cast<Grid *>(w + lay.grids);
    p.cell_count = reinterpret_cast<int *>(w + lay.cell_count);
    p.cell_start = reinterpret_cast<int *>(w + lay.cell_start);
    p.atom_cell = reinterpret_cast<int *>(w + lay.atom_cell);
    p.sorted = reinterpret_cast<double4 *>(w + lay.sorted);
    p.bad = reinterpret_cast<int *>(w + lay.bad);
    p.counts = a->counts;
    p.bounds_out = a->bounds;

    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const size_t n_cells_all = (size_t)p.S * p.max_cells + 1;
    int sms = 148, dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    MDB_CUDA_CHECK(cudaMemsetAsync(p.cell_count, 0, n_cells_all * 4, st));
    MDB_CUDA_CHECK(cudaMemsetAsync(p.counts, 0, (size_t)p.S * p.L * 8, st));
    overlap_init_kernel<<<(p.S * 6 + kOvThreads - 1) / kOvThreads, kOvThreads, 0, st>>>(p);
    int ref_blocks = (p.n_ref + kOvThreads - 1) / kOvThreads;
    const int per_origin = (sms * 8 + p.S - 1) / p.S;
    if (ref_blocks > per_origin) ref_blocks = per_origin;
    if (ref_blocks < 1) ref_blocks = 1;
    if (p.S > 65535) return fail(MDB_ERR_UNSUPPORTED, "more than 65535 origins in one call: split the origins");
    const dim3 ref_grid(ref_blocks, p.S);
    overlap_bounds_kernel<<<ref_grid, kOvThreads, 0, st>>>(p);
    overlap_grid_kernel<<<(p.S + 127) / 128, 128, 0, st>>>(p);
    overlap_count_kernel<<<ref_grid, kOvThreads, 0, st>>>(p);
    size_t temp = lay.scan_bytes;
    MDB_CUDA_CHECK(cub::DeviceScan::ExclusiveSum(w + lay.scan_temp, temp, p.cell_count, p.cell_start, (int)n_cells_all, st));
    MDB_CUDA_CHECK(cudaMemsetAsync(p.cell_count, 0, n_cells_all * 4, st));
    overlap_scatter_kernel<<<ref_grid, kOvThreads, 0, st>>>(p);
    const int chunks_per_pair = (p.n_atoms + kOvChunk * 32 - 1) / (kOvChunk * 32);
    const long long n_chunks = (long long)p.S * p.L * chunks_per_pair;
    if (n_chunks > 0) {
        const long long want = (long long)sms * 8;
        const long long blocks = (n_chunks + kOvThreads / 32 - 1) / (kOvThreads / 32);
        const int grid = (int)(blocks < want ? blocks : want);
        if (p.periodic) overlap_query_kernel<true><<<grid, kOvThreads, 0, st>>>(p, n_chunks, chunks_per_pair);
        else overlap_query_kernel<false><<<grid, kOvThreads, 0, st>>>(p, n_chunks, chunks_per_pair);
    }
    MDB_CUDA_CHECK(cudaGetLastError());
    if (a->tree_sizes) MDB_CUDA_CHECK(cudaMemcpyAsync(a->tree_sizes, p.tree_n, (size_t)p.S * 4, cudaMemcpyDeviceToDevice, st));
    if (a->out_of_box) MDB_CUDA_CHECK(cudaMemcpyAsync(a->out_of_box, p.bad, 4, cudaMemcpyDeviceToDevice, st));
    return MDB_OK;
}

extern "C" int mdb_overlap_launches(void) { return 7; }
