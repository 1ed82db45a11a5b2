// Declarations shared by the translation units of the correlation path:
//   corr_kernels.cu  -- the kernels; compiled four times (-DMDB_CORR_PART=0..3: direct-load kernels, bulk kernels
//                       without image counts, with byte counts, with packed counts) so that the instantiations
//                       build in parallel;
//   corr_host.cu     -- planning, the C ABI (mdb_shifted_corr, mdb_corr_plan_*), the finish kernel;
//   comm.cu          -- the peer-memory communicator the finish kernel uses.
#pragma once

#include "common.cuh"

namespace mdb {

constexpr int kThreads = 256;           // compute threads per CTA (both kernels)
constexpr int kLagChunk = 8;
constexpr int kWarps = kThreads / 32;
constexpr int kTileAtoms = 4 * kThreads;          // atoms per bulk tile
constexpr int kCoordTileBytes = kTileAtoms * 12;  // 12288
constexpr int kBulkThreads = kThreads + 32;       // + one producer warp
constexpr int kMaxStages = 8;
constexpr int kHistTailBytes = 256;               // scratch line of the histogram kernels (trash_slot) + alignment
#ifndef MDB_BULK_CTAS_PER_SM
#define MDB_BULK_CTAS_PER_SM 2
#endif

__host__ __device__ constexpr int img_tile_bytes(int img) {
    return img == IMG_BYTES ? kTileAtoms * 3 : (img == IMG_PACKED ? kTileAtoms : 0);
}

struct CorrDev {
    const void *coords;
    long long pitch;
    const void *images;      // rows in the format the kernel was instantiated for (bytes / packed words / int16)
    long long img_pitch;     // elements of that format per row
    const double *box;
    double box_const[3];     // used when box == nullptr: the same box for every row (kernel-parameter constants);
                             // divided by 4^d for packed counts (see packed_dim)
    int n_atoms;
    int n_groups;            // ceil(n_atoms / 4)
    int groups_per_block;
    int S, L;
    const int *rows;
    const uint8_t *select;   // [S][sel_pitch] 1 = atom takes part for origin s (sub-ensembles), or nullptr
    long long sel_pitch;
    int n_chunks, n_ablocks;
    int n_segments;          // > 1: the select bytes name one of n_segments sub-ensembles per atom (ISF, one pass)
    double q2;
    int f32d;                // float32 displacements (ISF alone; van Hove [+ alpha_2] with uniform bins): uniform
    float q2f;               // float(q^2) and the negated (scaled) box as the packed pairs (x,y) (z,x) (y,z):
    alignas(8) float neg_box2[6];   // kernel parameters = constant bank, read as 64-bit operands
    int leg_order;
    int n_bins;
    int bin_mode;
    double dx, inv_dx;
    float inv_dx_f;
    float hist_c1, hist_lim0;   // fast bin estimate: |frac - 0.5| < hist_lim0 + hist_c1 * (t - 0.5) decides a sample
    uint32_t ib_limit;          // 0x4B400000 + n_bins
    uint32_t magic23;           // 0x4B000000 (bits of 2^23) and the value 1 as REGISTER operands: see image_float /
    uint32_t one;               // red_shared_add_if in common.cuh
    const double *edges;
    double *partial;         // [S*L][n_ablocks][3]
    unsigned long long *hist;
    int no_shortcut;             // tuning.variant bit 2: always the general displacement formula (measurement)
    unsigned long long *stamp;   // plans: the first CTA stores %globaltimer here when it starts (timeline), or nullptr
};

__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

struct Plan {
    int n_groups, n_chunks, n_ablocks, groups_per_block;
    int img;             // IMG_* format the launch reads
    bool bulk;           // cp.async.bulk path
    bool vec;            // 128-bit loads in the fallback path
    int stages;
    size_t smem;
};

struct LaunchTiming {
    cudaEvent_t start = nullptr, stop = nullptr;     // both or none
};

// one per translation unit of corr_kernels.cu
int corr_launch_ldg(const CorrDev &p, const Plan &plan, int flags, int coord_type, cudaStream_t st, LaunchTiming tm);
int corr_launch_bulk_none(const CorrDev &p, const Plan &plan, int flags, cudaStream_t st, LaunchTiming tm);
int corr_launch_bulk_bytes(const CorrDev &p, const Plan &plan, int flags, cudaStream_t st, LaunchTiming tm);
int corr_launch_bulk_packed(const CorrDev &p, const Plan &plan, int flags, cudaStream_t st, LaunchTiming tm);

// The kernel instantiation that serves a request: the four single-purpose kernels, van Hove +
// non-Gaussian in one pass (BASELINE config 3), or the all-sums kernel (optionally with the histogram).
inline int kernel_flags(int flags) {
    constexpr int R2 = MDB_SUM_R2, R4 = MDB_SUM_R4, SC = MDB_SUM_SINC, H = MDB_HIST;
    if (flags == R2 || flags == SC || flags == (R2 | R4) || flags == H || flags == (R2 | R4 | H) ||
        flags == MDB_SUM_LEGENDRE)
        return flags;
    return (flags & H) ? (R2 | R4 | SC | H) : (R2 | R4 | SC);
}

// ---- communicator (comm.cu) as the finish kernel sees it ------------------------------------------------
struct CommDev {
    int rank, world;
    unsigned long long capacity;                 // 8-byte words per slot
    unsigned long long *inbox[MDB_COMM_MAX_RANKS];   // base of every rank's inbox: [2][world][capacity] words
    unsigned long long *flags[MDB_COMM_MAX_RANKS];   // every rank's flags: [2][world] sequence numbers
    long long timeout_cycles;
};
int comm_device_view(mdb_comm *comm, CommDev *out);     // MDB_OK when connected

}  // namespace mdb
