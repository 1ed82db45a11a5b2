// Vectors between atoms (reference: coordinates.vectors, coordinates.py:524-563) for the frames a schedule touches:
// the input rows of rotational_autocorrelation (correlation.py:223-236) built on the device, straight from the
// resident frame rows (+ image counts for nojump), so that the frames never travel to the host.
//
//   out[r][v] = pbc_diff( mean_i x[atoms_a[i][v]], mean_j x[atoms_b[j][v]], box ) [ / its norm ]
//
// in the arithmetic of the reference: float64 for nojump frames (x - n*box, pbc.py:235-247), float32 for raw frames,
// operation by operation (unfused): numpy's mean of up to 8 values is a sequential sum divided by the count,
// pbc_diff_rect is v = a - b; s = v / box; v = box * (s - rint(s)) (pbc.py:44-56), the norm is
// sqrt(x^2 + y^2 + z^2) summed left to right.
#include "common.cuh"

namespace mdb {
namespace {

struct VecDev {
    const float *coords;
    long long pitch;
    const void *images;
    int image_format;        // 0 = raw
    long long img_pitch;
    const double *box;       // per storage row, or nullptr: box_const
    double box_const[3];
    const int *rows;
    int n_rows;
    const int *atoms_a;
    int m_a;
    const int *atoms_b;
    int m_b;
    int n_vec;
    int normed;
    int use_pbc;
    double pbc_box[3];
    void *out;
    long long out_pitch;
};

__device__ __forceinline__ int count_of(const VecDev &p, long long row, int atom, int d) {
    const long long c = 3ll * atom + d;
    if (p.image_format == MDB_IMAGE_BYTES)
        return (int)static_cast<const uint8_t *>(p.images)[row * p.img_pitch + c] - MDB_IMAGE_BIAS;
    if (p.image_format == MDB_IMAGE_I16) return (int)static_cast<const int16_t *>(p.images)[row * p.img_pitch + c];
    if (p.image_format == MDB_IMAGE_PACKED)
        return (int)((static_cast<const uint32_t *>(p.images)[row * p.img_pitch + atom / 4] >> (8 * (atom & 3) + 2 * d)) & 3u) - 1;
    return 0;
}

__device__ __forceinline__ float sub_rn(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ double sub_rn(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ float div_rn(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ double div_rn(double a, double b) { return __ddiv_rn(a, b); }
__device__ __forceinline__ float sqrt_rn(float a) { return __fsqrt_rn(a); }
__device__ __forceinline__ double sqrt_rn(double a) { return __dsqrt_rn(a); }
__device__ __forceinline__ float rint_t(float a) { return rintf(a); }
__device__ __forceinline__ double rint_t(double a) { return rint(a); }

// position of one atom in the frame's own arithmetic: raw float32, or the float64 nojump frame x - n*box
template <typename T>
__device__ __forceinline__ T position(const VecDev &p, long long row, int atom, int d, const double *bx) {
    const float x = p.coords[row * p.pitch + 3ll * atom + d];
    if (sizeof(T) == 4) return (T)x;
    return (T)fma(-(double)count_of(p, row, atom, d), bx[d], (double)x);
}

template <typename T>
__device__ __forceinline__ T mean_position(const VecDev &p, long long row, const int *atoms, int m, int v, int d,
                                           const double *bx) {
    T acc = position<T>(p, row, atoms[v], d, bx);
    for (int i = 1; i < m; ++i) acc = add_rn(acc, position<T>(p, row, atoms[(long long)i * p.n_vec + v], d, bx));
    return m > 1 ? div_rn(acc, (T)m) : acc;
}

template <typename T>
__global__ void __launch_bounds__(256) vectors_kernel(const VecDev p) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y;
    if (v >= p.n_vec || r >= p.n_rows) return;
    const long long row = p.rows ? p.rows[r] : r;
    double bx[3];
    for (int d = 0; d < 3; ++d) bx[d] = p.box ? p.box[3 * row + d] : p.box_const[d];
    T vec[3];
    for (int d = 0; d < 3; ++d) {
        const T a = mean_position<T>(p, row, p.atoms_a, p.m_a, v, d, bx);
        const T b = mean_position<T>(p, row, p.atoms_b, p.m_b, v, d, bx);
        T diff = sub_rn(a, b);
        if (p.use_pbc) {
            const T box = (T)p.pbc_box[d];
            const T s = div_rn(diff, box);
            diff = mul_rn(box, sub_rn(s, rint_t(s)));
        }
        vec[d] = diff;
    }
    if (p.normed) {
        const T norm = sqrt_rn(add_rn(add_rn(mul_rn(vec[0], vec[0]), mul_rn(vec[1], vec[1])), mul_rn(vec[2], vec[2])));
        for (int d = 0; d < 3; ++d) vec[d] = div_rn(vec[d], norm);
    }
    T *out = static_cast<T *>(p.out) + (long long)r * p.out_pitch + 3ll * v;
    out[0] = vec[0];
    out[1] = vec[1];
    out[2] = vec[2];
}

}  // namespace
}  // namespace mdb

using namespace mdb;

extern "C" int mdb_vectors(const mdb_vectors_args *a, void *stream) {
    if (!a) return fail(MDB_ERR_ARGUMENT, "args is NULL");
    if (!a->coords || !a->atoms_a || !a->atoms_b || !a->out) return fail(MDB_ERR_ARGUMENT, "mdb_vectors: NULL pointer");
    if (a->n_rows < 0 || a->n_vec <= 0 || a->m_a < 1 || a->m_b < 1)
        return fail(MDB_ERR_ARGUMENT, "mdb_vectors: need n_vec > 0 and at least one atom per end");
    if (a->out_type != MDB_F32 && a->out_type != MDB_F64) return fail(MDB_ERR_ARGUMENT, "mdb_vectors: bad out_type");
    if (a->out_pitch < 3ll * a->n_vec) return fail(MDB_ERR_ARGUMENT, "mdb_vectors: out_pitch < 3*n_vec");
    if (a->images && a->out_type != MDB_F64)
        return fail(MDB_ERR_ARGUMENT, "mdb_vectors: nojump frames are float64 (pbc.py:235-247): out_type must be MDB_F64");
    if (a->images && !a->box && !(a->box_const[0] > 0 && a->box_const[1] > 0 && a->box_const[2] > 0))
        return fail(MDB_ERR_ARGUMENT, "mdb_vectors: images need box (per row) or a positive box_const");
    if (a->n_rows == 0) return MDB_OK;
    VecDev p;
    p.coords = a->coords;
    p.pitch = a->pitch;
    p.images = a->images;
    p.image_format = a->images ? a->image_format : 0;
    p.img_pitch = a->img_pitch;
    p.box = a->box;
    for (int d = 0; d < 3; ++d) {
        p.box_const[d] = a->box_const[d];
        p.pbc_box[d] = a->pbc_box[d];
    }
    p.rows = a->rows;
    p.n_rows = a->n_rows;
    p.atoms_a = a->atoms_a;
    p.m_a = a->m_a;
    p.atoms_b = a->atoms_b;
    p.m_b = a->m_b;
    p.n_vec = a->n_vec;
    p.normed = a->normed;
    p.use_pbc = a->use_pbc;
    p.out = a->out;
    p.out_pitch = a->out_pitch;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    for (int r0 = 0; r0 < a->n_rows; r0 += 65535) {
        VecDev q = p;
        q.rows = a->rows ? a->rows + r0 : nullptr;
        q.n_rows = a->n_rows - r0 < 65535 ? a->n_rows - r0 : 65535;
        const size_t esz = a->out_type == MDB_F32 ? 4 : 8;
        q.out = static_cast<char *>(a->out) + (size_t)r0 * a->out_pitch * esz;
        if (!a->rows && r0) return fail(MDB_ERR_UNSUPPORTED, "mdb_vectors: more than 65535 rows need a row table");
        const dim3 grid((unsigned)((a->n_vec + 255) / 256), (unsigned)q.n_rows);
        if (a->out_type == MDB_F32) vectors_kernel<float><<<grid, 256, 0, st>>>(q);
        else vectors_kernel<double><<<grid, 256, 0, st>>>(q);
        MDB_CUDA_CHECK(cudaGetLastError());
    }
    return MDB_OK;
}
