// gc_labels.cuh -- region adjacency graph of a label image, built on the device.
//
// Replaces the voxel-pair Python loops of medpy/graphcut/energy_label.py: boundary_stawiaski (:123-214, one
// set_nweight call per border voxel pair), boundary_stawiaski_directed (:217-342), __compute_edges_nd (:411-441,
// used by boundary_difference_of_means :33-120), scipy.ndimage.mean over the regions (:92) and the per-region sums
// of regional_atlas (:345-396).
//
// Everything is a keyed reduction over voxels or voxel pairs.  The reference accumulates with `r_cap += w`
// (graph.h:472-476) in a fixed order -- axis by axis, C order inside an axis -- and float64 addition is not
// associative, so the reduction here keeps that order: contributions are written in reference order (order
// preserving compaction: block counts -> exclusive scan -> in-block scan), stably sorted by key, and every key's run
// is then summed front to back by one thread.  The sums are therefore bit-identical to the reference's and
// independent of the launch geometry.
#pragma once
#include <cfloat>
#include <cstdint>

#include <cuda_runtime.h>

#include "gc_pairwise.cuh"

#define LAB_BLOCK 256

struct LabGeom {
    int nd;                 // canonical axes (1..4)
    long long n;            // voxels
    long long dim[4];
    long long stride[4];    // element strides, C order
};

// strided -> dense copy of an input array (any of the ABI's element types), logical C order
struct LabStrides { long long s[4]; };   // BYTE strides per canonical axis

template <typename E>
__global__ void __launch_bounds__(LAB_BLOCK) k_lab_gather(LabGeom G, const char* __restrict__ src, LabStrides st, E* __restrict__ dst)
{
    const long long p = (long long)blockIdx.x * LAB_BLOCK + threadIdx.x;
    if (p >= G.n) return;
    long long off = 0, r = p;
    for (int d = 0; d < G.nd; ++d) {
        const long long c = r / G.stride[d];
        r -= c * G.stride[d];
        off += c * st.s[d];
    }
    dst[p] = *reinterpret_cast<const E*>(src + off);
}

// ---------------------------------------------------------------------------------------------------------------
// __check_label_image (energy_label.py:444-456): ids must be exactly 1..K
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(LAB_BLOCK) k_lab_minmax(const int* __restrict__ labels, long long n, int* __restrict__ mm)
{
    int lo = INT32_MAX, hi = INT32_MIN;
    for (long long p = (long long)blockIdx.x * LAB_BLOCK + threadIdx.x; p < n; p += (long long)gridDim.x * LAB_BLOCK) {
        const int l = labels[p];
        lo = l < lo ? l : lo;
        hi = l > hi ? l : hi;
    }
    for (int o = 16; o > 0; o >>= 1) {
        const int a = __shfl_down_sync(0xffffffffu, lo, o), b = __shfl_down_sync(0xffffffffu, hi, o);
        lo = a < lo ? a : lo;
        hi = b > hi ? b : hi;
    }
    if ((threadIdx.x & 31) == 0) { atomicMin(&mm[0], lo); atomicMax(&mm[1], hi); }
}

// present[l-1] = 1 for every label that occurs (labels already known to lie in 1..K)
__global__ void __launch_bounds__(LAB_BLOCK) k_lab_presence(const int* __restrict__ labels, long long n, uint8_t* __restrict__ present)
{
    for (long long p = (long long)blockIdx.x * LAB_BLOCK + threadIdx.x; p < n; p += (long long)gridDim.x * LAB_BLOCK)
        present[labels[p] - 1] = 1;
}

__global__ void __launch_bounds__(LAB_BLOCK) k_lab_count_u8(const uint8_t* __restrict__ a, long long n, unsigned long long* __restrict__ count)
{
    unsigned long long c = 0;
    for (long long p = (long long)blockIdx.x * LAB_BLOCK + threadIdx.x; p < n; p += (long long)gridDim.x * LAB_BLOCK) c += a[p] ? 1u : 0u;
    for (int o = 16; o > 0; o >>= 1) c += __shfl_down_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(count, c);
}

// ---------------------------------------------------------------------------------------------------------------
// border voxel pairs, in the reference's order
// ---------------------------------------------------------------------------------------------------------------
// Item idx in [0, nd*n): axis d = idx / n (axis-major, like `for dim in range(ndim)`), voxel p = idx % n (C order
// inside the axis, like the sliced arrays).  The item is a border pair when p has a successor q = p + stride_d along
// d and the two labels differ.  `dup_first`: numpy.vectorize evaluates its function one extra time on the first
// element of its inputs to find the output type (no otypes given, energy_label.py:325-328), so the directed term
// adds the first pair of every axis TWICE when it is a border pair; the copy directly precedes the pair itself.
__device__ __forceinline__ unsigned lab_pair_items(const LabGeom& G, const int* __restrict__ labels, long long idx, int dup_first,
                                                   long long* p_out, long long* q_out)
{
    if (idx >= (long long)G.nd * G.n) return 0u;
    const int d = (int)(idx / G.n);
    const long long p = idx - (long long)d * G.n;
    const long long c = (p / G.stride[d]) % G.dim[d];
    if (c >= G.dim[d] - 1) return 0u;
    const long long q = p + G.stride[d];
    if (labels[p] == labels[q]) return 0u;
    *p_out = p;
    *q_out = q;
    return (dup_first && p == 0) ? 2u : 1u;
}

// exclusive prefix of `c` over the 256 threads of the block; *total = block sum
__device__ __forceinline__ unsigned lab_block_scan(unsigned c, unsigned* total)
{
    __shared__ unsigned wsum[LAB_BLOCK / 32];
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    unsigned incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= (unsigned)o) incl += t;
    }
    if (lane == 31u) wsum[warp] = incl;
    __syncthreads();
    unsigned base = 0, all = 0;
#pragma unroll
    for (unsigned w = 0; w < LAB_BLOCK / 32; ++w) {
        const unsigned s = wsum[w];
        if (w < warp) base += s;
        all += s;
    }
    *total = all;
    return base + incl - c;
}

__global__ void __launch_bounds__(LAB_BLOCK) k_lab_pair_count(LabGeom G, const int* __restrict__ labels, int dup_first,
                                                               unsigned* __restrict__ block_count)
{
    const long long idx = (long long)blockIdx.x * LAB_BLOCK + threadIdx.x;
    long long p, q;
    const unsigned c = lab_pair_items(G, labels, idx, dup_first, &p, &q);
    unsigned total;
    lab_block_scan(c, &total);
    if (threadIdx.x == 0) block_count[blockIdx.x] = total;
}

// exclusive scan of the block counts (one block, each thread owns a contiguous chunk); out[nb] = grand total
__global__ void __launch_bounds__(1024) k_lab_scan_blocks(const unsigned* __restrict__ cnt, long long nb, unsigned long long* __restrict__ out)
{
    __shared__ unsigned long long part[1024];
    const int tid = threadIdx.x;
    const long long chunk = (nb + 1023) / 1024;
    const long long b0 = (long long)tid * chunk;
    const long long b1 = b0 + chunk < nb ? b0 + chunk : nb;
    unsigned long long s = 0;
    for (long long b = b0; b < b1; ++b) s += cnt[b];
    part[tid] = s;
    __syncthreads();
    if (tid == 0) {
        unsigned long long acc = 0;
        for (int t = 0; t < 1024; ++t) { const unsigned long long v = part[t]; part[t] = acc; acc += v; }
        out[nb] = acc;
    }
    __syncthreads();
    unsigned long long acc = part[tid];
    for (long long b = b0; b < b1; ++b) { out[b] = acc; acc += cnt[b]; }
}

// g(max(|a|,|b|)) = (1/(1+x))^2 floored at DBL_MIN (energy_label.py:206-211,297-301), in the two arithmetics the
// reference ends up using (pinned by tests/golden/golden_labels_v1.npz):
//  * "native": numpy scalars of the gradient's dtype under numpy-2 promotion -- `1.0 + val` and `1.0 / (...)` stay
//    float32 for a float32 gradient, every other dtype goes to float64; numpy.abs wraps at the integer minimum.
//    Used by boundary_stawiaski and by the probing call numpy.vectorize makes on element 0.
//  * "pyfloat": the ufunc loop of numpy.vectorize runs over dtype=object copies, i.e. Python floats / ints: float64
//    arithmetic, abs() without wrap-around.  Used by boundary_stawiaski_directed.
// math.pow(r, 2) is taken as the correctly rounded r*r: exact for a float32 r; for a float64 r libm's pow is within
// one ulp of it (tests allow that ulp on float64 / integer gradients and demand equality on float32 ones).
template <typename E> struct LabAbs;
template <> struct LabAbs<float>   { __device__ static float  f(float x)   { return fabsf(x); } };
template <> struct LabAbs<double>  { __device__ static double f(double x)  { return fabs(x); } };
template <> struct LabAbs<uint8_t> { __device__ static uint8_t f(uint8_t x) { return x; } };
template <> struct LabAbs<int16_t> { __device__ static int16_t f(int16_t x) { return (int16_t)(x < 0 ? -x : x); } };
template <> struct LabAbs<int32_t> { __device__ static int32_t f(int32_t x) { return x < 0 ? (int32_t)(0u - (unsigned)x) : x; } };

template <typename E> struct LabIsF32 { static const bool v = false; };
template <> struct LabIsF32<float> { static const bool v = true; };

__device__ __forceinline__ double lab_floor_min(double w) { return (DBL_MIN > w) ? DBL_MIN : w; }   // max(w, float_info.min); NaN stays

template <typename E>
__device__ __forceinline__ double lab_weight_native(E a, E b)
{
    const E va = LabAbs<E>::f(a), vb = LabAbs<E>::f(b);
    const E val = vb > va ? vb : va;
    if (LabIsF32<E>::v) {
        const float s = __fadd_rn(1.0f, (float)val);
        const float r = __fdiv_rn(1.0f, s);
        return lab_floor_min(__dmul_rn((double)r, (double)r));
    }
    const double s = __dadd_rn(1.0, (double)val);
    const double r = __ddiv_rn(1.0, s);
    return lab_floor_min(__dmul_rn(r, r));
}

template <typename E>
__device__ __forceinline__ double lab_weight_pyfloat(E a, E b)
{
    const double va = fabs((double)a), vb = fabs((double)b);
    const double val = vb > va ? vb : va;
    const double r = __ddiv_rn(1.0, __dadd_rn(1.0, val));
    return lab_floor_min(__dmul_rn(r, r));
}

// MODE 0: adjacency only (weights unused), 1: boundary_stawiaski, 2: boundary_stawiaski_directed
// key = (lo << 32) | hi with lo < hi the 0-based node ids; wf accumulates cap(lo -> hi), wr cap(hi -> lo)
template <typename E, int MODE>
__global__ void __launch_bounds__(LAB_BLOCK) k_lab_pair_emit(LabGeom G, const int* __restrict__ labels, const E* __restrict__ grad,
                                                              double beta, int dark_to_light,
                                                              const unsigned long long* __restrict__ block_off,
                                                              unsigned long long* __restrict__ keys, double* __restrict__ wf,
                                                              double* __restrict__ wr)
{
    const long long idx = (long long)blockIdx.x * LAB_BLOCK + threadIdx.x;
    long long p = 0, q = 0;
    const unsigned c = lab_pair_items(G, labels, idx, MODE == 2, &p, &q);
    unsigned total;
    const unsigned ex = lab_block_scan(c, &total);
    if (!c) return;
    const int k1 = labels[p] - 1, k2 = labels[q] - 1;            // set_nweight(key1 - 1, key2 - 1, there, back)
    const bool fwd = k1 < k2;
    const unsigned long long lo = (unsigned long long)(fwd ? k1 : k2), hi = (unsigned long long)(fwd ? k2 : k1);
    const unsigned long long pos = block_off[blockIdx.x] + ex;
    for (unsigned r = 0; r < c; ++r) {
        keys[pos + r] = (lo << 32) | hi;
        if (MODE == 1) {
            wf[pos + r] = lab_weight_native<E>(grad[p], grad[q]);
        } else if (MODE == 2) {
            const E v1 = grad[p], v2 = grad[q];
            const bool probe = (c == 2u && r == 0u);             // the extra call numpy.vectorize makes on element 0
            const double w = probe ? lab_weight_native<E>(v1, v2) : lab_weight_pyfloat<E>(v1, v2);
            const double wb = __dadd_rn(w, beta);
            const double capped = (wb < 1.0) ? wb : 1.0;         // min(1, weight + beta)
            const bool first_gets_beta = dark_to_light ? !(v1 > v2) : (v1 > v2);
            const double there = first_gets_beta ? capped : w;
            const double back = first_gets_beta ? w : capped;
            wf[pos + r] = fwd ? there : back;
            wr[pos + r] = fwd ? back : there;
        }
    }
}

__global__ void __launch_bounds__(LAB_BLOCK) k_lab_iota(unsigned* __restrict__ a, long long m)
{
    const long long i = (long long)blockIdx.x * LAB_BLOCK + threadIdx.x;
    if (i < m) a[i] = (unsigned)i;
}

__global__ void __launch_bounds__(LAB_BLOCK) k_lab_permute(const double* __restrict__ src, const unsigned* __restrict__ perm, long long m,
                                                           double* __restrict__ dst)
{
    const long long i = (long long)blockIdx.x * LAB_BLOCK + threadIdx.x;
    if (i < m) dst[i] = src[perm[i]];
}

// runs of equal keys in the sorted key array: heads per block (-> exclusive scan -> output slot of every run, so the
// region pairs come out ordered by key without any further sort)
__global__ void __launch_bounds__(LAB_BLOCK) k_lab_seg_head_count(const unsigned long long* __restrict__ keys, long long m,
                                                                  unsigned* __restrict__ block_count)
{
    const long long i = (long long)blockIdx.x * LAB_BLOCK + threadIdx.x;
    const unsigned head = (i < m && (i == 0 || keys[i - 1] != keys[i])) ? 1u : 0u;
    unsigned total;
    lab_block_scan(head, &total);
    if (threadIdx.x == 0) block_count[blockIdx.x] = total;
}

// one thread per run of equal keys: front-to-back float64 sum (the order `r_cap += w` sees in the reference)
__global__ void __launch_bounds__(LAB_BLOCK) k_lab_seg_reduce(const unsigned long long* __restrict__ keys, const double* __restrict__ wf,
                                                              const double* __restrict__ wr, long long m,
                                                              const unsigned long long* __restrict__ block_off,
                                                              unsigned long long* __restrict__ out_key, double* __restrict__ out_f,
                                                              double* __restrict__ out_r)
{
    const long long i = (long long)blockIdx.x * LAB_BLOCK + threadIdx.x;
    const unsigned head = (i < m && (i == 0 || keys[i - 1] != keys[i])) ? 1u : 0u;
    unsigned total;
    const unsigned ex = lab_block_scan(head, &total);
    if (!head) return;
    const unsigned long long key = keys[i];
    double a = 0.0, b = 0.0;
    for (long long j = i; j < m && keys[j] == key; ++j) {
        if (wf) a = __dadd_rn(a, wf[j]);
        if (wr) b = __dadd_rn(b, wr[j]);
    }
    const unsigned long long pos = block_off[blockIdx.x] + ex;
    out_key[pos] = key;
    out_f[pos] = a;
    out_r[pos] = wr ? b : a;
}

// ---------------------------------------------------------------------------------------------------------------
// per-region sums (numpy.bincount(labels.ravel(), weights=image.ravel()) inside scipy.ndimage.mean; regional_atlas)
// ---------------------------------------------------------------------------------------------------------------
template <typename E, typename V>
__global__ void __launch_bounds__(LAB_BLOCK) k_lab_region_items(const int* __restrict__ labels, const E* __restrict__ values, long long n,
                                                                unsigned* __restrict__ keys, V* __restrict__ vals)
{
    const long long p = (long long)blockIdx.x * LAB_BLOCK + threadIdx.x;
    if (p >= n) return;
    keys[p] = (unsigned)(labels[p] - 1);
    vals[p] = (V)values[p];
}

// numpy.bincount(labels, weights): float64, front to back
__global__ void __launch_bounds__(LAB_BLOCK) k_lab_region_reduce(const unsigned* __restrict__ keys, const double* __restrict__ vals, long long n,
                                                                 double* __restrict__ sums, long long* __restrict__ counts)
{
    const long long i = (long long)blockIdx.x * LAB_BLOCK + threadIdx.x;
    if (i >= n) return;
    const unsigned key = keys[i];
    if (i > 0 && keys[i - 1] == key) return;
    double a = 0.0;
    long long j = i;
    for (; j < n && keys[j] == key; ++j) a = __dadd_rn(a, vals[j]);
    sums[key] = a;
    counts[key] = j - i;
}

template <typename V>
__global__ void __launch_bounds__(LAB_BLOCK) k_lab_region_reduce_pairwise(const unsigned* __restrict__ keys, const V* __restrict__ vals, long long n,
                                                                          double* __restrict__ sums, long long* __restrict__ counts)
{
    const long long i = (long long)blockIdx.x * LAB_BLOCK + threadIdx.x;
    if (i >= n) return;
    const unsigned key = keys[i];
    if (i > 0 && keys[i - 1] == key) return;
    long long j = i;
    while (j < n && keys[j] == key) ++j;
    sums[key] = (double)lab_pairwise_sum<V>(vals + i, j - i);
    counts[key] = j - i;
}

// flags[l-1] = 1 for every region with a marked voxel (numpy.unique(label_image[markers] - 1), generate.py:334-337)
__global__ void __launch_bounds__(LAB_BLOCK) k_lab_region_flags(const int* __restrict__ labels, const uint8_t* __restrict__ markers, long long n,
                                                                uint8_t* __restrict__ flags)
{
    for (long long p = (long long)blockIdx.x * LAB_BLOCK + threadIdx.x; p < n; p += (long long)gridDim.x * LAB_BLOCK)
        if (markers[p]) flags[labels[p] - 1] = 1;
}

// out[p] = per_region[label[p] - 1]: the relabel_map step of bin/medpy_graphcut_label.py:139-148
__global__ void __launch_bounds__(LAB_BLOCK) k_lab_apply(const int* __restrict__ labels, const uint8_t* __restrict__ per_region, long long n,
                                                         uint8_t* __restrict__ out)
{
    for (long long p = (long long)blockIdx.x * LAB_BLOCK + threadIdx.x; p < n; p += (long long)gridDim.x * LAB_BLOCK)
        out[p] = per_region[labels[p] - 1];
}
