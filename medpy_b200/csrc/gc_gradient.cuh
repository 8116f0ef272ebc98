// gc_gradient.cuh -- row f1 of SURVEY.md §8: the pre-step in front of every boundary_maximum_* run,
// bin/medpy_gradient.py:79-85 = scipy.ndimage.generic_gradient_magnitude(image, prewitt, output=float32).
//
// The arithmetic that script performs lives in SciPy (a third-party dependency of the reference; the version in this
// image is what the tests compare against): for every axis a
//     D_a = correlate1d(image, [-1, 0, 1], a)            -> stored as float32
//     D_a = correlate1d(D_a, [1, 1, 1], b)  for b != a, in increasing b, each pass stored as float32
// all with mode='reflect' (index -1 -> 0, n -> n-1), line arithmetic in double as NI_Correlate1D does
// (centre * w0 first, then (left + right) * w1, resp. (right - left) * w1), and finally
//     out = sqrt(D_0^2 + D_1^2 + ...)  with float32 multiplies, float32 adds in axis order and a float32 sqrt.
// One thread per voxel evaluates that expression tree directly (3^(ndim-1) derivative taps per axis, served by L1/L2);
// it reproduces SciPy bit for bit.
#pragma once
#include "gc_common.cuh"
#include "gc_terms.cuh"

template <int ND>
struct GradCtx {
    int dim[ND];
    long long stride[ND];
};

__device__ __forceinline__ int reflect_idx(int i, int n) { return i < 0 ? 0 : (i >= n ? n - 1 : i); }

// value after `LEVEL` smoothing passes of the derivative along axis A, at integer position c (already inside the lattice)
template <typename E, int ND, int A, int LEVEL>
struct GradEval {
    static __device__ __forceinline__ float at(const GradCtx<ND>& G, const E* __restrict__ img, int (&c)[ND])
    {
        // LEVEL-th smoothing axis = the LEVEL-th axis != A in increasing order (1-based)
        constexpr int B = (LEVEL - 1 < A) ? (LEVEL - 1) : LEVEL;
        const int keep = c[B];
        const float x0 = GradEval<E, ND, A, LEVEL - 1>::at(G, img, c);
        c[B] = reflect_idx(keep - 1, G.dim[B]);
        const float xm = GradEval<E, ND, A, LEVEL - 1>::at(G, img, c);
        c[B] = reflect_idx(keep + 1, G.dim[B]);
        const float xp = GradEval<E, ND, A, LEVEL - 1>::at(G, img, c);
        c[B] = keep;
        double t = __dmul_rn((double)x0, 1.0);
        t = __dadd_rn(t, __dmul_rn(__dadd_rn((double)xm, (double)xp), 1.0));
        return (float)t;
    }
};

template <typename E, int ND, int A>
struct GradEval<E, ND, A, 0> {
    static __device__ __forceinline__ float at(const GradCtx<ND>& G, const E* __restrict__ img, int (&c)[ND])
    {
        long long base = 0;
#pragma unroll
        for (int d = 0; d < ND; ++d) base += (long long)c[d] * G.stride[d];
        const int keep = c[A];
        const long long lo = base + (long long)(reflect_idx(keep - 1, G.dim[A]) - keep) * G.stride[A];
        const long long hi = base + (long long)(reflect_idx(keep + 1, G.dim[A]) - keep) * G.stride[A];
        const double centre = Elem<E>::val(img[base]);
        double t = __dmul_rn(centre, 0.0);
        t = __dadd_rn(t, __dmul_rn(__dsub_rn(Elem<E>::val(img[hi]), Elem<E>::val(img[lo])), 1.0));
        return (float)t;
    }
};

template <typename E, int ND, int A>
__device__ __forceinline__ float grad_axis(const GradCtx<ND>& G, const E* __restrict__ img, int (&c)[ND])
{
    return GradEval<E, ND, A, ND - 1>::at(G, img, c);
}

template <typename E, int ND>
__global__ void __launch_bounds__(256) k_gradient_magnitude(GradCtx<ND> G, long long n, const E* __restrict__ img, float* __restrict__ out)
{
    const long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n) return;
    int c[ND];
    long long r = v;
#pragma unroll
    for (int d = 0; d < ND; ++d) { c[d] = (int)(r / G.stride[d]); r -= (long long)c[d] * G.stride[d]; }
    float acc;
    {
        const float g0 = grad_axis<E, ND, 0>(G, img, c);
        acc = __fmul_rn(g0, g0);
    }
    if (ND > 1) { const float g = grad_axis<E, ND, (ND > 1 ? 1 : 0)>(G, img, c); acc = __fadd_rn(acc, __fmul_rn(g, g)); }
    if (ND > 2) { const float g = grad_axis<E, ND, (ND > 2 ? 2 : 0)>(G, img, c); acc = __fadd_rn(acc, __fmul_rn(g, g)); }
    if (ND > 3) { const float g = grad_axis<E, ND, (ND > 3 ? 3 : 0)>(G, img, c); acc = __fadd_rn(acc, __fmul_rn(g, g)); }
    out[v] = __fsqrt_rn(acc);
}
