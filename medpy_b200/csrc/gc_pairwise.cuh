// gc_pairwise.cuh -- numpy's pairwise summation, reproduced exactly (used by the regional_atlas reduction in
// gc_labels.cuh; energy_label.py:384-386 calls numpy.sum on the atlas values of one region).
//
// Plain inline functions, compiled for the device by nvcc and for the host by tests/emu/pairwise_emu.cpp, which checks
// them against numpy.sum bit for bit (float32 and float64, n = 1 .. 100003).
#pragma once

#if defined(__CUDACC__)
#define PW_HD __host__ __device__ __forceinline__
#else
#define PW_HD inline
#endif

template <typename V> PW_HD V lab_add(V a, V b);
template <> PW_HD float lab_add<float>(float a, float b)
{
#if defined(__CUDA_ARCH__)
    return __fadd_rn(a, b);
#else
    return a + b;
#endif
}
template <> PW_HD double lab_add<double>(double a, double b)
{
#if defined(__CUDA_ARCH__)
    return __dadd_rn(a, b);
#else
    return a + b;
#endif
}

// numpy/_core/src/umath/loops_utils.h.src, pairwise_sum: fewer than 8 elements front to back; up to 128 elements eight
// interleaved partial sums combined as ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)) plus a front-to-back tail; longer runs are
// halved (first half rounded down to a multiple of 8) and the two halves added.
template <typename V>
PW_HD V lab_pairwise_leaf(const V* a, long long n)      // n <= 128
{
    if (n < 8) {
        V res = (V)0;
        for (long long i = 0; i < n; ++i) res = lab_add<V>(res, a[i]);
        return res;
    }
    V r[8];
    for (int j = 0; j < 8; ++j) r[j] = a[j];
    long long i = 8;
    for (; i < n - (n % 8); i += 8) {
        for (int j = 0; j < 8; ++j) r[j] = lab_add<V>(r[j], a[i + j]);
    }
    V res = lab_add<V>(lab_add<V>(lab_add<V>(r[0], r[1]), lab_add<V>(r[2], r[3])),
                       lab_add<V>(lab_add<V>(r[4], r[5]), lab_add<V>(r[6], r[7])));
    for (; i < n; ++i) res = lab_add<V>(res, a[i]);
    return res;
}

template <typename V>
PW_HD V lab_pairwise_sum(const V* a, long long n)
{
    if (n <= 128) return lab_pairwise_leaf<V>(a, n);
    // explicit post-order walk of the halving tree (depth <= 25 for n < 2^31)
    long long fs[32], fn[32];
    int phase[32];
    V left[32];
    int sp = 0;
    fs[0] = 0; fn[0] = n; phase[0] = 0;
    V ret = (V)0;
    while (sp >= 0) {
        if (fn[sp] <= 128) { ret = lab_pairwise_leaf<V>(a + fs[sp], fn[sp]); --sp; continue; }
        long long n2 = fn[sp] / 2;
        n2 -= n2 % 8;
        if (phase[sp] == 0) {
            phase[sp] = 1;
            fs[sp + 1] = fs[sp]; fn[sp + 1] = n2; phase[sp + 1] = 0;
            ++sp;
        } else if (phase[sp] == 1) {
            left[sp] = ret;
            phase[sp] = 2;
            fs[sp + 1] = fs[sp] + n2; fn[sp + 1] = fn[sp] - n2; phase[sp + 1] = 0;
            ++sp;
        } else {
            ret = lab_add<V>(left[sp], ret);
            --sp;
        }
    }
    return ret;
}
