// tests/emu/pairwise_emu.cpp -- TEST INFRASTRUCTURE: medpy_b200/csrc/gc_pairwise.cuh compiled for the host, so that the
// summation the regional_atlas kernel uses can be compared with numpy.sum without a GPU.
#include "../../medpy_b200/csrc/gc_pairwise.cuh"

extern "C" float emu_pairwise_f32(const float* a, long long n) { return lab_pairwise_sum<float>(a, n); }
extern "C" double emu_pairwise_f64(const double* a, long long n) { return lab_pairwise_sum<double>(a, n); }
