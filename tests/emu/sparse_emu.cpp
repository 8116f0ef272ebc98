// tests/emu/sparse_emu.cpp -- TEST INFRASTRUCTURE.
// Compiles the per-node bodies of medpy_b200/csrc/gc_sparse.cuh (the very functions the CUDA kernels wrap) as plain
// host C++ and drives them with the same round structure as the device loop in gc_sparse_api.cu (sparse_solve), one
// node after the other.  It checks the LOGIC of the sparse push-relabel where no GPU is available; races and launch
// code are only exercised by the `-m gpu` tests.
#include <cstdint>
#include <vector>

#include "../../medpy_b200/csrc/gc_sparse.cuh"

extern "C" int emu_sparse_solve(int n, int m2, const int* row, const int* head, const int* sis, const double* cap_in,
                                const double* tr, int push_steps, int sweeps_per_round, uint8_t* mask_out,
                                double* absorbed_out, long long* rounds_out)
{
    std::vector<double> cap(cap_in, cap_in + m2), excess(n), sunk(n);
    std::vector<int> height(n);
    SparseState S{};
    S.n = n; S.m2 = m2; S.row = row; S.head = head; S.sis = sis; S.cap = cap.data(); S.tr = tr;
    S.excess = excess.data(); S.sunk = sunk.data(); S.height = height.data();
    for (int u = 0; u < n; ++u) sp_init_node(S, u);
    long long rounds = 0;
    for (;;) {
        for (int u = 0; u < n; ++u) sp_relabel_init_node(S, u);
        for (;;) {
            bool changed = false;
            for (int u = 0; u < n; ++u) changed |= sp_relax_node(S, u);
            if (!changed) break;
        }
        long long active = 0;
        for (int u = 0; u < n; ++u) active += sp_is_active(S, u) ? 1 : 0;
        if (!active) break;
        if (++rounds > 1000000) return -1;
        for (int s = 0; s < sweeps_per_round; ++s)
            for (int u = 0; u < n; ++u) sp_push_node(S, u, push_steps);
    }
    double a = 0.0;
    for (int u = 0; u < n; ++u) { mask_out[u] = height[u] >= SP_HINF ? 1 : 0; a += sunk[u]; }
    *absorbed_out = a;
    if (rounds_out) *rounds_out = rounds;
    return 0;
}
