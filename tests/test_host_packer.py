"""CPU test of the host-side marker bit packer (medpy_b200/csrc/host_pack.hpp): worker threads, ordered progress counter,
partial last word -- compiled on its own with g++ (no CUDA, no GPU)."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

HARNESS = r"""
#include <cstdio>
#include <cstdlib>
#include "host_pack.hpp"
int main()
{
    for (size_t n : {size_t(1), size_t(5), size_t(32), size_t(33), size_t(1000003), size_t(1) << 22, (size_t(1) << 22) + 17, size_t(3) << 21}) {
        std::vector<uint8_t> a(n + 64), b(n + 64);
        for (size_t i = 0; i < n; ++i) { a[i] = (rand() % 7 == 0); b[i] = (rand() % 5 == 0) ? 255 : 0; }
        const size_t words = (n + 31) / 32;
        std::vector<uint32_t> A(words, 0xdeadbeefu), B(words, 0xdeadbeefu);
        int64_t last = 0;
        {
            MarkerPacker p(a.data(), b.data(), A.data(), B.data(), n);
            for (;;) {      // the progress counter only ever grows and ends at `words`
                const int64_t r = p.ready.load();
                if (r < last) { printf("progress went backwards\n"); return 1; }
                last = r;
                if (r >= (int64_t)words) break;
            }
        }
        for (size_t i = 0; i < n; ++i) {
            const bool fa = (A[i >> 5] >> (i & 31)) & 1, fb = (B[i >> 5] >> (i & 31)) & 1;
            if (fa != (a[i] != 0) || fb != (b[i] != 0)) { printf("MISMATCH n=%zu i=%zu\n", n, i); return 1; }
        }
        if (n % 32) {       // bits beyond n in the last word are zero
            const uint32_t tail = A[words - 1] >> (n % 32);
            if (tail) { printf("tail bits set n=%zu\n", n); return 1; }
        }
    }
    // one plane absent
    {
        const size_t n = 100000;
        std::vector<uint8_t> a(n + 64, 1);
        std::vector<uint32_t> A((n + 31) / 32, 0);
        { MarkerPacker p(a.data(), nullptr, A.data(), nullptr, n); while (p.ready.load() < (int64_t)A.size()) {} }
        for (size_t i = 0; i < n; ++i) if (!((A[i >> 5] >> (i & 31)) & 1)) { printf("MISMATCH single plane\n"); return 1; }
    }
    printf("ok\n");
    return 0;
}
"""


def test_marker_packer(tmp_path):
    src = tmp_path / "packtest.cpp"
    src.write_text(HARNESS)
    exe = tmp_path / "packtest"
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-pthread", "-I", os.path.join(ROOT, "medpy_b200", "csrc"), str(src), "-o", str(exe)])
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and out.stdout.strip().endswith("ok"), out.stdout + out.stderr
