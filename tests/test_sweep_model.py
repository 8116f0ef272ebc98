"""CPU model of the warp-level row sweep of the global relabel (medpy_b200/csrc/gc_sweep.cuh: sweep_row_dir): the same three
steps -- sequential relaxation inside each lane's chunk, a Hillis-Steele scan over the lanes of the chunk summaries
f(c) = min(a, c + b) composed as (a2, b2) o (a1, b1) = (min(a2, a1 + b2), b1 + b2), application of the incoming label along
the open prefix -- written in numpy-free Python with the kernel's saturating arithmetic, against the obvious sequential
sweep.  It pins the ALGEBRA the kernel relies on (associativity of the composition, the early-exit rule of pass 2); the CUDA
code itself is checked on the GPU against the worklist BFS and BK (tests/test_gpu_round2.py)."""
import random

HINF = 0x3FFFFFFF


def inc(h):
    return HINF if h >= HINF else h + 1


def sequential(h, open_, fwd, carry):
    h = list(h)
    idx = range(len(h)) if fwd else range(len(h) - 1, -1, -1)
    prev = carry
    for x in idx:
        if open_[x]:
            c = inc(prev)
            if c < h[x]:
                h[x] = c
        prev = h[x]
    return h, prev


def warp_model(h, open_, fwd, carry_in, lanes=32):
    n = len(h)
    h = list(h)
    ln = (n + lanes - 1) // lanes
    a, b, nopen_l, cnt_l, c0_l = [0] * lanes, [0] * lanes, [0] * lanes, [0] * lanes, [0] * lanes
    for lane in range(lanes):
        c0 = lane * ln
        cnt = max(0, min(ln, n - c0))
        prev, nopen, chain = HINF, 0, True
        for j in range(cnt):
            x = c0 + j if fwd else c0 + cnt - 1 - j
            if open_[x]:
                c = inc(prev)
                if c < h[x]:
                    h[x] = c
            if chain:
                if open_[x]:
                    nopen += 1
                else:
                    chain = False
            prev = h[x]
        a[lane] = prev if cnt else HINF
        b[lane] = (cnt if nopen == cnt else HINF) if cnt else 0
        nopen_l[lane], cnt_l[lane], c0_l[lane] = nopen, cnt, c0
    # inclusive scan in sweep order
    o = 1
    while o < lanes:
        na, nb = list(a), list(b)
        for lane in range(lanes):
            src = lane - o if fwd else lane + o
            if 0 <= src < lanes:
                a1, b1 = a[src], b[src]
                t = HINF if (a1 >= HINF or b[lane] >= HINF) else min(a1 + b[lane], HINF)
                na[lane] = min(a[lane], t)
                nb[lane] = HINF if (b1 >= HINF or b[lane] >= HINF) else min(b1 + b[lane], HINF)
        a, b = na, nb
        o <<= 1
    for lane in range(lanes):
        prev_lane = lane - 1 if fwd else lane + 1
        if not (0 <= prev_lane < lanes):
            cin = carry_in
        else:
            pa, pb = a[prev_lane], b[prev_lane]
            t = HINF if (carry_in >= HINF or pb >= HINF) else min(carry_in + pb, HINF)
            cin = min(pa, t)
        if cin < HINF:
            c = cin
            for j in range(nopen_l[lane]):
                x = c0_l[lane] + j if fwd else c0_l[lane] + cnt_l[lane] - 1 - j
                c = inc(c)
                if c < h[x]:
                    h[x] = c
                else:
                    break
    last = lanes - 1 if fwd else 0
    t = HINF if (carry_in >= HINF or b[last] >= HINF) else min(carry_in + b[last], HINF)
    return h, min(a[last], t)


def test_warp_scan_equals_sequential_sweep():
    rng = random.Random(7)
    for trial in range(600):
        n = rng.choice([1, 2, 5, 31, 32, 33, 64, 100, 255, 256, 1000, 1024])
        p_open = rng.choice([0.0, 0.3, 0.8, 0.97, 1.0])
        style = rng.randrange(3)
        if style == 0:
            h = [rng.choice([1, HINF, HINF, HINF, rng.randrange(2, 50)]) for _ in range(n)]
        elif style == 1:
            h = [HINF] * n
            for _ in range(max(1, n // 40)):
                h[rng.randrange(n)] = 1
        else:
            h = [rng.randrange(1, 2000) for _ in range(n)]
        open_ = [rng.random() < p_open for _ in range(n)]
        carry = rng.choice([HINF, 1, 7, 300])
        for fwd in (True, False):
            # the first voxel in sweep direction has no predecessor inside a segment unless a carry comes in
            want, wlast = sequential(h, open_, fwd, carry)
            got, glast = warp_model(h, open_, fwd, carry)
            assert got == want, (trial, n, fwd)
            assert glast == wlast, (trial, n, fwd)


def test_composition_is_associative():
    rng = random.Random(3)

    def comp(f2, f1):     # first f1, then f2
        a1, b1 = f1
        a2, b2 = f2
        t = HINF if (a1 >= HINF or b2 >= HINF) else min(a1 + b2, HINF)
        return (min(a2, t), HINF if (b1 >= HINF or b2 >= HINF) else min(b1 + b2, HINF))

    def apply(f, c):
        a, b = f
        t = HINF if (c >= HINF or b >= HINF) else min(c + b, HINF)
        return min(a, t)

    for _ in range(2000):
        fs = [(rng.choice([HINF, rng.randrange(1, 500)]), rng.choice([HINF, 0, rng.randrange(1, 40)])) for _ in range(3)]
        left = comp(fs[2], comp(fs[1], fs[0]))
        right = comp(comp(fs[2], fs[1]), fs[0])
        for c in (HINF, 1, 17, 400):
            assert apply(left, c) == apply(right, c) == apply(fs[2], apply(fs[1], apply(fs[0], c)))
