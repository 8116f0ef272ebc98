// gc_tiles.cuh -- tile-resident solver kernels for the 3-D lattice (the production path; gc_solver.cuh keeps
// the per-voxel kernels used for 4-D lattices and as an A/B reference).
//
// The lattice is cut into 8x8x8 tiles; one 512-thread CTA owns one tile for the duration of a visit, keeps
// the tile's state in shared memory (heights with a 1-voxel halo; for the push kernel also the six residual
// capacity planes and the excess) and iterates there, so a visit costs one read and one write of the tile in
// HBM however many push/relabel or relaxation rounds it takes.  Work is driven by per-tile flags: a tile is
// visited only if it may hold work (an active voxel / a label that may still drop).
//
//  * k_relabel_init_tile / k_relabel_tile : exact backward BFS from the sink (global relabel).  Heights only
//    decrease during it, so tiles can run concurrently with benign races on the halo; a tile whose border
//    labels dropped flags its face neighbours for the next pass.
//  * k_push_tile : push/relabel discharge of one tile, several synchronous rounds in shared memory.  Tiles are
//    processed in two colours (3-D checkerboard): tiles of one colour are never face-adjacent, so a running tile
//    is the only writer of its own voxels and of its neighbours' border state, except that corner/edge
//    voxels of an idle tile can receive from up to three running tiles at once -> neighbour excess uses atomics.
#pragma once
#include "gc_common.cuh"

#define TILE 8
#define TILE_VOX 512
#define HALO_DIM 10
#define HALO_VOX 1000

struct Tiles {
    int nt[3];      // tiles along z, y, x
    int ntiles;
};

__device__ __forceinline__ int hidx(int z, int y, int x) { return (z * HALO_DIM + y) * HALO_DIM + x; }

// offsets in the halo cube for direction k (axis 0 = z slowest)
__device__ __forceinline__ int hoff(int k)
{
    const int s = (k >> 1) == 0 ? HALO_DIM * HALO_DIM : ((k >> 1) == 1 ? HALO_DIM : 1);
    return (k & 1) ? s : -s;
}

struct TileCtx {
    int tz, ty, tx;        // tile coordinates
    int lz, ly, lx;        // local coordinates of this thread's voxel
    int gz, gy, gx;        // global coordinates
    bool inb;              // voxel inside the lattice
    bool own;              // ... and owned (not a ghost plane of a z-slab)
    unsigned v;            // flat index (valid when inb)
};

__device__ __forceinline__ TileCtx tile_ctx(const Lattice& L, const Tiles& TL, int t)
{
    TileCtx c;
    c.tx = t % TL.nt[2];
    int r = t / TL.nt[2];
    c.ty = r % TL.nt[1];
    c.tz = r / TL.nt[1];
    const int tid = threadIdx.x;
    c.lx = tid & 7; c.ly = (tid >> 3) & 7; c.lz = tid >> 6;
    c.gz = c.tz * TILE + c.lz; c.gy = c.ty * TILE + c.ly; c.gx = c.tx * TILE + c.lx;
    c.inb = c.gz < L.dim[0] && c.gy < L.dim[1] && c.gx < L.dim[2];
    c.v = c.inb ? (unsigned)c.gz * L.stride[0] + (unsigned)c.gy * L.stride[1] + (unsigned)c.gx : 0u;
    c.own = c.inb && c.gz >= L.own0 && c.gz < L.own1;
    return c;
}

// cooperative load of heights (own voxel + the six halo faces) into the 10^3 cube; out-of-lattice -> HINF
__device__ __forceinline__ int load_heights(const Lattice& L, const TileCtx& c, const int* __restrict__ height, int* sh)
{
    const int tid = threadIdx.x;
    int h0 = c.inb ? height[c.v] : MGC_HINF;
    sh[hidx(c.lz + 1, c.ly + 1, c.lx + 1)] = h0;
    if (tid < 384) {
        const int face = tid >> 6, a = (tid >> 3) & 7, b = tid & 7;
        int z, y, x;   // local coordinates in [-1, 8]
        switch (face) {
            case 0: z = -1; y = a; x = b; break;
            case 1: z = TILE; y = a; x = b; break;
            case 2: z = a; y = -1; x = b; break;
            case 3: z = a; y = TILE; x = b; break;
            case 4: z = a; y = b; x = -1; break;
            default: z = a; y = b; x = TILE; break;
        }
        const int gz = c.tz * TILE + z, gy = c.ty * TILE + y, gx = c.tx * TILE + x;
        int h = MGC_HINF;
        if (gz >= 0 && gy >= 0 && gx >= 0 && gz < L.dim[0] && gy < L.dim[1] && gx < L.dim[2])
            h = height[(unsigned)gz * L.stride[0] + (unsigned)gy * L.stride[1] + (unsigned)gx];
        sh[hidx(z + 1, y + 1, x + 1)] = h;
    }
    return h0;
}

__device__ __forceinline__ void flag_tile(int* __restrict__ flags, int* __restrict__ counter, int t)
{
    if (atomicExch(&flags[t], 1) == 0 && counter) atomicAdd(counter, 1);
}

// ---------------------------------------------------------------------------------------------------
// global relabel, tile form
// ---------------------------------------------------------------------------------------------------
// init: residual bit mask, label 1 for voxels with a residual sink link else HINF; flag tiles that hold a voxel
// which still has to find its distance (unlabelled but with residual out-arcs).
template <typename T>
__global__ void __launch_bounds__(TILE_VOX) k_relabel_init_tile(Lattice L, Tiles TL, State<T> S, int* __restrict__ flags,
                                                                int* __restrict__ counter)
{
    const int t = blockIdx.x;
    const TileCtx c = tile_ctx(L, TL, t);
    int needs = 0;
    if (c.inb) {
        unsigned m = 0;
#pragma unroll
        for (int k = 0; k < 6; ++k)
            if (S.cap[k][c.v] > 0) m |= 1u << k;
        S.rmask[c.v] = (uint8_t)m;
        // ghost planes restart at HINF as well (see k_relabel_init)
        const int h = (c.own && (-S.tr[c.v]) - S.sink[c.v] > 0) ? 1 : MGC_HINF;
        S.height[c.v] = h;
        needs = (c.own && m != 0 && h == MGC_HINF) ? 1 : 0;
    }
    const int any = __syncthreads_or(needs);
    if (threadIdx.x == 0) {
        flags[t] = any;
        if (any) atomicAdd(counter, 1);
    }
}

// one visit: relax inside the tile until nothing changes, write back, wake the face neighbours whose halo changed
__global__ void __launch_bounds__(TILE_VOX) k_relabel_tile(Lattice L, Tiles TL, const uint8_t* __restrict__ rmask,
                                                           int* __restrict__ height, int* __restrict__ flag_cur,
                                                           int* __restrict__ flag_next, int* __restrict__ counter_next)
{
    __shared__ int sh[HALO_VOX];
    const int t = blockIdx.x;
    if (flag_cur[t] == 0) return;
    const TileCtx c = tile_ctx(L, TL, t);
    const int h0 = load_heights(L, c, height, sh);
    const unsigned m = c.own ? rmask[c.v] : 0u;
    __syncthreads();
    if (threadIdx.x == 0) flag_cur[t] = 0;
    const int me = hidx(c.lz + 1, c.ly + 1, c.lx + 1);
    int h = h0;
    for (;;) {
        int changed = 0;
        if (m && h > 1) {
            int best = h;
#pragma unroll
            for (int k = 0; k < 6; ++k)
                if (m & (1u << k)) { const int hw = sh[me + hoff(k)] + 1; best = hw < best ? hw : best; }
            if (best < h) { h = best; sh[me] = h; changed = 1; }
        }
        if (!__syncthreads_or(changed)) break;
    }
    if (h != h0) {
        height[c.v] = h;
        if (c.lz == 0 && c.tz > 0) flag_tile(flag_next, counter_next, t - TL.nt[1] * TL.nt[2]);
        if (c.lz == TILE - 1 && c.tz + 1 < TL.nt[0]) flag_tile(flag_next, counter_next, t + TL.nt[1] * TL.nt[2]);
        if (c.ly == 0 && c.ty > 0) flag_tile(flag_next, counter_next, t - TL.nt[2]);
        if (c.ly == TILE - 1 && c.ty + 1 < TL.nt[1]) flag_tile(flag_next, counter_next, t + TL.nt[2]);
        if (c.lx == 0 && c.tx > 0) flag_tile(flag_next, counter_next, t - 1);
        if (c.lx == TILE - 1 && c.tx + 1 < TL.nt[2]) flag_tile(flag_next, counter_next, t + 1);
    }
}

// ---------------------------------------------------------------------------------------------------
// push / relabel, tile form (one colour of the 3-D checkerboard per launch)
// ---------------------------------------------------------------------------------------------------
// grid: nt[0] * nt[1] * ceil(nt[2] / 2) blocks; block b of colour `color` maps to the tile whose x index has the
// parity that makes (tz + ty + tx) & 1 == color.
template <typename T>
__global__ void __launch_bounds__(TILE_VOX) k_push_tile(Lattice L, Tiles TL, State<T> S, int color, int iters,
                                                        int* __restrict__ tflag, int* __restrict__ n_still_active)
{
    __shared__ T s_cap[6 * TILE_VOX];
    __shared__ T s_exc[TILE_VOX];
    __shared__ int s_h[HALO_VOX];

    const int half = (TL.nt[2] + 1) >> 1;
    const int bx = blockIdx.x % half;
    const int r = blockIdx.x / half;
    const int tyy = r % TL.nt[1], tzz = r / TL.nt[1];
    const int txx = 2 * bx + ((tzz + tyy + color) & 1);
    if (txx >= TL.nt[2]) return;
    const int t = (tzz * TL.nt[1] + tyy) * TL.nt[2] + txx;
    if (tflag[t] == 0) return;

    const TileCtx c = tile_ctx(L, TL, t);
    const int tid = threadIdx.x;
    const int me = hidx(c.lz + 1, c.ly + 1, c.lx + 1);
    const int h0 = load_heights(L, c, S.height, s_h);
    T e0 = 0, c0[6], scap = 0, sf0 = 0;
#pragma unroll
    for (int k = 0; k < 6; ++k) { c0[k] = c.inb ? S.cap[k][c.v] : (T)0; s_cap[k * TILE_VOX + tid] = c0[k]; }
    if (c.inb) {
        e0 = S.excess[c.v];
        const T tr = S.tr[c.v];
        if (tr < 0) { scap = -tr; sf0 = S.sink[c.v]; }
    }
    s_exc[tid] = e0;
    T sf = sf0;
    int h = h0;
    __syncthreads();
    if (tid == 0) tflag[t] = 0;

    // neighbour bookkeeping: local index inside the tile or, for a halo voxel, its global index and tile
    for (int it = 0; it < iters; ++it) {
        const T e_in = s_exc[tid];
        int act = (c.own && e_in > 0 && h < MGC_HINF) ? 1 : 0;
        if (act) {
            T e = e_in, pushed = 0;
            if (scap > 0) {
                const T rr = scap - sf;
                if (rr > 0) {
                    T d;
                    if (e < rr) { d = e; sf += d; } else { d = rr; sf = scap; }
                    e -= d; pushed += d;
                }
            }
            int newh = h;
            if (e > 0) {
                T cc[6];
                int hn[6];
#pragma unroll
                for (int k = 0; k < 6; ++k) {
                    cc[k] = s_cap[k * TILE_VOX + tid];
                    hn[k] = cc[k] > 0 ? s_h[me + hoff(k)] : MGC_HINF;
                }
#pragma unroll 1
                for (int rep = 0; rep < 6; ++rep) {
                    int kb = -1, hb = MGC_HINF;
#pragma unroll
                    for (int k = 0; k < 6; ++k)
                        if (cc[k] > 0 && hn[k] < hb) { hb = hn[k]; kb = k; }
                    if (kb < 0) { newh = MGC_HINF; break; }
                    if (hb >= h) { newh = hb + 1; break; }
                    const T d = e < cc[kb] ? e : cc[kb];
                    // where does the arc lead?
                    const int ax = kb >> 1, sgn = (kb & 1) ? 1 : -1;
                    const int lc = (ax == 0 ? c.lz : (ax == 1 ? c.ly : c.lx)) + sgn;
                    atomicAdd(&s_cap[kb * TILE_VOX + tid], -d);
                    if (lc >= 0 && lc < TILE) {
                        const int wt = tid + sgn * (ax == 0 ? 64 : (ax == 1 ? 8 : 1));
                        atomicAdd(&s_cap[(kb ^ 1) * TILE_VOX + wt], d);
                        atomicAdd(&s_exc[wt], d);
                    } else {
                        const unsigned w = (unsigned)((int)c.v + dir_offset(L, kb));
                        atomicAdd(&S.cap[kb ^ 1][w], d);
                        atomicAdd(&S.excess[w], d);
                        const int nt_ = t + sgn * (ax == 0 ? TL.nt[1] * TL.nt[2] : (ax == 1 ? TL.nt[2] : 1));
                        tflag[nt_] = 1;
                    }
                    e -= d; pushed += d;
                    cc[kb] = 0;
                    if (!(e > 0)) break;
                }
            }
            if (newh != h) { h = newh; s_h[me] = h; }
            if (pushed > 0) atomicAdd(&s_exc[tid], -pushed);
        }
        if (!__syncthreads_or(act)) break;
    }

    // write back what changed (this CTA is the only writer of its own voxels during this launch)
    const T e1 = s_exc[tid];
    if (c.inb) {
        if (e1 != e0) S.excess[c.v] = e1;
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            const T ck = s_cap[k * TILE_VOX + tid];
            if (ck != c0[k]) S.cap[k][c.v] = ck;
        }
        if (h != h0) S.height[c.v] = h;
        if (sf != sf0) S.sink[c.v] = sf;
    }
    const int still = (c.own && e1 > 0 && h < MGC_HINF) ? 1 : 0;
    if (__syncthreads_or(still) && tid == 0) {
        tflag[t] = 1;
        if (n_still_active) atomicAdd(n_still_active, 1);
    }
}

__global__ void k_fill_int(int* __restrict__ p, int n, int val)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = val;
}
