// tile_map.cuh — (node, batch row) tiles that are visited in BOTH memory orders by one CTA of 256 threads.
//
// The reference's observation tensors are row-major [B, N, C] (node innermost among the sample's rows), everything the
// MPNN kernels produce is node-major with the batch row innermost (element (b, n) at n*B + b). A kernel that maps one
// thread to one (node, row) pair in either order reads or writes the other tensor with a stride of a whole sample
// (N*C*4 bytes: one sector and one TLB entry per lane). A CTA therefore owns a tile of kTilePairs pairs — Bp rows
// (the batch chunk rounded up to a power of two <= 32) x TN = kTilePairs/Bp consecutive nodes — walks it once with
// the NODE innermost (lanes = 32 consecutive nodes of one row: the [B, N, C] side is contiguous), once with the ROW
// innermost (lanes = the rows of one node: the node-major side is contiguous), and passes values between the two
// walks through shared memory (pitch TN+1: both walks are bank-conflict-free when Bp = 32).
#pragma once
#include <stdint.h>

namespace tarl {

constexpr int kTileThreads = 256;
constexpr int kTilePairs = 1024;
constexpr int kTileSmem = kTilePairs + 32;      // floats: Bp * (TN + 1) <= 1024 + 32

struct Tile {
    int n0, b0;        // first node / first batch row of the tile
    int TN, Bp, sh;    // nodes per tile, rows per tile (power of two), log2(Bp)
    int nrows;         // live rows: min(Bp, B - b0)
};

inline int tile_rows_pow2(int B) {
    int p = 1;
    while (p < B && p < 32) p <<= 1;
    return p;
}
// grid of a tiled kernel: x = node tiles, y = chunks of 32 batch rows
inline dim3 tile_grid(int N, int B) {
    const int TN = kTilePairs / tile_rows_pow2(B);
    return dim3((unsigned)((N + TN - 1) / TN), (unsigned)((B + 31) / 32));
}
inline int tile_count(int N, int B) {
    const dim3 g = tile_grid(N, B);
    return (int)(g.x * g.y);
}

__device__ __forceinline__ Tile tile_here(int B, int Bp) {
    Tile t;
    t.Bp = Bp;
    t.sh = 31 - __clz(Bp);
    t.TN = kTilePairs >> t.sh;
    t.n0 = blockIdx.x * t.TN;
    t.b0 = blockIdx.y * 32;
    t.nrows = min(Bp, B - t.b0);
    return t;
}
__device__ __forceinline__ int tile_slot(const Tile& t, int r, int j) { return r * (t.TN + 1) + j; }

// Walk with the NODE innermost: f(r, j) for every pair of the tile, the 32 lanes of a warp on 32 consecutive nodes j of
// one row r. Walk with the ROW innermost: f(r, j) with the Bp rows of node j on consecutive lanes.
template <typename F>
__device__ __forceinline__ void tile_walk_nodes(const Tile& t, F f) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int segs = t.TN >> 5;                                  // 32-node segments per row
    for (int u = warp; u < t.Bp * segs; u += kTileThreads / 32) {
        const int r = u / segs, j = (u - r * segs) * 32 + lane;
        f(r, j);
    }
}
template <typename F>
__device__ __forceinline__ void tile_walk_rows(const Tile& t, F f) {
    for (int p = threadIdx.x; p < kTilePairs; p += kTileThreads) f(p & (t.Bp - 1), p >> t.sh);
}

}  // namespace tarl
