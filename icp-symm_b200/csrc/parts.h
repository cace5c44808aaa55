// parts.h -- the canonical partition of a pair's queries into tiles and parts.  Plain C++ (host and device):
// included by common.cuh and by host/host_selftest.cpp, which checks its properties exhaustively on the CPU.
#pragma once
#include <stdint.h>
#ifdef __CUDACC__
#define ICP_HD __host__ __device__ __forceinline__
#else
#define ICP_HD inline
#endif

// ---- canonical order of the normal-equation sums -------------------------------------------------
// The queries of a pair are cut into 256-point tiles, the tiles into chunks of C consecutive tiles, and
// chunk c belongs to PART c mod 8: every part is an even sample of the whole curve (the cost of a far
// search varies 20x along it, so contiguous parts would leave the ranks of a sharded run unevenly loaded).
// K1 leaves one fp64 row per TILE (whichever block drew the tile), K2 adds the rows of a part in a fixed
// order and then the 8 part sums in part order.  Every step depends on the pair's own point count only --
// not on the grid, the batch the pair sits in, the scheduling of the tiles or the number of GPUs: a query
// shard is a run of whole parts (rank r of W owns parts [8r/W, 8(r+1)/W)), so a pair registered on 1, 2,
// 4 or 8 GPUs, alone or inside a batch, gives bit-identical sums, transforms and correspondences.
#define ICP_NPARTS 8
#define ICP_TILE 256                     // queries per tile = threads of a K1 block
ICP_HD uint32_t icp_tiles(uint32_t n) { return (n + ICP_TILE - 1u) / ICP_TILE; }
// tiles per chunk: a power of two, at most 32 (8192 queries), small enough that every part gets >= 8 chunks
ICP_HD uint32_t icp_part_chunk(uint32_t tiles) {
    uint32_t c = 32u;
    while (c > 1u && c * 64u > tiles) c >>= 1;
    return c;
}
// number of tiles of a part
ICP_HD uint32_t icp_part_tiles(uint32_t tiles, uint32_t part) {
    const uint32_t C = icp_part_chunk(tiles), round = ICP_NPARTS * C, full = tiles / round, rem = tiles - full * round;
    const uint32_t extra = rem > part * C ? (rem - part * C < C ? rem - part * C : C) : 0u;
    return full * C + extra;
}
// j-th tile of a part (j < icp_part_tiles)
ICP_HD uint32_t icp_part_tile(uint32_t tiles, uint32_t part, uint32_t j) {
    const uint32_t C = icp_part_chunk(tiles);
    return ((j / C) * ICP_NPARTS + part) * C + j % C;
}
// the tiles of a run of parts [part0, part0 + nparts) in curve order: work item w -> tile (>= tiles: none);
// in every round of 8 chunks the run owns nparts consecutive chunks
ICP_HD uint32_t icp_shard_tile(uint32_t tiles, uint32_t part0, uint32_t nparts, uint32_t w) {
    const uint32_t C = icp_part_chunk(tiles), run = nparts * C;
    return (w / run) * ICP_NPARTS * C + part0 * C + w % run;
}
// work items of such a run (monotone in `tiles`: a smaller pair of a batch never needs more than the largest)
ICP_HD uint32_t icp_shard_work(uint32_t tiles, uint32_t nparts) {
    const uint32_t C = icp_part_chunk(tiles), round = ICP_NPARTS * C;
    return ((tiles + round - 1u) / round) * nparts * C;
}

