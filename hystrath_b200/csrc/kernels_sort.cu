// kernels_sort.cu -- stage 2: rebuild cellOccupancy by a per-cell counting sort.
//
// Reference: dsmcCloud::buildCellOccupancy (DSMC/clouds/dsmcCloud.C:63-74) clears one
// DynamicList<dsmcParcel*> per cell and appends every parcel pointer in cloud-list order.
// Here occupancy is a CSR offset array over a physically reordered SoA cloud:
//   histogram (fused into the move kernel) -> exclusive scan -> index scatter (atomic cursor)
//   -> per-cell ordering of the scattered indices (restores list order => deterministic,
//      stable result) -> payload gather into the second buffer.
// All integer work: results are bit-exact by construction.
#include <algorithm>

#include "engine.h"

namespace dsmc {

namespace {
constexpr int SCAN_BLOCK = 512;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_BLOCK * SCAN_ITEMS;

__device__ __forceinline__ int warpInclusiveScan(int v) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int n = __shfl_up_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) >= o) v += n;
    }
    return v;
}

// exclusive scan of `v` over the block; returns exclusive prefix, total via *total
__device__ int blockExclusiveScan(int v, int* total) {
    __shared__ int warpSums[32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int inc = warpInclusiveScan(v);
    if (lane == 31) warpSums[w] = inc;
    __syncthreads();
    if (w == 0) {
        int s = lane < (blockDim.x >> 5) ? warpSums[lane] : 0;
        int si = warpInclusiveScan(s);
        warpSums[lane] = si - s;
        if (lane == 31) *total = si;
    }
    __syncthreads();
    int r = warpSums[w] + inc - v;
    __syncthreads();
    return r;
}
}  // namespace

__global__ void __launch_bounds__(SCAN_BLOCK) scanTileSums(const int32_t* __restrict__ in, int32_t n, int32_t* __restrict__ tileSums) {
    __shared__ int total;
    const int base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    int s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k)
        if (base + k < n) s += in[base + k];
    blockExclusiveScan(s, &total);
    if (threadIdx.x == 0) tileSums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(SCAN_BLOCK) scanTileOffsets(int32_t* tileSums, int32_t nTiles) {
    __shared__ int total;
    int carry = 0;
    for (int base = 0; base < nTiles; base += SCAN_BLOCK) {
        int idx = base + threadIdx.x;
        int v = idx < nTiles ? tileSums[idx] : 0;
        int ex = blockExclusiveScan(v, &total);
        if (idx < nTiles) tileSums[idx] = carry + ex;
        carry += total;
        __syncthreads();
    }
    if (threadIdx.x == 0) tileSums[nTiles] = carry;
}

__global__ void __launch_bounds__(SCAN_BLOCK) scanFinal(const int32_t* __restrict__ in, int32_t n, const int32_t* __restrict__ tileSums,
                                                        int32_t nTiles, int32_t* __restrict__ out, int32_t* __restrict__ out2) {
    __shared__ int total;
    const int base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    int v[SCAN_ITEMS];
    int s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        v[k] = (base + k < n) ? in[base + k] : 0;
        s += v[k];
    }
    int ex = blockExclusiveScan(s, &total) + tileSums[blockIdx.x];
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        if (base + k < n) {
            out[base + k] = ex;
            if (out2) out2[base + k] = ex;
        }
        ex += v[k];
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) out[n] = tileSums[nTiles];
}

int32_t scanScratchInts(int32_t n) { return (n + SCAN_TILE - 1) / SCAN_TILE + 2; }

// out[0..n] = exclusive scan of in[0..n) (out[n] = total); out2 (optional) receives a copy of out[0..n)
cudaError_t launchExclusiveScan(const int32_t* in, int32_t* out, int32_t* out2, int32_t n, int32_t* tileSums, cudaStream_t s) {
    const int nTiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    scanTileSums<<<nTiles, SCAN_BLOCK, 0, s>>>(in, n, tileSums);
    scanTileOffsets<<<1, SCAN_BLOCK, 0, s>>>(tileSums, nTiles);
    scanFinal<<<nTiles, SCAN_BLOCK, 0, s>>>(in, n, tileSums, nTiles, out, out2);
    return cudaGetLastError();
}

__global__ void histogramKernel(const int32_t* __restrict__ cell, int32_t n, int32_t* __restrict__ cellCount) {
    const int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int32_t c = cell[i];
    if (c >= 0) atomicAdd(&cellCount[c], 1);
}

cudaError_t launchHistogram(const int32_t* cell, int32_t n, int32_t* cellCount, cudaStream_t s) {
    if (n <= 0) return cudaSuccess;
    histogramKernel<<<(n + 255) / 256, 256, 0, s>>>(cell, n, cellCount);
    return cudaGetLastError();
}

// Warp-aggregated: the cloud enters cell-sorted and most parcels stay in their cell, so the 32 lanes of a warp fall into a handful
// of cells; the lanes of one cell send ONE atomic for the whole group and take consecutive slots in lane order.
namespace {
constexpr int SCATTER_PER_THREAD = 4;   // independent elements per thread: four reads, then four atomics in flight
}
__global__ void __launch_bounds__(256) scatterIndexKernel(const int32_t* __restrict__ cell, int32_t n, int32_t* __restrict__ cursor, int32_t* __restrict__ perm) {
    const int32_t i0 = blockIdx.x * (256 * SCATTER_PER_THREAD) + threadIdx.x;
    const int lane = threadIdx.x & 31;
    int32_t c[SCATTER_PER_THREAD], base[SCATTER_PER_THREAD];
    unsigned peers[SCATTER_PER_THREAD];
#pragma unroll
    for (int k = 0; k < SCATTER_PER_THREAD; ++k) { const int32_t i = i0 + 256 * k; c[k] = i < n ? cell[i] : -1; }
#pragma unroll
    for (int k = 0; k < SCATTER_PER_THREAD; ++k) {
        const unsigned live = __ballot_sync(0xffffffffu, c[k] >= 0);
        peers[k] = 0u; base[k] = 0;
        if (c[k] >= 0) {
            peers[k] = __match_any_sync(live, c[k]);
            if (lane == __ffs(peers[k]) - 1) base[k] = atomicAdd(&cursor[c[k]], __popc(peers[k]));
        }
    }
#pragma unroll
    for (int k = 0; k < SCATTER_PER_THREAD; ++k) {
        if (c[k] >= 0) {
            const int32_t b = __shfl_sync(peers[k], base[k], __ffs(peers[k]) - 1);
            perm[b + __popc(peers[k] & ((1u << lane) - 1u))] = i0 + 256 * k;
        }
    }
}

cudaError_t launchScatterIndex(const int32_t* cell, int32_t n, int32_t* cursor, int32_t* perm, cudaStream_t s) {
    if (n <= 0) return cudaSuccess;
    const int per = 256 * SCATTER_PER_THREAD;
    scatterIndexKernel<<<(n + per - 1) / per, 256, 0, s>>>(cell, n, cursor, perm);
    return cudaGetLastError();
}

// Order the indices of every cell ascending (= cloud-list order).  One warp per cell:
// cells of <= 32 parcels rank by shuffles, up to SEG_SMEM by a shared-memory rank sort.
namespace {
constexpr int SEG_WARPS = 8;
constexpr int SEG_SMEM = 1024;  // per warp
constexpr int SEG_CELLS = 4;    // cells a warp works on at a time
}

__global__ void __launch_bounds__(SEG_WARPS * 32) segmentSortKernel(const int32_t* __restrict__ cellOffset, int32_t nCells,
                                                                    int32_t* __restrict__ perm, DevCounters* counters, int32_t* __restrict__ bigList) {
    __shared__ __align__(16) int32_t sm[SEG_WARPS][SEG_SMEM + 32 * SEG_CELLS];   // rank-sort buffer of a medium cell | one row of 32 per small cell
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int32_t nWarps = gridDim.x * SEG_WARPS;
    // a warp takes SEG_CELLS consecutive cells at a time: their offsets are one read and their index lists are requested together, so a
    // cell costs a quarter of the two dependent round trips it would cost alone
    for (int32_t c0 = (blockIdx.x * SEG_WARPS + w) * SEG_CELLS; c0 < nCells; c0 += nWarps * SEG_CELLS) {
        const int32_t off = (lane <= SEG_CELLS && c0 + lane <= nCells) ? cellOffset[c0 + lane] : 0;
        int32_t b[SEG_CELLS], n[SEG_CELLS], v[SEG_CELLS];
#pragma unroll
        for (int k = 0; k < SEG_CELLS; ++k) {
            b[k] = __shfl_sync(0xffffffffu, off, k);
            const int32_t e = __shfl_sync(0xffffffffu, off, k + 1);
            n[k] = c0 + k < nCells ? e - b[k] : 0;
        }
#pragma unroll
        for (int k = 0; k < SEG_CELLS; ++k) v[k] = (n[k] > 1 && n[k] <= 32 && lane < n[k]) ? perm[b[k] + lane] : 0x7fffffff;
        // every lane compares its index with the 32 of its cell: eight broadcast 16-byte reads of shared memory instead of 32 shuffles
#pragma unroll
        for (int k = 0; k < SEG_CELLS; ++k) sm[w][SEG_SMEM + 32 * k + lane] = v[k];
        __syncwarp();
#pragma unroll
        for (int k = 0; k < SEG_CELLS; ++k) {
            if (n[k] <= 1) continue;
            if (n[k] <= 32) {
                int rank = 0;
                const int4* const row = reinterpret_cast<const int4*>(&sm[w][SEG_SMEM + 32 * k]);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int4 o = row[j];
                    rank += (o.x < v[k] ? 1 : 0) + (o.y < v[k] ? 1 : 0) + (o.z < v[k] ? 1 : 0) + (o.w < v[k] ? 1 : 0);
                }
                if (lane < n[k]) perm[b[k] + rank] = v[k];
            } else if (n[k] <= SEG_SMEM) {
                for (int i = lane; i < n[k]; i += 32) sm[w][i] = perm[b[k] + i];
                __syncwarp();
                for (int i = lane; i < n[k]; i += 32) {
                    const int32_t x = sm[w][i];
                    int rank = 0;
                    for (int j = 0; j < n[k]; ++j) rank += (sm[w][j] < x) ? 1 : 0;
                    perm[b[k] + rank] = x;
                }
                __syncwarp();
            } else {
                // very large cells (a heat bath in one cell: 1e5 parcels) go to bigSegmentSortKernel, one block each
                if (lane == 0) bigList[atomicAdd(&counters->bigSortCells, 1)] = c0 + k;
            }
        }
        __syncwarp();   // the rows are rewritten by the next group of cells
    }
}

// The index lists of the cells segmentSortKernel left: bitonic network with ascending comparators only (flip step + half cleaners),
// so that a length that is not a power of two needs no padding -- a partner beyond the end counts as +infinity and never swaps.
__global__ void __launch_bounds__(1024) bigSegmentSortKernel(const int32_t* __restrict__ cellOffset, const int32_t* __restrict__ bigList,
                                                             int32_t* __restrict__ giantList, int32_t* __restrict__ perm, DevCounters* counters) {
    const int32_t nBig = counters->bigSortCells;
    for (int32_t ib = blockIdx.x; ib < nBig; ib += gridDim.x) {
        const int32_t c = bigList[ib];
        int32_t* const v = perm + cellOffset[c];
        const int32_t n = cellOffset[c + 1] - cellOffset[c];
        if (n > GIANT_SORT) {   // 171 stages over global memory for 2^18 indices: left to giantSortKernel
            if (threadIdx.x == 0) giantList[atomicAdd(&counters->giantSortCells, 1)] = c;   // at most 2^31 / GIANT_SORT entries
            continue;
        }
        for (int32_t k = 2; (k >> 1) < n; k <<= 1) {
            for (int32_t i = threadIdx.x; i < n; i += blockDim.x) {
                const int32_t l = i ^ (k - 1);
                if (l > i && l < n) { const int32_t x = v[i], y = v[l]; if (x > y) { v[i] = y; v[l] = x; } }
            }
            __syncthreads();
            for (int32_t j = k >> 2; j > 0; j >>= 1) {
                for (int32_t i = threadIdx.x; i < n; i += blockDim.x) {
                    const int32_t l = i ^ j;
                    if (l > i && l < n) { const int32_t x = v[i], y = v[l]; if (x > y) { v[i] = y; v[l] = x; } }
                }
                __syncthreads();
            }
        }
    }
}

// A cell of more than GIANT_SORT parcels: its indices (distinct cloud positions) are marked in a bitmap over the range they span, and the
// bitmap is read back in ascending order -- O(range / 32 + n) instead of the bitonic network's O(n log^2 n) passes over global memory.
__global__ void __launch_bounds__(1024) giantSortKernel(const int32_t* __restrict__ cellOffset, const int32_t* __restrict__ giantList,
                                                        int32_t* __restrict__ perm, uint32_t* __restrict__ bitmap, int64_t wordsPerBlock, int32_t nGiant) {
    __shared__ int32_t sMin, sMax;
    __shared__ int32_t part[1024];
    uint32_t* const bm = bitmap + size_t(blockIdx.x) * size_t(wordsPerBlock);
    const int tid = threadIdx.x;
    for (int32_t g = blockIdx.x; g < nGiant; g += gridDim.x) {
        const int32_t c = giantList[g];
        int32_t* const v = perm + cellOffset[c];
        const int32_t n = cellOffset[c + 1] - cellOffset[c];
        if (tid == 0) { sMin = 0x7fffffff; sMax = -1; }
        __syncthreads();
        int32_t lo = 0x7fffffff, hi = -1;
        for (int32_t k = tid; k < n; k += 1024) { const int32_t x = v[k]; lo = min(lo, x); hi = max(hi, x); }
        atomicMin(&sMin, lo); atomicMax(&sMax, hi);
        __syncthreads();
        const int32_t base = sMin & ~31;
        const int32_t W = (sMax - base) / 32 + 1;
        for (int32_t w = tid; w < W; w += 1024) bm[w] = 0u;
        __syncthreads();
        for (int32_t k = tid; k < n; k += 1024) { const int32_t x = v[k] - base; atomicOr(&bm[x >> 5], 1u << (x & 31)); }
        __syncthreads();
        const int32_t chunk = (W + 1023) / 1024;
        const int32_t w0 = min(W, tid * chunk), w1 = min(W, w0 + chunk);
        int32_t cnt = 0;
        for (int32_t w = w0; w < w1; ++w) cnt += __popc(bm[w]);
        part[tid] = cnt;
        __syncthreads();
        for (int o = 1; o < 1024; o <<= 1) {   // inclusive scan of the chunk counts
            const int32_t add = tid >= o ? part[tid - o] : 0;
            __syncthreads();
            part[tid] += add;
            __syncthreads();
        }
        int32_t off = part[tid] - cnt;
        for (int32_t w = w0; w < w1; ++w) {
            uint32_t word = bm[w];
            while (word) { const int b = __ffs(word) - 1; word &= word - 1; v[off++] = base + w * 32 + b; }
        }
        __syncthreads();
    }
}

cudaError_t launchGiantSort(const int32_t* cellOffset, int32_t* perm, const int32_t* giantList, int32_t nGiant, uint32_t* bitmap,
                            int64_t wordsPerBlock, cudaStream_t s) {
    if (nGiant <= 0) return cudaSuccess;
    giantSortKernel<<<std::min(nGiant, GIANT_SORT_BLOCKS), 1024, 0, s>>>(cellOffset, giantList, perm, bitmap, wordsPerBlock, nGiant);
    return cudaGetLastError();
}

// bigList: nCells ints of scratch (the scatter cursors are free by now)
cudaError_t launchSegmentSort(const int32_t* cellOffset, int32_t nCells, int32_t* perm, DevCounters* c, int32_t* bigList, int32_t* giantList, cudaStream_t s) {
    int grid = (nCells + SEG_WARPS * SEG_CELLS - 1) / (SEG_WARPS * SEG_CELLS);
    if (grid > 148 * 16) grid = 148 * 16;
    if (grid < 1) grid = 1;
    cudaMemsetAsync(&c->bigSortCells, 0, 2 * sizeof(int32_t), s);   // bigSortCells, giantSortCells
    segmentSortKernel<<<grid, SEG_WARPS * 32, 0, s>>>(cellOffset, nCells, perm, c, bigList);
    bigSegmentSortKernel<<<std::min(nCells, 296), 1024, 0, s>>>(cellOffset, bigList, giantList, perm, c);
    return cudaGetLastError();
}

// cell id of output slot k: binary search in the CSR offsets is avoided by writing the cell
// array from a per-cell fill kernel instead.
__global__ void fillCellKernel(const int32_t* __restrict__ cellOffset, int32_t nCells, int32_t* __restrict__ cellOut) {
    const int lane = threadIdx.x & 31;
    const int32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int32_t nWarps = (gridDim.x * blockDim.x) >> 5;
    for (int32_t c = warp; c < nCells; c += nWarps) {
        const int32_t b = cellOffset[c], e = cellOffset[c + 1];
        for (int32_t k = b + lane; k < e; k += 32) cellOut[k] = c;
    }
}

__global__ void __launch_bounds__(256) gatherKernel(const __grid_constant__ ParcelArrays src, const __grid_constant__ ParcelArrays dst, const int32_t* __restrict__ perm,
                                                    const double* __restrict__ cellCentres, uint8_t* __restrict__ octKey, int32_t nOut,
                                                    int32_t nModes, int hasInternal) {
    const int32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nOut) return;
    const int32_t i = perm[k];
    const double x = src.px[i], y = src.py[i], z = src.pz[i];
    const int32_t cell = src.cell[i];
    dst.px[k] = x; dst.py[k] = y; dst.pz[k] = z;
    dst.ux[k] = src.ux[i]; dst.uy[k] = src.uy[i]; dst.uz[k] = src.uz[i];
    dst.cell[k] = cell;
    // the Cartesian sub-cell of noTimeCounter (noTimeCounter.C:124-138) while position and cell are in registers:
    // pos(relPos.x()) + 2*pos(relPos.y()) + 4*pos(relPos.z())
    {
        const double* cc = cellCentres + 3 * size_t(cell);
        const double rx = x - cc[0], ry = y - cc[1], rz = z - cc[2];
        octKey[k] = uint8_t((rx >= 0 ? 1 : 0) + 2 * (ry >= 0 ? 1 : 0) + 4 * (rz >= 0 ? 1 : 0));
    }
    dst.tet[k] = src.tet[i];
    dst.origId[k] = src.origId[i];
    dst.typeId[k] = src.typeId[i];
    if (hasInternal) {
        dst.erot[k] = src.erot[i];
        if (nModes > 0) dst.vib[0][k] = src.vib[0][i];
        if (nModes > 1) dst.vib[1][k] = src.vib[1][i];
        if (nModes > 2) dst.vib[2][k] = src.vib[2][i];
        dst.elevel[k] = src.elevel[i];
    }
    if (src.cls) dst.cls[k] = src.cls[i];
    if (src.origProc) dst.origProc[k] = src.origProc[i];
    if (src.rwf) dst.rwf[k] = src.rwf[i];
}

cudaError_t launchGather(const ParcelArrays& src, const ParcelArrays& dst, const int32_t* perm, const double* cellCentres,
                         uint8_t* octKey, int32_t nOut, int32_t nModes, bool hasInternal, cudaStream_t s) {
    if (nOut <= 0) return cudaSuccess;
    gatherKernel<<<(nOut + 255) / 256, 256, 0, s>>>(src, dst, perm, cellCentres, octKey, nOut, nModes, hasInternal ? 1 : 0);
    return cudaGetLastError();
}

}  // namespace dsmc
