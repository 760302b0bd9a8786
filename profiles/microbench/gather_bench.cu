// gather_bench.cu -- how the sm_100a L1 data path charges the move kernel's record gathers.
// Every lane of a warp picks one of the 24 tet records of two adjacent hex cells (like cell-sorted parcels do) and reads
// the whole 224-byte record; the next pick depends on the loaded data (like the tet walk).  Variants differ only in the
// table layout and the load width.  Prints cycles per warp-visit per SM at 16 resident warps/SM (the move kernel's occupancy).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a gather_bench.cu -o gather_bench && ./gather_bench
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

constexpr int TETS = 12;          // per cell
constexpr int WORDS = 28;         // doubles per record
constexpr int ITERS = 64;

__device__ __forceinline__ void ld4(const double* p, double& a, double& b, double& c, double& d) {
    asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p));
}
__device__ __forceinline__ void ld2(const double* p, double& a, double& b) {
    asm volatile("ld.global.nc.v2.f64 {%0,%1}, [%2];" : "=d"(a), "=d"(b) : "l"(p));
}
__device__ __forceinline__ double ld1(const double* p) {
    double a;
    asm volatile("ld.global.nc.f64 %0, [%1];" : "=d"(a) : "l"(p));
    return a;
}

// mode 0: AoS records, 7 x LDG.256           addr = (cell*12 + j)*28 + 4k
// mode 1: cell-transposed sectors, LDG.256   addr = cell*336 + k*48 + j*4
// mode 2: cell-transposed words, 28 x LDG.64 addr = cell*336 + w*12 + j
// mode 3: cell-transposed pairs, 14 x LDG.128 addr = cell*336 + p*24 + j*2
// mode 4: AoS from shared memory (window of the block's cells), 14 x LDS.128, record stride 29 doubles... (16-byte aligned: 30)
// mode 5: AoS from shared memory, 28 x LDS.64, stride 29 doubles
template <int MODE>
__global__ void __launch_bounds__(128, 4) bench(const double* __restrict__ tab, int nCells, double* out, int activeMask) {
    extern __shared__ double sm[];
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int cellsPerWarp = 16;
    int cell0 = (warp * cellsPerWarp) % (nCells - cellsPerWarp - 2);
    constexpr int SSTR = MODE == 4 ? 30 : 29;
    if (MODE >= 4) {
        // stage the block's window: 4 warps x 16 cells... keep it small: 24 records per warp region
        const int w = threadIdx.x >> 5;
        for (int r = lane; r < 24 * WORDS; r += 32) {
            const int rec = r / WORDS, wd = r % WORDS;
            sm[(w * 24 + rec) * SSTR + wd] = tab[(size_t(cell0) * TETS + rec) * WORDS + wd];
        }
        __syncthreads();
    }
    unsigned state = lane * 2654435761u + warp * 40503u;
    double acc = 0.0;
    const bool on = (activeMask >> lane) & 1;
    int cellShift = 0;
    for (int it = 0; it < ITERS; ++it) {
        state = state * 1664525u + 1013904223u;
        const int pick = (state >> 8) % 24;             // one of the 24 records of two adjacent cells
        const int cell = cell0 + cellShift + pick / TETS, j = pick % TETS;
        double v[WORDS];
        if (on) {
            if (MODE == 0) {
                const double* R = tab + (size_t(cell) * TETS + j) * WORDS;
#pragma unroll
                for (int k = 0; k < 7; ++k) ld4(R + 4 * k, v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
            } else if (MODE == 1) {
                const double* R = tab + size_t(cell) * 336 + j * 4;
#pragma unroll
                for (int k = 0; k < 7; ++k) ld4(R + 48 * k, v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
            } else if (MODE == 2) {
                const double* R = tab + size_t(cell) * 336 + j;
#pragma unroll
                for (int w = 0; w < 28; ++w) v[w] = ld1(R + 12 * w);
            } else if (MODE == 3) {
                const double* R = tab + size_t(cell) * 336 + j * 2;
#pragma unroll
                for (int p = 0; p < 14; ++p) ld2(R + 24 * p, v[2 * p], v[2 * p + 1]);
            } else if (MODE == 4) {
                const int w = threadIdx.x >> 5;
                const double2* R = reinterpret_cast<const double2*>(sm + (w * 24 + pick) * SSTR);
#pragma unroll
                for (int p = 0; p < 14; ++p) { double2 t = R[p]; v[2 * p] = t.x; v[2 * p + 1] = t.y; }
            } else {
                const int w = threadIdx.x >> 5;
                const double* R = sm + (w * 24 + pick) * SSTR;
#pragma unroll
                for (int q = 0; q < 28; ++q) v[q] = R[q];
            }
            double s = 0.0;
#pragma unroll
            for (int q = 0; q < WORDS; ++q) s += v[q];
            acc += s;
            state += (unsigned)(__double_as_longlong(s) & 3);   // the next pick depends on the data
        }
        if (MODE < 4 && (it & 7) == 7) cellShift = (cellShift + 1) % (cellsPerWarp - 2);  // drift through the warp's cells
        __syncwarp();
    }
    if (acc == 1.2345) out[0] = acc;
}

template <int MODE>
void run(const char* name, const double* tab, int nCells, double* out, int activeMask) {
    int dev = 0; cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, dev));
    const int blocks = p.multiProcessorCount * 4 * 8;   // 8 waves of 4 blocks per SM
    const size_t smem = MODE >= 4 ? size_t(4 * 24 * 30) * sizeof(double) : 0;
    for (int rep = 0; rep < 2; ++rep) {
        cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
        cudaEventRecord(a);
        bench<MODE><<<blocks, 128, smem>>>(tab, nCells, out, activeMask);
        cudaEventRecord(b);
        CK(cudaEventSynchronize(b));
        float ms; cudaEventElapsedTime(&ms, a, b);
        if (rep == 1) {
            const double warpVisits = double(blocks) * 4 * ITERS;
            const double cyc = ms * 1e-3 * 1.965e9 * p.multiProcessorCount;   // SM-cycles
            printf("%-44s active %2d: %8.3f ms  %7.1f SM-cycles per warp-visit  (%6.1f per active lane)\n", name, __builtin_popcount(activeMask), ms,
                   cyc / warpVisits, cyc / warpVisits / __builtin_popcount(activeMask));
        }
    }
}

int main() {
    const int nCells = 1 << 20;   // 1 Mi cells x 2688 B = 2.8 GB table
    const size_t nD = size_t(nCells) * TETS * WORDS;
    double* tab; CK(cudaMalloc(&tab, nD * 8));
    CK(cudaMemset(tab, 0, nD * 8));
    double* out; CK(cudaMalloc(&out, 8));
    const int masks[2] = {int(0xffffffffu), 0x000fffff};
    for (int mask : masks) {
        run<0>("AoS 224 B records, 7 x LDG.256", tab, nCells, out, mask);
        run<1>("cell-transposed sectors, 7 x LDG.256", tab, nCells, out, mask);
        run<3>("cell-transposed pairs, 14 x LDG.128", tab, nCells, out, mask);
        run<2>("cell-transposed words, 28 x LDG.64", tab, nCells, out, mask);
        run<4>("shared-memory records, 14 x LDS.128", tab, nCells, out, mask);
        run<5>("shared-memory records, 28 x LDS.64", tab, nCells, out, mask);
    }
    return 0;
}
