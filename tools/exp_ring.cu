// exp_ring.cu — what makes one iteration of the producer / consumer ring cost ~320 cycles?  Variants of the
// skeleton (no TMA, no MMA): barrier spacing, single-lane loops, test_wait polling, and the full-tap stage
// (2 boxes + 8 MMAs per barrier).  tools, not product.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I neurosymbolic-mcts_b200/csrc tools/exp_ring.cu -o tools/bin/exp_ring -lcuda
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "ptx.cuh"
using namespace cb;

struct P { int stages, iters, spacing /* u64 units between barriers */, lane0, mode, nmma, boxes, n; const CUtensorMap* map; long long* out; };

__device__ __forceinline__ bool test_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(ptx::smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void wait_v(uint64_t* bar, uint32_t parity, int mode) {
    if (mode == 1) { while (!test_wait(bar, parity)) {} }
    else if (mode == 2) { while (!ptx::mbar_try_wait_hint(bar, parity, 100000u)) {} }
    else ptx::mbar_wait(bar, parity);
}

__global__ void __launch_bounds__(128, 1) ring_kernel(const P p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* ring = smem;                       // 8 x 16 KB
    uint8_t* bbuf = ring + 8 * 16384;           // 32 KB
    uint64_t* bars = reinterpret_cast<uint64_t*>(bbuf + 32768);   // up to 8 KB of barrier space
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 1000);
    uint64_t* w_full = bars;
    uint64_t* w_empty = bars + 16 * p.spacing;
    uint64_t* done = bars + 40 * p.spacing;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 32768 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(bbuf)[i] = 0;
    if (threadIdx.x == 0) {
        for (int s = 0; s < p.stages; ++s) { ptx::mbar_init(&w_full[s * p.spacing], 1); ptx::mbar_init(&w_empty[s * p.spacing], 1); }
        ptx::mbar_init(done, 1);
        ptx::fence_barrier_init();
    }
    if (warp == 1) ptx::tmem_alloc<256>(tmem_ptr);
    ptx::fence_proxy_async_smem();
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem = *tmem_ptr;
    if (warp == 0) {
        int ws = 0; uint32_t ph = 0; int row = 0;
        if (p.lane0) {
            if (lane == 0)
                for (int it = 0; it < p.iters; ++it) {
                    wait_v(&w_empty[ws * p.spacing], ph ^ 1, p.mode);
                    if (p.boxes) {
                        ptx::mbar_arrive_expect_tx(&w_full[ws * p.spacing], 16384 * p.boxes);
                        for (int b = 0; b < p.boxes; ++b) { ptx::tma_load_2d(ring + ((ws * p.boxes + b) & 7) * 16384, p.map, &w_full[ws * p.spacing], 0, row); row = (row + 128) & 32767; }
                    } else ptx::mbar_arrive(&w_full[ws * p.spacing]);
                    if (++ws == p.stages) { ws = 0; ph ^= 1; }
                }
        } else {
            for (int it = 0; it < p.iters; ++it) {
                wait_v(&w_empty[ws * p.spacing], ph ^ 1, p.mode);
                if (ptx::elect_one()) {
                    if (p.boxes) {
                        ptx::mbar_arrive_expect_tx(&w_full[ws * p.spacing], 16384 * p.boxes);
                        for (int b = 0; b < p.boxes; ++b) { ptx::tma_load_2d(ring + ((ws * p.boxes + b) & 7) * 16384, p.map, &w_full[ws * p.spacing], 0, row); row = (row + 128) & 32767; }
                    } else ptx::mbar_arrive(&w_full[ws * p.spacing]);
                }
                __syncwarp();
                if (++ws == p.stages) { ws = 0; ph ^= 1; }
            }
        }
    } else if (warp == 1) {
        int ws = 0; uint32_t ph = 0;
        const uint64_t a0 = ptx::umma_desc_sw128(ptx::smem_u32(ring)), b0 = ptx::umma_desc_sw128(ptx::smem_u32(bbuf));
        const uint32_t a_lo = (uint32_t)a0, a_hi = (uint32_t)(a0 >> 32), b_lo = (uint32_t)b0, b_hi = (uint32_t)(b0 >> 32);
        const uint32_t idesc = ptx::umma_idesc_f16(1, 128, p.n);
        long long t0 = clock64();
        if (p.lane0) {
            if (lane == 0)
                for (int it = 0; it < p.iters; ++it) {
                    wait_v(&w_full[ws * p.spacing], ph, p.mode);
                    if (p.nmma) {
                        ptx::tc_fence_after();
                        for (int k = 0; k < p.nmma; ++k) ptx::umma_f16_lohi(tmem, a_lo + ((ws * p.boxes) & 7) * 1024 + 2 * (k & 3), a_hi, b_lo + 2 * (k & 3), b_hi, idesc, 1u);
                        ptx::umma_commit(&w_empty[ws * p.spacing]);
                    } else ptx::mbar_arrive(&w_empty[ws * p.spacing]);
                    if (++ws == p.stages) { ws = 0; ph ^= 1; }
                }
        } else {
            for (int it = 0; it < p.iters; ++it) {
                wait_v(&w_full[ws * p.spacing], ph, p.mode);
                if (ptx::elect_one()) {
                    if (p.nmma) {
                        for (int k = 0; k < p.nmma; ++k) ptx::umma_f16_lohi(tmem, a_lo + ((ws * p.boxes) & 7) * 1024 + 2 * (k & 3), a_hi, b_lo + 2 * (k & 3), b_hi, idesc, 1u);
                        ptx::umma_commit(&w_empty[ws * p.spacing]);
                    } else ptx::mbar_arrive(&w_empty[ws * p.spacing]);
                }
                __syncwarp();
                if (++ws == p.stages) { ws = 0; ph ^= 1; }
            }
        }
        __syncwarp();
        if (ptx::elect_one()) ptx::umma_commit(done);
        __syncwarp();
        ptx::mbar_wait(done, 0);
        long long t1 = clock64();
        if (lane == 0) p.out[blockIdx.x] = t1 - t0;
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) { ptx::tc_fence_after(); ptx::tmem_dealloc<256>(tmem); }
}

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main() {
    void* fn = nullptr; cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    void* w = nullptr; CK(cudaMalloc(&w, 32768 * 128)); CK(cudaMemset(w, 0, 32768 * 128));
    CUtensorMap hm; const cuuint64_t dims[2] = {64, 32768}; const cuuint64_t str[1] = {128}; const cuuint32_t box[2] = {64, 128}; const cuuint32_t es[2] = {1, 1};
    if (((EncodeTiledFn)fn)(&hm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, w, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) return 1;
    CUtensorMap* dm; CK(cudaMalloc(&dm, sizeof(hm))); CK(cudaMemcpy(dm, &hm, sizeof(hm), cudaMemcpyHostToDevice));
    long long* d; CK(cudaMalloc(&d, 148 * 8));
    const int smem = 1024 + 8 * 16384 + 32768 + 8192 + 64;
    CK(cudaFuncSetAttribute(ring_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const int iters = 2048;
    auto run = [&](const char* name, int stages, int spacing, int lane0, int mode, int nmma, int boxes, int n) {
        P p{stages, iters, spacing, lane0, mode, nmma, boxes, n, dm, d};
        long long c = 0;
        for (int rep = 0; rep < 2; ++rep) { ring_kernel<<<1, 128, smem>>>(p); CK(cudaDeviceSynchronize()); CK(cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost)); }
        printf("%-64s stages %d: %7.1f cycles / iteration\n", name, stages, (double)c / iters);
    };
    for (int stages : {2, 4, 8}) {
        run("skeleton: warp loops, adjacent barriers, try_wait", stages, 1, 0, 0, 0, 0, 64);
        run("skeleton: warp loops, barriers 128 B apart", stages, 16, 0, 0, 0, 0, 64);
        run("skeleton: lane-0 loops, adjacent barriers", stages, 1, 1, 0, 0, 0, 64);
        run("skeleton: lane-0 loops, barriers 128 B apart", stages, 16, 1, 0, 0, 0, 64);
        run("skeleton: lane-0 loops, test_wait poll", stages, 1, 1, 1, 0, 0, 64);
        run("skeleton: lane-0 loops, try_wait with a suspend hint", stages, 1, 1, 2, 0, 0, 64);
    }
    for (int n : {64, 128, 256}) {
        char nm[128];
        snprintf(nm, sizeof nm, "1 box + 4 MMA N=%d, warp loops", n);   run(nm, 8, 1, 0, 0, 4, 1, n);
        snprintf(nm, sizeof nm, "1 box + 4 MMA N=%d, lane-0 loops", n); run(nm, 8, 1, 1, 0, 4, 1, n);
        snprintf(nm, sizeof nm, "2 boxes + 8 MMA N=%d, warp loops", n);  run(nm, 4, 1, 0, 0, 8, 2, n);
        snprintf(nm, sizeof nm, "2 boxes + 8 MMA N=%d, lane-0 loops", n); run(nm, 4, 1, 1, 0, 8, 2, n);
        snprintf(nm, sizeof nm, "4 boxes + 16 MMA N=%d, lane-0 loops", n); run(nm, 2, 1, 1, 0, 16, 4, n);
    }
    return 0;
}
