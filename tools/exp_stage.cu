// exp_stage.cu — the product's weight-stage loop in isolation: producer warp = TMA ring (expect_tx + 1 or 2 boxes of
// 16 KB per barrier), issuer warp = wait(w_full) + NM MMAs (uniform-register form, compile-time unrolled, M = 128,
// N template) + tcgen05.commit(w_empty).  Cycles per 16 KB of weights (= 4 MMAs) for N = 64 / 128 / 192 / 256.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I neurosymbolic-mcts_b200/csrc tools/exp_stage.cu -o tools/bin/exp_stage -lcuda
#include <cstdio>
#include <cstdlib>
#include "ptx.cuh"
using namespace cb;

struct P { const CUtensorMap* map; int iters; long long* out; };

template <int N, int BOXES, int STAGES, int SKEW, int M = 128>
__global__ void __launch_bounds__(128, 1) stage_kernel(const P p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* ring = smem;                                  // STAGES x BOXES x 16 KB  (<= 128 KB)
    uint8_t* bbuf = ring + 8 * 16384;                      // 64 KB B operand
    uint64_t* w_full = reinterpret_cast<uint64_t*>(bbuf + 65536);
    uint64_t* w_empty = w_full + 8;
    uint64_t* done = w_empty + 8;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(done + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 65536 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(bbuf)[i] = 0;
    if (threadIdx.x == 0) {
        for (int s = 0; s < 8; ++s) { ptx::mbar_init(&w_full[s], 1); ptx::mbar_init(&w_empty[s], 1); }
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
        for (int it = 0; it < p.iters; ++it) {
            ptx::mbar_wait(&w_empty[ws], ph ^ 1);
            if (ptx::elect_one()) {
                ptx::mbar_arrive_expect_tx(&w_full[ws], M * 128 * BOXES);
#pragma unroll
                for (int b = 0; b < BOXES; ++b) { ptx::tma_load_2d(ring + (ws * BOXES + b) * 16384, M == 64 ? p.map + 1 : p.map, &w_full[ws], 0, row); row = (row + M) & 32767; }
            }
            __syncwarp();
            if (++ws == STAGES) { ws = 0; ph ^= 1; }
        }
    } else if (warp == 1) {
        int ws = 0; uint32_t ph = 0;
        const uint64_t a0 = ptx::umma_desc_sw128(ptx::smem_u32(ring)), b0 = ptx::umma_desc_sw128(ptx::smem_u32(bbuf));
        const uint32_t a_lo0 = (uint32_t)a0, a_hi = (uint32_t)(a0 >> 32), b_lo = (uint32_t)b0, b_hi = (uint32_t)(b0 >> 32);
        constexpr uint32_t idesc = ptx::umma_idesc_f16(1, M, N);
        long long t0 = clock64();
        if constexpr (SKEW == 0) {
        for (int it = 0; it < p.iters; ++it) {
            ptx::mbar_wait(&w_full[ws], ph);
            if (ptx::elect_one()) {
                const uint32_t a_lo = a_lo0 + ws * BOXES * 1024;
#pragma unroll
                for (int k = 0; k < 4 * BOXES; ++k) ptx::umma_f16_lohi(tmem, a_lo + (k >> 2) * 1024 + 2 * (k & 3), a_hi, b_lo + 2 * (k & 3), b_hi, idesc, 1u);
                ptx::umma_commit(&w_empty[ws]);
            }
            __syncwarp();
            if (++ws == STAGES) { ws = 0; ph ^= 1; }
        }
        } else {
        // skewed: the wait for stage s + 1 sits before the last SKEW MMAs of stage s, so the pipe has work during the wait
        ptx::mbar_wait(&w_full[0], 0);
        for (int it = 0; it < p.iters; ++it) {
            int wn = ws + 1; uint32_t phn = ph; if (wn == STAGES) { wn = 0; phn ^= 1; }
            const uint32_t a_lo = a_lo0 + ws * BOXES * 1024;
            if (ptx::elect_one()) {
#pragma unroll
                for (int k = 0; k < 4 * BOXES - SKEW; ++k) ptx::umma_f16_lohi(tmem, a_lo + (k >> 2) * 1024 + 2 * (k & 3), a_hi, b_lo + 2 * (k & 3), b_hi, idesc, 1u);
            }
            __syncwarp();
            if (it + 1 < p.iters) ptx::mbar_wait(&w_full[wn], phn);
            if (ptx::elect_one()) {
#pragma unroll
                for (int k = 4 * BOXES - SKEW; k < 4 * BOXES; ++k) ptx::umma_f16_lohi(tmem, a_lo + (k >> 2) * 1024 + 2 * (k & 3), a_hi, b_lo + 2 * (k & 3), b_hi, idesc, 1u);
                ptx::umma_commit(&w_empty[ws]);
            }
            __syncwarp();
            ws = wn; ph = phn;
        }
        }
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
static P g_p; static long long* g_d;
template <int N, int BOXES, int STAGES, int SKEW = 0, int M = 128>
void run() {
    const int smem = 1024 + 8 * 16384 + 65536 + 256;
    CK(cudaFuncSetAttribute(stage_kernel<N, BOXES, STAGES, SKEW, M>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    P p = g_p; p.iters = 2048 / BOXES;
    long long c = 0;
    for (int rep = 0; rep < 2; ++rep) { stage_kernel<N, BOXES, STAGES, SKEW, M><<<1, 128, smem>>>(p); CK(cudaDeviceSynchronize()); CK(cudaMemcpy(&c, g_d, 8, cudaMemcpyDeviceToHost)); }
    printf("M=%3d N=%3d, skew %d, %d x 16 KB + %2d MMAs per barrier, %d slots: %7.1f cycles per box of M x 64 weights (MMA floor %d, smem floor %d)\n", M, N, SKEW, BOXES, 4 * BOXES, STAGES,
           (double)c / p.iters / BOXES, 4 * (N / 2), 4 * (M + N) * 32 / 128);
}
int main() {
    void* fn = nullptr; cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    void* w = nullptr; CK(cudaMalloc(&w, 32768 * 128)); CK(cudaMemset(w, 0, 32768 * 128));
    CUtensorMap hm[2]; const cuuint64_t dims[2] = {64, 32768}; const cuuint64_t str[1] = {128}; const cuuint32_t es[2] = {1, 1};
    for (int i = 0; i < 2; ++i) {
        const cuuint32_t box[2] = {64, (cuuint32_t)(i ? 64 : 128)};
        if (((EncodeTiledFn)fn)(&hm[i], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, w, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) return 1;
    }
    CUtensorMap* dm; CK(cudaMalloc(&dm, sizeof(hm))); CK(cudaMemcpy(dm, hm, sizeof(hm), cudaMemcpyHostToDevice));
    CK(cudaMalloc(&g_d, 148 * 8));
    g_p = P{dm, 0, g_d};
    run<64, 1, 4>(); run<64, 2, 4>(); run<64, 4, 2>();
    run<128, 1, 4>(); run<128, 2, 4>(); run<256, 1, 4>();
    // M = 64: a CTA pair splitting the output channels (each CTA streams half of every weight tile)
    run<64, 1, 4, 0, 64>(); run<64, 2, 4, 0, 64>(); run<64, 4, 2, 0, 64>();
    run<128, 1, 4, 0, 64>(); run<128, 2, 4, 0, 64>(); run<128, 4, 2, 0, 64>();
    run<256, 2, 4, 0, 64>();
    return 0;
}
