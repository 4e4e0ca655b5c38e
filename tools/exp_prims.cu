// exp_prims.cu — cost of the synchronisation primitives of the forward kernel in isolation (one warp, one CTA):
// cycles per iteration of a loop that does nothing else.  tools, not product.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I neurosymbolic-mcts_b200/csrc tools/exp_prims.cu -o tools/bin/exp_prims
#include <cstdio>
#include <cstdlib>
#include "ptx.cuh"
using namespace cb;

// TRYWAIT on a complete phase from nw warps at once: is the unit shared by the warps of an SM?
__global__ void __launch_bounds__(512, 1) trywait_kernel(int iters, int poll, long long* out) {
    __shared__ __align__(16) uint64_t bar;
    __shared__ volatile uint32_t flag;
    if (threadIdx.x == 0) { ptx::mbar_init(&bar, 1); ptx::fence_barrier_init(); ptx::mbar_arrive(&bar); flag = 1; }
    __syncthreads();
    long long t0 = clock64();
    uint32_t acc = 0;
    if (poll) { for (int i = 0; i < iters; ++i) { uint32_t v; asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(v) : "r"(ptx::smem_u32((const void*)&flag)) : "memory"); acc += v; } }
    else { for (int i = 0; i < iters; ++i) ptx::mbar_wait(&bar, 0); }
    long long t1 = clock64();
    if (threadIdx.x == 0) { out[0] = t1 - t0; out[1] = acc; }
}

__global__ void __launch_bounds__(64, 1) prims_kernel(int mode, int iters, long long* out) {
    __shared__ __align__(16) uint64_t bars[16];
    __shared__ uint32_t tmem_ptr;
    __shared__ __align__(1024) uint8_t buf[32768];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 32768 / 4; i += 64) reinterpret_cast<uint32_t*>(buf)[i] = 0;
    if (threadIdx.x == 0) {
        ptx::mbar_init(&bars[0], 1);          // completed once below: waits on parity 0 succeed at once
        ptx::mbar_init(&bars[1], (1u << 20) - 1);   // arrive target that never completes
        ptx::mbar_init(&bars[2], 1);          // ping
        ptx::mbar_init(&bars[3], 1);          // pong
        ptx::mbar_init(&bars[4], 1);
        ptx::fence_barrier_init();
        ptx::mbar_arrive(&bars[0]);
    }
    if (warp == 0) ptx::tmem_alloc<256>(&tmem_ptr);
    ptx::fence_proxy_async_smem();
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem = tmem_ptr;
    const uint64_t a0 = ptx::umma_desc_sw128(ptx::smem_u32(buf));
    const uint64_t b0 = ptx::umma_desc_sw128(ptx::smem_u32(buf + 16384));
    const uint32_t a_lo = (uint32_t)a0, a_hi = (uint32_t)(a0 >> 32), b_lo = (uint32_t)b0, b_hi = (uint32_t)(b0 >> 32);
    const uint32_t idesc = ptx::umma_idesc_f16(1, 128, 64);
    long long t0 = clock64();
    if (warp == 0) {
        switch (mode) {
        case 0: for (int i = 0; i < iters; ++i) ptx::mbar_wait(&bars[0], 0); break;                     // try_wait, complete
        case 1: for (int i = 0; i < iters; ++i) { if (ptx::elect_one()) ptx::mbar_arrive(&bars[1]); __syncwarp(); } break;
        case 2: for (int i = 0; i < iters; ++i) { ptx::mbar_wait(&bars[0], 0); if (ptx::elect_one()) ptx::mbar_arrive(&bars[1]); __syncwarp(); } break;
        case 3: for (int i = 0; i < iters; ++i) { if (lane == 0) ptx::mbar_arrive(&bars[1]); __syncwarp(); } break;
        case 4:   // ping-pong with warp 1: round trip of two plain arrives
            for (int i = 0; i < iters; ++i) {
                if (lane == 0) ptx::mbar_arrive(&bars[2]);
                ptx::mbar_wait(&bars[3], i & 1);
            }
            break;
        case 5:   // commit round trip: commit (nothing pending) -> wait on it
            for (int i = 0; i < iters; ++i) {
                if (ptx::elect_one()) ptx::umma_commit(&bars[4]);
                __syncwarp();
                ptx::mbar_wait(&bars[4], i & 1);
            }
            break;
        case 6:   // 4 MMAs (N = 64, uniform descriptors) + commit, waited every iteration: MMA latency + commit
            for (int i = 0; i < iters; ++i) {
                if (ptx::elect_one()) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) ptx::umma_f16_lohi(tmem, a_lo + 2 * k, a_hi, b_lo + 2 * k, b_hi, idesc, 1u);
                    ptx::umma_commit(&bars[4]);
                }
                __syncwarp();
                ptx::mbar_wait(&bars[4], i & 1);
            }
            break;
        case 7:   // issue only: 4 MMAs + commit to a barrier nobody waits on (count 2^20)
            for (int i = 0; i < iters; ++i) {
                if (ptx::elect_one()) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) ptx::umma_f16_lohi(tmem, a_lo + 2 * k, a_hi, b_lo + 2 * k, b_hi, idesc, 1u);
                    ptx::umma_commit(&bars[1]);
                }
                __syncwarp();
            }
            break;
        case 8:   // issue only: 16 MMAs per commit
            for (int i = 0; i < iters; ++i) {
                if (ptx::elect_one()) {
#pragma unroll
                    for (int k = 0; k < 16; ++k) ptx::umma_f16_lohi(tmem, a_lo + 2 * (k & 3), a_hi, b_lo + 2 * (k & 3), b_hi, idesc, 1u);
                    ptx::umma_commit(&bars[1]);
                }
                __syncwarp();
            }
            break;
        case 9:   // commits only, nobody waits
            for (int i = 0; i < iters; ++i) { if (ptx::elect_one()) ptx::umma_commit(&bars[1]); __syncwarp(); }
            break;
        case 10:  // tmem ld + wait
            for (int i = 0; i < iters; ++i) { uint32_t r[8]; ptx::tmem_ld_32x8(tmem, r); ptx::tmem_ld_wait(); if (r[0] == 0x12345u) out[1] = 1; }
            break;
        case 11:  // named barrier among 64 threads (both warps)
            break;
        }
    } else {
        if (mode == 4) {
            for (int i = 0; i < iters; ++i) {
                ptx::mbar_wait(&bars[2], i & 1);
                if (lane == 0) ptx::mbar_arrive(&bars[3]);
            }
        }
    }
    if (mode == 11) for (int i = 0; i < iters; ++i) ptx::named_bar_sync(1, 64);
    if (mode >= 6 && mode <= 9 && warp == 0) {   // drain
        if (ptx::elect_one()) ptx::umma_commit(&bars[2]);
        __syncwarp();
        ptx::mbar_wait(&bars[2], 0);
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) out[0] = t1 - t0;
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 0) { ptx::tc_fence_after(); ptx::tmem_dealloc<256>(tmem); }
}

int main() {
    long long* d; cudaMalloc(&d, 16);
    const char* names[12] = {"try_wait on a complete phase", "elect + arrive + syncwarp", "try_wait + elect + arrive + syncwarp", "lane0 arrive + syncwarp",
                             "ping-pong round trip (2 arrives, 2 waits)", "commit (nothing pending) -> wait", "4 MMA N=64 + commit -> wait",
                             "4 MMA N=64 + commit, issue only", "16 MMA N=64 + commit, issue only", "commit only, issue", "tmem ld x8 + wait", "named barrier 64 thr"};
    const int iters = 4096;
    for (int mode = 0; mode < 12; ++mode) {
        long long c = 0;
        for (int rep = 0; rep < 2; ++rep) {
            prims_kernel<<<1, 64>>>(mode, iters, d);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("mode %d: %s\n", mode, cudaGetErrorString(e)); return 1; }
            cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
        }
        printf("%-44s %7.1f cycles / iteration\n", names[mode], (double)c / iters);
    }
    for (int poll = 0; poll < 2; ++poll)
        for (int nw : {1, 2, 4, 8, 16}) {
            long long c = 0;
            for (int rep = 0; rep < 2; ++rep) { trywait_kernel<<<1, nw * 32>>>(iters, poll, d); cudaDeviceSynchronize(); cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost); }
            printf("%s from %2d warps at once: %7.1f cycles / iteration\n", poll ? "ld.acquire.cta.shared poll" : "try_wait on a complete phase", nw, (double)c / iters);
        }
    return 0;
}
