// exp_ring2.cu — follow-up to exp_ring.cu: where do the ~330 cycles of a ring iteration come from?
//  A  one warp, self-ring: arrive(full[s+1]); wait(full[s])            (phase flips, no second warp)
//  B  two warps, each its own self-ring                                  (cross-warp interference?)
//  C  ring, producer delayed by `delay` ALU cycles per iteration         (consumer always finds the barrier pending)
//  D  ring, consumer delayed                                             (producer always pending, consumer never waits)
//  E  ring where the consumer polls a shared-memory counter published by the producer (no mbarrier on that side)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I neurosymbolic-mcts_b200/csrc tools/exp_ring2.cu -o tools/bin/exp_ring2
#include <cstdio>
#include "ptx.cuh"
using namespace cb;

__device__ __forceinline__ void spin(int cycles) { const long long t = clock64(); while (clock64() - t < cycles) {} }

__global__ void __launch_bounds__(64, 1) k(int mode, int iters, int stages, int delay, long long* out) {
    __shared__ __align__(16) uint64_t full[16], empty[16], full2[16];
    __shared__ volatile uint32_t cnt_full, cnt_empty;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < 16; ++s) { ptx::mbar_init(&full[s], 1); ptx::mbar_init(&empty[s], 1); ptx::mbar_init(&full2[s], 1); }
        cnt_full = 0; cnt_empty = 0;
        ptx::fence_barrier_init();
    }
    __syncthreads();
    long long t0 = clock64();
    if (mode == 0 || mode == 1) {          // self-ring(s)
        if (warp == 0 || mode == 1) {
            uint64_t* b = warp ? full2 : full;
            if (lane == 0) {
                int s = 0; uint32_t ph = 0;
                ptx::mbar_arrive(&b[0]);
                for (int i = 0; i < iters; ++i) {
                    const int s1 = s + 1 == stages ? 0 : s + 1;
                    ptx::mbar_arrive(&b[s1]);
                    ptx::mbar_wait(&b[s], ph);
                    s = s1; if (s == 0) ph ^= 1;
                }
            }
        }
    } else if (mode == 2 || mode == 3 || mode == 4) {
        if (warp == 0) {                   // producer
            if (lane == 0) {
                int s = 0; uint32_t ph = 0;
                for (int i = 0; i < iters; ++i) {
                    ptx::mbar_wait(&empty[s], ph ^ 1);
                    if (mode == 2) spin(delay);
                    ptx::mbar_arrive(&full[s]);
                    if (++s == stages) { s = 0; ph ^= 1; }
                }
            }
        } else {
            if (lane == 0) {
                int s = 0; uint32_t ph = 0;
                for (int i = 0; i < iters; ++i) {
                    ptx::mbar_wait(&full[s], ph);
                    if (mode == 3) spin(delay);
                    ptx::mbar_arrive(&empty[s]);
                    if (++s == stages) { s = 0; ph ^= 1; }
                }
            }
        }
    } else if (mode == 5) {                // both directions through shared-memory counters
        if (warp == 0) {
            if (lane == 0) {
                for (int i = 0; i < iters; ++i) {
                    uint32_t e; do { asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(e) : "r"(ptx::smem_u32((const void*)&cnt_empty)) : "memory"); } while ((int)(i - e) >= stages);
                    asm volatile("st.release.cta.shared.u32 [%0], %1;" :: "r"(ptx::smem_u32((const void*)&cnt_full)), "r"(i + 1) : "memory");
                }
            }
        } else {
            if (lane == 0) {
                for (int i = 0; i < iters; ++i) {
                    uint32_t f; do { asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(f) : "r"(ptx::smem_u32((const void*)&cnt_full)) : "memory"); } while ((int)(f - i) <= 0);
                    asm volatile("st.release.cta.shared.u32 [%0], %1;" :: "r"(ptx::smem_u32((const void*)&cnt_empty)), "r"(i + 1) : "memory");
                }
            }
        }
    } else if (mode == 6) {                // producer side mbarrier (as TMA needs), consumer side: one wait covers a batch of `delay` stages
        if (warp == 0) {
            if (lane == 0) {
                int s = 0; uint32_t ph = 0;
                for (int i = 0; i < iters; ++i) {
                    ptx::mbar_wait(&empty[s], ph ^ 1);
                    ptx::mbar_arrive(&full[s]);
                    if (++s == stages) { s = 0; ph ^= 1; }
                }
            }
        } else {
            if (lane == 0) {
                int s = 0; uint32_t ph = 0;
                for (int i = 0; i < iters; ++i) {
                    if (!ptx::mbar_try_wait(&full[s], ph)) ptx::mbar_wait(&full[s], ph);
                    ptx::mbar_arrive(&empty[s]);
                    if (++s == stages) { s = 0; ph ^= 1; }
                }
            }
        }
    }
    __syncwarp();
    long long t1 = clock64();
    if (lane == 0) out[warp] = t1 - t0;
}

int main() {
    long long* d; cudaMalloc(&d, 16);
    const int iters = 4096;
    auto run = [&](const char* name, int mode, int stages, int delay) {
        long long c[2] = {0, 0};
        for (int rep = 0; rep < 2; ++rep) { k<<<1, 64>>>(mode, iters, stages, delay, d); cudaError_t e = cudaDeviceSynchronize(); if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); exit(1); } cudaMemcpy(c, d, 16, cudaMemcpyDeviceToHost); }
        printf("%-58s stages %d delay %4d: %7.1f / %7.1f cycles per iteration (warp 0 / 1)\n", name, stages, delay, (double)c[0] / iters, (double)c[1] / iters);
    };
    for (int st : {2, 8}) {
        run("A self-ring, one warp", 0, st, 0);
        run("B self-rings, two warps", 1, st, 0);
        run("ring, lane-0 loops", 4, st, 0);
        for (int dl : {100, 200, 400, 800}) run("C ring, producer delayed", 2, st, dl);
        for (int dl : {100, 200, 400, 800}) run("D ring, consumer delayed", 3, st, dl);
        run("E ring through shared-memory counters (no mbarrier)", 5, st, 0);
    }
    return 0;
}
