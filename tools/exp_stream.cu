// exp_stream.cu — micro-experiments behind the small-batch forward (tools, not product):
// what does ONE weight stage cost a CTA when nothing else is in the way?
//   mode 0  skeleton: producer arrives on w_full without loading, consumer releases the slot with a plain
//           mbarrier.arrive                                     -> the two-warp ring hand-off itself
//   mode 1  skeleton, consumer releases with tcgen05.commit (no MMA pending)  -> what the commit adds
//   mode 2  TMA stream (16 KB stage: 128 rows x 64 ch, SWIZZLE_128B) + plain release -> L2 -> SM rate of one SM
//   mode 3  TMA stream + 4 MMAs (M = 128, N = n, K = 16) + commit release    -> the product's stage
//   mode 4  no ring at all: MMAs from resident operands + one commit per stage, the issuer alone
// for grid = 1 / 16 / 148 CTAs and ring depth 2 / 4 / 8 / 12.  Prints cycles per stage and B/clk per SM.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I neurosymbolic-mcts_b200/csrc tools/exp_stream.cu -o gpurun_out/exp_stream -lcuda
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda_bf16.h>
#include "ptx.cuh"

using namespace cb;

constexpr int kStageBytes = 16384;
constexpr int kMaxStages = 12;

struct Params {
    const CUtensorMap* map;   // [rows_total][64] bf16, box {64, rows_per_stage}
    int stages;               // ring depth
    int iters;                // stages to stream
    int mode;
    int n;                    // MMA N
    int rows_total;
    int rows_per_stage;       // 128 (16 KB) or 64 (8 KB)
    int boxes;                // TMA boxes per stage (mode 5: raw stream rate, stage = boxes x 16 KB, ring = stages / boxes slots)
    long long* cyc;           // [grid]
};

__global__ void __launch_bounds__(128, 1) stream_kernel(const Params p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* ring = smem;                                    // [stages][16 KB]
    uint8_t* bbuf = ring + kMaxStages * kStageBytes;         // 32 KB B operand (zeros)
    uint64_t* w_full = reinterpret_cast<uint64_t*>(bbuf + 32768);
    uint64_t* w_empty = w_full + kMaxStages;
    uint64_t* done = w_empty + kMaxStages;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(done + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 32768 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(bbuf)[i] = 0;
    if (threadIdx.x == 0) {
        for (int s = 0; s < kMaxStages; ++s) { ptx::mbar_init(&w_full[s], 1); ptx::mbar_init(&w_empty[s], 1); }
        ptx::mbar_init(done, 1);
        ptx::fence_barrier_init();
    }
    if (warp == 1) ptx::tmem_alloc<256>(tmem_ptr);
    ptx::fence_proxy_async_smem();
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    const int stage_bytes = p.rows_per_stage * 128;
    long long t0 = 0, t1 = 0;
    if (warp == 0) {
        if (p.mode != 4) {
            int ws = 0; uint32_t ph = 0;
            int row = (blockIdx.x * 1024) % p.rows_total;
            for (int it = 0; it < p.iters; ++it) {
                ptx::mbar_wait(&w_empty[ws], ph ^ 1);
                if (ptx::elect_one()) {
                    if (p.mode < 2) ptx::mbar_arrive(&w_full[ws]);
                    else if (p.mode == 5) {
                        ptx::mbar_arrive_expect_tx(&w_full[ws], stage_bytes * p.boxes);
                        for (int b = 0; b < p.boxes; ++b) {
                            ptx::tma_load_2d(ring + (ws * p.boxes + b) * kStageBytes, p.map, &w_full[ws], 0, row);
                            row += p.rows_per_stage; if (row >= p.rows_total) row = 0;
                        }
                    } else {
                        ptx::mbar_arrive_expect_tx(&w_full[ws], stage_bytes);
                        ptx::tma_load_2d(ring + ws * kStageBytes, p.map, &w_full[ws], 0, row);
                    }
                }
                __syncwarp();
                row += p.rows_per_stage; if (row >= p.rows_total) row = 0;
                if (++ws == p.stages) { ws = 0; ph ^= 1; }
            }
        }
    } else if (warp == 1) {
        int ws = 0; uint32_t ph = 0;
        const uint64_t a0 = ptx::umma_desc_sw128(ptx::smem_u32(ring));
        const uint64_t b0 = ptx::umma_desc_sw128(ptx::smem_u32(bbuf));
        const uint32_t idesc = ptx::umma_idesc_f16(1, 128, p.n);
        t0 = clock64();
        for (int it = 0; it < p.iters; ++it) {
            if (p.mode != 4) ptx::mbar_wait(&w_full[ws], ph);
            if (ptx::elect_one()) {
                if (p.mode >= 3) {
                    const uint64_t a = a0 + static_cast<uint64_t>((ws * kStageBytes) >> 4);
#pragma unroll
                    for (int k = 0; k < 4; ++k) ptx::umma_f16(tmem_base, a + 2 * k, b0 + 2 * k, idesc, 1u);
                }
                if (p.mode == 0 || p.mode == 2 || p.mode == 5) ptx::mbar_arrive(&w_empty[ws]);
                else ptx::umma_commit(&w_empty[ws]);
            }
            __syncwarp();
            if (++ws == p.stages) { ws = 0; ph ^= 1; }
        }
        if (ptx::elect_one()) ptx::umma_commit(done);
        __syncwarp();
        ptx::mbar_wait(done, 0);
        t1 = clock64();
        if (lane == 0) p.cyc[blockIdx.x] = t1 - t0;
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) { ptx::tc_fence_after(); ptx::tmem_dealloc<256>(tmem_base); }
}

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    EncodeTiledFn encode = (EncodeTiledFn)fn;
    const int rows_total = 32768;   // 4 MB of 128-byte rows: the weights of the 6x128 net
    void* w = nullptr;
    CK(cudaMalloc(&w, (size_t)rows_total * 128));
    CK(cudaMemset(w, 0, (size_t)rows_total * 128));
    CUtensorMap h_map[2];
    for (int i = 0; i < 2; ++i) {
        const cuuint64_t dims[2] = {64, (cuuint64_t)rows_total};
        const cuuint64_t strides[1] = {128};
        const cuuint32_t box[2] = {64, (cuuint32_t)(i ? 64 : 128)};
        const cuuint32_t estr[2] = {1, 1};
        CUresult r = encode(&h_map[i], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, w, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
    }
    CUtensorMap* d_map = nullptr;
    CK(cudaMalloc(&d_map, sizeof(h_map)));
    CK(cudaMemcpy(d_map, h_map, sizeof(h_map), cudaMemcpyHostToDevice));
    long long* d_cyc = nullptr;
    CK(cudaMalloc(&d_cyc, 148 * sizeof(long long)));
    const int smem = 1024 + kMaxStages * kStageBytes + 32768 + 512;
    CK(cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const char* names[5] = {"skeleton, plain release", "skeleton, commit release", "TMA stream, plain release", "TMA + 4 MMA + commit", "MMA + commit, no ring"};
    const int iters = 2048;
    for (int half = 0; half < 2; ++half) {
        for (int mode = 0; mode < 5; ++mode) {
            if (half && mode != 2 && mode != 3) continue;
            for (int n : {64, 128, 256}) {
                if (mode < 3 && n != 64) continue;
                for (int grid : {1, 16, 148}) {
                    if ((mode < 2 || mode == 4) && grid != 1) continue;
                    printf("%-26s stage %2d KB n=%3d grid=%3d :", names[mode], half ? 8 : 16, n, grid);
                    for (int stages : {2, 4, 8, 12}) {
                        if (mode == 4 && stages != 2) continue;
                        Params p{d_map + half, stages, iters, mode, n, rows_total, half ? 64 : 128, 1, d_cyc};
                        for (int rep = 0; rep < 2; ++rep) {
                            stream_kernel<<<grid, 128, smem>>>(p);
                            CK(cudaDeviceSynchronize());
                        }
                        std::vector<long long> c(grid);
                        CK(cudaMemcpy(c.data(), d_cyc, grid * sizeof(long long), cudaMemcpyDeviceToHost));
                        long long mx = 0;
                        for (long long v : c) mx = v > mx ? v : mx;
                        const double per = (double)mx / iters;
                        printf("  ring %2d: %6.1f cyc/stage", stages, per);
                        if (mode == 2 || mode == 3) printf(" (%5.1f B/clk)", (half ? 8192.0 : 16384.0) / per);
                    }
                    printf("\n");
                }
            }
        }
    }
    // raw stream rate: several 16 KB boxes per barrier, 12 boxes of ring in total
    for (int grid : {1, 16, 148}) {
        for (int boxes : {1, 2, 3, 4, 6}) {
            const int slots = 12 / boxes;
            Params p{d_map, slots, iters / boxes, 5, 64, rows_total, 128, boxes, d_cyc};
            for (int rep = 0; rep < 2; ++rep) { stream_kernel<<<grid, 128, smem>>>(p); CK(cudaDeviceSynchronize()); }
            std::vector<long long> c(grid);
            CK(cudaMemcpy(c.data(), d_cyc, grid * sizeof(long long), cudaMemcpyDeviceToHost));
            long long mx = 0;
            for (long long v : c) mx = v > mx ? v : mx;
            const double per = (double)mx / (iters / boxes);
            printf("raw stream grid=%3d: %d x 16 KB per barrier, %d slots: %7.1f cyc/stage = %5.1f B/clk per SM\n", grid, boxes, slots, per, boxes * 16384.0 / per);
        }
    }
    return 0;
}
