// policy_umma.cuh — policy head on the tensor cores: p_head 1x1 conv (HID -> 73 planes) as a
// tcgen05 GEMM, then the full-row log-sum-exp over the 4672 logits of each board, and either
// all probabilities (compat, src/neural_net.rs:279-284) or only the legal moves gathered and
// renormalised in list order with the uniform fallback (src/tensor.rs:184-213).
//
//   logits[128 pixels of a pair][80] = P[128][HID] * Wp[80][HID]^T      (73 real planes, 7 zero)
//
// P is the post-ReLU policy feature map written by p_conv in the pair-interleaved layout
// (conv_umma.cuh); its [128 rows][64 ch] chunks land in shared memory through the same TMA
// box the conv kernel stores with, which is already the K-major SWIZZLE_128B A operand.  Wp is
// loaded once per CTA.  Logits never leave the SM: TMEM -> registers -> shared (fp32, both
// boards of the pair: 37 KB) -> softmax / gather.
//
// Warp roles (192 threads, persistent over pairs): warp 0 TMA producer, warp 1 MMA issuer +
// TMEM owner (2 accumulators of 128 columns, 80 used), warps 2-5 epilogue.
#pragma once
#include "conv_umma.cuh"
#include "kernels_simt.cuh"

namespace cb {

template <int HID>
struct PolicyCfg {
    static constexpr int kN = 80;                          // 73 planes padded to a legal UMMA N
    static constexpr int kKc = HID / 64;
    static constexpr int kWChunkBytes = kN * 128;          // 10240 (1024-aligned)
    static constexpr int kWBytes = kKc * kWChunkBytes;
    static constexpr int kPChunkBytes = 128 * 128;
    static constexpr int kPBytes = kKc * kPChunkBytes;     // one pair of boards
    static constexpr int kStages = 2;
    static constexpr int kLogitFloats = 2 * CB_POLICY_SIZE;
    static constexpr int kMiscBytes = 256 + 4 * (80 + 32 + 256);  // barriers | bias | reductions | gather
    static constexpr int kSmemBytes = 1024 + kWBytes + kStages * kPBytes + 4 * kLogitFloats + kMiscBytes;
    static constexpr int kThreads = 192;
    static constexpr int kTmemCols = 256;
    static_assert(kSmemBytes <= 232448, "shared memory budget");
};

template <int HID, bool BF16>
__global__ void __launch_bounds__(PolicyCfg<HID>::kThreads, 1)
policy_umma_kernel(const __grid_constant__ CUtensorMap tm_p, const __grid_constant__ CUtensorMap tm_wp,
                   const float* __restrict__ bias, int num_tiles, int B, PolicyOut po) {
    using Cfg = PolicyCfg<HID>;
    constexpr int kStages = Cfg::kStages;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* w_smem = smem;
    uint8_t* p_smem = w_smem + Cfg::kWBytes;
    float* s_logit = reinterpret_cast<float*>(p_smem + kStages * Cfg::kPBytes);   // [2 boards][4672]
    uint8_t* misc = reinterpret_cast<uint8_t*>(s_logit + Cfg::kLogitFloats);
    uint64_t* w_full = reinterpret_cast<uint64_t*>(misc);
    uint64_t* p_full = w_full + 1;              // [kStages]
    uint64_t* p_empty = p_full + kStages;       // [kStages]
    uint64_t* tmem_full = p_empty + kStages;    // [2]
    uint64_t* tmem_empty = tmem_full + 2;       // [2]
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);
    float* s_bias = reinterpret_cast<float*>(misc + 256);  // [80]
    float* s_red = s_bias + 80;                            // [4 warps][2 boards] x {max, sum}, total
    float* s_g = s_red + 32;                               // [256] gathered legal-move probabilities

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        ptx::prefetch_tensormap(&tm_p);
        ptx::prefetch_tensormap(&tm_wp);
        ptx::mbar_init(w_full, 1);
        for (int s = 0; s < kStages; ++s) {
            ptx::mbar_init(&p_full[s], 1);
            ptx::mbar_init(&p_empty[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            ptx::mbar_init(&tmem_full[a], 1);
            ptx::mbar_init(&tmem_empty[a], 128);
        }
        ptx::fence_barrier_init();
    }
    if (warp == 1) ptx::tmem_alloc<Cfg::kTmemCols>(tmem_ptr);
    for (int i = threadIdx.x; i < 80; i += blockDim.x) s_bias[i] = i < 73 ? bias[i] : 0.0f;
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0) {
        // ------------------------------------------------------ TMA producer
        if (ptx::elect_one()) {
            ptx::mbar_arrive_expect_tx(w_full, Cfg::kWBytes);
#pragma unroll
            for (int kc = 0; kc < Cfg::kKc; ++kc)
                ptx::tma_load_2d(w_smem + kc * Cfg::kWChunkBytes, &tm_wp, w_full, kc * 64, 0);
        }
        __syncwarp();
        int st = 0;
        uint32_t ph = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            ptx::mbar_wait(&p_empty[st], ph ^ 1);
            if (ptx::elect_one()) {
                ptx::mbar_arrive_expect_tx(&p_full[st], Cfg::kPBytes);
#pragma unroll
                for (int kc = 0; kc < Cfg::kKc; ++kc)
                    ptx::tma_load_5d(p_smem + st * Cfg::kPBytes + kc * Cfg::kPChunkBytes, &tm_p, &p_full[st], kc * 64,
                                     0, 0, 0, tile);
            }
            __syncwarp();
            if (++st == kStages) { st = 0; ph ^= 1; }
        }
    } else if (warp == 1) {
        // -------------------------------------------------------- MMA issuer
        constexpr uint32_t idesc = ptx::umma_idesc_f16(BF16 ? 1 : 0, 128, Cfg::kN);
        ptx::mbar_wait(w_full, 0);
        int st = 0, it = 0;
        uint32_t ph = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
            const int acc = it & 1;
            const uint32_t acc_phase = (it >> 1) & 1;
            ptx::mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
            ptx::mbar_wait(&p_full[st], ph);
            ptx::tc_fence_after();
            if (ptx::elect_one()) {
                const uint32_t d_tmem = tmem_base + acc * 128;
#pragma unroll
                for (int kc = 0; kc < Cfg::kKc; ++kc) {
                    const uint64_t a_desc =
                        ptx::umma_desc_sw128(ptx::smem_u32(p_smem + st * Cfg::kPBytes + kc * Cfg::kPChunkBytes));
                    const uint64_t b_desc = ptx::umma_desc_sw128(ptx::smem_u32(w_smem + kc * Cfg::kWChunkBytes));
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        ptx::umma_f16(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kc | k) != 0 ? 1u : 0u);
                }
                ptx::umma_commit(&p_empty[st]);
                ptx::umma_commit(&tmem_full[acc]);
            }
            __syncwarp();
            if (++st == kStages) { st = 0; ph ^= 1; }
        }
    } else {
        // ---------------------------------------------------------- epilogue
        const int quad = warp & 3;
        const int row = quad * 32 + lane;        // (y, b2, x) = (row>>4, (row>>3)&1, row&7)
        const int et = threadIdx.x - 64;
        const int ew = et >> 5;
        const int b2 = (lane >> 3) & 1;
        const int px = (row >> 4) * 8 + (row & 7);
        float* my_logits = s_logit + b2 * CB_POLICY_SIZE + px * 73;
        int it = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
            const int acc = it & 1;
            const uint32_t acc_phase = (it >> 1) & 1;
            ptx::mbar_wait(&tmem_full[acc], acc_phase);
            ptx::tc_fence_after();
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * 128;
            float mx = -INFINITY;
#pragma unroll
            for (int c0 = 0; c0 < 96; c0 += 32) {
                uint32_t r[32];
                ptx::tmem_ld_32x32(taddr + c0, r);
                ptx::tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const int c = c0 + j;
                    if (c < 73) {
                        const float v = __uint_as_float(r[j]) + s_bias[c];
                        my_logits[c] = v;
                        mx = fmaxf(mx, v);
                    }
                }
            }
            ptx::tc_fence_before();
            ptx::mbar_arrive(&tmem_empty[acc]);
            // per-board row max over the 16 lanes of this warp that share the board bit, then over warps
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 4));
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 16));
            if ((lane & 23) == 0) s_red[ew * 2 + b2] = mx;
            ptx::named_bar_sync(1, 128);
            float bmax[2];
#pragma unroll
            for (int b = 0; b < 2; ++b)
                bmax[b] = fmaxf(fmaxf(s_red[b], s_red[2 + b]), fmaxf(s_red[4 + b], s_red[6 + b]));
            // sum of exp over each board's 4672 logits (any thread may read any logit now)
            float se[2] = {0.0f, 0.0f};
#pragma unroll
            for (int b = 0; b < 2; ++b)
                for (int i = et; i < CB_POLICY_SIZE; i += 128) se[b] += __expf(s_logit[b * CB_POLICY_SIZE + i] - bmax[b]);
#pragma unroll
            for (int b = 0; b < 2; ++b) {
#pragma unroll
                for (int d = 16; d >= 1; d >>= 1) se[b] += __shfl_xor_sync(0xffffffffu, se[b], d);
            }
            if (lane == 0) {
                s_red[8 + ew * 2 + 0] = se[0];
                s_red[8 + ew * 2 + 1] = se[1];
            }
            ptx::named_bar_sync(1, 128);
            float lse[2];
#pragma unroll
            for (int b = 0; b < 2; ++b)
                lse[b] = bmax[b] + __logf(s_red[8 + b] + s_red[10 + b] + s_red[12 + b] + s_red[14 + b]);

#pragma unroll 1
            for (int b = 0; b < 2; ++b) {
                const int board = tile * 2 + b;
                if (board >= B) break;
                const float* lg = s_logit + b * CB_POLICY_SIZE;
                if (po.probs) {
                    float* o = po.probs + static_cast<size_t>(board) * CB_POLICY_SIZE;
                    for (int i = et; i < CB_POLICY_SIZE; i += 128) o[i] = __expf(lg[i] - lse[b]);
                }
                if (po.move_idx) {
                    const uint32_t lo = po.move_cnt ? static_cast<uint32_t>(board) * CB_MAX_MOVES : po.move_off[board];
                    const uint32_t hi = po.move_cnt ? lo + po.move_cnt[board] : po.move_off[board + 1];
                    const int n = static_cast<int>(hi - lo);
                    // gathered values -> registers (n <= 256 = 2 per thread), total via an ordered
                    // sequential sum by one thread, exactly the reference's loop (src/tensor.rs:186-201)
                    float g0 = 0.0f, g1 = 0.0f;
                    if (et < n) {
                        const int idx = po.move_idx[lo + et];
                        g0 = idx < CB_POLICY_SIZE ? __expf(lg[idx] - lse[b]) : 0.0f;
                    }
                    if (et + 128 < n) {
                        const int idx = po.move_idx[lo + et + 128];
                        g1 = idx < CB_POLICY_SIZE ? __expf(lg[idx] - lse[b]) : 0.0f;
                    }
                    if (et < n) s_g[et] = g0;
                    if (et + 128 < n) s_g[et + 128] = g1;
                    ptx::named_bar_sync(1, 128);
                    if (et == 0) {
                        float t = 0.0f;
                        for (int i = 0; i < n; ++i) t += s_g[i];
                        s_red[16] = t;
                    }
                    ptx::named_bar_sync(1, 128);
                    const float total = s_red[16];
                    if (et < n) po.priors[lo + et] = total > 0.0f ? g0 / total : 1.0f / static_cast<float>(n);
                    if (et + 128 < n) po.priors[lo + et + 128] = total > 0.0f ? g1 / total : 1.0f / static_cast<float>(n);
                }
            }
            // logits of this pair fully consumed before the next tile overwrites them
            ptx::named_bar_sync(1, 128);
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc<Cfg::kTmemCols>(tmem_base);
    }
}

}  // namespace cb
