// conv_umma.cuh — 3x3 convolution on 8x8 boards as an implicit GEMM on the
// 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM, operands
// staged by TMA), with the whole layer tail fused into the epilogue.
//
// What it computes (python/model.py:113, :30-51 ResBlock, :8-26 SEBlock, :119):
//   MODE_RELU    out = relu(conv3x3(in) * bn_scale + bn_bias)          (start conv, conv1, p_conv)
//   MODE_SE_RES  o   = conv3x3(in) * bn_scale + bn_bias                (conv2 + bn2)
//                s   = sigmoid(W2 relu(W1 mean_hw(o)))                 (squeeze-excite)
//                out = relu(o * s + residual)                          (block output)
//                v_pre = <out, v_conv.weight>  (optional, fp32)        (value 1x1, :128)
// BatchNorm scale is folded into the packed weights at load time; bias here.
//
// GEMM view: one tile = 2 whole boards = 128 pixels (M = 128 = one UMMA),
// N = HID output channels (all of them, so SE's per-board channel mean and the
// residual fuse into the epilogue), K = 9 taps x C_in.  For each tap the A
// operand is the activation tensor shifted by (dy, dx): a 4-D TMA box
// {64 ch, 8, 8, 2 boards} fetched at coordinates (c0, dx, dy, b0) with dx, dy in
// {-1, 0, 1}; the hardware zero-fills out-of-range rows/columns, which IS the
// conv padding.  The box lands in shared memory as 128 rows x 128 B with the
// 128-byte swizzle, exactly the K-major SWIZZLE_128B canonical UMMA layout.
//
// Activations are NHWC 16-bit [B][8][8][C] in HBM (L2-resident at B = 1024:
// 16.8 MB per layer); weights are [tap][C_out][C_in] 16-bit (295 KB per layer).
//
// Warp roles (192 threads, 1 CTA per SM, persistent over tiles):
//   warp 0     TMA producer (one lane)      smem ring: full[]/empty[] mbarriers
//   warp 1     TMEM allocator + MMA issuer  accumulator ring of 2 in TMEM
//   warps 2-5  epilogue: tcgen05.ld -> fused tail -> swizzled smem tile -> TMA store
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "ptx.cuh"

namespace cb {

enum { MODE_RELU = 0, MODE_SE_RES = 1 };

template <int HID>
struct ConvCfg {
    static constexpr int kStages = HID == 128 ? 5 : 3;
    static constexpr int kABytes = 128 * 128;  // 128 pixels x 64 channels x 2 B
    static constexpr int kWBytes = HID * 128;  // HID rows x 64 channels x 2 B
    static constexpr int kStageBytes = kABytes + kWBytes;
    static constexpr int kOutBytes = 128 * HID * 2;
    static constexpr int kTmemCols = 2 * HID;  // two accumulators
    static constexpr int kSE = HID / 16;
    // barriers + tmem ptr (256) | bias HID | v_w HID | part 4*HID | scale 2*HID | hid 2*kSE   (floats)
    static constexpr int kMiscBytes = 256 + 4 * (HID + HID + 4 * HID + 2 * HID + 2 * kSE);
    static constexpr int kSmemBytes = 1024 + kStages * kStageBytes + kOutBytes + kMiscBytes;
    static constexpr int kThreads = 192;
};

struct ConvParams {
    int num_tiles;          // ceil(B / 2)
    int kc_in;              // C_in / 64
    const float* bias;      // [HID] folded BN bias
    const float* se_w1;     // [HID/16][HID]   (MODE_SE_RES)
    const float* se_w2;     // [HID][HID/16]
    const void* residual;   // NHWC 16-bit [B][64][HID]
    const float* v_w;       // [HID] value 1x1 weights or nullptr
    float* v_pre;           // [B*64] or nullptr
};

template <bool BF16>
__device__ __forceinline__ uint32_t pack2(float a, float b) {
    if constexpr (BF16) {
        __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
        return *reinterpret_cast<uint32_t*>(&v);
    } else {
        __half2 v = __floats2half2_rn(a, b);
        return *reinterpret_cast<uint32_t*>(&v);
    }
}
template <bool BF16>
__device__ __forceinline__ float2 unpack2(uint32_t u) {
    if constexpr (BF16) {
        return __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&u));
    } else {
        return __half22float2(*reinterpret_cast<__half2*>(&u));
    }
}

// Sum each of 32 per-thread values across the 32 lanes of a warp; lane l ends
// with the total of v[l].  31 shuffles (butterfly transpose-reduce).
__device__ __forceinline__ float warp_column_sums(float (&v)[32], int lane) {
#pragma unroll
    for (int h = 16; h >= 1; h >>= 1) {
        const bool upper = (lane & h) != 0;
#pragma unroll
        for (int i = 0; i < h; ++i) {
            const float keep = upper ? v[i + h] : v[i];
            const float send = upper ? v[i] : v[i + h];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, h);
        }
    }
    return v[0];
}

template <int HID, bool BF16, int MODE>
__global__ void __launch_bounds__(ConvCfg<HID>::kThreads, 1)
conv3x3_umma_kernel(const __grid_constant__ CUtensorMap tm_in, const __grid_constant__ CUtensorMap tm_w,
                    const __grid_constant__ CUtensorMap tm_out, const ConvParams p) {
    using Cfg = ConvCfg<HID>;
    constexpr int kStages = Cfg::kStages;
    constexpr int kSE = Cfg::kSE;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* stage_base = smem;
    uint8_t* out_tile = smem + kStages * Cfg::kStageBytes;
    uint8_t* misc = out_tile + Cfg::kOutBytes;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(misc);  // [kStages]
    uint64_t* empty_bar = full_bar + kStages;                // [kStages]
    uint64_t* tmem_full = empty_bar + kStages;               // [2]
    uint64_t* tmem_empty = tmem_full + 2;                    // [2]
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);
    float* s_bias = reinterpret_cast<float*>(misc + 256);
    float* s_vw = s_bias + HID;
    float* s_part = s_vw + HID;       // [4][HID] per-warp column sums
    float* s_scale = s_part + 4 * HID;  // [2][HID]
    float* s_hid = s_scale + 2 * HID;   // [2][kSE]

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int ksteps = 9 * p.kc_in;

    if (threadIdx.x == 0) {
        ptx::prefetch_tensormap(&tm_in);
        ptx::prefetch_tensormap(&tm_w);
        ptx::prefetch_tensormap(&tm_out);
        for (int s = 0; s < kStages; ++s) {
            ptx::mbar_init(&full_bar[s], 1);
            ptx::mbar_init(&empty_bar[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            ptx::mbar_init(&tmem_full[a], 1);
            ptx::mbar_init(&tmem_empty[a], 128);
        }
        ptx::fence_barrier_init();
    }
    if (warp == 1) ptx::tmem_alloc<Cfg::kTmemCols>(tmem_ptr);
    for (int i = threadIdx.x; i < HID; i += blockDim.x) {
        s_bias[i] = p.bias[i];
        s_vw[i] = p.v_w ? p.v_w[i] : 0.0f;
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0) {
        // ------------------------------------------------------ TMA producer
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
                const int b0 = tile * 2;
                for (int tap = 0; tap < 9; ++tap) {
                    const int dy = tap / 3 - 1, dx = tap % 3 - 1;
                    for (int kc = 0; kc < p.kc_in; ++kc) {
                        ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
                        uint8_t* a_dst = stage_base + stage * Cfg::kStageBytes;
                        uint8_t* w_dst = a_dst + Cfg::kABytes;
                        ptx::mbar_arrive_expect_tx(&full_bar[stage], Cfg::kStageBytes);
                        ptx::tma_load_4d(a_dst, &tm_in, &full_bar[stage], kc * 64, dx, dy, b0);
                        ptx::tma_load_2d(w_dst, &tm_w, &full_bar[stage], kc * 64, tap * HID);
                        if (++stage == kStages) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // -------------------------------------------------------- MMA issuer
        if (lane == 0) {
            constexpr uint32_t idesc = ptx::umma_idesc_f16(BF16 ? 1 : 0, 128, HID);
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
                const int acc = it & 1;
                const uint32_t acc_phase = (it >> 1) & 1;
                ptx::mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
                ptx::tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * HID;
                for (int ks = 0; ks < ksteps; ++ks) {
                    ptx::mbar_wait(&full_bar[stage], phase);
                    ptx::tc_fence_after();
                    const uint32_t a_addr = ptx::smem_u32(stage_base + stage * Cfg::kStageBytes);
                    const uint64_t a_desc = ptx::umma_desc_sw128(a_addr);
                    const uint64_t b_desc = ptx::umma_desc_sw128(a_addr + Cfg::kABytes);
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        // advance 16 elements (32 B) along K inside the swizzle atom: +2 in the
                        // 16-byte-unit start-address field
                        ptx::umma_f16(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (ks | k) != 0 ? 1u : 0u);
                    }
                    ptx::umma_commit(&empty_bar[stage]);
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
                ptx::umma_commit(&tmem_full[acc]);
            }
        }
    } else {
        // ---------------------------------------------------------- epilogue
        const int quad = warp & 3;            // TMEM lane quarter this warp may touch
        const int row = quad * 32 + lane;     // pixel row of the tile == TMEM lane
        const int et = threadIdx.x - 64;      // 0..127
        const bool leader = (et == 0);
        const uint32_t swz = static_cast<uint32_t>(row & 7);
        uint8_t* my_row = out_tile + row * 128;
        int it = 0;
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
            const int acc = it & 1;
            const uint32_t acc_phase = (it >> 1) & 1;
            // staging tile free again? (previous TMA store has read it)
            if (it > 0) {
                if (leader) ptx::tma_store_wait_read0();
                ptx::named_bar_sync(1, 128);
            }
            ptx::mbar_wait(&tmem_full[acc], acc_phase);
            ptx::tc_fence_after();
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * HID;

            if constexpr (MODE == MODE_RELU) {
#pragma unroll 1
                for (int c0 = 0; c0 < HID; c0 += 32) {
                    uint32_t r[32];
                    ptx::tmem_ld_32x32(taddr + c0, r);
                    ptx::tmem_ld_wait();
                    uint8_t* dst = my_row + (c0 >> 6) * Cfg::kABytes;
                    const uint32_t q0 = (c0 & 63) >> 3;
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        uint4 o;
                        uint32_t* ow = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const int c = i * 8 + j * 2;
                            const float a = fmaxf(__uint_as_float(r[c]) + s_bias[c0 + c], 0.0f);
                            const float b = fmaxf(__uint_as_float(r[c + 1]) + s_bias[c0 + c + 1], 0.0f);
                            ow[j] = pack2<BF16>(a, b);
                        }
                        *reinterpret_cast<uint4*>(dst + (((q0 + i) ^ swz) << 4)) = o;
                    }
                }
            } else {
                // pass 1: per-board channel sums of o = acc + bias
#pragma unroll 1
                for (int c0 = 0; c0 < HID; c0 += 32) {
                    uint32_t r[32];
                    ptx::tmem_ld_32x32(taddr + c0, r);
                    ptx::tmem_ld_wait();
                    float v[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]) + s_bias[c0 + j];
                    const float cs = warp_column_sums(v, lane);
                    s_part[quad * HID + c0 + lane] = cs;
                }
                ptx::named_bar_sync(1, 128);
                // squeeze FC: hid[b][j] = relu(<W1[j], mean[b]>)
                {
                    const int ew = et >> 5;
                    for (int o = ew; o < 2 * kSE; o += 4) {
                        const int b = o / kSE, j = o % kSE;
                        float s = 0.0f;
                        for (int c = lane; c < HID; c += 32)
                            s += __ldg(&p.se_w1[j * HID + c]) * (s_part[(2 * b) * HID + c] + s_part[(2 * b + 1) * HID + c]);
#pragma unroll
                        for (int d = 16; d >= 1; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
                        if (lane == 0) s_hid[b * kSE + j] = fmaxf(s * (1.0f / 64.0f), 0.0f);
                    }
                }
                ptx::named_bar_sync(1, 128);
                // excite FC + sigmoid
                for (int o = et; o < 2 * HID; o += 128) {
                    const int b = o / HID, c = o % HID;
                    float s = 0.0f;
#pragma unroll
                    for (int j = 0; j < kSE; ++j) s += __ldg(&p.se_w2[c * kSE + j]) * s_hid[b * kSE + j];
                    s_scale[b * HID + c] = 1.0f / (1.0f + __expf(-s));
                }
                ptx::named_bar_sync(1, 128);
                // pass 2: scale, residual, relu, (value 1x1), pack
                const float* scale = s_scale + (quad >> 1) * HID;
                const uint4* res = reinterpret_cast<const uint4*>(
                    reinterpret_cast<const uint16_t*>(p.residual) + (static_cast<size_t>(tile) * 128 + row) * HID);
                float vacc = 0.0f;
#pragma unroll 1
                for (int c0 = 0; c0 < HID; c0 += 32) {
                    uint4 rr[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) rr[i] = __ldg(res + (c0 >> 3) + i);
                    uint32_t r[32];
                    ptx::tmem_ld_32x32(taddr + c0, r);
                    ptx::tmem_ld_wait();
                    uint8_t* dst = my_row + (c0 >> 6) * Cfg::kABytes;
                    const uint32_t q0 = (c0 & 63) >> 3;
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        uint4 o;
                        uint32_t* ow = reinterpret_cast<uint32_t*>(&o);
                        const uint32_t* rw = reinterpret_cast<const uint32_t*>(&rr[i]);
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const int c = i * 8 + j * 2;
                            const float2 rv = unpack2<BF16>(rw[j]);
                            float a = (__uint_as_float(r[c]) + s_bias[c0 + c]) * scale[c0 + c] + rv.x;
                            float b = (__uint_as_float(r[c + 1]) + s_bias[c0 + c + 1]) * scale[c0 + c + 1] + rv.y;
                            a = fmaxf(a, 0.0f);
                            b = fmaxf(b, 0.0f);
                            vacc += a * s_vw[c0 + c] + b * s_vw[c0 + c + 1];
                            ow[j] = pack2<BF16>(a, b);
                        }
                        *reinterpret_cast<uint4*>(dst + (((q0 + i) ^ swz) << 4)) = o;
                    }
                }
                if (p.v_pre) p.v_pre[static_cast<size_t>(tile) * 128 + row] = vacc;
            }
            // accumulator drained -> MMA warp may overwrite it
            ptx::tc_fence_before();
            ptx::mbar_arrive(&tmem_empty[acc]);
            // publish the tile: generic-proxy writes -> async proxy, then one thread stores
            ptx::fence_proxy_async_smem();
            ptx::named_bar_sync(1, 128);
            if (leader) {
#pragma unroll
                for (int kc = 0; kc < HID / 64; ++kc)
                    ptx::tma_store_4d(&tm_out, out_tile + kc * Cfg::kABytes, kc * 64, 0, 0, tile * 2);
                ptx::tma_store_commit();
            }
        }
        if (leader) ptx::tma_store_wait0();
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc<Cfg::kTmemCols>(tmem_base);
    }
}

}  // namespace cb
