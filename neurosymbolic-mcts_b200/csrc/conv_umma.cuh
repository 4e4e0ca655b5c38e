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
// Data layout in HBM.  Activations are 16-bit, "pair-interleaved NHWC":
//   act[pair][y][b2][x][C]   board = 2*pair + b2, (y, x) = input row / column
// so one tile = one board pair = 128 consecutive pixel rows m = (y*2 + b2)*8 + x of C
// channels each.  Weights are [tap][C_out][C_in] 16-bit with the BN scale folded.
//
// GEMM view: M = 128 pixels of a pair (one UMMA), N = HID output channels (all of them,
// so SE's per-board channel mean and the residual fuse into the epilogue), K = 9 taps x
// C_in.  The A operand of tap (dy, dx) is the activation shifted by (dy, dx):
//   - dx: a 5-D TMA box {64 ch, 8 x, 2 b2, 10 y, 1 pair} fetched at x-coordinate dx
//     (hardware zero-fill of the out-of-range column = conv padding), y from -1 to 8
//     (zero-filled halo rows).  The box lands as 160 rows x 128 B with the 128-byte
//     swizzle: the K-major SWIZZLE_128B canonical UMMA layout, rows ordered (y, b2, x).
//   - dy: because y is the OUTER row index and each (y, b2) is one 8-row swizzle atom,
//     the three dy taps are just the same slab read at byte offset (dy+1)*2048 —
//     1024-aligned, so the swizzle phase is unchanged.  3 slab loads serve 9 taps.
// One round of a CTA processes kTiles pairs against the same weight stages, so each
// weight stage (HID x 64 ch) read from L2 feeds kTiles x 4 MMAs.
//
// Warp roles (64 + kTiles*128 threads, 1 CTA per SM, persistent over rounds):
//   warp 0       TMA producer (one lane): slab ring a_full/a_empty, weight ring w_full/w_empty
//   warp 1       TMEM allocator + MMA issuer; 2 accumulator sets x kTiles tiles in TMEM
//   warps 2..    one 4-warp epilogue group per tile of the round:
//                tcgen05.ld -> fused tail -> swizzled smem tile -> TMA store
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "ptx.cuh"

namespace cb {

enum { MODE_RELU = 0, MODE_SE_RES = 1 };

template <int HID>
struct ConvCfg {
    static constexpr int kTiles = HID == 128 ? 2 : 1;     // board pairs per round
    static constexpr int kSlabBytes = 160 * 128;           // [10 y][2 b2][8 x][64 ch] 16-bit
    static constexpr int kAStageBytes = kTiles * kSlabBytes;
    static constexpr int kAStages = 2;
    static constexpr int kWBytes = HID * 128;              // HID rows x 64 channels
    static constexpr int kWStages = HID == 128 ? 4 : 3;
    static constexpr int kKcBytes = 128 * 128;             // one 64-channel chunk of an output tile
    static constexpr int kOutBytes = 128 * HID * 2;        // per tile
    static constexpr int kTmemCols = 512;                  // 2 sets x kTiles x HID fp32 columns
    static constexpr int kSE = HID / 16;
    static constexpr int kThreads = 64 + kTiles * 128;
    // per epilogue group: part [4 quads][2 boards][HID] | scale [2][HID] | hid [2][kSE]  (floats)
    static constexpr int kGroupFloats = 8 * HID + 2 * HID + 2 * kSE;
    // barriers + tmem ptr (256 B) | bias HID | v_w HID | groups
    static constexpr int kMiscBytes = 256 + 4 * (2 * HID + kTiles * kGroupFloats);
    static constexpr int kSmemBytes =
        1024 + kAStages * kAStageBytes + kWStages * kWBytes + kTiles * kOutBytes + kMiscBytes;
    static_assert(2 * kTiles * HID == kTmemCols, "TMEM budget");
    static_assert(kSmemBytes <= 232448, "shared memory budget (227 KB)");
};

struct ConvParams {
    int num_tiles;          // board pairs = ceil(B / 2)
    int kc_in;              // C_in / 64
    const float* bias;      // [HID] folded BN bias
    const float* se_w1;     // [HID/16][HID]   (MODE_SE_RES)
    const float* se_w2;     // [HID][HID/16]
    const void* residual;   // pair-interleaved 16-bit [pairs][128][HID]
    const float* v_w;       // [HID] value 1x1 weights or nullptr
    float* v_pre;           // [pairs*128] or nullptr
    int dbg;                // experiment switches (0 in production)
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

// One activation value <-> its 16-bit storage form.
template <bool BF16>
__device__ __forceinline__ uint16_t to16bits(float v) {
    if constexpr (BF16) {
        __nv_bfloat16 b = __float2bfloat16_rn(v);
        return *reinterpret_cast<uint16_t*>(&b);
    } else {
        __half h = __float2half_rn(v);
        return *reinterpret_cast<uint16_t*>(&h);
    }
}
template <bool BF16>
__device__ __forceinline__ float from16bits(uint16_t u) {
    if constexpr (BF16) {
        return __bfloat162float(*reinterpret_cast<__nv_bfloat16*>(&u));
    } else {
        return __half2float(*reinterpret_cast<__half*>(&u));
    }
}

// Per-board column sums inside one warp.  Lane bit 3 is the board (b2) of the lane's
// pixel row, so the butterfly runs over lane bits 4, 2, 1, 0 only: 32 per-thread values
// shrink to 2, summed over the 16 lanes that share the board bit.  On return v[0], v[1]
// are the totals of columns col and col+1, col = 16*bit4 + 8*bit2 + 4*bit1 + 2*bit0.
__device__ __forceinline__ int warp_board_column_sums(float (&v)[32], int lane) {
    {
        const bool upper = (lane & 16) != 0;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const float keep = upper ? v[i + 16] : v[i];
            const float send = upper ? v[i] : v[i + 16];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
        }
    }
#pragma unroll
    for (int h = 4; h >= 1; h >>= 1) {
        const bool upper = (lane & h) != 0;
        const int n = h * 2;  // values kept after this step: 8, 4, 2
#pragma unroll
        for (int i = 0; i < n; ++i) {
            const float keep = upper ? v[i + n] : v[i];
            const float send = upper ? v[i] : v[i + n];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, h);
        }
    }
    return ((lane >> 4) & 1) * 16 + ((lane >> 2) & 1) * 8 + ((lane >> 1) & 1) * 4 + (lane & 1) * 2;
}

template <int HID, bool BF16, int MODE>
__global__ void __launch_bounds__(ConvCfg<HID>::kThreads, 1)
conv3x3_umma_kernel(const __grid_constant__ CUtensorMap tm_in, const __grid_constant__ CUtensorMap tm_w,
                    const __grid_constant__ CUtensorMap tm_out, const ConvParams p) {
    using Cfg = ConvCfg<HID>;
    constexpr int kTiles = Cfg::kTiles;
    constexpr int kAStages = Cfg::kAStages;
    constexpr int kWStages = Cfg::kWStages;
    constexpr int kSE = Cfg::kSE;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // keep the pointer derived from the __shared__ array (no integer round trip) so the compiler
    // emits LDS/STS instead of generic LD/ST for everything below
    uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* a_base = smem;
    uint8_t* w_base = a_base + kAStages * Cfg::kAStageBytes;
    uint8_t* out_base = w_base + kWStages * Cfg::kWBytes;
    uint8_t* misc = out_base + kTiles * Cfg::kOutBytes;
    uint64_t* a_full = reinterpret_cast<uint64_t*>(misc);   // [kAStages]
    uint64_t* a_empty = a_full + kAStages;                  // [kAStages]
    uint64_t* w_full = a_empty + kAStages;                  // [kWStages]
    uint64_t* w_empty = w_full + kWStages;                  // [kWStages]
    uint64_t* tmem_full = w_empty + kWStages;               // [2 sets][kTiles]
    uint64_t* tmem_empty = tmem_full + 2 * kTiles;          // [2 sets][kTiles]
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2 * kTiles);
    float* s_bias = reinterpret_cast<float*>(misc + 256);
    float* s_vw = s_bias + HID;
    float* s_groups = s_vw + HID;

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int num_rounds = (p.num_tiles + kTiles - 1) / kTiles;

    if (threadIdx.x == 0) {
        ptx::prefetch_tensormap(&tm_in);
        ptx::prefetch_tensormap(&tm_w);
        ptx::prefetch_tensormap(&tm_out);
        for (int s = 0; s < kAStages; ++s) {
            ptx::mbar_init(&a_full[s], 1);
            ptx::mbar_init(&a_empty[s], 1);
        }
        for (int s = 0; s < kWStages; ++s) {
            ptx::mbar_init(&w_full[s], 1);
            ptx::mbar_init(&w_empty[s], 1);
        }
        for (int a = 0; a < 2 * kTiles; ++a) {
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
        // The whole warp runs the (uniform) control flow and waits; one elected lane issues,
        // so addresses stay in uniform registers (no per-instruction waterfall loops).
        int as = 0, ws = 0;
        uint32_t aph = 0, wph = 0;
        for (int round = blockIdx.x; round < num_rounds; round += gridDim.x) {
            const int tile0 = round * kTiles;
            for (int dx = -1; dx <= 1; ++dx) {
                for (int kc = 0; kc < p.kc_in; ++kc) {
                    ptx::mbar_wait(&a_empty[as], aph ^ 1);
                    if (ptx::elect_one()) {
                        if (p.dbg & 8) {
                            ptx::mbar_arrive(&a_full[as]);
                        } else {
                            ptx::mbar_arrive_expect_tx(&a_full[as], Cfg::kAStageBytes);
#pragma unroll
                            for (int t = 0; t < kTiles; ++t)
                                ptx::tma_load_5d(a_base + as * Cfg::kAStageBytes + t * Cfg::kSlabBytes, &tm_in,
                                                 &a_full[as], kc * 64, dx, 0, -1, tile0 + t);
                        }
                    }
                    __syncwarp();
                    if (++as == kAStages) { as = 0; aph ^= 1; }
                    for (int dy = -1; dy <= 1; ++dy) {
                        const int tap = (dy + 1) * 3 + (dx + 1);
                        ptx::mbar_wait(&w_empty[ws], wph ^ 1);
                        if (ptx::elect_one()) {
                            if (p.dbg & 8) {
                                ptx::mbar_arrive(&w_full[ws]);
                            } else {
                                ptx::mbar_arrive_expect_tx(&w_full[ws], Cfg::kWBytes);
                                ptx::tma_load_2d(w_base + ws * Cfg::kWBytes, &tm_w, &w_full[ws], kc * 64, tap * HID);
                            }
                        }
                        __syncwarp();
                        if (++ws == kWStages) { ws = 0; wph ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // -------------------------------------------------------- MMA issuer
        // Warp-uniform control flow; the elected lane issues the MMAs and their commits.
        constexpr uint32_t idesc = ptx::umma_idesc_f16(BF16 ? 1 : 0, 128, HID);
        int as = 0, ws = 0;
        uint32_t aph = 0, wph = 0;
        int it = 0;
        for (int round = blockIdx.x; round < num_rounds; round += gridDim.x, ++it) {
            const int set = it & 1;
            const uint32_t set_phase = (it >> 1) & 1;
#pragma unroll
            for (int t = 0; t < kTiles; ++t) ptx::mbar_wait(&tmem_empty[set * kTiles + t], set_phase ^ 1);
            ptx::tc_fence_after();
            uint32_t accum = 0;  // first MMA of the round overwrites the accumulators
            for (int dxi = 0; dxi < 3; ++dxi) {
                for (int kc = 0; kc < p.kc_in; ++kc) {
                    ptx::mbar_wait(&a_full[as], aph);
                    const uint32_t slab0 = ptx::smem_u32(a_base + as * Cfg::kAStageBytes);
                    for (int dyi = 0; dyi < 3; ++dyi) {
                        ptx::mbar_wait(&w_full[ws], wph);
                        ptx::tc_fence_after();
                        if (ptx::elect_one()) {
                            const uint64_t b_desc = ptx::umma_desc_sw128(ptx::smem_u32(w_base + ws * Cfg::kWBytes));
#pragma unroll
                            for (int t = 0; t < kTiles; ++t) {
                                const uint32_t d_tmem = tmem_base + (set * kTiles + t) * HID;
                                // tap dy = rows [(dy+1)*16, (dy+1)*16 + 128) of the slab
                                const uint64_t a_desc = ptx::umma_desc_sw128(slab0 + t * Cfg::kSlabBytes + dyi * 2048);
                                ptx::umma_f16(d_tmem, a_desc, b_desc, idesc, accum);
#pragma unroll
                                for (int k = 1; k < 4; ++k)
                                    ptx::umma_f16(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, 1u);
                            }
                            ptx::umma_commit(&w_empty[ws]);
                            if (dyi == 2) ptx::umma_commit(&a_empty[as]);
                        }
                        __syncwarp();
                        accum = 1;
                        if (++ws == kWStages) { ws = 0; wph ^= 1; }
                    }
                    if (++as == kAStages) { as = 0; aph ^= 1; }
                }
            }
            if (ptx::elect_one()) {
#pragma unroll
                for (int t = 0; t < kTiles; ++t) ptx::umma_commit(&tmem_full[set * kTiles + t]);
            }
            __syncwarp();
        }
    } else {
        // ---------------------------------------------------------- epilogue
        const int g = (warp - 2) >> 2;        // epilogue group == tile slot of the round
        const int quad = warp & 3;            // TMEM lane quarter this warp may touch
        const int row = quad * 32 + lane;     // pixel row of the tile == TMEM lane; (y, b2, x) = (row>>4, (row>>3)&1, row&7)
        const int et = (threadIdx.x - 64) & 127;
        const int ew = et >> 5;
        const bool leader = (et == 0);
        const int bar_id = 1 + g;
        const int b2 = (lane >> 3) & 1;
        const uint32_t swz = static_cast<uint32_t>(row & 7);
        uint8_t* out_tile = out_base + g * Cfg::kOutBytes;
        uint8_t* my_row = out_tile + row * 128;
        float* s_part = s_groups + g * Cfg::kGroupFloats;   // [4 quads][2 boards][HID]
        float* s_scale = s_part + 8 * HID;                  // [2][HID]
        float* s_hid = s_scale + 2 * HID;                   // [2][kSE]
        // SE weights this thread needs, fetched once per launch (L1 is ~7 KB next to 221 KB of
        // shared memory, so per-tile __ldg chains would each pay an L2 round trip).
        //   squeeze FC: warp ew computes hid[b][j] for j in {ew, ew+4, ...}; lane covers c = lane + 32 i
        //   excite FC : thread et computes scale[b][c] for c in {et, et+128, ...}
        [[maybe_unused]] float w1r[kSE / 4][HID / 32];
        [[maybe_unused]] float w2r[HID / 128][kSE];
        if constexpr (MODE == MODE_SE_RES) {
#pragma unroll
            for (int jj = 0; jj < kSE / 4; ++jj)
#pragma unroll
                for (int i = 0; i < HID / 32; ++i) w1r[jj][i] = __ldg(&p.se_w1[(ew + 4 * jj) * HID + lane + 32 * i]);
#pragma unroll
            for (int cc = 0; cc < HID / 128; ++cc)
#pragma unroll
                for (int j = 0; j < kSE; ++j) w2r[cc][j] = __ldg(&p.se_w2[(et + 128 * cc) * kSE + j]);
        }
        int it = 0;
        bool stored = false;
        for (int round = blockIdx.x; round < num_rounds; round += gridDim.x, ++it) {
            const int set = it & 1;
            const uint32_t set_phase = (it >> 1) & 1;
            const int tile = round * kTiles + g;
            const int slot = set * kTiles + g;
            const bool valid = tile < p.num_tiles;
            // staging tile free again? (this group's previous TMA store has read it)
            if (stored) {
                if (leader) ptx::tma_store_wait_read0();
                ptx::named_bar_sync(bar_id, 128);
            }
            // residual prefetch (latency hides behind the accumulator wait / SE phases)
            [[maybe_unused]] uint4 rr[MODE == MODE_SE_RES && HID == 128 ? 16 : 1];
            [[maybe_unused]] const uint4* res = nullptr;
            if constexpr (MODE == MODE_SE_RES) {
                res = reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(p.residual) +
                                                     (static_cast<size_t>(valid ? tile : 0) * 128 + row) * HID);
                if constexpr (HID == 128) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) rr[i] = __ldg(res + i);
                }
            }
            ptx::mbar_wait(&tmem_full[slot], set_phase);
            ptx::tc_fence_after();
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + slot * HID;

            if (!valid || (p.dbg & 16)) {  // padding tile of an odd round: drain protocol only
                ptx::tc_fence_before();
                ptx::mbar_arrive(&tmem_empty[slot]);
                continue;
            }
            if constexpr (MODE == MODE_RELU) {
#pragma unroll 1
                for (int c0 = 0; c0 < HID; c0 += 32) {
                    uint32_t r[32];
                    ptx::tmem_ld_32x32(taddr + c0, r);
                    ptx::tmem_ld_wait();
                    uint8_t* dst = my_row + (c0 >> 6) * Cfg::kKcBytes;
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
                for (int c0 = 0; c0 < ((p.dbg & 1) ? 0 : HID); c0 += 32) {
                    uint32_t r[32];
                    ptx::tmem_ld_32x32(taddr + c0, r);
                    ptx::tmem_ld_wait();
                    float v[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]) + s_bias[c0 + j];
                    const int col = warp_board_column_sums(v, lane);
                    *reinterpret_cast<float2*>(&s_part[(quad * 2 + b2) * HID + c0 + col]) = make_float2(v[0], v[1]);
                }
                ptx::named_bar_sync(bar_id, 128);
                // squeeze FC: hid[b][j] = relu(<W1[j], mean[b]>)
                if (!(p.dbg & 2)) {
#pragma unroll
                    for (int b = 0; b < 2; ++b) {
                        float m[HID / 32];
#pragma unroll
                        for (int i = 0; i < HID / 32; ++i) {
                            const int c = lane + 32 * i;
                            m[i] = s_part[(0 * 2 + b) * HID + c] + s_part[(1 * 2 + b) * HID + c] +
                                   s_part[(2 * 2 + b) * HID + c] + s_part[(3 * 2 + b) * HID + c];
                        }
#pragma unroll
                        for (int jj = 0; jj < kSE / 4; ++jj) {
                            float s = 0.0f;
#pragma unroll
                            for (int i = 0; i < HID / 32; ++i) s = fmaf(w1r[jj][i], m[i], s);
#pragma unroll
                            for (int d = 16; d >= 1; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
                            if (lane == 0) s_hid[b * kSE + ew + 4 * jj] = fmaxf(s * (1.0f / 64.0f), 0.0f);
                        }
                    }
                }
                ptx::named_bar_sync(bar_id, 128);
                // excite FC + sigmoid
#pragma unroll
                for (int cc = 0; cc < HID / 128; ++cc) {
#pragma unroll
                    for (int b = 0; b < 2; ++b) {
                        float s = 0.0f;
                        if (!(p.dbg & 2)) {
#pragma unroll
                            for (int j = 0; j < kSE; ++j) s = fmaf(w2r[cc][j], s_hid[b * kSE + j], s);
                        }
                        s_scale[b * HID + et + 128 * cc] = 1.0f / (1.0f + __expf(-s));
                    }
                }
                ptx::named_bar_sync(bar_id, 128);
                // pass 2: scale, residual, relu, (value 1x1), pack
                const float* scale = s_scale + b2 * HID;
                float vacc = 0.0f;
#pragma unroll
                for (int c0 = 0; c0 < HID; c0 += 32) {
                    uint4 rl[4];
                    if constexpr (HID == 128) {
#pragma unroll
                        for (int i = 0; i < 4; ++i) rl[i] = rr[(c0 >> 3) + i];
                    } else {
#pragma unroll
                        for (int i = 0; i < 4; ++i) rl[i] = __ldg(res + (c0 >> 3) + i);
                    }
                    uint32_t r[32];
                    ptx::tmem_ld_32x32(taddr + c0, r);
                    ptx::tmem_ld_wait();
                    uint8_t* dst = my_row + (c0 >> 6) * Cfg::kKcBytes;
                    const uint32_t q0 = (c0 & 63) >> 3;
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        uint4 o;
                        uint32_t* ow = reinterpret_cast<uint32_t*>(&o);
                        const uint32_t* rw = reinterpret_cast<const uint32_t*>(&rl[i]);
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
            ptx::mbar_arrive(&tmem_empty[slot]);
            // publish the tile: generic-proxy writes -> async proxy, then one thread stores
            ptx::fence_proxy_async_smem();
            ptx::named_bar_sync(bar_id, 128);
            if (leader) {
#pragma unroll
                for (int kc = 0; kc < HID / 64; ++kc)
                    ptx::tma_store_5d(&tm_out, out_tile + kc * Cfg::kKcBytes, kc * 64, 0, 0, 0, tile);
                ptx::tma_store_commit();
            }
            stored = true;
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
