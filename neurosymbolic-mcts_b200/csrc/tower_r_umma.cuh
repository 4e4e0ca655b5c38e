// tower_r_umma.cuh — the WHOLE leaf evaluation of the 128-channel net as ONE persistent kernel:
// planes from packed bitboards -> 14 fused 3x3 convolutions (BN, ReLU, squeeze-excite, residual) ->
// policy head (1x1 conv, softmax over the legal moves) and value head (tanh(v_logit + k dM)), with the
// activations RESIDENT IN SHARED MEMORY between layers (no global round trip, no halo'd slab reloads).
//
//     D^T[c_out = 128 lanes][pixels = N columns] += W_tap[c_out][c_in] * X_tap[pixels][c_in]^T
//
// A CTA owns a UNIT of up to 8 boards split into two CHAINS of up to 4 boards; chain c accumulates
// in TMEM columns [256c, 256c + 64 nb).  The boards of a chain live in one shared-memory buffer as
// 16-bit K-major SWIZZLE_128B rows (one pixel = 128 B = 64 channels, two 64-channel halves):
//
//     row(y, b, x) = ((y + halo) * nb + b) * 9 + 1 + x            (b < nb boards, 9 rows per group)
//
// Row 0 of every 9-row group is a zero row: it is x = -1 of its own group and x = 8 of the group
// before it, so a dx-shifted tap is just the same buffer read from row offset (1 + dx) with an
// 8-row-group stride (SBO) of 9 rows.  The UMMA swizzle is a function of the absolute shared-memory
// address (tools/exp_desc.cu), so the start address needs no 1024-byte alignment.  A dy shift is a
// start offset of nb groups.  For an even board count there are no y-halo rows at all: the dy = -1
// (+1) taps are issued as N = 56 nb MMAs that only touch output rows y >= 1 (y <= 6), i.e. an
// accumulator column offset of 8 nb (0) — the padding rows are never multiplied.  An odd board
// count (56 nb is not a legal N) keeps two zero y-halo group rows and uses N = 64 nb for every tap.
//
// The epilogue writes layer L's 16-bit output straight back into the chain's buffer (all MMAs of
// the layer have completed by then), which makes it the B operand of layer L + 1, and presets the
// accumulator with layer L + 1's folded-BN bias (tcgen05.st) so that every MMA accumulates.  Only the
// weights stream from L2 (TMA ring, 16 KB per tap and 64-channel half); the block input needed by
// the residual add is parked by each epilogue thread in a thread-private global scratch (coalesced
// 16-byte stores, read back by the same thread); the first layer's input planes are encoded from
// the packed bitboards in-kernel (src/tensor.rs:21-77).  The p_head 1x1 conv is a 15th, one-tap
// "layer" on the same buffers; its epilogue is heads_epilogue() below, the value head is
// value_phase().  Results do not depend on how a batch is split into units (tests: ragged sizes
// are bit-identical to the full batch).
//
// Roles (576 threads): warp 0 weight producer, warp 1 MMA issuer (warp-uniform control flow: UTCHMMA
// takes descriptors from uniform registers), 16 epilogue warps = 4 board teams of 128 threads; a
// thread owns ONE output channel (TMEM lane) of ONE board (8 column groups of 8 pixels).
// Work-item order on every role: for layer: for chain; chain c's epilogue runs under chain 1 - c's MMAs.
// mbarriers: w_full / w_empty (weight ring), tmem_full[c] (MMA -> epilogue), chain_ready[c] (epilogue ->
// MMA: buffer written, accumulator preset).  All waits are bounded (a protocol bug traps).
#pragma once
#include "kernels_simt.cuh"
#include "tower_umma.cuh"

namespace cb {

struct TowerRCfg {
    static constexpr int HID = 128;
    static constexpr int kGroupRows = 9;
    static constexpr int kHalfBytes = 37 * 1024;            // >= (32 * 9 + 1) * 128, 1024-aligned
    static constexpr int kChainBytes = 2 * kHalfBytes;
    static constexpr int kWBytes = 128 * 128;
    static constexpr int kWStages = 4;
    static constexpr int kParts = 4;
    static constexpr int kEpiThreads = 128 * kParts;        // 16 epilogue warps
    static constexpr int kThreads = 64 + kEpiThreads;
    // barriers (256 B) | part [4 boards][128] (+ spare) | hid [4][8] | vpart [4 lane quarters][4 boards][64 pixels]
    static constexpr int kMiscBytes = 256 + 4 * (kParts * 512 + 32 + 1024);
    static constexpr int kSmemBytes = 1024 + 2 * kChainBytes + kWStages * kWBytes + kMiscBytes;
    static constexpr int kScratchPerCta = 2 * 4 * 8 * 128;  // uint4 entries: [chain][board slot][y][channel]
    static_assert(kSmemBytes <= 232448, "shared memory budget (227 KB)");
};

enum { MODE_POLICY = 2 };        // p_head 1x1 conv: one tap, epilogue = softmax / legal-move gather + value head

struct TowerRLayer {
    int w_map;                   // tensor-map table index of this layer's weights
    int kc_in;                   // C_in / 64
    int mode;                    // MODE_RELU / MODE_SE_RES / MODE_POLICY
    int save_res;                // 1: the output is a block input, park it for the residual add
    int ntaps;                   // 9 (3x3 conv) or 1 (1x1 conv: only the centre tap of the tap order)
    int w_row_off;               // added to tap * 128 to form the weight row coordinate (-512 for the 1x1 conv)
    int w_lo_off;                // SPLIT: rows between a weight tile and the tile of its low halves (taps * 128)
    const float* bias;
    const float* se_w1;
    const float* se_w2;
    const float* v_w;
    float* v_pre;                // plain [board][64] layout
};

struct TowerRParams {
    const cb_board* boards;
    const CUtensorMap* maps;     // weights of every layer (conv tower, then p_head), box {64 ch, 128 rows}
    const CUtensorMap* maps_half; // the same tensors with a 64-row box (cluster mode: each CTA loads half a stage)
    const TowerRLayer* layers;
    PolicyOut po;                // policy outputs (full probabilities and / or priors of the legal moves)
    ValueParams vp;              // value head weights and outputs (v_pre is produced and consumed in-kernel)
    int num_layers;              // conv layers + 1 (the policy head)
    int num_boards;
    int num_units;
    uint4* res_scratch;          // [grid][TowerRCfg::kScratchPerCta]
    float* dbg_backbone;         // optional f32 NCHW [board][128][64] dump of the backbone output
    long long* trace;            // debug timeline (CTA 0), or nullptr
    unsigned long long* cta_times;   // debug: [grid][4] globaltimer ns at kernel entry / first MMA / last epilogue / exit
    int dbg;                     // CB_DBG_R bits: 1 no MMA issue, 2 no p_conv store, 4 no residual scratch, 8 no plane bits
};

// boards [b0, b0 + n0) form chain 0 and [b0 + n0, b0 + n0 + n1) chain 1 of a unit
__device__ __forceinline__ void unit_split(const TowerRParams& p, int unit, int& b0, int& n0, int& n1) {
    const int s = static_cast<int>((static_cast<long long>(unit) * p.num_boards) / p.num_units);
    const int e = static_cast<int>((static_cast<long long>(unit + 1) * p.num_boards) / p.num_units);
    b0 = s;
    const int n = e - s;
    // up to 4 boards fit one chain (256 accumulator columns): a single chain pays MMA + epilogue per layer
    // (~7 us) but two small chains pay two issue-bound MMA phases (~8.4 us); from 5 boards on the two chains
    // hide each other's epilogue
    n0 = n <= 4 ? n : (n + 1) >> 1;
    n1 = n - n0;
}

__device__ __forceinline__ uint64_t umma_desc_sw128_sbo(uint32_t smem_addr, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
    d |= static_cast<uint64_t>(1) << 16;
    d |= static_cast<uint64_t>(sbo_bytes >> 4) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(2) << 61;
    return d;
}

// Sum each of 16 per-thread values across the 32 lanes of a warp; lanes l and l + 16 end with the
// total of v[l & 15].
__device__ __forceinline__ float warp_transpose_sum16(float (&v)[16], int lane) {
#pragma unroll
    for (int h = 8; h >= 1; h >>= 1) {
        const bool upper = (lane & h) != 0;
#pragma unroll
        for (int i = 0; i < h; ++i) {
            const float keep = upper ? v[i + h] : v[i];
            const float send = upper ? v[i] : v[i + h];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, h);
        }
    }
    return v[0] + __shfl_xor_sync(0xffffffffu, v[0], 16);
}

// Sum each of 8 per-thread values across the 32 lanes of a warp; every lane ends with the total of v[lane & 7].
__device__ __forceinline__ float warp_transpose_sum8(float (&v)[8], int lane) {
#pragma unroll
    for (int h = 4; h >= 1; h >>= 1) {
        const bool upper = (lane & h) != 0;
#pragma unroll
        for (int i = 0; i < h; ++i) {
            const float keep = upper ? v[i + h] : v[i];
            const float send = upper ? v[i] : v[i + h];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, h);
        }
    }
    float t = v[0] + __shfl_xor_sync(0xffffffffu, v[0], 8);
    return t + __shfl_xor_sync(0xffffffffu, t, 16);
}

// max(lo, 0), max(hi, 0) rounded to the 16-bit activation type, packed (lo in the low half).
template <bool BF16>
__device__ __forceinline__ uint32_t pack_relu16(float lo, float hi) {
    uint32_t d;
    if constexpr (BF16) asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
    else asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
    return d;
}

// lo, hi rounded to fp16 and packed (lo in the low half), no clamp: the low halves of the split mode are signed
__device__ __forceinline__ uint32_t pack_f16x2(float lo, float hi) {
    uint32_t d;
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
    return d;
}
__device__ __forceinline__ float2 unpack_f16x2(uint32_t v) {
    return __half22float2(*reinterpret_cast<const __half2*>(&v));
}

__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

__device__ __forceinline__ void trace_stamp(long long* trace, int item, int slot) {
    if (trace && blockIdx.x == 0) trace[item * 8 + slot] = clock64();
}

template <bool V>
struct BoolC { static constexpr bool v = V; };
template <int V>
struct IntC { static constexpr int v = V; };


__device__ __forceinline__ void team_bar(int id) { asm volatile("bar.sync %0, 128;" ::"r"(id) : "memory"); }

// Policy + value heads of ONE board, run by its 128-thread board team (the four epilogue warps that
// share a board slot; tix = the thread's index in the team = its TMEM lane = the p_head plane).
//   logits[73 planes][64 pixels] (accumulator, bias preset) -> shared memory as the reference's flat
//   [src * 73 + plane] row (python/model.py:120-125) -> full-row log-sum-exp -> all probabilities
//   (src/neural_net.rs:279-284) and / or the legal moves gathered and renormalised in list order with
//   the uniform fallback (src/tensor.rs:184-213);  value head BN(1)+ReLU, cat q, FC 65->256, FC 256->1,
//   k, tanh fusion (python/model.py:128-132,109-110; src/mcts/tactical_mcts.rs:762-765,787-790).
__device__ __forceinline__ void heads_epilogue(const PolicyOut& po, int board, int bs, int nb, int tix, int lane, int quadl,
                                               uint32_t tcol, float* lg, float* s_red, float* s_g) {
    const int bar = 2 + bs;
    float mx = -INFINITY;
#pragma unroll
    for (int y = 0; y < 8; ++y) {
        uint32_t r[8];
        ptx::tmem_ld_32x8(tcol + y * nb * 8, r);
        ptx::tmem_ld_wait();
        if (tix < 73) {
#pragma unroll
            for (int x = 0; x < 8; ++x) {
                const float v = __uint_as_float(r[x]);
                lg[(y * 8 + x) * 73 + tix] = v;
                mx = fmaxf(mx, v);
            }
        }
    }
    ptx::tc_fence_before();
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, d));
    if (lane == 0) s_red[quadl] = mx;
    team_bar(bar);
    const float bmax = fmaxf(fmaxf(s_red[0], s_red[1]), fmaxf(s_red[2], s_red[3]));
    // 4672 = 36 * 128 + 64: four independent partial sums, combined in a fixed order
    float se4[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
    for (int i = 0; i < 36; ++i) se4[i & 3] += __expf(lg[tix + 128 * i] - bmax);
    if (tix < 64) se4[0] += __expf(lg[tix + 128 * 36] - bmax);
    float se = (se4[0] + se4[1]) + (se4[2] + se4[3]);
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) se += __shfl_xor_sync(0xffffffffu, se, d);
    if (lane == 0) s_red[4 + quadl] = se;
    team_bar(bar);
    const float lse = bmax + __logf((s_red[4] + s_red[5]) + (s_red[6] + s_red[7]));
    if (po.probs) {
        float* o = po.probs + static_cast<size_t>(board) * CB_POLICY_SIZE;
        for (int i = tix; i < CB_POLICY_SIZE; i += 128) o[i] = __expf(lg[i] - lse);
    }
    if (po.move_idx) {
        uint32_t lo;
        int n;
        if (po.move_cnt) {
            lo = static_cast<uint32_t>(board) * CB_MAX_MOVES;
            n = po.move_cnt[board];
        } else {
            lo = po.move_off[board];
            n = static_cast<int>(po.move_off[board + 1] - lo);
        }
        // gathered values -> registers (n <= 256 = 2 per thread), total via an ordered sequential sum by one
        // thread, exactly the reference's loop (src/tensor.rs:186-201)
        float g0 = 0.0f, g1 = 0.0f;
        if (tix < n) {
            const int idx = po.move_idx[lo + tix];
            g0 = idx < CB_POLICY_SIZE ? __expf(lg[idx] - lse) : 0.0f;
            s_g[tix] = g0;
        }
        if (tix + 128 < n) {
            const int idx = po.move_idx[lo + tix + 128];
            g1 = idx < CB_POLICY_SIZE ? __expf(lg[idx] - lse) : 0.0f;
            s_g[tix + 128] = g1;
        }
        team_bar(bar);
        if (tix == 0) {
            // strictly sequential adds in list order; the loads are batched so their latency overlaps
            float t = 0.0f;
            int i = 0;
            for (; i + 8 <= n; i += 8) {
                const float4 a = *reinterpret_cast<const float4*>(s_g + i), b = *reinterpret_cast<const float4*>(s_g + i + 4);
                t += a.x; t += a.y; t += a.z; t += a.w; t += b.x; t += b.y; t += b.z; t += b.w;
            }
            for (; i < n; ++i) t += s_g[i];
            s_red[8] = t;
        }
        team_bar(bar);
        const float total = s_red[8];
        if (tix < n) po.priors[lo + tix] = total > 0.0f ? g0 / total : 1.0f / static_cast<float>(n);
        if (tix + 128 < n) po.priors[lo + tix + 128] = total > 0.0f ? g1 / total : 1.0f / static_cast<float>(n);
    }
}

// Value head of every board of the unit (both chains: up to 8 boards), run by the whole epilogue team once both
// chains' v_pre rows are complete.  Thread = (hidden unit j of 256, half of the 65 inputs): every FC weight is
// loaded once per CTA and applied to all 8 boards; the second half's partial sums go through shared memory.
//   BN(1)+ReLU, cat q, FC 65->256 + ReLU, FC 256->1, k, tanh fusion
//   (python/model.py:128-132,109-110; src/mcts/tactical_mcts.rs:762-765,787-790)
// Summation order per board is fixed: (fc_b + inputs 0..32) + inputs 33..64, then a shuffle tree over 32 hidden
// units and the 8 warps in order.
__device__ __forceinline__ void value_phase(const TowerRParams& p, int et, int lane, int b0, int nb0, int nb1, float* s_part,
                                            float* s_xT, float* s_red) {
    const int j = et & 255, half = et >> 8, vw8 = (et >> 5) & 7;
    // x[e][slot]: slot = chain * 4 + board; e < 64 relu(bn(v_conv)), e = 64 the material scalar q; empty slots read 0
    for (int i = et; i < 65 * 8; i += TowerRCfg::kEpiThreads) {
        const int e = i >> 3, slot = i & 7, c = slot >> 2, b = slot & 3;
        float x = 0.0f;
        if (b < (c ? nb1 : nb0)) {
            const int board = b0 + (c ? nb0 : 0) + b;
            if (e < 64) x = fmaxf((__ldcg(p.vp.v_pre + board * 64 + e) + p.vp.conv_bias) * p.vp.bn_scale + p.vp.bn_bias, 0.0f);
            else x = p.vp.q ? p.vp.q[board] : 0.0f;
        }
        s_xT[i] = x;
    }
    ptx::named_bar_sync(1, TowerRCfg::kEpiThreads);
    const float fb = half ? 0.0f : __ldg(p.vp.fc_b + j);
    float h[8] = {fb, fb, fb, fb, fb, fb, fb, fb};
    const int e0 = half ? 33 : 0;
#pragma unroll
    for (int i = 0; i < 33; ++i) {
        const int e = e0 + i;
        if (e < 65) {
            const float w = __ldg(p.vp.fc_wt + e * 256 + j);
            const float4 xa = *reinterpret_cast<const float4*>(s_xT + e * 8), xb = *reinterpret_cast<const float4*>(s_xT + e * 8 + 4);
            h[0] = fmaf(w, xa.x, h[0]); h[1] = fmaf(w, xa.y, h[1]); h[2] = fmaf(w, xa.z, h[2]); h[3] = fmaf(w, xa.w, h[3]);
            h[4] = fmaf(w, xb.x, h[4]); h[5] = fmaf(w, xb.y, h[5]); h[6] = fmaf(w, xb.z, h[6]); h[7] = fmaf(w, xb.w, h[7]);
        }
    }
    if (half) {
#pragma unroll
        for (int s8 = 0; s8 < 8; ++s8) s_part[s8 * 256 + j] = h[s8];
    }
    ptx::named_bar_sync(1, TowerRCfg::kEpiThreads);
    if (!half) {
        const float ow = __ldg(p.vp.out_w + j);
#pragma unroll
        for (int s8 = 0; s8 < 8; ++s8) {
            float t = fmaxf(h[s8] + s_part[s8 * 256 + j], 0.0f) * ow;
#pragma unroll
            for (int d = 16; d >= 1; d >>= 1) t += __shfl_xor_sync(0xffffffffu, t, d);
            if (lane == 0) s_red[vw8 * 8 + s8] = t;
        }
    }
    ptx::named_bar_sync(1, TowerRCfg::kEpiThreads);
    if (et < 8) {
        const int c = et >> 2, b = et & 3;
        if (b < (c ? nb1 : nb0)) {
            float v = p.vp.out_b;
#pragma unroll
            for (int w = 0; w < 8; ++w) v += s_red[w * 8 + et];
            const int board = b0 + (c ? nb0 : 0) + b;
            const float q = p.vp.q ? p.vp.q[board] : 0.0f;
            p.vp.v_logit[board] = v;
            p.vp.k_out[board] = p.vp.k;
            p.vp.v_final[board] = p.vp.material_on ? tanhf(v + p.vp.k * q) : tanhf(v);
        }
    }
}

// DBG = true: honour the CB_DBG_R experiment switches, the timeline stamps and the backbone dump (debug
// entry points only); the product instantiation compiles all of that out of the hot loops.
// CL2 = true (opt-in, CB_CLUSTER=1): launched as clusters of two CTAs (the two SMs of a TPC) that stream the
// weights TOGETHER: each CTA loads half of every weight stage and TMA multicasts it into both CTAs' rings, so the
// L2 -> SM weight traffic — a quarter of the kernel's power draw, and the kernel is power-capped — is halved.  A
// ring slot is refilled when both CTAs' MMAs have released it (commit multicast to both w_empty barriers).
// Requires the two CTAs to walk the same stage sequence (same number of chains, one unit each): the host
// checks.  Measured: +5 % SM clock, but the lock-step on a 4-stage ring costs 10 % more cycles: not the default.
//
// DEEP = true (launches whose units all have ONE chain, i.e. at most 4 boards per CTA: batches of up to 4 boards per
// SM): the second chain's activation buffer is idle, so the weight ring grows into it (128 KB, contiguous with the
// regular ring) and a stage becomes a WHOLE TAP: both 64-channel halves (2 x 16 KB, one barrier, one commit).  What
// bounds such launches is the hand-off of a ring stage, not the stream or the MMAs (tools/exp_stage.cu, one SM alone:
// wait + 4 MMAs + commit per 16 KB costs 349 cycles at N = 64 against an MMA floor of 192; 8 MMAs per barrier cost
// 273 per 16 KB; N >= 192 runs at the MMA floor either way; one SM pulls up to 125 B/clk through TMA).
//
// SPLIT = true (CB_DTYPE_FP32 on the tensor cores): every operand is a PAIR of fp16 numbers, x = x_hi + x_lo with
// x_hi = fp16(x), x_lo = fp16(x - x_hi) (22 significant bits), and a product is three MMAs into the same fp32
// accumulator: W_hi X_hi + W_hi X_lo + W_lo X_hi (the dropped W_lo X_lo term is 2^-22 relative).  The unit is ONE chain of up
// to 4 boards: chain buffer 0 holds X_hi, chain buffer 1 X_lo (same row layout); the weight tensor carries the low halves
// `w_lo_off` rows behind the high ones and streams as two stages per tap and 64-channel half; the epilogue writes both
// halves and parks the block input as a pair.  Measured against the reference's fp32 outputs: see DESIGN section 5.
template <bool BF16, bool DBG, bool CL2, bool DEEP = false, bool SPLIT = false>
__global__ void __launch_bounds__(TowerRCfg::kThreads, 1) tower_r_umma_kernel(const TowerRParams p) {
    static_assert(!SPLIT || (!BF16 && !CL2 && !DEEP), "the split mode is fp16 pairs on the plain single-CTA ring");
    const int dbg = DBG ? p.dbg : 0;
    long long* const trace = DBG ? p.trace : nullptr;
    unsigned long long* const cta_times = DBG ? p.cta_times : nullptr;
    float* const dbg_backbone = DBG ? p.dbg_backbone : nullptr;
    using Cfg = TowerRCfg;
    constexpr int HID = 128;
    constexpr int kWStages = Cfg::kWStages, kSE = 8;
    constexpr int kStageBytes = DEEP ? 2 * Cfg::kWBytes : Cfg::kWBytes;   // DEEP: a stage is a whole tap (both halves)
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* chain_base = smem;                                // [2][2 halves][kHalfBytes]
    uint8_t* w_base = chain_base + 2 * Cfg::kChainBytes - (DEEP ? Cfg::kWStages * Cfg::kWBytes : 0);
    uint8_t* misc = w_base + kWStages * kStageBytes;
    static_assert(Cfg::kChainBytes >= Cfg::kWStages * Cfg::kWBytes, "the deep ring lives in the idle chain's buffer");
    uint64_t* w_full = reinterpret_cast<uint64_t*>(misc);      // [kWStages]
    uint64_t* w_empty = w_full + kWStages;                     // [kWStages]
    uint64_t* tmem_full = w_empty + kWStages;                  // [2 chains]   (20 barriers + the TMEM pointer < 256 B)
    uint64_t* chain_ready = tmem_full + 2;                     // [2 chains] buffer written + TMEM drained
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(chain_ready + 2);
    float* s_part = reinterpret_cast<float*>(misc + 256);      // [kParts][4 boards][128 ch]
    float* s_hid = s_part + Cfg::kParts * 512;                 // [4][8]
    float* s_vpart = s_hid + 32;                               // [4 lane quarters][256 pixels]

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    if (cta_times && threadIdx.x == 0) cta_times[blockIdx.x * 4 + 0] = globaltimer_ns();
    // tap order: the first tap of a layer must cover every accumulator column (dy = 0)
    constexpr int kTapOrder[9] = {4, 3, 5, 1, 0, 2, 7, 6, 8};
    constexpr unsigned long long kTapOrderPacked = 0x867201534ULL;   // the same, one nibble per tap (runtime-indexed uses)

    if (threadIdx.x == 0) {
        // this CTA is resident: a kernel launched as a programmatic dependent (the tree kernel of the other game
        // pool, cb_selfplay_device with two pools) may now take the SMs this grid leaves free; no-op otherwise
        ptx::grid_launch_dependents();
        for (int s = 0; s < kWStages; ++s) { ptx::mbar_init(&w_full[s], 1); ptx::mbar_init(&w_empty[s], CL2 ? 2 : 1); }
        for (int c = 0; c < 2; ++c) {
            ptx::mbar_init(&tmem_full[c], 1);
            ptx::mbar_init(&chain_ready[c], Cfg::kEpiThreads / 32);   // one arrival per epilogue warp
        }
        ptx::fence_barrier_init();
    }
    if (warp == 1) ptx::tmem_alloc<512>(tmem_ptr);
    ptx::tc_fence_before();
    __syncthreads();
    if constexpr (CL2) ptx::cluster_sync_all();      // the peer's barriers are initialised before anything targets them
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0) {
        // ------------------------------------------------------ weight producer
        int ws = 0;
        uint32_t wph = 0;
        for (int unit = blockIdx.x; unit < p.num_units; unit += gridDim.x) {
            int b0, nbc[2];
            unit_split(p, unit, b0, nbc[0], nbc[1]);
            for (int L = 0; L < p.num_layers; ++L) {
                const TowerRLayer* ly = p.layers + L;
                const CUtensorMap* tm_w = p.maps + ly->w_map;
                const int kc_in = ly->kc_in, ntaps = ly->ntaps, w_row_off = ly->w_row_off;
                for (int c = 0; c < 2; ++c) {
                    if (nbc[c] == 0) continue;
#pragma unroll 1
                    for (int t = 0; t < ntaps; ++t) {
                        const int tap = static_cast<int>((kTapOrderPacked >> (4 * t)) & 15u);
                        if constexpr (DEEP) {
                            ptx::mbar_wait(&w_empty[ws], wph ^ 1);
                            if (ptx::elect_one()) {
                                if (dbg & 16) {
                                    ptx::mbar_arrive(&w_full[ws]);
                                } else {
                                    ptx::mbar_arrive_expect_tx(&w_full[ws], kc_in * Cfg::kWBytes);
                                    for (int kc = 0; kc < kc_in; ++kc)
                                        ptx::tma_load_2d(w_base + ws * kStageBytes + kc * Cfg::kWBytes, tm_w, &w_full[ws], kc * 64, tap * HID + w_row_off);
                                }
                            }
                            __syncwarp();
                            if (++ws == kWStages) { ws = 0; wph ^= 1; }
                        } else
                        for (int kc = 0; kc < kc_in; ++kc) {
                            ptx::mbar_wait(&w_empty[ws], wph ^ 1);
                            if (ptx::elect_one()) {
                                if (dbg & 16) {
                                    ptx::mbar_arrive(&w_full[ws]);
                                } else if constexpr (CL2) {
                                    // my 64 rows of the stage, delivered to both CTAs; the other 64 arrive from the peer
                                    const int r64 = static_cast<int>(ptx::cluster_ctarank()) * 64;
                                    ptx::mbar_arrive_expect_tx(&w_full[ws], Cfg::kWBytes);
                                    ptx::tma_load_2d_mc(w_base + ws * Cfg::kWBytes + r64 * 128, p.maps_half + ly->w_map, &w_full[ws],
                                                        kc * 64, tap * HID + w_row_off + r64, 3);
                                } else {
                                    ptx::mbar_arrive_expect_tx(&w_full[ws], Cfg::kWBytes);
                                    ptx::tma_load_2d(w_base + ws * Cfg::kWBytes, tm_w, &w_full[ws], kc * 64, tap * HID + w_row_off);
                                }
                            }
                            __syncwarp();
                            if (++ws == kWStages) { ws = 0; wph ^= 1; }
                            if constexpr (SPLIT) {
                                ptx::mbar_wait(&w_empty[ws], wph ^ 1);
                                if (ptx::elect_one()) {
                                    ptx::mbar_arrive_expect_tx(&w_full[ws], Cfg::kWBytes);
                                    ptx::tma_load_2d(w_base + ws * Cfg::kWBytes, tm_w, &w_full[ws], kc * 64,
                                                     tap * HID + w_row_off + ly->w_lo_off);
                                }
                                __syncwarp();
                                if (++ws == kWStages) { ws = 0; wph ^= 1; }
                            }
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // -------------------------------------------------------- MMA issuer
        // Warp-uniform control flow and compile-time taps keep descriptors / idesc / TMEM addresses in
        // uniform registers (UTCHMMA takes them from there: a divergent issuer pays 6 R2UR per MMA, which
        // caps the issue rate at ~150 cycles per MMA); the MMAs themselves are issued by one elected lane.
        int ws = 0;
        uint32_t wph = 0, ready_ph0 = 0, ready_ph1 = 0;
        // descriptor halves: low = start address >> 4 | LBO field, high = SBO | version | SWIZZLE_128B
        const uint64_t a_desc0 = ptx::umma_desc_sw128(ptx::smem_u32(w_base));
        const uint64_t b_desc0 = umma_desc_sw128_sbo(ptx::smem_u32(chain_base), Cfg::kGroupRows * 128);
        const uint32_t a_lo0 = static_cast<uint32_t>(a_desc0), a_hi = static_cast<uint32_t>(a_desc0 >> 32);
        const uint32_t b_lo0 = static_cast<uint32_t>(b_desc0), b_hi = static_cast<uint32_t>(b_desc0 >> 32);
        for (int unit = blockIdx.x; unit < p.num_units; unit += gridDim.x) {
            int b0, nbc0, nbc1;
            unit_split(p, unit, b0, nbc0, nbc1);
            for (int L = 0; L < p.num_layers; ++L) {
                const int kc_in = p.layers[L].kc_in, ntaps = p.layers[L].ntaps;
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    const int nb = c ? nbc1 : nbc0;
                    if (nb == 0) continue;
                    const int halo = nb & 1;
                    if (c == 0) { ptx::mbar_wait(&chain_ready[0], ready_ph0); ready_ph0 ^= 1; }
                    else { ptx::mbar_wait(&chain_ready[1], ready_ph1); ready_ph1 ^= 1; }
                    ptx::tc_fence_after();
                    const uint32_t b_lo_c = b_lo0 + c * (Cfg::kChainBytes >> 4);
                    const uint32_t d_tmem = tmem_base + c * 256;
                    if (lane == 0) trace_stamp(trace, L * 2 + c, 2);
                    if (cta_times && lane == 0 && L == 0 && c == 0 && unit == blockIdx.x) cta_times[blockIdx.x * 4 + 1] = globaltimer_ns();
#pragma unroll
                    for (int t = 0; t < 9; ++t) {
                        if (t >= ntaps) break;
                        const int tap = kTapOrder[t];
                        const int dy = tap / 3 - 1, dx = tap % 3 - 1;
                        int start_row, n, dcol;
                        if (halo) {
                            start_row = (1 + dy) * nb * 9 + 1 + dx;
                            n = 64 * nb;
                            dcol = 0;
                        } else {
                            start_row = (dy > 0 ? nb * 9 : 0) + 1 + dx;
                            n = dy == 0 ? 64 * nb : 56 * nb;
                            dcol = dy < 0 ? 8 * nb : 0;
                        }
                        const uint32_t idesc = ptx::umma_idesc_f16(BF16 ? 1 : 0, 128, n);
                        const uint32_t b_lo_t = b_lo_c + start_row * 8;      // 128-byte rows in 16-byte units
                        const uint32_t d_t = d_tmem + dcol;
                        if constexpr (DEEP) {
                            ptx::mbar_wait(w_full + ws, wph);
                            if (ptx::elect_one()) {
                                for (int kc = 0; kc < kc_in && !(dbg & 1); ++kc) {
                                    const uint32_t a_lo = a_lo0 + ws * (kStageBytes >> 4) + kc * (Cfg::kWBytes >> 4);
                                    const uint32_t b_lo = b_lo_t + kc * (Cfg::kHalfBytes >> 4);
#pragma unroll
                                    for (int k = 0; k < 4; ++k)
                                        ptx::umma_f16_lohi(d_t, a_lo + 2 * k, a_hi, b_lo + 2 * k, b_hi, idesc, 1u);   // bias preset
                                }
                                ptx::umma_commit(w_empty + ws);
                            }
                            __syncwarp();
                            if (++ws == kWStages) { ws = 0; wph ^= 1; }
                        } else
                        for (int kc = 0; kc < kc_in; ++kc) {
                            ptx::mbar_wait(w_full + ws, wph);
                            if (ptx::elect_one()) {
                                const uint32_t a_lo = a_lo0 + ws * (Cfg::kWBytes >> 4);
                                const uint32_t b_lo = b_lo_t + kc * (Cfg::kHalfBytes >> 4);
                                if (!(dbg & 1)) {
#pragma unroll
                                    for (int k = 0; k < 4; ++k)
                                        ptx::umma_f16_lohi(d_t, a_lo + 2 * k, a_hi, b_lo + 2 * k, b_hi, idesc, 1u);   // bias preset
                                    if constexpr (SPLIT) {   // W_hi X_lo: the low halves live in chain buffer 1
#pragma unroll
                                        for (int k = 0; k < 4; ++k)
                                            ptx::umma_f16_lohi(d_t, a_lo + 2 * k, a_hi, b_lo + (Cfg::kChainBytes >> 4) + 2 * k, b_hi, idesc, 1u);
                                    }
                                }
                                if constexpr (CL2) ptx::umma_commit_mc(w_empty + ws, 3);   // release the slot in both CTAs
                                else ptx::umma_commit(w_empty + ws);
                            }
                            __syncwarp();
                            if (++ws == kWStages) { ws = 0; wph ^= 1; }
                            if constexpr (SPLIT) {       // W_lo X_hi
                                ptx::mbar_wait(w_full + ws, wph);
                                if (ptx::elect_one()) {
                                    const uint32_t a_lo = a_lo0 + ws * (Cfg::kWBytes >> 4);
                                    const uint32_t b_lo = b_lo_t + kc * (Cfg::kHalfBytes >> 4);
#pragma unroll
                                    for (int k = 0; k < 4; ++k)
                                        ptx::umma_f16_lohi(d_t, a_lo + 2 * k, a_hi, b_lo + 2 * k, b_hi, idesc, 1u);
                                    ptx::umma_commit(w_empty + ws);
                                }
                                __syncwarp();
                                if (++ws == kWStages) { ws = 0; wph ^= 1; }
                            }
                        }
                    }
                    if (ptx::elect_one()) ptx::umma_commit(&tmem_full[c]);
                    __syncwarp();
                    if (lane == 0) trace_stamp(trace, L * 2 + c, 3);
                }
            }
        }
    } else {
        // ---------------------------------------------------------- epilogue (16 warps)
        // Thread = (output channel ch = TMEM lane, board slot bs = part): it owns the 8 rows y = 0..7 of
        // one board, i.e. 8 groups of 8 accumulator columns.  Board sums for the squeeze-excite block are
        // therefore thread-local and summed in a fixed order (y, then x), independent of how many boards
        // the chain holds or where a board sits in the batch.
        constexpr int kEpi = Cfg::kEpiThreads;
        const int et = threadIdx.x - 64;          // 0..kEpi-1
        const int ew = et >> 5;                   // 0..15
        const int quadl = warp & 3;               // TMEM lane quarter this warp may touch
        const int bs = ew >> 2;                   // board slot inside the chain
        const int ch = quadl * 32 + lane;         // my output channel == TMEM lane
        const uint32_t kc_off = static_cast<uint32_t>(ch >> 6) * Cfg::kHalfBytes;
        const uint32_t cq = static_cast<uint32_t>((ch & 63) >> 3);
        const uint32_t cb2 = static_cast<uint32_t>(ch & 7) * 2;
        const bool odd = (lane & 1) != 0;
        // byte selectors that merge my half-word with my pair partner's, even channel in the low half
        const uint32_t sel_lo = odd ? 0x1054u : 0x5410u, sel_hi = odd ? 0x3276u : 0x7632u;
        uint4* my_res = p.res_scratch + static_cast<size_t>(blockIdx.x) * Cfg::kScratchPerCta + bs * 8 * HID + ch;
        uint32_t full_ph = 0;                     // bit c: phase of tmem_full[c]
        for (int unit = blockIdx.x; unit < p.num_units; unit += gridDim.x) {
            int b0, nbc0, nbc1;
            unit_split(p, unit, b0, nbc0, nbc1);
            // ---- input planes of both chains, straight from the packed bitboards
            // (the heads of the previous unit used the chain buffers as logit storage)
            if (unit != static_cast<int>(blockIdx.x)) ptx::named_bar_sync(1, kEpi);
#pragma unroll 1
            for (int c = 0; c < 2; ++c) {
                const int nb = c ? nbc1 : nbc0;
                if (nb == 0) continue;
                const int halo = nb & 1;
                const int groups = (8 + 2 * halo) * nb;
                const int cb0 = b0 + (c ? nbc0 : 0);
                uint8_t* buf = chain_base + c * Cfg::kChainBytes;
                for (int r = et; r <= groups * 9; r += kEpi) {
                    const int g = r / 9, xr = r - g * 9;
                    const int yy = g / nb - halo, b = g - (g / nb) * nb;
                    const bool data = xr >= 1 && g < groups && yy >= 0 && yy < 8;
                    uint32_t msk = 0;
                    if (data && !(dbg & 8)) msk = pixel_plane_mask(p.boards + cb0 + b, yy, xr - 1);
                    uint4* row0 = reinterpret_cast<uint4*>(buf + r * 128);
                    uint4* row1 = reinterpret_cast<uint4*>(buf + Cfg::kHalfBytes + r * 128);
                    const uint32_t one = BF16 ? 0x3F80u : 0x3C00u;
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        uint4 v = make_uint4(0, 0, 0, 0);
                        if (i < 3) {
                            const uint32_t m8 = msk >> (8 * i);
                            v.x = ((m8 >> 0) & 1u ? one : 0u) | ((m8 >> 1) & 1u ? (one << 16) : 0u);
                            v.y = ((m8 >> 2) & 1u ? one : 0u) | ((m8 >> 3) & 1u ? (one << 16) : 0u);
                            v.z = ((m8 >> 4) & 1u ? one : 0u) | ((m8 >> 5) & 1u ? (one << 16) : 0u);
                            v.w = ((m8 >> 6) & 1u ? one : 0u) | ((m8 >> 7) & 1u ? (one << 16) : 0u);
                        }
                        row0[i ^ (r & 7)] = v;
                    }
                    if (!data) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) row1[i ^ (r & 7)] = make_uint4(0, 0, 0, 0);   // (rotated: no bank conflicts)
                    }
                    if constexpr (SPLIT) {   // X_lo: zero planes (0 / 1 are exact), zero padding rows
                        uint4* lo0 = reinterpret_cast<uint4*>(buf + Cfg::kChainBytes + r * 128);
                        uint4* lo1 = reinterpret_cast<uint4*>(buf + Cfg::kChainBytes + Cfg::kHalfBytes + r * 128);
#pragma unroll
                        for (int i = 0; i < 8; ++i) lo0[i ^ (r & 7)] = make_uint4(0, 0, 0, 0);
                        if (!data) {
#pragma unroll
                            for (int i = 0; i < 8; ++i) lo1[i ^ (r & 7)] = make_uint4(0, 0, 0, 0);
                        }
                    }
                }
                // accumulator preset: the folded-BN bias of the first layer (every MMA accumulates)
                if (bs < nb) {
                    const uint32_t bias0 = __float_as_uint(__ldg(p.layers[0].bias + ch));
                    const uint32_t tcol = tmem_base + (static_cast<uint32_t>(quadl * 32) << 16) + c * 256 + bs * 8;
#pragma unroll
                    for (int y = 0; y < 8; ++y) ptx::tmem_st_32x8_bcast(tcol + y * nb * 8, bias0);
                    ptx::tmem_st_wait();
                }
                ptx::tc_fence_before();
                ptx::fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(&chain_ready[c]);
            }

            for (int L = 0; L < p.num_layers; ++L) {
                const TowerRLayer* ly = p.layers + L;
                const int mode = ly->mode;
                const bool last_layer = L == p.num_layers - 1;   // the policy head (MODE_POLICY)
#pragma unroll 1
                for (int c = 0; c < 2; ++c) {
                    const int nb = c ? nbc1 : nbc0;
                    if (nb == 0) continue;
                    const int halo = nb & 1;
                    // Thread -> (board slot bsx, rows [y0, y0 + ny)).  Normally a thread owns all 8 rows of board slot bs.
                    // DEEP (single-chain units, the small-batch launches, where nothing hides the epilogue): with one or
                    // two boards in the chain the board teams that would idle take a share of the rows instead
                    // (1 board: 2 rows per team; 2 boards: 4 rows).  Sums stay in the full-board order (see pass 1).
                    int bsx = bs, y0 = 0, ny = 8;
                    if constexpr (DEEP) {
                        if (nb == 1) { bsx = 0; y0 = 2 * bs; ny = 2; }
                        else if (nb == 2) { bsx = bs & 1; y0 = 4 * (bs >> 1); ny = 4; }
                    }
                    const bool rowsplit = DEEP && nb <= 2;
                    const bool active = bsx < nb;     // this thread's board slot exists in the chain
                    const int cb0 = b0 + (c ? nbc0 : 0);
                    uint8_t* buf = chain_base + c * Cfg::kChainBytes;
                    uint4* const my_res_x = my_res + (bsx - bs) * 8 * HID;
                    constexpr int kFc1 = 32 / (kEpi / 32);   // squeeze-FC outputs per warp
                    float w1r[kFc1][4], w2r[kSE];
                    uint4 rr[8];
                    if (mode == MODE_SE_RES) {
#pragma unroll
                        for (int i = 0; i < kFc1; ++i)
#pragma unroll
                            for (int k = 0; k < 4; ++k)
                                w1r[i][k] = __ldg(&ly->se_w1[((ew * kFc1 + i) & 7) * HID + lane + 32 * k]);
#pragma unroll
                        for (int j = 0; j < kSE; ++j) w2r[j] = __ldg(&ly->se_w2[ch * kSE + j]);
                        // block input of my (channel, board): parked by this same thread earlier
                        if (active && !(dbg & 4)) {
#pragma unroll
                            for (int yy = 0; yy < 8; ++yy) {
                                if (yy >= ny) break;
                                rr[yy] = __ldcg(my_res_x + (c * 32 + y0 + yy) * HID);
                            }
                        }
                    }
                    const float bias_next = last_layer ? 0.0f : __ldg(p.layers[L + 1].bias + ch);   // p_head: padded to 128
                    // (everything the epilogue needs from the layer table is read BEFORE the wait: behind it, each of
                    // these loads is an L2 round trip in front of the next layer's first MMA)
                    const bool want_v = (mode == MODE_SE_RES) && ly->v_pre != nullptr;
                    const float vw = want_v ? __ldg(ly->v_w + ch) : 0.0f;
                    const bool save_res = ly->save_res != 0 && !(dbg & 4);
                    if (dbg & 64) ptx::mbar_wait(&tmem_full[c], (full_ph >> c) & 1u);
                    else ptx::mbar_wait_parked(&tmem_full[c], (full_ph >> c) & 1u);
                    full_ph ^= 1u << c;
                    ptx::tc_fence_after();
                    if (et == 0) trace_stamp(trace, L * 2 + c, 4);
                    // my board's 8 column groups: column (y * nb + bs) * 8
                    const uint32_t tcol_p = tmem_base + (static_cast<uint32_t>(quadl * 32) << 16) + c * 256 + bs * 8;
                    const uint32_t tcol = tcol_p + (bsx - bs) * 8;
                    if (mode == MODE_POLICY) {
                        if (bs < nb)     // a whole board team (128 threads) per board
                            heads_epilogue(p.po, cb0 + bs, bs, nb, ch, lane, quadl, tcol_p,
                                           reinterpret_cast<float*>(buf) + bs * CB_POLICY_SIZE, s_part + 512 + bs * 32,
                                           s_vpart + bs * 256);
                        if (et == 0) trace_stamp(trace, L * 2 + c, 5);
                        continue;
                    }

                    float scale = 1.0f;
                    if (mode == MODE_SE_RES) {
                        // pass 1: my channel's sum over my board (the bias is already in the accumulator)
                        float sb = 0.0f;
                        if (active) {
#pragma unroll(DEEP ? 1 : 4)
                            for (int yy = 0; yy < ny; yy += 2) {
                                const int y = y0 + yy;
                                uint32_t r0[8], r1[8];
                                ptx::tmem_ld_32x8(tcol + y * nb * 8, r0);
                                ptx::tmem_ld_32x8(tcol + (y + 1) * nb * 8, r1);
                                ptx::tmem_ld_wait();
                                float g0 = 0.0f, g1 = 0.0f;
#pragma unroll
                                for (int x = 0; x < 8; ++x) { g0 += __uint_as_float(r0[x]); g1 += __uint_as_float(r1[x]); }
                                if (rowsplit) {
                                    // row sums go out one by one: the board sum is formed below in the order y = 0..7
                                    s_part[(bsx * 8 + y) * HID + ch] = g0;
                                    s_part[(bsx * 8 + y + 1) * HID + ch] = g1;
                                } else {
                                    sb += g0;
                                    sb += g1;
                                }
                            }
                        }
                        if (!rowsplit) s_part[bs * HID + ch] = sb;
                        ptx::named_bar_sync(1, kEpi);
                        if (et == 0) trace_stamp(trace, L * 2 + c, 0);
#pragma unroll
                        for (int i = 0; i < kFc1; ++i) {
                            const int o = ew * kFc1 + i, b = o >> 3, j = o & 7;
                            float s = 0.0f;
                            if (rowsplit) {
                                if (b < nb) {
#pragma unroll
                                    for (int k = 0; k < 4; ++k) {
                                        float cs = 0.0f;
#pragma unroll
                                        for (int y = 0; y < 8; ++y) cs += s_part[(b * 8 + y) * HID + lane + 32 * k];
                                        s = fmaf(w1r[i][k], cs, s);
                                    }
                                }
                            } else
#pragma unroll
                            for (int k = 0; k < 4; ++k) s = fmaf(w1r[i][k], s_part[b * HID + lane + 32 * k], s);
#pragma unroll
                            for (int d = 16; d >= 1; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
                            if (lane == 0) s_hid[b * kSE + j] = fmaxf(s * (1.0f / 64.0f), 0.0f);
                        }
                        ptx::named_bar_sync(1, kEpi);
                        if (et == 0) trace_stamp(trace, L * 2 + c, 1);
                        float s = 0.0f;
#pragma unroll
                        for (int j = 0; j < kSE; ++j) s = fmaf(w2r[j], s_hid[bsx * kSE + j], s);
                        scale = 1.0f / (1.0f + __expf(-s));
                    }
                    // main pass: (SE scale + residual), ReLU, 16-bit, transpose-store into the chain buffer
                    const bool dump = want_v && dbg_backbone != nullptr;
                    // The main pass is instantiated per (SE + residual, park the output, value conv) combination so that
                    // its 8 x 60 instructions carry no per-group runtime branches (the epilogue warps share their
                    // schedulers with the MMA issuer).
                    auto main_pass = [&](auto se_c, auto save_c, auto wantv_c, auto ny_c) {
                        constexpr bool kSe = decltype(se_c)::v, kSave = decltype(save_c)::v, kWantV = decltype(wantv_c)::v;
                        constexpr int kNy = decltype(ny_c)::v;     // rows per thread: compile-time, so that a short pass is short CODE
                                                                   // (the epilogue of a single-chain unit runs from a cold instruction cache)
                        const uint32_t buf_ch = ptx::smem_u32(buf) + kc_off + (cb2 & ~2u);   // the even channel of my pair
                        // SPLIT: the low halves of the block input, fetched one row ahead (the high halves sit in rr)
                        uint4 rl = make_uint4(0, 0, 0, 0);
                        if constexpr (SPLIT && kSe) rl = __ldcg(my_res_x + (32 + y0) * HID);
#pragma unroll
                        for (int yy = 0; yy < kNy; ++yy) {
                            const int y = y0 + yy;
                            [[maybe_unused]] const uint4 rl_cur = rl;
                            if constexpr (SPLIT && kSe) { if (yy + 1 < kNy) rl = __ldcg(my_res_x + (32 + y + 1) * HID); }
                            uint32_t r[8];
                            ptx::tmem_ld_32x8(tcol + y * nb * 8, r);
                            ptx::tmem_ld_wait();
                            if (DBG && !kSe && yy == 0 && et == 0) trace_stamp(trace, L * 2 + c, 0);
                            float v[8];
#pragma unroll
                            for (int x = 0; x < 8; ++x) v[x] = __uint_as_float(r[x]);
                            if constexpr (kSe) {
                                const uint32_t w[4] = {rr[yy].x, rr[yy].y, rr[yy].z, rr[yy].w};
                                if constexpr (SPLIT) {
                                    const uint32_t wl[4] = {rl_cur.x, rl_cur.y, rl_cur.z, rl_cur.w};
#pragma unroll
                                    for (int q = 0; q < 4; ++q) {
                                        const float2 a = unpack_f16x2(w[q]), b = unpack_f16x2(wl[q]);
                                        v[2 * q] = fmaf(v[2 * q], scale, a.x + b.x);
                                        v[2 * q + 1] = fmaf(v[2 * q + 1], scale, a.y + b.y);
                                    }
                                } else {
#pragma unroll
                                    for (int x = 0; x < 8; ++x) {
                                        const uint32_t h = (x & 1) ? (w[x >> 1] >> 16) : (w[x >> 1] & 0xFFFFu);
                                        v[x] = fmaf(v[x], scale, from16bits<BF16>(static_cast<uint16_t>(h)));
                                    }
                                }
                            }
                            uint32_t pk[4];
#pragma unroll
                            for (int q = 0; q < 4; ++q) pk[q] = pack_relu16<BF16>(v[2 * q], v[2 * q + 1]);
                            uint32_t pl[4] = {0, 0, 0, 0};    // SPLIT: what the fp16 rounding of max(v, 0) left over
                            if constexpr (SPLIT) {
#pragma unroll
                                for (int q = 0; q < 4; ++q) {
                                    const float2 h2 = unpack_f16x2(pk[q]);
                                    pl[q] = pack_f16x2(fmaxf(v[2 * q], 0.0f) - h2.x, fmaxf(v[2 * q + 1], 0.0f) - h2.y);
                                }
                            }
                            // rows of group (y, bs): ((y + halo) * nb + bs) * 9 + 1 + x, swizzled by absolute address.
                            // Lane pairs (channels ch, ch ^ 1) trade halves of the group so that every store is one
                            // 32-bit word = two adjacent channels of one pixel: the even lane writes pixels 0..3, the
                            // odd lane pixels 4..7 (rows 4 apart flip the swizzle half: no bank conflicts).
                            const uint32_t row0 = static_cast<uint32_t>(((y + halo) * nb + bsx) * 9 + 1) + (odd ? 4u : 0u);
                            const uint32_t o0 = __shfl_xor_sync(0xffffffffu, odd ? pk[0] : pk[2], 1);
                            const uint32_t o1 = __shfl_xor_sync(0xffffffffu, odd ? pk[1] : pk[3], 1);
                            const uint32_t m0 = odd ? pk[2] : pk[0], m1 = odd ? pk[3] : pk[1];
                            const uint32_t wd[4] = {__byte_perm(m0, o0, sel_lo), __byte_perm(m0, o0, sel_hi),
                                                    __byte_perm(m1, o1, sel_lo), __byte_perm(m1, o1, sel_hi)};
#pragma unroll
                            for (int x = 0; x < 4; ++x) {
                                const uint32_t row = row0 + x;
                                const uint32_t addr = buf_ch + row * 128 + ((cq ^ (row & 7)) << 4);
                                asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(wd[x]) : "memory");
                            }
                            if constexpr (SPLIT) {   // the same transpose-store for the low halves, into chain buffer 1
                                const uint32_t p0 = __shfl_xor_sync(0xffffffffu, odd ? pl[0] : pl[2], 1);
                                const uint32_t p1 = __shfl_xor_sync(0xffffffffu, odd ? pl[1] : pl[3], 1);
                                const uint32_t n0 = odd ? pl[2] : pl[0], n1 = odd ? pl[3] : pl[1];
                                const uint32_t wl[4] = {__byte_perm(n0, p0, sel_lo), __byte_perm(n0, p0, sel_hi),
                                                        __byte_perm(n1, p1, sel_lo), __byte_perm(n1, p1, sel_hi)};
#pragma unroll
                                for (int x = 0; x < 4; ++x) {
                                    const uint32_t row = row0 + x;
                                    const uint32_t addr = buf_ch + Cfg::kChainBytes + row * 128 + ((cq ^ (row & 7)) << 4);
                                    asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(wl[x]) : "memory");
                                }
                            }
                            if constexpr (kSave) {
                                __stcg(my_res_x + (c * 32 + y) * HID, make_uint4(pk[0], pk[1], pk[2], pk[3]));
                                if constexpr (SPLIT) __stcg(my_res_x + (32 + y) * HID, make_uint4(pl[0], pl[1], pl[2], pl[3]));
                            }
                            if constexpr (kWantV) {
                                // value 1x1 conv: sum over this warp's 32 channels for each of the 8 pixels
                                float tv[8];
#pragma unroll
                                for (int x = 0; x < 8; ++x) tv[x] = fmaxf(v[x], 0.0f) * vw;
                                const float t = warp_transpose_sum8(tv, lane);
                                if (lane < 8) s_vpart[quadl * 256 + bsx * 64 + y * 8 + lane] = t;
                            }
                            if (dump) {
#pragma unroll
                                for (int x = 0; x < 8; ++x) {
                                    const uint16_t h16 = static_cast<uint16_t>((x & 1) ? (pk[x >> 1] >> 16) : (pk[x >> 1] & 0xFFFFu));
                                    const uint16_t l16 = static_cast<uint16_t>((x & 1) ? (pl[x >> 1] >> 16) : (pl[x >> 1] & 0xFFFFu));
                                    dbg_backbone[(static_cast<size_t>(cb0 + bsx) * HID + ch) * 64 + y * 8 + x] =
                                        from16bits<BF16>(h16) + (SPLIT ? from16bits<false>(l16) : 0.0f);
                                }
                            }
                            // accumulator preset for the next layer of this chain
                            ptx::tmem_st_32x8_bcast(tcol + y * nb * 8, __float_as_uint(bias_next));
                            if (DBG && !kSe && yy == 0 && et == 0) trace_stamp(trace, L * 2 + c, 1);
                        }
                        if (DBG && et == 0) trace_stamp(trace, L * 2 + c, 7);
                        ptx::tmem_st_wait();
                    };
                    auto main_pass_rows = [&](auto ny_c) {
                        if (mode == MODE_SE_RES) {
                            if (want_v) main_pass(BoolC<true>{}, BoolC<false>{}, BoolC<true>{}, ny_c);
                            else if (save_res) main_pass(BoolC<true>{}, BoolC<true>{}, BoolC<false>{}, ny_c);
                            else main_pass(BoolC<true>{}, BoolC<false>{}, BoolC<false>{}, ny_c);
                        } else {
                            if (save_res) main_pass(BoolC<false>{}, BoolC<true>{}, BoolC<false>{}, ny_c);
                            else main_pass(BoolC<false>{}, BoolC<false>{}, BoolC<false>{}, ny_c);
                        }
                    };
                    if (active && !(dbg & 32)) {
                        if constexpr (DEEP) {
                            if (ny == 2) main_pass_rows(IntC<2>{});
                            else if (ny == 4) main_pass_rows(IntC<4>{});
                            else main_pass_rows(IntC<8>{});
                        } else {
                            main_pass_rows(IntC<8>{});
                        }
                    }
                    if (et == 0) trace_stamp(trace, L * 2 + c, 6);
                    ptx::tc_fence_before();
                    ptx::fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) ptx::mbar_arrive(&chain_ready[c]);
                    if (et == 0) trace_stamp(trace, L * 2 + c, 5);
                    if (!DEEP && L == p.num_layers - 2 && c == 0) {
                        // both chains' v_pre rows are complete (last SE layer): value heads of the whole unit,
                        // hidden under chain 1's p_conv MMAs
                        // (scratch: the SE partial sums and the v_conv staging are dead by now)
                        value_phase(p, et, lane, b0, nbc0, nbc1, s_part, s_vpart, s_vpart + 528);
                    }
                    if (want_v) {
                        ptx::named_bar_sync(1, kEpi);
                        // fixed summation order over the four lane quarters keeps the result independent of batch composition
                        if (et < 64 * nb) {
                            const int n = et;     // board slot n >> 6, pixel n & 63
                            const float v = (s_vpart[n] + s_vpart[256 + n]) + (s_vpart[512 + n] + s_vpart[768 + n]);
                            ly->v_pre[(cb0 + (n >> 6)) * 64 + (n & 63)] = v;
                        }
                    }
                    // s_part / s_hid / s_vpart of this item are dead for every thread before the next item writes them
                    if (mode == MODE_SE_RES) ptx::named_bar_sync(1, kEpi);
                    // single-chain units: the value heads run right behind the value conv, under the p_conv MMAs
                    if (DEEP && want_v) {
                        value_phase(p, et, lane, b0, nbc0, nbc1, s_part, s_vpart, s_vpart + 528);
                        ptx::named_bar_sync(1, kEpi);   // its scratch is read to the end before the heads reuse it
                    }
                }
            }
        }
        if (cta_times && et == 0) cta_times[blockIdx.x * 4 + 2] = globaltimer_ns();
    }
    ptx::tc_fence_before();
    __syncthreads();
    if constexpr (CL2) ptx::cluster_sync_all();      // no CTA leaves while its peer may still signal its barriers
    if (warp == 1) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc<512>(tmem_base);
    }
    if (cta_times && threadIdx.x == 0) cta_times[blockIdx.x * 4 + 3] = globaltimer_ns();
}

}  // namespace cb
