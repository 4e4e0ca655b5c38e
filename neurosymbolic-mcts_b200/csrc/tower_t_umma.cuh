// tower_t_umma.cuh — fused persistent conv tower for the 128-channel net, TRANSPOSED GEMM:
//
//     D^T[c_out = 128 lanes][pixels = 256 columns] += W_tap[c_out][c_in] * X_tap[pixels][c_in]^T
//
// Why: in SS mode both UMMA operands stream from shared memory at 128 B/clk.  With pixels on M
// and the 128 output channels on N (tower_umma.cuh) every M=128,N=128,K=16 MMA reads 8 KB and
// sustains ~98 cycles against a 64-cycle floor.  Putting the weights on M and FOUR boards
// (256 pixels) on N makes each MMA read 4 KB (weights) + 8 KB (activations) for twice the work:
// 12 KB per 128-cycle MMA, ~140 cycles measured for this shape => 1.4x more tensor throughput.
//
// Data layout: "quad-interleaved" activations act[quad][y][b4][x][C] (board = 4*quad + b4); one
// 5-D TMA box {64 ch, 8 x, 4 b4, 10 y, 1 quad} at (x = dx, y = -1) is the halo'd, dx-shifted slab
// of 320 rows x 128 B (K-major SWIZZLE_128B); tap dy is the slab at byte offset (dy+1)*4096.
// A chain = one quad; a CTA owns two chains (TMEM columns [0,256) and [256,512)).
//
// Epilogue (8 warps, one team alternating between the chains): a thread owns ONE output channel
// (its TMEM lane) and walks the pixel columns: folded-BN bias / ReLU are per-thread scalars, the
// squeeze-excite channel means are thread-local sums (no shuffles), the excite scale is a
// per-thread register.  The bf16 result is transposed on the way out: each element is written to
// its (pixel row, channel) slot of the swizzled staging tile (st.shared.u16; the 32 lanes of a
// warp fill 64 contiguous bytes), the residual is TMA-loaded into that same tile beforehand and
// consumed in place.  The last layer (p_conv) is written in the PAIR layout the policy kernel reads.
#pragma once
#include "tower_umma.cuh"

namespace cb {

enum { LAYOUT_PLAIN = 0, LAYOUT_PAIR = 1, LAYOUT_QUAD = 2 };

struct TowerTCfg {
    static constexpr int HID = 128;
    static constexpr int kSlabBytes = 320 * 128;            // [10 y][4 b4][8 x][64 ch] 16-bit
    static constexpr int kAStages = 2;
    static constexpr int kWBytes = 128 * 128;
    static constexpr int kWStages = 4;
    static constexpr int kKcBytes = 256 * 128;              // one 64-channel chunk of a quad's output tile
    static constexpr int kOutBytes = 2 * kKcBytes;          // 64 KB staging (one chain at a time)
    static constexpr int kTmemCols = 512;
    static constexpr int kParts = 4;                        // column quarters: 4 epilogue warps per TMEM lane quarter
    static constexpr int kEpiThreads = 128 * kParts;        // 16 epilogue warps
    static constexpr int kThreads = 96 + kEpiThreads;       // producer, MMA, store warp + epilogue
    // barriers (256 B) | part [kParts][4][128] | hid [4][8] | vpart [4 lane quarters][256 pixels]
    static constexpr int kMiscBytes = 256 + 4 * (kParts * 512 + 32 + 1024);
    static constexpr int kSmemBytes = 1024 + kAStages * kSlabBytes + kWStages * kWBytes + kOutBytes + kMiscBytes;
    static_assert(kSmemBytes <= 232448, "shared memory budget (227 KB)");
};

struct TowerTLayer {
    int in_map, w_map, out_map, res_map;  // tensor-map table indices (res_map: quad box-8 map of the residual)
    int kc_in;
    int mode;
    int out_pair;                // 1: write the output in pair layout (two TMA stores per 64-ch chunk)
    int pad;
    const float* bias;
    const float* se_w1;
    const float* se_w2;
    const float* v_w;
    float* v_pre;                // plain [board][64] layout
};

struct TowerTParams {
    const CUtensorMap* maps;
    const TowerTLayer* layers;
    int num_layers;
    int num_quads;
    int num_boards;
    long long* trace;            // debug timeline (CTA 0): [item][8] SM clock stamps, or nullptr
};

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

// Sum each of 32 per-thread values across the 32 lanes of a warp; lane l ends with the total of
// v[l].  31 shuffles (butterfly transpose-reduce) instead of 32 x 5.
__device__ __forceinline__ float warp_transpose_sum(float (&v)[32], int lane) {
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

// trace slots per work item: 0 producer got act_ready, 1 producer issued last load, 2 MMA start,
// 3 MMA last issue, 4 epilogue saw tmem_full, 5 epilogue published the tile, 6 store issued, 7 store complete
__device__ __forceinline__ void trace_stamp(const TowerTParams& p, int item, int slot) {
    if (p.trace && blockIdx.x == 0) p.trace[item * 8 + slot] = clock64();
}

template <bool BF16>
__global__ void __launch_bounds__(TowerTCfg::kThreads, 1) tower_t_umma_kernel(const TowerTParams p) {
    using Cfg = TowerTCfg;
    constexpr int HID = 128;
    constexpr int kAStages = Cfg::kAStages, kWStages = Cfg::kWStages, kSE = 8;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* a_base = smem;
    uint8_t* w_base = a_base + kAStages * Cfg::kSlabBytes;
    uint8_t* out_tile = w_base + kWStages * Cfg::kWBytes;
    uint8_t* misc = out_tile + Cfg::kOutBytes;
    uint64_t* a_full = reinterpret_cast<uint64_t*>(misc);   // [2]
    uint64_t* a_empty = a_full + kAStages;                  // [2]
    uint64_t* w_full = a_empty + kAStages;                  // [4]
    uint64_t* w_empty = w_full + kWStages;                  // [4]
    uint64_t* tmem_full = w_empty + kWStages;               // [2 chains]
    uint64_t* tmem_empty = tmem_full + 2;                   // [2 chains]
    uint64_t* act_ready = tmem_empty + 2;                   // [2 chains]
    uint64_t* res_full = act_ready + 2;                     // [1] residual tile landed in the staging tile
    uint64_t* stage_ready = res_full + 1;                   // [1] epilogue team -> store warp
    uint64_t* stage_free = stage_ready + 1;                 // [1] store warp -> epilogue team
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(stage_free + 1);
    float* s_part = reinterpret_cast<float*>(misc + 256);   // [kParts][4 boards][128 ch]
    float* s_hid = s_part + Cfg::kParts * 512;              // [4][8]
    float* s_vpart = s_hid + 32;                            // [4 lane quarters][256 pixels]

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int num_units = (p.num_quads + 1) / 2;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kAStages; ++s) { ptx::mbar_init(&a_full[s], 1); ptx::mbar_init(&a_empty[s], 1); }
        for (int s = 0; s < kWStages; ++s) { ptx::mbar_init(&w_full[s], 1); ptx::mbar_init(&w_empty[s], 1); }
        for (int c = 0; c < 2; ++c) {
            ptx::mbar_init(&tmem_full[c], 1);
            ptx::mbar_init(&tmem_empty[c], Cfg::kEpiThreads);
            ptx::mbar_init(&act_ready[c], 1);
        }
        ptx::mbar_init(res_full, 1);
        ptx::mbar_init(stage_ready, Cfg::kEpiThreads);
        ptx::mbar_init(stage_free, 1);
        ptx::fence_barrier_init();
    }
    if (warp == 1) ptx::tmem_alloc<Cfg::kTmemCols>(tmem_ptr);
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0) {
        // ------------------------------------------------------ TMA producer
        int as = 0, ws = 0;
        uint32_t aph = 0, wph = 0, ready_ph[2] = {0, 0};
        for (int unit = blockIdx.x; unit < num_units; unit += gridDim.x) {
            for (int L = 0; L < p.num_layers; ++L) {
                const TowerTLayer* ly = p.layers + L;
                const CUtensorMap* tm_in = p.maps + ly->in_map;
                const CUtensorMap* tm_w = p.maps + ly->w_map;
                const int kc_in = ly->kc_in;
                for (int c = 0; c < 2; ++c) {
                    const int quad = unit * 2 + c;
                    if (quad >= p.num_quads) continue;
                    if (L > 0) {
                        ptx::mbar_wait(&act_ready[c], ready_ph[c]);
                        ready_ph[c] ^= 1;
                        ptx::fence_proxy_async_all();
                    }
                    if (lane == 0) trace_stamp(p, L * 2 + c, 0);
                    for (int dx = -1; dx <= 1; ++dx) {
                        for (int kc = 0; kc < kc_in; ++kc) {
                            ptx::mbar_wait(&a_empty[as], aph ^ 1);
                            if (ptx::elect_one()) {
                                ptx::mbar_arrive_expect_tx(&a_full[as], Cfg::kSlabBytes);
                                ptx::tma_load_5d(a_base + as * Cfg::kSlabBytes, tm_in, &a_full[as], kc * 64, dx, 0, -1, quad);
                            }
                            __syncwarp();
                            if (++as == kAStages) { as = 0; aph ^= 1; }
                            for (int dy = -1; dy <= 1; ++dy) {
                                const int tap = (dy + 1) * 3 + (dx + 1);
                                ptx::mbar_wait(&w_empty[ws], wph ^ 1);
                                if (ptx::elect_one()) {
                                    ptx::mbar_arrive_expect_tx(&w_full[ws], Cfg::kWBytes);
                                    ptx::tma_load_2d(w_base + ws * Cfg::kWBytes, tm_w, &w_full[ws], kc * 64, tap * HID);
                                }
                                __syncwarp();
                                if (++ws == kWStages) { ws = 0; wph ^= 1; }
                            }
                        }
                    }
                    if (lane == 0) trace_stamp(p, L * 2 + c, 1);
                }
            }
        }
    } else if (warp == 1) {
        // -------------------------------------------------------- MMA issuer
        constexpr uint32_t idesc = ptx::umma_idesc_f16(BF16 ? 1 : 0, 128, 256);
        int as = 0, ws = 0;
        uint32_t aph = 0, wph = 0, slot_ph[2] = {0, 0};
        for (int unit = blockIdx.x; unit < num_units; unit += gridDim.x) {
            for (int L = 0; L < p.num_layers; ++L) {
                const int kc_in = p.layers[L].kc_in;
                for (int c = 0; c < 2; ++c) {
                    if (unit * 2 + c >= p.num_quads) continue;
                    ptx::mbar_wait(&tmem_empty[c], slot_ph[c] ^ 1);
                    slot_ph[c] ^= 1;
                    ptx::tc_fence_after();
                    const uint32_t d_tmem = tmem_base + c * 256;
                    uint32_t accum = 0;
                    if (lane == 0) trace_stamp(p, L * 2 + c, 2);
                    for (int dxi = 0; dxi < 3; ++dxi) {
                        for (int kc = 0; kc < kc_in; ++kc) {
                            ptx::mbar_wait(&a_full[as], aph);
                            const uint32_t slab = ptx::smem_u32(a_base + as * Cfg::kSlabBytes);
                            for (int dyi = 0; dyi < 3; ++dyi) {
                                ptx::mbar_wait(&w_full[ws], wph);
                                ptx::tc_fence_after();
                                if (ptx::elect_one()) {
                                    // A = weights [128 c_out][64 c_in], B = 256 pixel rows of the slab for tap dy
                                    const uint64_t a_desc = ptx::umma_desc_sw128(ptx::smem_u32(w_base + ws * Cfg::kWBytes));
                                    const uint64_t b_desc = ptx::umma_desc_sw128(slab + dyi * 4096);
                                    ptx::umma_f16(d_tmem, a_desc, b_desc, idesc, accum);
#pragma unroll
                                    for (int k = 1; k < 4; ++k)
                                        ptx::umma_f16(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, 1u);
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
                    if (ptx::elect_one()) ptx::umma_commit(&tmem_full[c]);
                    __syncwarp();
                    if (lane == 0) trace_stamp(p, L * 2 + c, 3);
                }
            }
        }
    } else if (warp == 2) {
        // -------------------------------------------------------- store warp
        // Owns the staging tile's traffic with global memory so the epilogue team never waits for a
        // store to COMPLETE: residual prefetch into the tile, TMA store out of it, and the
        // act_ready signal to the producer once the layer output is visible in global memory.
        // Per item j (in the common work-item order):
        //   tile free -> [residual prefetch for j] -> stage_free -> finish item j-1 (wait_group 0, act_ready)
        //   -> wait stage_ready(j) -> TMA store(j) -> wait_group.read (tile free again)
        uint32_t ready_ph = 0;
        bool have_prev = false, prev_last = false;
        int prev_c = 0, prev_item = 0;
        for (int unit = blockIdx.x; unit < num_units; unit += gridDim.x) {
            for (int L = 0; L < p.num_layers; ++L) {
                const TowerTLayer* ly = p.layers + L;
                for (int c = 0; c < 2; ++c) {
                    const int quad = unit * 2 + c;
                    if (quad >= p.num_quads) continue;
                    if (ptx::elect_one()) {
                        if (ly->mode == MODE_SE_RES) {
                            ptx::mbar_arrive_expect_tx(res_full, Cfg::kOutBytes);
                            ptx::tma_load_5d(out_tile, p.maps + ly->res_map, res_full, 0, 0, 0, 0, quad);
                            ptx::tma_load_5d(out_tile + Cfg::kKcBytes, p.maps + ly->res_map, res_full, 64, 0, 0, 0, quad);
                        }
                        if (have_prev) {
                            ptx::mbar_arrive(stage_free);
                            ptx::tma_store_wait0();      // item j-1 complete in global memory
                            trace_stamp(p, prev_item, 7);
                            ptx::fence_proxy_async_all();
                            if (!prev_last) ptx::mbar_arrive(&act_ready[prev_c]);
                        }
                    }
                    __syncwarp();
                    ptx::mbar_wait(stage_ready, ready_ph);   // epilogue wrote the tile (and fenced it)
                    ready_ph ^= 1;
                    if (ptx::elect_one()) {
                        const CUtensorMap* tm_out = p.maps + ly->out_map;
                        if (ly->out_pair) {
#pragma unroll
                            for (int pr = 0; pr < 2; ++pr)
#pragma unroll
                                for (int kc = 0; kc < 2; ++kc)
                                    ptx::tma_store_5d(tm_out, out_tile + kc * Cfg::kKcBytes + pr * 128 * 128, kc * 64, 0, 0, 0,
                                                      quad * 2 + pr);
                        } else {
#pragma unroll
                            for (int kc = 0; kc < 2; ++kc)
                                ptx::tma_store_5d(tm_out, out_tile + kc * Cfg::kKcBytes, kc * 64, 0, 0, 0, quad);
                        }
                        ptx::tma_store_commit();
                        trace_stamp(p, L * 2 + c, 6);
                        ptx::tma_store_wait_read0();     // the tile may be overwritten
                    }
                    __syncwarp();
                    have_prev = true;
                    prev_c = c;
                    prev_item = L * 2 + c;
                    prev_last = (L == p.num_layers - 1);
                }
            }
        }
        if (have_prev && ptx::elect_one()) ptx::tma_store_wait0();
        __syncwarp();
    } else {
        // ---------------------------------------------------------- epilogue (16 warps)
        constexpr int kParts = Cfg::kParts, kCols = 256 / kParts, kEpi = Cfg::kEpiThreads;
        const int et = threadIdx.x - 96;          // 0..kEpi-1
        const int ew = et >> 5;                   // 0..15
        const int quadl = warp & 3;               // TMEM lane quarter this warp may touch
        const int part = (warp - 3) >> 2;         // which kCols pixel columns this warp handles
        const int ch = quadl * 32 + lane;         // my output channel == TMEM lane
        // byte offset of my channel inside a 128-byte row chunk, before the row-dependent swizzle
        const uint32_t kc_off = static_cast<uint32_t>(ch >> 6) * Cfg::kKcBytes;
        const uint32_t cq = static_cast<uint32_t>((ch & 63) >> 3);
        const uint32_t cb2 = static_cast<uint32_t>(ch & 7) * 2;
        uint32_t full_ph[2] = {0, 0}, res_ph = 0, free_ph = 0;
        bool first_item = true;
        for (int unit = blockIdx.x; unit < num_units; unit += gridDim.x) {
            for (int L = 0; L < p.num_layers; ++L) {
                const TowerTLayer* ly = p.layers + L;
                const int mode = ly->mode;
                const bool out_pair = ly->out_pair != 0;
                for (int c = 0; c < 2; ++c) {
                    const int quad = unit * 2 + c;
                    if (quad >= p.num_quads) continue;
                    const float bias = __ldg(ly->bias + ch);
                    constexpr int kFc1 = 32 / (kEpi / 32);   // squeeze-FC outputs per warp
                    float w1r[kFc1][4], w2r[kSE];
                    if (mode == MODE_SE_RES) {
                        // (the store warp prefetches the residual tile into the staging tile)
                        // squeeze FC: warp ew owns outputs o = ew*kFc1 + i -> (board o>>3, unit o&7)
#pragma unroll
                        for (int i = 0; i < kFc1; ++i)
#pragma unroll
                            for (int k = 0; k < 4; ++k)
                                w1r[i][k] = __ldg(&ly->se_w1[((ew * kFc1 + i) & 7) * HID + lane + 32 * k]);
#pragma unroll
                        for (int j = 0; j < kSE; ++j) w2r[j] = __ldg(&ly->se_w2[ch * kSE + j]);
                    }
                    ptx::mbar_wait(&tmem_full[c], full_ph[c]);
                    full_ph[c] ^= 1;
                    ptx::tc_fence_after();
                    if (et == 0) trace_stamp(p, L * 2 + c, 4);
                    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quadl * 32) << 16) + c * 256 + part * kCols;

                    float scale[4] = {1.0f, 1.0f, 1.0f, 1.0f};
                    if (mode == MODE_SE_RES) {
                        // pass 1: my channel's sum over this warp's pixels of each of the 4 boards
                        float sb[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll 1
                        for (int c0 = 0; c0 < kCols; c0 += 32) {
                            uint32_t r[32];
                            ptx::tmem_ld_32x32(taddr + c0, r);
                            ptx::tmem_ld_wait();
#pragma unroll
                            for (int j = 0; j < 32; ++j) sb[(j >> 3) & 3] += __uint_as_float(r[j]) + bias;
                        }
#pragma unroll
                        for (int b = 0; b < 4; ++b) s_part[(part * 4 + b) * HID + ch] = sb[b];
                        ptx::named_bar_sync(1, kEpi);
#pragma unroll
                        for (int i = 0; i < kFc1; ++i) {
                            const int o = ew * kFc1 + i, b = o >> 3, j = o & 7;
                            float s = 0.0f;
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                const int cc = lane + 32 * k;
                                float m = 0.0f;
#pragma unroll
                                for (int q = 0; q < kParts; ++q) m += s_part[(q * 4 + b) * HID + cc];
                                s = fmaf(w1r[i][k], m, s);
                            }
#pragma unroll
                            for (int d = 16; d >= 1; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
                            if (lane == 0) s_hid[b * kSE + j] = fmaxf(s * (1.0f / 64.0f), 0.0f);
                        }
                        ptx::named_bar_sync(1, kEpi);
#pragma unroll
                        for (int b = 0; b < 4; ++b) {
                            float s = 0.0f;
#pragma unroll
                            for (int j = 0; j < kSE; ++j) s = fmaf(w2r[j], s_hid[b * kSE + j], s);
                            scale[b] = 1.0f / (1.0f + __expf(-s));
                        }
                        ptx::mbar_wait(res_full, res_ph);   // residual tile has landed (implies the tile was free)
                        res_ph ^= 1;
                    }
                    // staging tile free again (previous store has read it)?  One stage_free phase per item
                    // after the first; SE items observe it transitively through res_full but must consume it.
                    if (!first_item) {
                        ptx::mbar_wait(stage_free, free_ph);
                        free_ph ^= 1;
                    }
                    first_item = false;
                    // main pass: bias (+ SE scale + residual), ReLU, transpose-store as 16-bit
                    const float vw = (mode == MODE_SE_RES && ly->v_w) ? __ldg(ly->v_w + ch) : 0.0f;
                    const bool want_v = (mode == MODE_SE_RES) && ly->v_pre != nullptr;
#pragma unroll 1
                    for (int c0 = 0; c0 < kCols; c0 += 32) {
                        uint32_t r[32];
                        ptx::tmem_ld_32x32(taddr + c0, r);
                        ptx::tmem_ld_wait();
                        const int n0 = part * kCols + c0;     // first pixel column: y = n0 >> 5, b4/x from j
                        const int y = n0 >> 5;
                        float tv[32];
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            const int b4 = (j >> 3) & 3, x = j & 7;
                            // staging row: quad order (y, b4, x) or, for the policy kernel, pair order
                            const int row = out_pair ? ((b4 >> 1) * 128 + ((y * 2 + (b4 & 1)) << 3) + x) : (n0 + j);
                            uint8_t* ptr = out_tile + kc_off + row * 128 + ((cq ^ static_cast<uint32_t>(x)) << 4) + cb2;
                            float v = __uint_as_float(r[j]) + bias;
                            if (mode == MODE_SE_RES) {
                                v = v * scale[b4] + from16bits<BF16>(*reinterpret_cast<uint16_t*>(ptr));
                            }
                            v = fmaxf(v, 0.0f);
                            *reinterpret_cast<uint16_t*>(ptr) = to16bits<BF16>(v);
                            tv[j] = v * vw;
                        }
                        if (want_v) {
                            // value 1x1 conv: sum over this warp's 32 channels for each of the 32 pixels
                            const float t = warp_transpose_sum(tv, lane);
                            s_vpart[quadl * 256 + n0 + lane] = t;
                        }
                    }
                    ptx::tc_fence_before();
                    ptx::mbar_arrive(&tmem_empty[c]);
                    // publish the tile to the store warp: generic-proxy writes -> async proxy, then arrive
                    ptx::fence_proxy_async_smem();
                    ptx::mbar_arrive(stage_ready);
                    if (et == 0) trace_stamp(p, L * 2 + c, 5);
                    if (want_v) {
                        ptx::named_bar_sync(1, kEpi);
                        // pixel column n = (y, b4, x) -> plain [board][y*8+x]; fixed summation order over the
                        // four lane quarters keeps the result independent of batch composition
                        if (et < 256) {
                            const int n = et, b4 = (n >> 3) & 3, board = quad * 4 + b4;
                            const float v = (s_vpart[n] + s_vpart[256 + n]) + (s_vpart[512 + n] + s_vpart[768 + n]);
                            if (board < p.num_boards) ly->v_pre[board * 64 + (n >> 5) * 8 + (n & 7)] = v;
                        }
                    }
                    // s_part / s_hid / s_vpart of this item are dead for every thread before the next item writes them
                    ptx::named_bar_sync(1, kEpi);
                }
            }
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
