// tower_umma.cuh — the whole convolution tower (start conv, residual blocks, p_conv) as ONE
// persistent kernel.
//
// A 3x3 convolution only mixes pixels of the same board, so a board pair can be taken through
// every layer without ever synchronising with another CTA: the output tile of layer L is stored
// (TMA) to the pair's rows of an HBM/L2 activation buffer and fetched back as the halo'd, dx-
// shifted slabs of layer L+1 by the same CTA.  Per-layer launches (conv_umma.cuh) pay ~7 us of
// launch / prologue / pipeline-fill / tail for ~15 us of work at batch 1024; here those costs are
// paid once per forward.
//
// Each CTA owns two independent CHAINS of kTiles board pairs.  Chain c accumulates in TMEM set c,
// so while chain 0's epilogue (TMEM -> BN/ReLU/SE/residual -> smem -> TMA store) and its store ->
// load turnaround run, the tensor core is busy with chain 1's MMAs of the same layer, and vice
// versa.  Work-item order on every role: for layer: for chain.
//
// Synchronisation (all mbarriers, CTA-local):
//   a_full/a_empty, w_full/w_empty   TMA producer <-> MMA issuer (slab ring, weight ring)
//   tmem_full/tmem_empty [chain][tile] MMA issuer <-> epilogue group
//   act_ready[chain]                  epilogue leaders (after cp.async.bulk.wait_group 0: the layer's
//                                     output is complete in global memory) -> TMA producer
// The per-layer math and data layout are exactly those of conv_umma.cuh.
#pragma once
#include "conv_umma.cuh"

namespace cb {

struct TowerLayer {
    int in_map, w_map, out_map;  // indices into the tensor-map table
    int kc_in;                   // C_in / 64
    int mode;                    // MODE_RELU / MODE_SE_RES
    int pad;
    const float* bias;
    const float* se_w1;
    const float* se_w2;
    const void* residual;
    const float* v_w;
    float* v_pre;
};

struct TowerParams {
    const CUtensorMap* maps;     // global memory, 64-byte aligned entries
    const TowerLayer* layers;    // global memory
    int num_layers;
    int num_tiles;               // board pairs
};

template <int HID, bool BF16>
__global__ void __launch_bounds__(ConvCfg<HID>::kThreads, 1) tower_umma_kernel(const TowerParams p) {
    using Cfg = ConvCfg<HID>;
    constexpr int kTiles = Cfg::kTiles;
    constexpr int kAStages = Cfg::kAStages;
    constexpr int kWStages = Cfg::kWStages;
    constexpr int kSE = Cfg::kSE;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* a_base = smem;
    uint8_t* w_base = a_base + kAStages * Cfg::kAStageBytes;
    uint8_t* out_base = w_base + kWStages * Cfg::kWBytes;
    uint8_t* misc = out_base + kTiles * Cfg::kOutBytes;
    uint64_t* a_full = reinterpret_cast<uint64_t*>(misc);   // [kAStages]
    uint64_t* a_empty = a_full + kAStages;                  // [kAStages]
    uint64_t* w_full = a_empty + kAStages;                  // [kWStages]
    uint64_t* w_empty = w_full + kWStages;                  // [kWStages]
    uint64_t* tmem_full = w_empty + kWStages;               // [2 chains][kTiles]
    uint64_t* tmem_empty = tmem_full + 2 * kTiles;          // [2 chains][kTiles]
    uint64_t* act_ready = tmem_empty + 2 * kTiles;          // [2 chains]
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(act_ready + 2);
    // per epilogue group: bias HID | part 8*HID | scale 2*HID | hid 2*kSE  (v_w is read from global)
    float* s_groups = reinterpret_cast<float*>(misc + 256);
    constexpr int kGroupFloats = HID + Cfg::kGroupFloats;
    static_assert(256 + 4 * kTiles * kGroupFloats <= Cfg::kMiscBytes, "misc budget");

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int tiles_per_unit = 2 * kTiles;
    const int num_units = (p.num_tiles + tiles_per_unit - 1) / tiles_per_unit;

    if (threadIdx.x == 0) {
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
        ptx::mbar_init(&act_ready[0], kTiles);
        ptx::mbar_init(&act_ready[1], kTiles);
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
                const TowerLayer* ly = p.layers + L;
                const CUtensorMap* tm_in = p.maps + ly->in_map;
                const CUtensorMap* tm_w = p.maps + ly->w_map;
                const int kc_in = ly->kc_in;
                for (int c = 0; c < 2; ++c) {
                    const int tile0 = (unit * 2 + c) * kTiles;
                    if (tile0 >= p.num_tiles) continue;  // empty chain
                    if (L > 0) {
                        // this chain's previous layer is complete in global memory
                        ptx::mbar_wait(&act_ready[c], ready_ph[c]);
                        ready_ph[c] ^= 1;
                        ptx::fence_proxy_async_all();
                    }
                    for (int dx = -1; dx <= 1; ++dx) {
                        for (int kc = 0; kc < kc_in; ++kc) {
                            ptx::mbar_wait(&a_empty[as], aph ^ 1);
                            if (ptx::elect_one()) {
                                ptx::mbar_arrive_expect_tx(&a_full[as], Cfg::kAStageBytes);
#pragma unroll
                                for (int t = 0; t < kTiles; ++t)
                                    ptx::tma_load_5d(a_base + as * Cfg::kAStageBytes + t * Cfg::kSlabBytes, tm_in,
                                                     &a_full[as], kc * 64, dx, 0, -1, tile0 + t);
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
                }
            }
        }
    } else if (warp == 1) {
        // -------------------------------------------------------- MMA issuer
        constexpr uint32_t idesc = ptx::umma_idesc_f16(BF16 ? 1 : 0, 128, HID);
        int as = 0, ws = 0;
        uint32_t aph = 0, wph = 0, slot_ph[2] = {0, 0};
        for (int unit = blockIdx.x; unit < num_units; unit += gridDim.x) {
            for (int L = 0; L < p.num_layers; ++L) {
                const int kc_in = p.layers[L].kc_in;
                for (int c = 0; c < 2; ++c) {
                    if ((unit * 2 + c) * kTiles >= p.num_tiles) continue;
#pragma unroll
                    for (int t = 0; t < kTiles; ++t) ptx::mbar_wait(&tmem_empty[c * kTiles + t], slot_ph[c] ^ 1);
                    slot_ph[c] ^= 1;
                    ptx::tc_fence_after();
                    uint32_t accum = 0;
                    for (int dxi = 0; dxi < 3; ++dxi) {
                        for (int kc = 0; kc < kc_in; ++kc) {
                            ptx::mbar_wait(&a_full[as], aph);
                            const uint32_t slab0 = ptx::smem_u32(a_base + as * Cfg::kAStageBytes);
                            for (int dyi = 0; dyi < 3; ++dyi) {
                                ptx::mbar_wait(&w_full[ws], wph);
                                ptx::tc_fence_after();
                                if (ptx::elect_one()) {
                                    const uint64_t b_desc =
                                        ptx::umma_desc_sw128(ptx::smem_u32(w_base + ws * Cfg::kWBytes));
#pragma unroll
                                    for (int t = 0; t < kTiles; ++t) {
                                        const uint32_t d_tmem = tmem_base + (c * kTiles + t) * HID;
                                        const uint64_t a_desc =
                                            ptx::umma_desc_sw128(slab0 + t * Cfg::kSlabBytes + dyi * 2048);
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
                        for (int t = 0; t < kTiles; ++t) ptx::umma_commit(&tmem_full[c * kTiles + t]);
                    }
                    __syncwarp();
                }
            }
        }
    } else {
        // ---------------------------------------------------------- epilogue
        const int g = (warp - 2) >> 2;        // epilogue group == tile slot within a chain
        const int quad = warp & 3;
        const int row = quad * 32 + lane;     // (y, b2, x) = (row>>4, (row>>3)&1, row&7)
        const int et = (threadIdx.x - 64) & 127;
        const int ew = et >> 5;
        const bool leader = (et == 0);
        const int bar_id = 1 + g;
        const int b2 = (lane >> 3) & 1;
        const uint32_t swz = static_cast<uint32_t>(row & 7);
        uint8_t* out_tile = out_base + g * Cfg::kOutBytes;
        uint8_t* my_row = out_tile + row * 128;
        float* s_bias = s_groups + g * kGroupFloats;   // [HID]
        float* s_part = s_bias + HID;                  // [4 quads][2 boards][HID]
        float* s_scale = s_part + 8 * HID;             // [2][HID]
        float* s_hid = s_scale + 2 * HID;              // [2][kSE]
        uint32_t full_ph[2] = {0, 0};
        for (int unit = blockIdx.x; unit < num_units; unit += gridDim.x) {
            for (int L = 0; L < p.num_layers; ++L) {
                const TowerLayer* ly = p.layers + L;
                const int mode = ly->mode;
                const bool last_layer = (L == p.num_layers - 1);
                for (int c = 0; c < 2; ++c) {
                    if ((unit * 2 + c) * kTiles >= p.num_tiles) continue;
                    const int tile = (unit * 2 + c) * kTiles + g;
                    const int slot = c * kTiles + g;
                    const bool valid = tile < p.num_tiles;
                    // per-item constants; the previous item of this group is past its last use of s_bias
                    // (it ended with a named barrier before the store) so it can be rewritten here
                    s_bias[et] = __ldg(ly->bias + et);
                    if constexpr (HID == 256) s_bias[et + 128] = __ldg(ly->bias + et + 128);
                    [[maybe_unused]] float w1r[kSE / 4][HID / 32];
                    [[maybe_unused]] float w2r[HID / 128][kSE];
                    [[maybe_unused]] uint4 rr[HID == 128 ? 16 : 1];
                    const uint4* res = nullptr;
                    if (mode == MODE_SE_RES) {
#pragma unroll
                        for (int jj = 0; jj < kSE / 4; ++jj)
#pragma unroll
                            for (int i = 0; i < HID / 32; ++i)
                                w1r[jj][i] = __ldg(&ly->se_w1[(ew + 4 * jj) * HID + lane + 32 * i]);
#pragma unroll
                        for (int cc = 0; cc < HID / 128; ++cc)
#pragma unroll
                            for (int j = 0; j < kSE; ++j) w2r[cc][j] = __ldg(&ly->se_w2[(et + 128 * cc) * kSE + j]);
                        // residual was written by this CTA's own TMA store two layers ago: read it at L2
                        // (ld.global.cg), never through the non-coherent path
                        res = reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(ly->residual) +
                                                             (static_cast<size_t>(valid ? tile : 0) * 128 + row) * HID);
                        if constexpr (HID == 128) {
#pragma unroll
                            for (int i = 0; i < 16; ++i) rr[i] = __ldcg(res + i);
                        }
                    }
                    ptx::named_bar_sync(bar_id, 128);  // s_bias visible to the group
                    ptx::mbar_wait(&tmem_full[slot], full_ph[c]);
                    full_ph[c] ^= 1;
                    ptx::tc_fence_after();
                    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + slot * HID;

                    if (!valid) {  // padding tile of a ragged chain: protocol only
                        ptx::tc_fence_before();
                        ptx::mbar_arrive(&tmem_empty[slot]);
                        if (leader && !last_layer) ptx::mbar_arrive(&act_ready[c]);
                        continue;
                    }
                    if (mode == MODE_RELU) {
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
                                    const int cc = i * 8 + j * 2;
                                    const float a = fmaxf(__uint_as_float(r[cc]) + s_bias[c0 + cc], 0.0f);
                                    const float b = fmaxf(__uint_as_float(r[cc + 1]) + s_bias[c0 + cc + 1], 0.0f);
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
                            const int col = warp_board_column_sums(v, lane);
                            *reinterpret_cast<float2*>(&s_part[(quad * 2 + b2) * HID + c0 + col]) = make_float2(v[0], v[1]);
                        }
                        ptx::named_bar_sync(bar_id, 128);
                        // squeeze FC
#pragma unroll
                        for (int b = 0; b < 2; ++b) {
                            float m[HID / 32];
#pragma unroll
                            for (int i = 0; i < HID / 32; ++i) {
                                const int cc = lane + 32 * i;
                                m[i] = s_part[(0 * 2 + b) * HID + cc] + s_part[(1 * 2 + b) * HID + cc] +
                                       s_part[(2 * 2 + b) * HID + cc] + s_part[(3 * 2 + b) * HID + cc];
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
                        ptx::named_bar_sync(bar_id, 128);
                        // excite FC + sigmoid
#pragma unroll
                        for (int cc = 0; cc < HID / 128; ++cc) {
#pragma unroll
                            for (int b = 0; b < 2; ++b) {
                                float s = 0.0f;
#pragma unroll
                                for (int j = 0; j < kSE; ++j) s = fmaf(w2r[cc][j], s_hid[b * kSE + j], s);
                                s_scale[b * HID + et + 128 * cc] = 1.0f / (1.0f + __expf(-s));
                            }
                        }
                        ptx::named_bar_sync(bar_id, 128);
                        // pass 2: scale, residual, relu, (value 1x1), pack
                        const float* scale = s_scale + b2 * HID;
                        const float* vw = ly->v_w;
                        float vacc = 0.0f;
#pragma unroll
                        for (int c0 = 0; c0 < HID; c0 += 32) {
                            uint4 rl[4];
                            if constexpr (HID == 128) {
#pragma unroll
                                for (int i = 0; i < 4; ++i) rl[i] = rr[(c0 >> 3) + i];
                            } else {
#pragma unroll
                                for (int i = 0; i < 4; ++i) rl[i] = __ldcg(res + (c0 >> 3) + i);
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
                                    const int cc = i * 8 + j * 2;
                                    const float2 rv = unpack2<BF16>(rw[j]);
                                    float a = (__uint_as_float(r[cc]) + s_bias[c0 + cc]) * scale[c0 + cc] + rv.x;
                                    float b = (__uint_as_float(r[cc + 1]) + s_bias[c0 + cc + 1]) * scale[c0 + cc + 1] + rv.y;
                                    a = fmaxf(a, 0.0f);
                                    b = fmaxf(b, 0.0f);
                                    if (vw) vacc += a * __ldg(vw + c0 + cc) + b * __ldg(vw + c0 + cc + 1);
                                    ow[j] = pack2<BF16>(a, b);
                                }
                                *reinterpret_cast<uint4*>(dst + (((q0 + i) ^ swz) << 4)) = o;
                            }
                        }
                        if (ly->v_pre) ly->v_pre[static_cast<size_t>(tile) * 128 + row] = vacc;
                    }
                    // accumulator drained -> MMA warp may overwrite it
                    ptx::tc_fence_before();
                    ptx::mbar_arrive(&tmem_empty[slot]);
                    // publish the tile: generic-proxy writes -> async proxy, then one thread stores
                    ptx::fence_proxy_async_smem();
                    ptx::named_bar_sync(bar_id, 128);
                    if (leader) {
                        const CUtensorMap* tm_out = p.maps + ly->out_map;
#pragma unroll
                        for (int kc = 0; kc < HID / 64; ++kc)
                            ptx::tma_store_5d(tm_out, out_tile + kc * Cfg::kKcBytes, kc * 64, 0, 0, 0, tile);
                        ptx::tma_store_commit();
                        // full completion (not just smem read): the next layer's loads depend on it
                        ptx::tma_store_wait0();
                        ptx::fence_proxy_async_all();
                        if (!last_layer) ptx::mbar_arrive(&act_ready[c]);
                    }
                    // the staging tile is rewritten only after the leader has passed wait0:
                    ptx::named_bar_sync(bar_id, 128);
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
