// tower_r2_umma.cuh — the forward of the 256-channel net (BASELINE config 4) on CTA PAIRS: the conv tower, p_conv and the
// policy head of up to 8 boards per pair, activations resident in shared memory, weights split across the pair.
//
//     D[c_out = 256 rows over the pair][pixels = 256 columns] += W_tap[c_out][c_in] * X_tap[pixels][c_in]^T
//     tcgen05.mma.cta_group::2, M = 256, N = 256, K = 16
//
// Why a pair: with 256 output channels a single CTA needs two M = 128 accumulators per board, so only 2 boards fit a
// chain (4 boards x 256 channels of fp32 = all 512 TMEM columns) and every 16 KB weight tile is used for N = 128 pixels
// only — 64 B/clk of weight stream per SM plus 128 B/clk of operand fetch, beyond the 128 B/clk shared-memory port (what
// the global-round-trip tower_umma_kernel<256> runs into as well).  As a pair, CTA r loads the weight rows of output
// channels [128 r, 128 r + 128) only and the tensor cores of both SMs use them against the pixels of BOTH CTAs: the
// weight stream per SM is that of the 128-channel kernel (32 B/clk) and the activations never leave the chip.
//
// A pair owns a UNIT of up to 8 boards = two CHAINS of 4; of a chain's 4 boards CTA r holds boards 2r, 2r + 1 (all 256
// input channels) in one shared-memory buffer of four 64-channel QUARTERS, each K-major SWIZZLE_128B rows:
//
//     row(q, y, b, x) = q * 162 + ((y + 1) * 2 + b) * 9 + 1 + x          (b < 2 local boards, 9 rows per group)
//
// Row 0 of every 9-row group is zero (x = -1 / x = 8 padding, as in tower_r_umma.cuh); each quarter starts with one zero
// band of 2 groups = the y = -1 halo of this quarter and the y = 8 halo of the quarter before it (one more band closes the
// buffer), so every tap is one N = 256 MMA per K step: start row (1 + dy) * 18 + 1 + dx, 8-row-group stride 9 rows.
// (The N = 56 nb trick of the 128-channel kernel does not compose with the pair: the hardware takes the first N / 2
// accumulator columns from CTA 0's operand and the rest from CTA 1's.)  Accumulator of chain c in each CTA's TMEM:
// columns [256 c, 256 c + 256), column = r * 128 + (y * 2 + b) * 8 + x, lane = the CTA's output channel.
//
// The epilogue thread (output channel, board of the chain) is the 128-channel kernel's: BN bias preset in the accumulator,
// squeeze-excite, residual, ReLU, 16-bit pack.  What is new: a board lives in ONE CTA but its output channels are produced
// by BOTH, so half of every CTA's epilogue stores go to the peer's buffer through distributed shared memory
// (st.shared::cluster); the squeeze-excite board sums are exchanged the same way (one cluster-scope mbarrier per chain);
// chain_ready lives in the leader CTA and counts one elected arrival per CTA.  The block input for the residual add is
// parked per thread in global scratch as fp16 whatever the operand type (bf16 operands keep an 11-bit skip path).
// Weights: a ring of 3 x 16 KB per CTA; each CTA's TMA signals the LEADER's w_full barrier (cta_group::2 load), the
// leader's MMA warp releases ring slots and publishes accumulators with multicast commits.
// The 1x1 value conv is accumulated into v_pre with one atomicAdd per CTA and pixel (two addends: order-free); the value
// head's FC layers run in value_head_kernel after this launch.  The policy head is the 23rd, one-tap layer with p_head
// padded to 256 rows: its 73 planes are lanes 0..72 of CTA 0, which runs the softmax / legal-move gather of
// tower_r_umma.cuh for the chain's 4 boards.
//
// Reference: python/model.py:54-145 with hidden_dim = 256, num_blocks = 10.
#pragma once
#include "tower_r_umma.cuh"

namespace cb {

struct TowerR2Cfg {
    static constexpr int HID = 256;
    static constexpr int kNbLocal = 2;                          // boards per CTA and chain
    static constexpr int kQuarterRows = 9 * kNbLocal * 9;       // halo band + 8 data group rows, 9 rows per group
    static constexpr int kChainRows = 4 * kQuarterRows + kNbLocal * 9 + 1;
    static constexpr int kChainBytes = 84 * 1024;               // >= kChainRows * 128, 1024-aligned
    static constexpr int kWBytes = 128 * 128;
    static constexpr int kWStages = 3;
    static constexpr int kEpiThreads = 512;
    static constexpr int kThreads = 64 + kEpiThreads;
    static constexpr int kSE = 16;
    // barriers (256 B) | s_part [4 boards][256] | s_hid [4][16] | s_vpart [4 lane quarters][4 boards][64] | s_red [4][32]
    static constexpr int kMiscBytes = 256 + 4096 + 256 + 4096 + 512;
    static constexpr int kSmemBytes = 1024 + 2 * kChainBytes + kWStages * kWBytes + kMiscBytes;
    static constexpr int kScratchPerCta = 2 * 4 * 8 * 128;      // uint4 entries: [chain][board][y][channel]
    static_assert(kChainRows * 128 <= kChainBytes, "chain buffer");
    static_assert(4 * CB_POLICY_SIZE * 4 <= kChainBytes, "the logits of a chain's 4 boards fit its buffer");
    static_assert(kSmemBytes <= 232448, "shared memory budget (227 KB)");
};

struct TowerR2Params {
    const cb_board* boards;
    const CUtensorMap* maps;     // weights of every layer (conv tower, then the padded p_head), box {64 ch, 128 rows}
    const TowerRLayer* layers;   // same table as the 128-channel kernel (kc_in = C_in / 64)
    PolicyOut po;
    float* v_pre;                // [board][64], zeroed before the launch
    int num_layers;              // conv layers + 1 (the policy head)
    int num_boards;
    int num_units;               // units [0, full_units) hold 8 boards (two chains), the tail units 4 (one chain)
    int full_units;
    uint4* res_scratch;          // [grid][TowerR2Cfg::kScratchPerCta]
    float* dbg_backbone;         // optional f32 NCHW [board][256][64] dump of the backbone output
};

namespace ptx {
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void st_cluster_u32(uint32_t addr, uint32_t v) {
    asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ void st_cluster_f32(uint32_t addr, float v) {
    asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
// arrive on an mbarrier given by its shared::cluster address (any CTA of the cluster), release at cluster scope
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait_cluster(uint64_t* bar, uint32_t parity, uint32_t hint_ns) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity), "r"(hint_ns)
        : "memory");
    return ok != 0;
}
// bounded wait with cluster-scope acquire (the arrivals come from both CTAs of the pair)
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait_cluster(bar, parity, 20000u)) {
        if (++spins > (1u << 20)) __trap();
    }
}
// 2-D tile load of a CTA pair: the data lands in THIS CTA's shared memory, the transaction bytes are counted on the
// LEADER CTA's mbarrier at the same offset (cute::SM100_TMA_2SM_LOAD_2D: the peer bit of the barrier address cleared).
__device__ __forceinline__ void tma_load_2d_2sm(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
    const uint32_t mbar = smem_u32(bar) & 0xFEFFFFFFu;
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(mbar), "r"(c0), "r"(c1)
        : "memory");
}
template <uint32_t kCols>
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t* smem_result) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)), "n"(kCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <uint32_t kCols>
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols) : "memory");
}
// D[tmem of both CTAs] += A[smem of both] * B[smem of both]^T, issued by one thread of the leader CTA
__device__ __forceinline__ void umma_f16_2sm_lohi(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                                  uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
        "mov.b64 da, {%1, %2};\n\t"
        "mov.b64 db, {%3, %4};\n\t"
        "setp.ne.b32 p, %6, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %5, p;\n\t}" ::"r"(d_tmem),
        "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive on the mbarrier at this offset in the CTAs of `mask` once all MMAs issued so far by this thread have retired
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                     smem_u32(bar)),
                 "h"(mask)
                 : "memory");
}
}  // namespace ptx

template <bool BF16>
__global__ void __launch_bounds__(TowerR2Cfg::kThreads, 1) tower_r2_umma_kernel(const TowerR2Params p) {
    using Cfg = TowerR2Cfg;
    constexpr int HID = Cfg::HID, kSE = Cfg::kSE, kWStages = Cfg::kWStages, kEpi = Cfg::kEpiThreads;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* chain_base = smem;                                 // [2 chains][kChainBytes]
    uint8_t* w_base = chain_base + 2 * Cfg::kChainBytes;
    uint8_t* misc = w_base + kWStages * Cfg::kWBytes;
    uint64_t* w_full = reinterpret_cast<uint64_t*>(misc);       // [kWStages]   (used in the leader)
    uint64_t* w_empty = w_full + kWStages;                      // [kWStages]
    uint64_t* tmem_full = w_empty + kWStages;                   // [2 chains]
    uint64_t* chain_ready = tmem_full + 2;                      // [2 chains]   (used in the leader: one arrival per CTA)
    uint64_t* se_bar = chain_ready + 2;                         // [2 chains]   the peer's board sums have landed
    uint64_t* se_free = se_bar + 2;                             // the peer has read the sums I sent it for the last item
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(se_free + 1);
    float* s_part = reinterpret_cast<float*>(misc + 256);       // [4 boards][256 channels]
    float* s_hid = s_part + 4 * HID;                            // [4][16]
    float* s_vpart = s_hid + 4 * kSE;                           // [4 lane quarters][4 boards][64 pixels]
    float* s_red = s_vpart + 1024;                              // [4 boards][32]

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = ptx::cluster_ctarank();               // 0 = leader
    const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
    constexpr int kTapOrder[9] = {4, 3, 5, 1, 0, 2, 7, 6, 8};

    if (threadIdx.x == 0) {
        for (int s = 0; s < kWStages; ++s) { ptx::mbar_init(&w_full[s], 1); ptx::mbar_init(&w_empty[s], 1); }
        for (int c = 0; c < 2; ++c) {
            ptx::mbar_init(&tmem_full[c], 1);
            ptx::mbar_init(&chain_ready[c], 2);
            ptx::mbar_init(&se_bar[c], 1);
        }
        ptx::mbar_init(se_free, 1);
        ptx::fence_barrier_init();
    }
    if (warp == 1) ptx::tmem_alloc_2sm<512>(tmem_ptr);
    ptx::tc_fence_before();
    __syncthreads();
    ptx::cluster_sync_all();      // both CTAs' barriers are initialised before anything targets them
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    // first board of a unit; chains of a unit that hold at least one board.  When the last round of 8-board units would
    // leave more than half of the pairs idle, the host cuts those units in two: twice as many pairs each run one chain
    // (its epilogue is then not hidden: 0.79 of a full unit's time for half the boards, still a gain of 5 % at batch 2048.
    // Measured and dropped: running such a chain as two sub-chains of one board per CTA — N = 128 MMAs over every other
    // 9-row group, 8-row-group stride 18 rows — so that the epilogues hide again: the N = 128 MMAs of a pair are bound by
    // the shared-memory port (4 KB of weights re-read per 64 cycles) and the launch got 1.8 % slower, not 4 % faster)
    auto unit_b0 = [&](int unit) { return unit < p.full_units ? unit * 8 : p.full_units * 8 + (unit - p.full_units) * 4; };
    auto chain_live = [&](int unit, int c) { return (c == 0 || unit < p.full_units) && unit_b0(unit) + c * 4 < p.num_boards; };

    if (warp == 0) {
        // ------------------------------------------------------ weight producer (both CTAs: each its 128 rows)
        int ws = 0;
        uint32_t wph = 0;
        for (int unit = pair; unit < p.num_units; unit += n_pairs) {
            for (int L = 0; L < p.num_layers; ++L) {
                const TowerRLayer* ly = p.layers + L;
                const CUtensorMap* tm_w = p.maps + ly->w_map;
                const int kc_in = ly->kc_in, ntaps = ly->ntaps, w_row_off = ly->w_row_off;
                for (int c = 0; c < 2; ++c) {
                    if (!chain_live(unit, c)) continue;
#pragma unroll 1
                    for (int t = 0; t < ntaps; ++t) {
                        const int tap = kTapOrder[t];
                        for (int kc = 0; kc < kc_in; ++kc) {
                            ptx::mbar_wait(&w_empty[ws], wph ^ 1);
                            if (ptx::elect_one()) {
                                if (rank == 0) ptx::mbar_arrive_expect_tx(&w_full[ws], 2 * Cfg::kWBytes);   // both CTAs' tiles
                                ptx::tma_load_2d_2sm(w_base + ws * Cfg::kWBytes, tm_w, &w_full[ws], kc * 64,
                                                     tap * HID + w_row_off + static_cast<int>(rank) * 128);
                            }
                            __syncwarp();
                            if (++ws == kWStages) { ws = 0; wph ^= 1; }
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // -------------------------------------------------------- MMA issuer (leader CTA only)
        if (rank == 0) {
            int ws = 0;
            uint32_t wph = 0, ready_ph0 = 0, ready_ph1 = 0;
            const uint64_t a_desc0 = ptx::umma_desc_sw128(ptx::smem_u32(w_base));
            const uint64_t b_desc0 = umma_desc_sw128_sbo(ptx::smem_u32(chain_base), 9 * 128);
            const uint32_t a_lo0 = static_cast<uint32_t>(a_desc0), a_hi = static_cast<uint32_t>(a_desc0 >> 32);
            const uint32_t b_lo0 = static_cast<uint32_t>(b_desc0), b_hi = static_cast<uint32_t>(b_desc0 >> 32);
            const uint32_t idesc = ptx::umma_idesc_f16(BF16 ? 1 : 0, 256, 256);
            for (int unit = pair; unit < p.num_units; unit += n_pairs) {
                for (int L = 0; L < p.num_layers; ++L) {
                    const int kc_in = p.layers[L].kc_in, ntaps = p.layers[L].ntaps;
#pragma unroll
                    for (int c = 0; c < 2; ++c) {
                        if (!chain_live(unit, c)) continue;
                        if (c == 0) { ptx::mbar_wait_cluster(&chain_ready[0], ready_ph0); ready_ph0 ^= 1; }
                        else { ptx::mbar_wait_cluster(&chain_ready[1], ready_ph1); ready_ph1 ^= 1; }
                        ptx::tc_fence_after();
                        const uint32_t b_lo_c = b_lo0 + c * (Cfg::kChainBytes >> 4);
                        const uint32_t d_tmem = tmem_base + c * 256;
#pragma unroll
                        for (int t = 0; t < 9; ++t) {
                            if (t >= ntaps) break;
                            const int tap = kTapOrder[t];
                            const int dy = tap / 3 - 1, dx = tap % 3 - 1;
                            const int start_row = (1 + dy) * Cfg::kNbLocal * 9 + 1 + dx;
                            const uint32_t b_lo_t = b_lo_c + start_row * 8;      // 128-byte rows in 16-byte units
                            for (int kc = 0; kc < kc_in; ++kc) {
                                ptx::mbar_wait(w_full + ws, wph);
                                ptx::tc_fence_after();
                                if (ptx::elect_one()) {
                                    const uint32_t a_lo = a_lo0 + ws * (Cfg::kWBytes >> 4);
                                    const uint32_t b_lo = b_lo_t + kc * (Cfg::kQuarterRows * 8);
#pragma unroll
                                    for (int k = 0; k < 4; ++k)
                                        ptx::umma_f16_2sm_lohi(d_tmem, a_lo + 2 * k, a_hi, b_lo + 2 * k, b_hi, idesc, 1u);   // bias preset
                                    ptx::umma_commit_2sm(w_empty + ws, 3);      // the slot is free in both CTAs
                                }
                                __syncwarp();
                                if (++ws == kWStages) { ws = 0; wph ^= 1; }
                            }
                        }
                        if (ptx::elect_one()) ptx::umma_commit_2sm(&tmem_full[c], 3);
                        __syncwarp();
                    }
                }
            }
        }
    } else {
        // ---------------------------------------------------------- epilogue (16 warps in each CTA)
        // Thread = (output channel ch of this CTA's 128 = TMEM lane, board bs of the chain's 4).  Boards 2r', 2r' + 1 live
        // in CTA r': the store goes to my own buffer or through distributed shared memory to the peer's.
        const int et = threadIdx.x - 64;          // 0..511
        const int ew = et >> 5;                   // 0..15
        const int quadl = warp & 3;               // TMEM lane quarter this warp may touch
        const int bs = ew >> 2;                   // board of the chain
        const int ch = quadl * 32 + lane;         // my TMEM lane
        const int g = static_cast<int>(rank) * 128 + ch;   // my output channel of the 256
        const uint32_t dst_rank = static_cast<uint32_t>(bs >> 1);
        const bool remote = dst_rank != rank;
        const int bl = bs & 1;                    // the board's slot in its CTA
        const uint32_t q_row0 = static_cast<uint32_t>(g >> 6) * Cfg::kQuarterRows;
        const uint32_t cq = static_cast<uint32_t>((g & 63) >> 3);
        const uint32_t cb2 = static_cast<uint32_t>(g & 7) * 2;
        const bool odd = (lane & 1) != 0;
        const uint32_t sel_lo = odd ? 0x1054u : 0x5410u, sel_hi = odd ? 0x3276u : 0x7632u;
        uint4* my_res = p.res_scratch + static_cast<size_t>(blockIdx.x) * Cfg::kScratchPerCta + bs * 8 * 128 + ch;
        // where my stores go: the chain buffers of the CTA that holds my board, as shared::cluster addresses
        const uint32_t dst_chain0 = remote ? ptx::mapa_u32(ptx::smem_u32(chain_base), dst_rank) : ptx::smem_u32(chain_base);
        const uint32_t peer_part = ptx::mapa_u32(ptx::smem_u32(s_part), rank ^ 1u);
        const uint32_t peer_se_bar = ptx::mapa_u32(ptx::smem_u32(se_bar), rank ^ 1u);
        const uint32_t peer_se_free = ptx::mapa_u32(ptx::smem_u32(se_free), rank ^ 1u);
        uint32_t free_ph = 0;                     // phase of se_free
        const uint32_t leader_ready = ptx::mapa_u32(ptx::smem_u32(chain_ready), 0u);
        uint32_t full_ph = 0, se_ph = 0;          // bit c: phase of tmem_full[c] / se_bar[c]

        for (int unit = pair; unit < p.num_units; unit += n_pairs) {
            const int b0 = unit_b0(unit);
            // ---- a clean buffer (zero rows, halo bands; the heads of the previous unit used it as logit storage), then the
            // input planes of my CTA's boards in quarter 0, straight from the packed bitboards
            if (unit != pair) ptx::named_bar_sync(1, kEpi);
#pragma unroll 1
            for (int c = 0; c < 2; ++c) {
                if (!chain_live(unit, c)) continue;
                uint8_t* buf = chain_base + c * Cfg::kChainBytes;
                for (int i = et; i < Cfg::kChainBytes / 16; i += kEpi) reinterpret_cast<uint4*>(buf)[i] = make_uint4(0, 0, 0, 0);
            }
            ptx::named_bar_sync(1, kEpi);
#pragma unroll 1
            for (int c = 0; c < 2; ++c) {
                if (!chain_live(unit, c)) continue;
                uint8_t* buf = chain_base + c * Cfg::kChainBytes;
                const int cb0 = b0 + c * 4 + static_cast<int>(rank) * 2;      // my CTA's first board of this chain
                for (int r = et; r < 8 * 2 * 8; r += kEpi) {                   // (y, b, x)
                    const int y = r >> 4, b = (r >> 3) & 1, x = r & 7;
                    if (cb0 + b >= p.num_boards) continue;
                    const uint32_t msk = pixel_plane_mask(p.boards + cb0 + b, y, x);
                    const uint32_t row = static_cast<uint32_t>(((y + 1) * 2 + b) * 9 + 1 + x);
                    uint4* row0 = reinterpret_cast<uint4*>(buf + row * 128);
                    const uint32_t one = BF16 ? 0x3F80u : 0x3C00u;
#pragma unroll
                    for (int i = 0; i < 3; ++i) {
                        const uint32_t m8 = msk >> (8 * i);
                        uint4 v;
                        v.x = ((m8 >> 0) & 1u ? one : 0u) | ((m8 >> 1) & 1u ? (one << 16) : 0u);
                        v.y = ((m8 >> 2) & 1u ? one : 0u) | ((m8 >> 3) & 1u ? (one << 16) : 0u);
                        v.z = ((m8 >> 4) & 1u ? one : 0u) | ((m8 >> 5) & 1u ? (one << 16) : 0u);
                        v.w = ((m8 >> 6) & 1u ? one : 0u) | ((m8 >> 7) & 1u ? (one << 16) : 0u);
                        row0[i ^ (row & 7)] = v;
                    }
                }
                // accumulator preset: the folded-BN bias of the first layer (every MMA accumulates)
                {
                    const uint32_t bias0 = __float_as_uint(__ldg(p.layers[0].bias + g));
                    const uint32_t tcol = tmem_base + (static_cast<uint32_t>(quadl * 32) << 16) + c * 256 + (bs >> 1) * 128 + bl * 8;
#pragma unroll
                    for (int y = 0; y < 8; ++y) ptx::tmem_st_32x8_bcast(tcol + y * 16, bias0);
                    ptx::tmem_st_wait();
                }
                ptx::tc_fence_before();
                ptx::fence_proxy_async_all();
                ptx::named_bar_sync(1, kEpi);
                if (et == 0) ptx::mbar_arrive_cluster(leader_ready + c * 8);
            }

            for (int L = 0; L < p.num_layers; ++L) {
                const TowerRLayer* ly = p.layers + L;
                const int mode = ly->mode;
                const bool last_layer = L == p.num_layers - 1;   // the policy head (MODE_POLICY)
#pragma unroll 1
                for (int c = 0; c < 2; ++c) {
                    if (!chain_live(unit, c)) continue;
                    const int board = b0 + c * 4 + bs;
                    const bool active = board < p.num_boards;
                    uint8_t* buf = chain_base + c * Cfg::kChainBytes;
                    float w2r[kSE];
                    uint4 rr[8];
                    const float bias_next = last_layer ? 0.0f : __ldg(p.layers[L + 1].bias + g);   // p_head: padded to 256
                    // (layer-table fields are read before the wait: behind it each is an L2 round trip on the epilogue's path)
                    const bool want_v = (mode == MODE_SE_RES) && ly->v_pre != nullptr;
                    const float vw = want_v ? __ldg(ly->v_w + g) : 0.0f;
                    const bool save_res = ly->save_res != 0;
                    ptx::mbar_wait_parked(&tmem_full[c], (full_ph >> c) & 1u);
                    full_ph ^= 1u << c;
                    ptx::tc_fence_after();
                    const uint32_t tcol = tmem_base + (static_cast<uint32_t>(quadl * 32) << 16) + c * 256 + (bs >> 1) * 128 + bl * 8;
                    if (mode == MODE_POLICY) {
                        // the 73 planes are lanes 0..72 of the leader; the peer's lanes are padding
                        if (rank == 0 && active)
                            heads_epilogue(p.po, board, bs, 2, ch, lane, quadl, tcol,
                                           reinterpret_cast<float*>(buf) + bs * CB_POLICY_SIZE, s_red + bs * 32, s_vpart + bs * 256);
                        continue;
                    }

                    float scale = 1.0f;
                    if (mode == MODE_SE_RES) {
                        // pass 1: my channel's sum over my board (the bias is already in the accumulator), for both CTAs
                        float sb = 0.0f;
                        if (active) {
#pragma unroll
                            for (int y = 0; y < 8; y += 2) {
                                uint32_t r0[8], r1[8];
                                ptx::tmem_ld_32x8(tcol + y * 16, r0);
                                ptx::tmem_ld_32x8(tcol + (y + 1) * 16, r1);
                                ptx::tmem_ld_wait();
                                float g0 = 0.0f, g1 = 0.0f;
#pragma unroll
                                for (int x = 0; x < 8; ++x) { g0 += __uint_as_float(r0[x]); g1 += __uint_as_float(r1[x]); }
                                sb += g0;
                                sb += g1;
                            }
                        }
                        // excite weights and the block input of my (channel, board), parked by this same thread earlier: their L2
                        // latency runs under the exchange of the board sums
#pragma unroll
                        for (int j = 0; j < kSE; ++j) w2r[j] = __ldg(&ly->se_w2[g * kSE + j]);
                        if (active) {
#pragma unroll
                            for (int y = 0; y < 8; ++y) rr[y] = __ldcg(my_res + (c * 32 + y) * 128);
                        }
                        // (the peer must have finished reading what I sent it for the item before: its squeeze FC is long
                        // over by now, the wait only makes that formal)
                        ptx::mbar_wait_cluster(se_free, free_ph ^ 1u);
                        free_ph ^= 1u;
                        s_part[bs * HID + g] = sb;
                        ptx::st_cluster_f32(peer_part + static_cast<uint32_t>(bs * HID + g) * 4u, sb);
                        ptx::named_bar_sync(1, kEpi);
                        if (et == 0) ptx::mbar_arrive_cluster(peer_se_bar + c * 8);
                        ptx::mbar_wait_cluster(&se_bar[c], (se_ph >> c) & 1u);
                        se_ph ^= 1u << c;
                        // squeeze FC: 4 boards x 16 outputs over 256 channels; warp ew takes 4 of the 64
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const int o = ew * 4 + i, b = o >> 4, j = o & 15;
                            float s = 0.0f;
#pragma unroll
                            for (int k = 0; k < 8; ++k) s = fmaf(__ldg(&ly->se_w1[j * HID + lane + 32 * k]), s_part[b * HID + lane + 32 * k], s);
#pragma unroll
                            for (int d = 16; d >= 1; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
                            if (lane == 0) s_hid[b * kSE + j] = fmaxf(s * (1.0f / 64.0f), 0.0f);
                        }
                        ptx::named_bar_sync(1, kEpi);
                        if (et == 0) ptx::mbar_arrive_cluster(peer_se_free);     // the peer may overwrite my copy of its sums
                        float s = 0.0f;
#pragma unroll
                        for (int j = 0; j < kSE; ++j) s = fmaf(w2r[j], s_hid[bs * kSE + j], s);
                        scale = 1.0f / (1.0f + __expf(-s));
                    }
                    // main pass: (SE scale + residual), ReLU, 16-bit, transpose-store into the buffer of the CTA that holds
                    // my board
                    const bool dump = want_v && p.dbg_backbone != nullptr;
                    auto main_pass = [&](auto se_c, auto save_c, auto wantv_c, auto remote_c) {
                        constexpr bool kSe = decltype(se_c)::v, kSave = decltype(save_c)::v, kWantV = decltype(wantv_c)::v,
                                       kRemote = decltype(remote_c)::v;
                        // address (shared::cluster if remote, else shared::cta) of the even channel of my lane pair in row 0 of
                        // the destination chain buffer
                        const uint32_t buf_ch = dst_chain0 + c * Cfg::kChainBytes + (cb2 & ~2u);
#pragma unroll
                        for (int y = 0; y < 8; ++y) {
                            uint32_t r[8];
                            ptx::tmem_ld_32x8(tcol + y * 16, r);
                            ptx::tmem_ld_wait();
                            float v[8];
#pragma unroll
                            for (int x = 0; x < 8; ++x) v[x] = __uint_as_float(r[x]);
                            if constexpr (kSe) {
                                const uint32_t w[4] = {rr[y].x, rr[y].y, rr[y].z, rr[y].w};
#pragma unroll
                                for (int x = 0; x < 8; ++x) {
                                    const uint32_t h = (x & 1) ? (w[x >> 1] >> 16) : (w[x >> 1] & 0xFFFFu);
                                    v[x] = fmaf(v[x], scale, from16bits<false>(static_cast<uint16_t>(h)));   // the skip path is fp16
                                }
                            }
                            uint32_t pk[4];
#pragma unroll
                            for (int qq = 0; qq < 4; ++qq) pk[qq] = pack_relu16<BF16>(v[2 * qq], v[2 * qq + 1]);
                            const uint32_t row0 = q_row0 + static_cast<uint32_t>(((y + 1) * 2 + bl) * 9 + 1) + (odd ? 4u : 0u);
                            const uint32_t o0 = __shfl_xor_sync(0xffffffffu, odd ? pk[0] : pk[2], 1);
                            const uint32_t o1 = __shfl_xor_sync(0xffffffffu, odd ? pk[1] : pk[3], 1);
                            const uint32_t m0 = odd ? pk[2] : pk[0], m1 = odd ? pk[3] : pk[1];
                            const uint32_t wd[4] = {__byte_perm(m0, o0, sel_lo), __byte_perm(m0, o0, sel_hi),
                                                    __byte_perm(m1, o1, sel_lo), __byte_perm(m1, o1, sel_hi)};
#pragma unroll
                            for (int x = 0; x < 4; ++x) {
                                const uint32_t row = row0 + x;
                                const uint32_t addr = buf_ch + row * 128 + ((cq ^ (row & 7)) << 4);
                                if constexpr (kRemote) ptx::st_cluster_u32(addr, wd[x]);
                                else asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(wd[x]) : "memory");
                            }
                            if constexpr (kSave) {
                                uint32_t hk[4];
#pragma unroll
                                for (int qq = 0; qq < 4; ++qq) hk[qq] = pack_relu16<false>(v[2 * qq], v[2 * qq + 1]);
                                __stcg(my_res + (c * 32 + y) * 128, make_uint4(hk[0], hk[1], hk[2], hk[3]));
                            }
                            if constexpr (kWantV) {
                                float tv[8];
#pragma unroll
                                for (int x = 0; x < 8; ++x) tv[x] = fmaxf(v[x], 0.0f) * vw;
                                const float t = warp_transpose_sum8(tv, lane);
                                if (lane < 8) s_vpart[quadl * 256 + bs * 64 + y * 8 + lane] = t;
                            }
                            if (dump) {
#pragma unroll
                                for (int x = 0; x < 8; ++x) {
                                    const uint16_t h16 = static_cast<uint16_t>((x & 1) ? (pk[x >> 1] >> 16) : (pk[x >> 1] & 0xFFFFu));
                                    p.dbg_backbone[(static_cast<size_t>(board) * HID + g) * 64 + y * 8 + x] = from16bits<BF16>(h16);
                                }
                            }
                            ptx::tmem_st_32x8_bcast(tcol + y * 16, __float_as_uint(bias_next));
                        }
                        ptx::tmem_st_wait();
                    };
                    auto dispatch = [&](auto remote_c) {
                        if (mode == MODE_SE_RES) {
                            if (want_v) main_pass(BoolC<true>{}, BoolC<false>{}, BoolC<true>{}, remote_c);
                            else if (save_res) main_pass(BoolC<true>{}, BoolC<true>{}, BoolC<false>{}, remote_c);
                            else main_pass(BoolC<true>{}, BoolC<false>{}, BoolC<false>{}, remote_c);
                        } else {
                            if (save_res) main_pass(BoolC<false>{}, BoolC<true>{}, BoolC<false>{}, remote_c);
                            else main_pass(BoolC<false>{}, BoolC<false>{}, BoolC<false>{}, remote_c);
                        }
                    };
                    if (active) {
                        if (remote) dispatch(BoolC<true>{});
                        else dispatch(BoolC<false>{});
                    } else {
                        // an empty board slot still re-arms its accumulator columns (they are multiplied every layer)
#pragma unroll
                        for (int y = 0; y < 8; ++y) ptx::tmem_st_32x8_bcast(tcol + y * 16, __float_as_uint(bias_next));
                        ptx::tmem_st_wait();
                    }
                    if (want_v) {
                        ptx::named_bar_sync(1, kEpi);
                        // my CTA's 128 channels: fixed order over the four lane quarters, then one add per CTA into v_pre
                        if (et < 256) {
                            const int n = et;     // board n >> 6 of the chain, pixel n & 63
                            const int bd = b0 + c * 4 + (n >> 6);
                            if (bd < p.num_boards) {
                                const float v = (s_vpart[n] + s_vpart[256 + n]) + (s_vpart[512 + n] + s_vpart[768 + n]);
                                atomicAdd(p.v_pre + bd * 64 + (n & 63), v);
                            }
                        }
                    }
                    ptx::tc_fence_before();
                    ptx::fence_proxy_async_all();
                    ptx::named_bar_sync(1, kEpi);
                    if (et == 0 && !last_layer) ptx::mbar_arrive_cluster(leader_ready + c * 8);
                }
            }
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::cluster_sync_all();      // no CTA leaves while its peer may still write its buffers or signal its barriers
    if (warp == 1) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc_2sm<512>(tmem_base);
    }
}

}  // namespace cb
