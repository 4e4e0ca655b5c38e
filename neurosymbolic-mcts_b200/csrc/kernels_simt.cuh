// kernels_simt.cuh — the non-tensor-core kernels of the leaf-evaluation path:
//   encode_*          packed bitboards -> network input planes   (src/tensor.rs:21-77)
//   pst_eval_kernel   static PeSTO stand-pat term of dM          (src/eval.rs:93-122)
//   conv3x3_fp32 / se_res_fp32   exact-mode (CB_DTYPE_FP32) tower on CUDA cores
//   policy_head_kernel  1x1 conv 73 planes + full-row log-sum-exp + legal gather
//                       (python/model.py:120-125, src/neural_net.rs:279-284, src/tensor.rs:184-213)
//   value_head_kernel   BN(1)+ReLU, cat q, FC 65->256, FC 256->1, k, tanh fusion
//                       (python/model.py:128-132,109-110; src/mcts/tactical_mcts.rs:762-765,787-790)
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>

#include "../../include/caissa_b200.h"
#include "pesto_tables.h"

namespace cb {

// Row of (board b, pixel px = y*8 + x) in an activation buffer.  layout 0: plain NHWC [b][px]
// (exact-mode tower); 1: pair-interleaved act[pair][y][b2][x] (conv_umma.cuh / tower_umma.cuh);
// 2: quad-interleaved act[quad][y][b4][x] (tower_t_umma.cuh).
__host__ __device__ __forceinline__ int act_row(int b, int px, int layout) {
    if (layout == 1) return (b >> 1) * 128 + (((px >> 3) * 2 + (b & 1)) << 3) + (px & 7);
    if (layout == 2) return (b >> 2) * 256 + (((px >> 3) * 4 + (b & 3)) << 3) + (px & 7);
    return b * 64 + px;
}

// ------------------------------------------------------------------ encoder
// 17-bit plane mask of one input pixel (row, col) of the side-to-move view.
// bit p set <=> planes[p][row][col] == 1.0 in the reference's board_to_planes.
__device__ __forceinline__ uint32_t pixel_plane_mask(const cb_board* __restrict__ b, int row, int col) {
    const bool white = b->w_to_move != 0;
    const int sq = (white ? (7 - row) : row) * 8 + col;
    const int stm = white ? 0 : 1;
    uint32_t m = 0;
#pragma unroll
    for (int pt = 0; pt < 6; ++pt) {
        m |= static_cast<uint32_t>((__ldg(&b->pieces[stm * 6 + pt]) >> sq) & 1ull) << pt;
        m |= static_cast<uint32_t>((__ldg(&b->pieces[(1 - stm) * 6 + pt]) >> sq) & 1ull) << (6 + pt);
    }
    if (b->en_passant == sq) m |= 1u << 12;  // 0xFF never equals a square
    const uint32_t c = b->castling;          // WK=1 WQ=2 BK=4 BQ=8
    const uint32_t rel = white ? c : (((c >> 2) & 3u) | ((c & 3u) << 2));
    m |= rel << 13;
    return m;
}

// NCHW f32 [B][17][8][8] — the reference's Vec<f32> layout (cb_encode hook).
__global__ void encode_nchw_f32_kernel(const cb_board* __restrict__ boards, int B, float* __restrict__ out) {
    const int gid = blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= B * 64) return;
    const int b = gid >> 6, px = gid & 63;
    const uint32_t m = pixel_plane_mask(boards + b, px >> 3, px & 7);
    float* o = out + static_cast<size_t>(b) * CB_PLANES + px;
#pragma unroll
    for (int p = 0; p < 17; ++p) o[p * 64] = (m >> p) & 1u ? 1.0f : 0.0f;
}

// NHWC f32 [B][64][17] — input of the exact-mode tower.
__global__ void encode_nhwc_f32_kernel(const cb_board* __restrict__ boards, int B, float* __restrict__ out) {
    const int gid = blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= B * 64) return;
    const int b = gid >> 6, px = gid & 63;
    const uint32_t m = pixel_plane_mask(boards + b, px >> 3, px & 7);
    float* o = out + static_cast<size_t>(gid) * 17;
#pragma unroll
    for (int p = 0; p < 17; ++p) o[p] = (m >> p) & 1u ? 1.0f : 0.0f;
}

// Group-interleaved 16-bit planes [Bpad/G][8 y][G boards][8 x][64] with G = 2 (pair) or 4 (quad):
// channels 0..16 are the planes, 17..63 zero (K padding of the start conv to one 64-channel
// TMA/UMMA chunk).  8 threads per pixel row, one 16-byte store each => fully coalesced 512 B per
// warp instruction.  Boards at index >= B (batch padded to a whole group) are written as zeros.
template <bool BF16>
__global__ void encode_tiled16_kernel(const cb_board* __restrict__ boards, int B, int Bpad, int group_log2,
                                      uint4* __restrict__ out) {
    const int gid = blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= Bpad * 512) return;
    const int j = gid & 7, grow = gid >> 3;          // global pixel row = grp*(64*G) + (y*G + bi)*8 + x
    const int G = 1 << group_log2;
    const int m = grow & (64 * G - 1);
    const int b = (grow >> (6 + group_log2)) * G + ((m >> 3) & (G - 1));
    const int y = m >> (3 + group_log2), x = m & 7;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (j < 3 && b < B) {
        const uint32_t msk = pixel_plane_mask(boards + b, y, x) >> (j * 8);
        const uint32_t one = BF16 ? 0x3F80u : 0x3C00u;
        uint32_t w[4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
            w[i] = ((msk >> (2 * i)) & 1u ? one : 0u) | ((msk >> (2 * i + 1)) & 1u ? (one << 16) : 0u);
        v = make_uint4(w[0], w[1], w[2], w[3]);
    }
    out[gid] = v;
}

// ------------------------------------------------------------- static PeSTO
__global__ void pst_eval_kernel(const cb_board* __restrict__ boards, int B, int32_t* __restrict__ cp) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const cb_board* bd = boards + b;
    int mg[2] = {0, 0}, eg[2] = {0, 0}, phase = 0;
    for (int color = 0; color < 2; ++color)
        for (int pt = 0; pt < 6; ++pt) {
            uint64_t bb = bd->pieces[color * 6 + pt];
            while (bb) {
                const int sq = __ffsll(static_cast<long long>(bb)) - 1;
                const int t = color == 0 ? (sq ^ 56) : sq;
                mg[color] += CB_MG_VALUE[pt] + CB_MG_PST[pt][t];
                eg[color] += CB_EG_VALUE[pt] + CB_EG_PST[pt][t];
                phase += CB_PHASE_INC[pt];
                bb &= bb - 1;
            }
        }
    const int mgp = phase < 24 ? phase : 24;
    const int score = ((mg[0] - mg[1]) * mgp + (eg[0] - eg[1]) * (24 - mgp)) / 24;  // truncating, like i32 '/'
    cp[b] = bd->w_to_move ? score : -score;
}

// ------------------------------------------------- exact-mode tower (fp32)
// One CTA per board.  in NHWC f32 [B][64][CIN]; w [9][CIN][HID] (BN folded);
// out NHWC f32 [B][64][HID] = (relu?)(conv + bias).
template <int HID>
__global__ void __launch_bounds__(256)
conv3x3_fp32_kernel(const float* __restrict__ in, int cin, const float* __restrict__ w, const float* __restrict__ bias,
                    float* __restrict__ out, int relu) {
    extern __shared__ float s_in[];  // [10][10][cin], zero halo
    constexpr int kGroups = 256 / HID;       // pixel groups across the CTA
    constexpr int kPix = 64 / kGroups;       // pixels per thread
    const int b = blockIdx.x;
    for (int i = threadIdx.x; i < 100 * cin; i += 256) {
        const int pos = i / cin, c = i % cin;
        const int y = pos / 10 - 1, x = pos % 10 - 1;
        s_in[i] = (y >= 0 && y < 8 && x >= 0 && x < 8) ? in[(static_cast<size_t>(b) * 64 + y * 8 + x) * cin + c] : 0.0f;
    }
    __syncthreads();
    const int co = threadIdx.x % HID, g = threadIdx.x / HID;
    float acc[kPix];
#pragma unroll
    for (int i = 0; i < kPix; ++i) acc[i] = 0.0f;
    for (int tap = 0; tap < 9; ++tap) {
        const int ky = tap / 3, kx = tap % 3;
        for (int c = 0; c < cin; ++c) {
            const float wv = __ldg(&w[(static_cast<size_t>(tap) * cin + c) * HID + co]);
#pragma unroll
            for (int i = 0; i < kPix; ++i) {
                const int px = g * kPix + i;
                acc[i] = fmaf(s_in[((px >> 3) + ky) * 10 * cin + ((px & 7) + kx) * cin + c], wv, acc[i]);
            }
        }
    }
    const float bv = bias[co];
#pragma unroll
    for (int i = 0; i < kPix; ++i) {
        float v = acc[i] + bv;
        if (relu) v = fmaxf(v, 0.0f);
        out[(static_cast<size_t>(b) * 64 + g * kPix + i) * HID + co] = v;
    }
}

// Squeeze-excite + residual + ReLU (+ value 1x1) for the exact-mode tower.
// One CTA per board, HID threads (thread = channel).
template <int HID>
__global__ void __launch_bounds__(HID)
se_res_fp32_kernel(const float* __restrict__ o, const float* __restrict__ res, const float* __restrict__ w1,
                   const float* __restrict__ w2, float* __restrict__ out, const float* __restrict__ v_w,
                   float* __restrict__ v_pre) {
    constexpr int kSE = HID / 16;
    __shared__ float s_mean[HID];
    __shared__ float s_hid[kSE];
    __shared__ float s_red[64];
    const int b = blockIdx.x, c = threadIdx.x;
    const size_t base = static_cast<size_t>(b) * 64 * HID;
    float sum = 0.0f;
    for (int px = 0; px < 64; ++px) sum += o[base + px * HID + c];
    s_mean[c] = sum * (1.0f / 64.0f);
    if (c < 64) s_red[c] = 0.0f;
    __syncthreads();
    if (c < kSE) {
        float s = 0.0f;
        for (int i = 0; i < HID; ++i) s += w1[c * HID + i] * s_mean[i];
        s_hid[c] = fmaxf(s, 0.0f);
    }
    __syncthreads();
    float s = 0.0f;
#pragma unroll
    for (int j = 0; j < kSE; ++j) s += w2[c * kSE + j] * s_hid[j];
    const float scale = 1.0f / (1.0f + expf(-s));
    const float vw = v_w ? v_w[c] : 0.0f;
    for (int px = 0; px < 64; ++px) {
        const float v = fmaxf(o[base + px * HID + c] * scale + res[base + px * HID + c], 0.0f);
        out[base + px * HID + c] = v;
        if (v_pre) {
            float t = v * vw;
#pragma unroll
            for (int d = 16; d >= 1; d >>= 1) t += __shfl_xor_sync(0xffffffffu, t, d);
            if ((c & 31) == 0) atomicAdd(&s_red[px], t);
        }
    }
    __syncthreads();
    if (v_pre && c < 64) v_pre[static_cast<size_t>(b) * 64 + c] = s_red[c];
}

// -------------------------------------------------------------- policy head
template <typename T>
__device__ __forceinline__ float to_f32(T v);
template <>
__device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <>
__device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <>
__device__ __forceinline__ float to_f32<__half>(__half v) { return __half2float(v); }

struct PolicyOut {
    float* probs;              // [B][4672] or nullptr (compat mode, = exp(log_softmax))
    const uint16_t* move_idx;  // CSR of legal policy indices or nullptr
    const uint32_t* move_off;  // [B+1]
    float* priors;             // [move_off[B]]
    const uint16_t* move_cnt;  // nullable: fixed-stride lists instead of CSR (board b: move_idx[b*256 .. +move_cnt[b]),
                               // priors at the same offsets) — the layout the device-side tree kernels write
};

constexpr int kPolicyThreads = 320;  // 4 pixel groups x 73 planes = 292 active in the GEMV phase

// One CTA per board.  p: NHWC [B][64][HID] post-ReLU policy features; wt: [HID][73] (p_head
// transposed), bias [73].  Logits stay in shared memory; the 4672-wide log-sum-exp is taken
// over the whole row (so that the reference's "sum == 0 -> uniform" fallback is reproducible),
// then either all probabilities are written or only the legal moves are gathered and
// renormalised in list order.
template <typename T, int HID>
__global__ void __launch_bounds__(kPolicyThreads)
policy_head_kernel(const T* __restrict__ p, const float* __restrict__ wt, const float* __restrict__ bias, PolicyOut po,
                   int pair_layout) {
    extern __shared__ float sm[];
    float* s_p = sm;                      // [64][HID]
    float* s_w = s_p + 64 * HID;          // [HID][73]
    float* s_logit = s_w + HID * 73;      // [4672]
    __shared__ float s_red[16];
    __shared__ float s_total;
    const int b = blockIdx.x, tid = threadIdx.x;
    for (int i = tid; i < 64 * HID; i += kPolicyThreads) {
        const int px = i / HID, c = i % HID;
        s_p[i] = to_f32<T>(p[static_cast<size_t>(act_row(b, px, pair_layout)) * HID + c]);
    }
    for (int i = tid; i < HID * 73; i += kPolicyThreads) s_w[i] = wt[i];
    __syncthreads();
    if (tid < 292) {
        const int c = tid % 73, g = tid / 73;
        float acc[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[i] = 0.0f;
        const float* pp = s_p + g * 16 * HID;
        for (int k = 0; k < HID; ++k) {
            const float wv = s_w[k * 73 + c];
#pragma unroll
            for (int i = 0; i < 16; ++i) acc[i] = fmaf(pp[i * HID + k], wv, acc[i]);
        }
        const float bv = bias[c];
#pragma unroll
        for (int i = 0; i < 16; ++i) s_logit[(g * 16 + i) * 73 + c] = acc[i] + bv;
    }
    __syncthreads();
    // row max
    float mx = -INFINITY;
    for (int i = tid; i < CB_POLICY_SIZE; i += kPolicyThreads) mx = fmaxf(mx, s_logit[i]);
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, d));
    if ((tid & 31) == 0) s_red[tid >> 5] = mx;
    __syncthreads();
    mx = s_red[0];
#pragma unroll
    for (int i = 1; i < kPolicyThreads / 32; ++i) mx = fmaxf(mx, s_red[i]);
    __syncthreads();
    float se = 0.0f;
    for (int i = tid; i < CB_POLICY_SIZE; i += kPolicyThreads) se += expf(s_logit[i] - mx);
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) se += __shfl_xor_sync(0xffffffffu, se, d);
    if ((tid & 31) == 0) s_red[tid >> 5] = se;
    __syncthreads();
    se = 0.0f;
#pragma unroll
    for (int i = 0; i < kPolicyThreads / 32; ++i) se += s_red[i];
    const float lse = mx + logf(se);

    if (po.probs) {
        float* o = po.probs + static_cast<size_t>(b) * CB_POLICY_SIZE;
        for (int i = tid; i < CB_POLICY_SIZE; i += kPolicyThreads) o[i] = expf(s_logit[i] - lse);
    }
    if (po.move_idx) {
        const uint32_t lo = po.move_cnt ? static_cast<uint32_t>(b) * CB_MAX_MOVES : po.move_off[b];
        const uint32_t hi = po.move_cnt ? lo + po.move_cnt[b] : po.move_off[b + 1];
        const int n = static_cast<int>(hi - lo);
        __syncthreads();
        // gather into the (now free) feature area, then a sequential f32 sum in list
        // order exactly like the reference's loop (src/tensor.rs:186-201)
        float* g = s_p;
        for (int i = tid; i < n; i += kPolicyThreads) {
            const int idx = po.move_idx[lo + i];
            g[i] = idx < CB_POLICY_SIZE ? expf(s_logit[idx] - lse) : 0.0f;
        }
        __syncthreads();
        if (tid == 0) {
            float t = 0.0f;
            for (int i = 0; i < n; ++i) t += g[i];
            s_total = t;
        }
        __syncthreads();
        const float total = s_total;
        for (int i = tid; i < n; i += kPolicyThreads)
            po.priors[lo + i] = total > 0.0f ? g[i] / total : 1.0f / static_cast<float>(n);
    }
}

// --------------------------------------------------------------- value head
struct ValueParams {
    const float* v_pre;    // [B][64] <x, v_conv.weight> without bias
    const float* q;        // [B] material scalar (nullptr => 0)
    float conv_bias;       // v_conv.bias
    float bn_scale;        // v_bn folded
    float bn_bias;
    const float* fc_wt;    // [65][256] v_fc.weight transposed
    const float* fc_b;     // [256]
    const float* out_w;    // [256]
    float out_b;
    float k;               // 0.47 * softplus(k_logit)
    int material_on;
    int pair_layout;       // layout code of v_pre rows, see act_row()
    int B;
    float* v_logit;        // [B]
    float* v_final;        // [B] tanh(v_logit + k q) or tanh(v_logit)
    float* k_out;          // [B]
};

constexpr int kValueBoardsPerCta = 8;

__global__ void __launch_bounds__(256) value_head_kernel(const ValueParams vp) {
    __shared__ float s_x[kValueBoardsPerCta][65];
    __shared__ float s_red[kValueBoardsPerCta][8];
    const int j = threadIdx.x;
    const int b0 = blockIdx.x * kValueBoardsPerCta;
    for (int i = j; i < kValueBoardsPerCta * 65; i += 256) {
        const int bl = i / 65, e = i % 65, b = b0 + bl;
        float v = 0.0f;
        if (b < vp.B) {
            if (e < 64)
                v = fmaxf((vp.v_pre[act_row(b, e, vp.pair_layout)] + vp.conv_bias) * vp.bn_scale + vp.bn_bias, 0.0f);
            else
                v = vp.q ? vp.q[b] : 0.0f;
        }
        s_x[bl][e] = v;
    }
    float w[65];
#pragma unroll
    for (int e = 0; e < 65; ++e) w[e] = __ldg(&vp.fc_wt[e * 256 + j]);
    const float fb = vp.fc_b[j], ow = vp.out_w[j];
    __syncthreads();
#pragma unroll 1
    for (int bl = 0; bl < kValueBoardsPerCta; ++bl) {
        float h = fb;
#pragma unroll
        for (int e = 0; e < 65; ++e) h = fmaf(w[e], s_x[bl][e], h);
        float t = fmaxf(h, 0.0f) * ow;
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) t += __shfl_xor_sync(0xffffffffu, t, d);
        if ((j & 31) == 0) s_red[bl][j >> 5] = t;
    }
    __syncthreads();
    if (j < kValueBoardsPerCta && b0 + j < vp.B) {
        float v = vp.out_b;
#pragma unroll
        for (int i = 0; i < 8; ++i) v += s_red[j][i];
        const float q = vp.q ? vp.q[b0 + j] : 0.0f;
        vp.v_logit[b0 + j] = v;
        vp.k_out[b0 + j] = vp.k;
        vp.v_final[b0 + j] = vp.material_on ? tanhf(v + vp.k * q) : tanhf(v);
    }
}

// activation rows -> f32 NCHW [B][HID][64] (debug/parity hook), any input type / layout.
template <typename T>
__global__ void act_to_nchw_f32_kernel(const T* __restrict__ in, int hid, int total, int pair_layout,
                                       float* __restrict__ out) {
    const int gid = blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= total) return;
    const int c = gid % hid, px = (gid / hid) & 63, b = gid / (hid * 64);
    out[gid / (hid * 64) * hid * 64 + c * 64 + px] =
        to_f32<T>(in[static_cast<size_t>(act_row(b, px, pair_layout)) * hid + c]);
}

}  // namespace cb
