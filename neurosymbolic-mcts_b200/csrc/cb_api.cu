// cb_api.cu — the C ABI of libcaissa_b200.so (include/caissa_b200.h): handle,
// weight folding / packing, device buffers, TMA descriptors, forward-pass
// orchestration and CUDA-graph replay.  sm_100a only; no CPU fallback.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <string.h>

#include <stdlib.h>

#include <chrono>
#include <map>
#include <new>
#include <string>
#include <algorithm>
#include <vector>

#include "../../include/caissa_b200.h"
#include "conv_umma.cuh"
#include "kernels_simt.cuh"
#include "policy_umma.cuh"
#include "tower_umma.cuh"
#include "tower_r_umma.cuh"
#include "tower_r2_umma.cuh"
#include "mcts_device.cuh"

namespace {

using namespace cb;

constexpr float kBnEps = 1e-5f;  // torch BatchNorm2d default (python/model.py uses the default)
constexpr size_t kFlushBytes = 256ull << 20;

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct ConvLayer {
    int cin = 0;          // logical input channels (17 or HID)
    int cin_pad = 0;      // padded to a multiple of 64 for the tensor-core path
    void* w16 = nullptr;  // [9][HID][cin_pad] 16-bit, BN scale folded
    float* w32 = nullptr; // [9][cin][HID] fp32, BN scale folded (exact mode)
    void* w16s = nullptr; // CB_DTYPE_FP32 on the tensor cores: [2][9][HID][cin_pad] fp16, high halves then low halves
    float* bias = nullptr;
    CUtensorMap tm_w;
    CUtensorMap tm_w_half;   // 64-row box (cluster mode of the resident tower)
    CUtensorMap tm_w128;     // 128-row box: one CTA's share of a stage in the resident kernels (all of it for 128 channels)
};

}  // namespace

struct cb_handle {
    int device = 0, blocks = 0, hid = 0, dtype = 0;
    int sm_count = 0;
    bool loaded = false;
    cudaStream_t stream = nullptr;
    cudaStream_t stream2 = nullptr;             // side branch: value head runs beside the policy head
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr, ev3 = nullptr;
    std::string err;
    EncodeTiledFn encode_tiled = nullptr;

    // weights
    std::vector<ConvLayer> convs;  // start, (conv1, conv2) x blocks, p_conv
    std::vector<float*> se_w1, se_w2;
    float* p_wt = nullptr;   // [HID][73]
    float* p_bias = nullptr; // [73]
    void* p_w16 = nullptr;   // [80][HID] 16-bit (rows 73..79 zero), tensor-core policy head
    CUtensorMap tm_pw;
    float* v_w = nullptr;    // [HID]
    float* fc_wt = nullptr;  // [65][256]
    float* fc_b = nullptr;
    float* out_w = nullptr;
    float v_conv_bias = 0, v_bn_scale = 1, v_bn_bias = 0, out_b = 0, k = 0;

    // device buffers (capacity cap boards, even)
    int cap = 0;
    size_t csr_cap = 0;
    // inputs live in ONE device block and ONE pinned block (boards | q | move_off | move_idx) and the
    // outputs in another (v_logit | v_final | k | priors) so a full batch moves with one copy each way
    uint8_t* d_in = nullptr;
    uint8_t* d_out = nullptr;
    uint8_t* h_in = nullptr;
    uint8_t* h_out = nullptr;
    size_t in_idx_off = 0, out_pri_off = 0;   // byte offsets of move_idx / priors inside the blocks
    cb_board* d_boards = nullptr;
    float* d_q = nullptr;
    void* d_x0 = nullptr;       // encoder output
    void* d_act[3] = {nullptr, nullptr, nullptr};
    float* d_tmp32 = nullptr;   // exact mode: pre-SE conv2 output
    float* d_vpre = nullptr;
    float* d_probs = nullptr;   // [cap][4672], allocated on first compat call
    int probs_cap = 0;
    uint16_t* d_move_idx = nullptr;
    uint32_t* d_move_off = nullptr;
    float* d_priors = nullptr;
    float* d_vlogit = nullptr;
    float* d_vfinal = nullptr;
    float* d_k = nullptr;
    float* d_scratch = nullptr; // debug / encode hook output
    size_t scratch_floats = 0;
    void* d_flush = nullptr;
    CUtensorMap tm_x0, tm_act_ld[3], tm_act_st[3];
    CUtensorMap* d_maps = nullptr;    // fused tower: tensor-map table (x0, act loads, act stores, weights)
    cb::TowerLayer* d_layers = nullptr;
    bool tower_ready = false;
    bool use_fused = true;
    long long* d_trace = nullptr;   // debug timeline of CTA 0 of the resident tower (cb_debug_tower_trace)
    // smem-resident fused tower for the 128-channel net (tower_r_umma.cuh)
    void* p_w16r = nullptr;              // p_head weights padded to [128][HID] 16-bit (rows 73.. zero)
    float* p_bias_r = nullptr;           // p_head bias padded to 128
    CUtensorMap tm_pwr, tm_pwr_half;
    CUtensorMap* d_maps_r_half = nullptr;
    bool use_cluster = false;
    CUtensorMap* d_maps_r = nullptr;
    cb::TowerRLayer* d_layers_r = nullptr;
    uint4* d_res_scratch = nullptr;
    float* d_dbg_backbone = nullptr;     // set only inside cb_debug_backbone
    unsigned long long* d_cta_times = nullptr;   // set only inside cb_debug_cta_times
    bool tower_r_ready = false;
    bool use_resident = true;
    bool use_split = true;               // CB_DTYPE_FP32 handles: fp16-pair operands on the tensor cores (CB_FP32_EXACT=1: CUDA cores)
    void* p_w16s = nullptr;              // split mode: p_head [2][HID][HID] fp16 (high halves, low halves)
    bool chess_tables_uploaded = false;  // attack tables of the device-side tree kernels

    // pinned staging
    cb_board* h_boards = nullptr;
    float* h_q = nullptr;
    uint16_t* h_move_idx = nullptr;
    uint32_t* h_move_off = nullptr;
    float* h_priors = nullptr;
    float* h_small = nullptr;   // v_logit | v_final | k, 3*cap

    // cb_eval_leaves_submit / _wait: the batch in flight (-1: none)
    int pending_B = -1;
    size_t pending_priors = 0;

    // resident state
    int staged_B = 0;
    size_t staged_moves = 0;

    std::map<uint64_t, cudaGraphExec_t> graphs;
    std::map<uint64_t, int> graph_seen;          // how often a forward shape was requested (run_forward)
    cb_timing timing = {};
    bool use_graphs = true;
    int layer_limit = 0;  // debug: stop the tower after this many conv launches (0 = all)
};

namespace {

#define CB_CUDA(h, expr)                                                                        \
    do {                                                                                        \
        cudaError_t _e = (expr);                                                                \
        if (_e != cudaSuccess) {                                                                \
            (h)->err = std::string(#expr) + ": " + cudaGetErrorString(_e);                      \
            return CB_ECUDA;                                                                    \
        }                                                                                       \
    } while (0)

int fail(cb_handle* h, int code, const std::string& msg) {
    if (h) h->err = msg;
    return code;
}

size_t elem_bytes(const cb_handle* h) { return h->dtype == CB_DTYPE_FP32 ? 4 : 2; }

uint16_t to16(float v, int dtype) {
    if (dtype == CB_DTYPE_BF16) {
        __nv_bfloat16 b = __float2bfloat16_rn(v);
        uint16_t u;
        memcpy(&u, &b, 2);
        return u;
    }
    __half hv = __float2half_rn(v);
    uint16_t u;
    memcpy(&u, &hv, 2);
    return u;
}

void free_dev(void* p) {
    if (p) cudaFree(p);
}

void drop_graphs(cb_handle* h) {
    for (auto& kv : h->graphs) cudaGraphExecDestroy(kv.second);
    h->graphs.clear();
}

// 5-D map over the pair-interleaved activation act[pair][y][b2][x][C]; box_y = 10 for the
// halo'd slab loads (fetched at y = -1), 8 for residual-free output stores.
int make_act_map(cb_handle* h, CUtensorMap* tm, void* base, int channels, int boards, int box_y, int group = 2) {
    const cuuint64_t C = (cuuint64_t)channels, G = (cuuint64_t)group;
    const cuuint64_t dims[5] = {C, 8, G, 8, (cuuint64_t)boards / G};
    const cuuint64_t strides[4] = {C * 2, C * 16, C * 16 * G, C * 128 * G};
    const cuuint32_t box[5] = {64, 8, (cuuint32_t)group, (cuuint32_t)box_y, 1};
    const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    const CUtensorMapDataType dt =
        h->dtype == CB_DTYPE_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
    CUresult r = h->encode_tiled(tm, dt, 5, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                 CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(h, CB_ECUDA, "cuTensorMapEncodeTiled(activation) failed: " + std::to_string((int)r));
    return CB_OK;
}

int make_weight_map(cb_handle* h, ConvLayer& L) {
    const cuuint64_t dims[2] = {(cuuint64_t)L.cin_pad, (cuuint64_t)9 * h->hid};
    const cuuint64_t strides[1] = {(cuuint64_t)L.cin_pad * 2};
    const cuuint32_t box[2] = {64, (cuuint32_t)h->hid};
    const cuuint32_t estr[2] = {1, 1};
    const CUtensorMapDataType dt =
        h->dtype == CB_DTYPE_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
    CUresult r = h->encode_tiled(&L.tm_w, dt, 2, L.w16, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                 CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(h, CB_ECUDA, "cuTensorMapEncodeTiled(weights) failed: " + std::to_string((int)r));
    const cuuint32_t box_half[2] = {64, 64};
    r = h->encode_tiled(&L.tm_w_half, dt, 2, L.w16, dims, strides, box_half, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(h, CB_ECUDA, "cuTensorMapEncodeTiled(half weights) failed: " + std::to_string((int)r));
    const cuuint32_t box128[2] = {64, 128};
    r = h->encode_tiled(&L.tm_w128, dt, 2, L.w16, dims, strides, box128, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(h, CB_ECUDA, "cuTensorMapEncodeTiled(128-row weights) failed: " + std::to_string((int)r));
    return CB_OK;
}

// Split mode: one tensor [2 * taps * HID][cin_pad] (high halves, then low halves), 128-row boxes.
int make_split_map(cb_handle* h, CUtensorMap* tm, void* w, int cin_pad, int taps) {
    const cuuint64_t dims[2] = {(cuuint64_t)cin_pad, (cuuint64_t)2 * taps * h->hid};
    const cuuint64_t strides[1] = {(cuuint64_t)cin_pad * 2};
    const cuuint32_t box[2] = {64, 128};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = h->encode_tiled(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, w, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                 CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(h, CB_ECUDA, "cuTensorMapEncodeTiled(split weights) failed: " + std::to_string((int)r));
    return CB_OK;
}

void free_buffers(cb_handle* h) {
    drop_graphs(h);
    free_dev(h->d_in); free_dev(h->d_out); free_dev(h->d_x0);
    h->d_in = nullptr; h->d_out = nullptr;
    for (int i = 0; i < 3; ++i) { free_dev(h->d_act[i]); h->d_act[i] = nullptr; }
    free_dev(h->d_tmp32); free_dev(h->d_vpre); free_dev(h->d_probs);
    free_dev(h->d_scratch);
    h->d_boards = nullptr; h->d_q = nullptr; h->d_x0 = nullptr; h->d_tmp32 = nullptr; h->d_vpre = nullptr;
    h->d_probs = nullptr; h->d_move_idx = nullptr; h->d_move_off = nullptr; h->d_priors = nullptr;
    h->d_vlogit = nullptr; h->d_vfinal = nullptr; h->d_k = nullptr; h->d_scratch = nullptr;
    if (h->h_in) cudaFreeHost(h->h_in);
    if (h->h_out) cudaFreeHost(h->h_out);
    h->h_in = nullptr; h->h_out = nullptr;
    h->h_boards = nullptr; h->h_q = nullptr; h->h_move_idx = nullptr; h->h_move_off = nullptr;
    h->h_priors = nullptr; h->h_small = nullptr;
    h->cap = 0; h->csr_cap = 0; h->probs_cap = 0; h->scratch_floats = 0; h->staged_B = 0;
}

int build_tower_tables(cb_handle* h);

// Grow device + pinned buffers to hold B boards (rounded up to a whole 2-board tile).
int ensure_capacity(cb_handle* h, int B) {
    const int need = (B + 3) & ~3;
    if (need <= h->cap) return CB_OK;
    int cap = h->cap ? h->cap : 64;
    while (cap < need) cap *= 2;
    free_buffers(h);
    const size_t eb = elem_bytes(h);
    const size_t csr = (size_t)cap * CB_MAX_MOVES;
    const size_t off_q = (size_t)cap * sizeof(cb_board), off_mo = off_q + (size_t)cap * 4;
    const size_t off_mi = (off_mo + ((size_t)cap + 1) * 4 + 15) & ~(size_t)15, in_bytes = off_mi + csr * 2;
    const size_t off_pri = (size_t)cap * 12, out_bytes = off_pri + csr * 4;
    CB_CUDA(h, cudaMalloc(&h->d_in, in_bytes));
    CB_CUDA(h, cudaMemset(h->d_in, 0, in_bytes));
    CB_CUDA(h, cudaMalloc(&h->d_out, out_bytes));
    CB_CUDA(h, cudaHostAlloc(&h->h_in, in_bytes, cudaHostAllocDefault));
    CB_CUDA(h, cudaHostAlloc(&h->h_out, out_bytes, cudaHostAllocDefault));
    memset(h->h_in, 0, in_bytes);
    h->in_idx_off = off_mi;
    h->out_pri_off = off_pri;
    h->d_boards = (cb_board*)h->d_in;            h->h_boards = (cb_board*)h->h_in;
    h->d_q = (float*)(h->d_in + off_q);          h->h_q = (float*)(h->h_in + off_q);
    h->d_move_off = (uint32_t*)(h->d_in + off_mo); h->h_move_off = (uint32_t*)(h->h_in + off_mo);
    h->d_move_idx = (uint16_t*)(h->d_in + off_mi); h->h_move_idx = (uint16_t*)(h->h_in + off_mi);
    h->d_vlogit = (float*)h->d_out;              h->h_small = (float*)h->h_out;
    h->d_vfinal = h->d_vlogit + cap;
    h->d_k = h->d_vlogit + 2 * (size_t)cap;
    h->d_priors = (float*)(h->d_out + off_pri);  h->h_priors = (float*)(h->h_out + off_pri);
    CB_CUDA(h, cudaMalloc(&h->d_x0, (size_t)cap * 64 * (h->dtype == CB_DTYPE_FP32 ? 17 * 4 : 64 * 2)));
    for (int i = 0; i < 3; ++i) CB_CUDA(h, cudaMalloc(&h->d_act[i], (size_t)cap * 64 * h->hid * eb));
    if (h->dtype == CB_DTYPE_FP32) CB_CUDA(h, cudaMalloc(&h->d_tmp32, (size_t)cap * 64 * h->hid * 4));
    CB_CUDA(h, cudaMalloc(&h->d_vpre, (size_t)cap * 64 * 4));
    h->cap = cap;
    h->csr_cap = csr;
    if (h->dtype == CB_DTYPE_FP32 && h->hid == 128) {
        free_dev(h->d_res_scratch);
        h->d_res_scratch = nullptr;
        CB_CUDA(h, cudaMalloc(&h->d_res_scratch, (size_t)h->sm_count * TowerRCfg::kScratchPerCta * sizeof(uint4)));
        int rc;
        if (h->loaded && (rc = build_tower_tables(h))) return rc;
    }
    if (h->dtype != CB_DTYPE_FP32) {
        int rc = make_act_map(h, &h->tm_x0, h->d_x0, 64, cap, 10);
        if (rc) return rc;
        for (int i = 0; i < 3; ++i) {
            if ((rc = make_act_map(h, &h->tm_act_ld[i], h->d_act[i], h->hid, cap, 10))) return rc;
            if ((rc = make_act_map(h, &h->tm_act_st[i], h->d_act[i], h->hid, cap, 8))) return rc;
        }
        if (h->hid == 128 || h->hid == 256) {
            free_dev(h->d_res_scratch);
            h->d_res_scratch = nullptr;
            CB_CUDA(h, cudaMalloc(&h->d_res_scratch, (size_t)h->sm_count * TowerRCfg::kScratchPerCta * sizeof(uint4)));
        }
        if (h->loaded && (rc = build_tower_tables(h))) return rc;
    }
    return CB_OK;
}

// Device-side tables of the fused tower kernel.  Rebuilt whenever buffers or weights change.
int build_tower_tables(cb_handle* h) {
    h->tower_ready = false;
    if (!h->cap || h->convs.empty()) return CB_OK;
    const int nl = (int)h->convs.size();
    if (h->dtype == CB_DTYPE_FP32) {
        // fp16-pair mode of the resident tower (128 channels): the map table points at the split weight tensors
        h->tower_r_ready = false;
        if (h->hid != 128 || !h->p_w16s) return CB_OK;
        std::vector<CUtensorMap> mr(nl + 1);
        int rc;
        for (int l = 0; l < nl; ++l)
            if ((rc = make_split_map(h, &mr[l], h->convs[l].w16s, h->convs[l].cin_pad, 9))) return rc;
        if ((rc = make_split_map(h, &mr[nl], h->p_w16s, h->hid, 1))) return rc;
        std::vector<TowerRLayer> lr(nl + 1);
        memset(lr.data(), 0, sizeof(TowerRLayer) * (nl + 1));
        for (int l = 0; l < nl; ++l) {
            lr[l].w_map = l;
            lr[l].kc_in = l == 0 ? 1 : h->hid / 64;
            lr[l].mode = (l >= 1 && l < nl - 1 && (l & 1) == 0) ? MODE_SE_RES : MODE_RELU;
            lr[l].bias = h->convs[l].bias;
            lr[l].ntaps = 9;
            lr[l].w_lo_off = 9 * h->hid;
        }
        lr[nl].w_map = nl; lr[nl].kc_in = h->hid / 64; lr[nl].mode = MODE_POLICY; lr[nl].bias = h->p_bias_r;
        lr[nl].ntaps = 1; lr[nl].w_row_off = -4 * h->hid; lr[nl].w_lo_off = h->hid;
        lr[0].save_res = 1;
        for (int i = 0; i < h->blocks; ++i) {
            TowerRLayer& b = lr[2 + 2 * i];
            b.se_w1 = h->se_w1[i]; b.se_w2 = h->se_w2[i];
            b.save_res = i < h->blocks - 1;
            if (i == h->blocks - 1) { b.v_w = h->v_w; b.v_pre = h->d_vpre; }
        }
        free_dev(h->d_maps_r); free_dev(h->d_layers_r);
        h->d_maps_r = nullptr; h->d_layers_r = nullptr;
        CB_CUDA(h, cudaMalloc(&h->d_maps_r, mr.size() * sizeof(CUtensorMap)));
        CB_CUDA(h, cudaMalloc(&h->d_layers_r, lr.size() * sizeof(TowerRLayer)));
        CB_CUDA(h, cudaMemcpy(h->d_maps_r, mr.data(), mr.size() * sizeof(CUtensorMap), cudaMemcpyHostToDevice));
        CB_CUDA(h, cudaMemcpy(h->d_layers_r, lr.data(), lr.size() * sizeof(TowerRLayer), cudaMemcpyHostToDevice));
        h->tower_r_ready = true;
        return CB_OK;
    }
    std::vector<CUtensorMap> maps(7 + nl);
    maps[0] = h->tm_x0;
    for (int i = 0; i < 3; ++i) { maps[1 + i] = h->tm_act_ld[i]; maps[4 + i] = h->tm_act_st[i]; }
    for (int l = 0; l < nl; ++l) maps[7 + l] = h->convs[l].tm_w;
    std::vector<TowerLayer> ly(nl);
    memset(ly.data(), 0, sizeof(TowerLayer) * nl);
    int cur = 0;
    ly[0].in_map = 0; ly[0].w_map = 7; ly[0].out_map = 4 + 0; ly[0].kc_in = 1; ly[0].mode = MODE_RELU;
    ly[0].bias = h->convs[0].bias;
    for (int i = 0; i < h->blocks; ++i) {
        const int nxt = cur == 0 ? 2 : 0;
        TowerLayer& a = ly[1 + 2 * i];
        a.in_map = 1 + cur; a.w_map = 7 + 1 + 2 * i; a.out_map = 4 + 1; a.kc_in = h->hid / 64; a.mode = MODE_RELU;
        a.bias = h->convs[1 + 2 * i].bias;
        TowerLayer& b = ly[2 + 2 * i];
        b.in_map = 1 + 1; b.w_map = 7 + 2 + 2 * i; b.out_map = 4 + nxt; b.kc_in = h->hid / 64; b.mode = MODE_SE_RES;
        b.bias = h->convs[2 + 2 * i].bias;
        b.se_w1 = h->se_w1[i]; b.se_w2 = h->se_w2[i]; b.residual = h->d_act[cur];
        if (i == h->blocks - 1) { b.v_w = h->v_w; b.v_pre = h->d_vpre; }
        cur = nxt;
    }
    TowerLayer& pl = ly[nl - 1];
    pl.in_map = 1 + cur; pl.w_map = 7 + nl - 1; pl.out_map = 4 + 1; pl.kc_in = h->hid / 64; pl.mode = MODE_RELU;
    pl.bias = h->convs[nl - 1].bias;
    free_dev(h->d_maps); free_dev(h->d_layers);
    h->d_maps = nullptr; h->d_layers = nullptr;
    CB_CUDA(h, cudaMalloc(&h->d_maps, maps.size() * sizeof(CUtensorMap)));
    CB_CUDA(h, cudaMalloc(&h->d_layers, ly.size() * sizeof(TowerLayer)));
    CB_CUDA(h, cudaMemcpy(h->d_maps, maps.data(), maps.size() * sizeof(CUtensorMap), cudaMemcpyHostToDevice));
    CB_CUDA(h, cudaMemcpy(h->d_layers, ly.data(), ly.size() * sizeof(TowerLayer), cudaMemcpyHostToDevice));
    h->tower_ready = true;
    if (h->hid == 128 || h->hid == 256) {
        // resident towers (tower_r_umma.cuh for 128 channels, tower_r2_umma.cuh on CTA pairs for 256): map table = weights
        // of the nl conv layers (128-row boxes), then the padded p_head; the layer table ends with the policy head as a
        // one-tap layer
        h->tower_r_ready = false;
        std::vector<CUtensorMap> mr(nl + 1);
        for (int l = 0; l < nl; ++l) mr[l] = h->convs[l].tm_w128;
        mr[nl] = h->tm_pwr;
        std::vector<TowerRLayer> lr(nl + 1);
        memset(lr.data(), 0, sizeof(TowerRLayer) * (nl + 1));
        for (int l = 0; l < nl; ++l) {
            lr[l].w_map = l;
            lr[l].kc_in = l == 0 ? 1 : h->hid / 64;
            lr[l].mode = (l >= 1 && l < nl - 1 && (l & 1) == 0) ? MODE_SE_RES : MODE_RELU;
            lr[l].bias = h->convs[l].bias;
            lr[l].ntaps = 9;
        }
        lr[nl].w_map = nl; lr[nl].kc_in = h->hid / 64; lr[nl].mode = MODE_POLICY; lr[nl].bias = h->p_bias_r;
        lr[nl].ntaps = 1; lr[nl].w_row_off = -4 * h->hid;
        lr[0].save_res = 1;
        for (int i = 0; i < h->blocks; ++i) {
            TowerRLayer& b = lr[2 + 2 * i];
            b.se_w1 = h->se_w1[i]; b.se_w2 = h->se_w2[i];
            b.save_res = i < h->blocks - 1;
            if (i == h->blocks - 1) { b.v_w = h->v_w; b.v_pre = h->d_vpre; }
        }
        std::vector<CUtensorMap> mrh(nl + 1);
        for (int l = 0; l < nl; ++l) mrh[l] = h->convs[l].tm_w_half;
        mrh[nl] = h->tm_pwr_half;
        free_dev(h->d_maps_r); free_dev(h->d_layers_r); free_dev(h->d_maps_r_half);
        h->d_maps_r = nullptr; h->d_layers_r = nullptr; h->d_maps_r_half = nullptr;
        CB_CUDA(h, cudaMalloc(&h->d_maps_r_half, mrh.size() * sizeof(CUtensorMap)));
        CB_CUDA(h, cudaMemcpy(h->d_maps_r_half, mrh.data(), mrh.size() * sizeof(CUtensorMap), cudaMemcpyHostToDevice));
        CB_CUDA(h, cudaMalloc(&h->d_maps_r, mr.size() * sizeof(CUtensorMap)));
        CB_CUDA(h, cudaMalloc(&h->d_layers_r, lr.size() * sizeof(TowerRLayer)));
        CB_CUDA(h, cudaMemcpy(h->d_maps_r, mr.data(), mr.size() * sizeof(CUtensorMap), cudaMemcpyHostToDevice));
        CB_CUDA(h, cudaMemcpy(h->d_layers_r, lr.data(), lr.size() * sizeof(TowerRLayer), cudaMemcpyHostToDevice));
        h->tower_r_ready = true;
    }
    return CB_OK;
}

// Which tower runs for this handle right now: 0 per-layer launches, 1 the pair-layout fused tower (256-channel nets;
// 128-channel ones with CB_TOWER_PAIR=1; 256-channel ones with it too), 3 the shared-memory-resident single-launch forward
// (128-channel nets), 4 the resident forward on CTA pairs (256-channel nets: tower + policy head, then value_head_kernel).
int fused_kind(const cb_handle* h) {
    if (!h->use_fused || h->layer_limit != 0) return 0;
    // CB_DTYPE_FP32: the resident forward with fp16-pair operands (128 channels), else the CUDA-core layers
    if (h->dtype == CB_DTYPE_FP32) return (h->hid == 128 && h->use_split && h->use_resident && h->tower_r_ready) ? 3 : 0;
    if (h->hid == 128 && h->use_resident && h->tower_r_ready) return 3;
    if (h->hid == 256 && h->use_resident && h->tower_r_ready) return 4;
    return h->tower_ready ? 1 : 0;
}

ValueParams value_params(const cb_handle* h, int B, int material_on, bool use_q, int pair_layout) {
    ValueParams vp = {};
    vp.v_pre = h->d_vpre;
    vp.q = use_q ? h->d_q : nullptr;
    vp.conv_bias = h->v_conv_bias;
    vp.bn_scale = h->v_bn_scale;
    vp.bn_bias = h->v_bn_bias;
    vp.fc_wt = h->fc_wt;
    vp.fc_b = h->fc_b;
    vp.out_w = h->out_w;
    vp.out_b = h->out_b;
    vp.k = h->k;
    vp.material_on = material_on;
    vp.pair_layout = pair_layout;   // layout of v_pre rows: the resident tower emits plain order
    vp.B = B;
    vp.v_logit = h->d_vlogit;
    vp.v_final = h->d_vfinal;
    vp.k_out = h->d_k;
    return vp;
}

// Units of the resident tower: boards are spread evenly over the SMs, at most 8 per unit; batches of up to
// one board per SM run one board (one chain) per CTA, which is the shortest latency (104 us for any B <= 148).
int tower_r_units(const cb_handle* h, int B) {
    const int rounds = (B + 8 * h->sm_count - 1) / (8 * h->sm_count);
    const int even = h->sm_count * rounds;
    return even < B ? even : B;
}

// One launch = the whole forward: planes, conv tower, policy head and value head.
// `src` (default: h itself): the handle whose input block (boards, q) is evaluated — a match evaluates one
// set of leaves with two models; `stream`: where to launch (default: h's own stream).
// `board0`: evaluate the boards [board0, board0 + B) of the input block, results at the same offsets (`po` is passed
// already offset); `unit_cap` > 0 bounds the grid (a two-pool self-play run leaves SMs to the other pool's tree kernel).
int launch_tower_r(cb_handle* h, int B, const PolicyOut& po, int material_on, bool use_q, const cb_handle* src = nullptr,
                   cudaStream_t stream = nullptr, int board0 = 0, int unit_cap = 0) {
    if (!src) src = h;
    if (!stream) stream = h->stream;
    TowerRParams tp;
    tp.boards = src->d_boards + board0;
    tp.maps = h->d_maps_r;
    tp.maps_half = h->d_maps_r_half;
    tp.layers = h->d_layers_r;
    tp.po = po;
    tp.vp = value_params(h, B, material_on, use_q, 0);
    tp.vp.q = use_q ? src->d_q + board0 : nullptr;
    tp.vp.v_logit += board0;
    tp.vp.v_final += board0;
    tp.vp.k_out += board0;
    tp.num_layers = (int)h->convs.size() + 1;
    tp.num_boards = B;
    const bool split = h->dtype == CB_DTYPE_FP32;   // fp16-pair operands: one chain of at most 4 boards per unit
    tp.num_units = tower_r_units(h, B);
    if (split) {
        const int even = h->sm_count * ((B + 4 * h->sm_count - 1) / (4 * h->sm_count));
        tp.num_units = even < B ? even : B;
    } else if (unit_cap > 0 && tp.num_units > unit_cap && (long long)unit_cap * 8 >= B) {
        tp.num_units = unit_cap;
    }
    tp.res_scratch = h->d_res_scratch;
    tp.dbg_backbone = h->d_dbg_backbone;
    tp.trace = h->d_trace;
    tp.cta_times = h->d_cta_times;
    const char* dbg_env = getenv("CB_DBG_R");
    tp.dbg = dbg_env ? atoi(dbg_env) : 0;
    const int grid = tp.num_units < h->sm_count ? tp.num_units : h->sm_count;
    // debug instantiation only when an experiment switch, the timeline, the per-CTA timers or the backbone dump is on
    const bool debug = tp.dbg != 0 || tp.trace || tp.cta_times || tp.dbg_backbone;
    // cluster mode (pairs of CTAs share the weight stream): one unit per CTA, an even grid, and every unit with
    // the same number of chains (all of 5..8 boards, or all of 1..4), so both CTAs of a pair walk the same stages
    const int n_lo = B / tp.num_units, n_hi = (B + tp.num_units - 1) / tp.num_units;
    const bool cl2 = !split && h->use_cluster && !debug && tp.num_units == grid && (grid & 1) == 0 && (n_lo >= 5 || n_hi <= 4);
    // every unit a single chain (at most 4 boards per CTA): the deep weight ring (latency-bound small batches)
    const bool deep = !split && !cl2 && n_hi <= 4 && !getenv("CB_NO_DEEP_RING");
    void (*kern)(const TowerRParams);
    if (split)
        kern = debug ? tower_r_umma_kernel<false, true, false, false, true> : tower_r_umma_kernel<false, false, false, false, true>;
    else if (h->dtype == CB_DTYPE_BF16)
        kern = cl2 ? tower_r_umma_kernel<true, false, true>
             : debug ? (deep ? tower_r_umma_kernel<true, true, false, true> : tower_r_umma_kernel<true, true, false>)
             : deep ? tower_r_umma_kernel<true, false, false, true> : tower_r_umma_kernel<true, false, false>;
    else
        kern = cl2 ? tower_r_umma_kernel<false, false, true>
             : debug ? (deep ? tower_r_umma_kernel<false, true, false, true> : tower_r_umma_kernel<false, true, false>)
             : deep ? tower_r_umma_kernel<false, false, false, true> : tower_r_umma_kernel<false, false, false>;
    CB_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, TowerRCfg::kSmemBytes));
    if (cl2) {
        cudaLaunchConfig_t lc = {};
        lc.gridDim = dim3(grid);
        lc.blockDim = dim3(TowerRCfg::kThreads);
        lc.dynamicSmemBytes = TowerRCfg::kSmemBytes;
        lc.stream = stream;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = 2;
        at[0].val.clusterDim.y = 1;
        at[0].val.clusterDim.z = 1;
        lc.attrs = at;
        lc.numAttrs = 1;
        CB_CUDA(h, cudaLaunchKernelEx(&lc, kern, tp));
    } else {
        kern<<<grid, TowerRCfg::kThreads, TowerRCfg::kSmemBytes, stream>>>(tp);
    }
    CB_CUDA(h, cudaGetLastError());
    return CB_OK;
}

// 256-channel nets: the resident forward on CTA pairs (tower, p_conv, policy head) followed by the value head's FC layers.
int launch_tower_r2(cb_handle* h, int B, const PolicyOut& po, int material_on, bool use_q) {
    TowerR2Params tp = {};
    tp.boards = h->d_boards;
    tp.maps = h->d_maps_r;
    tp.layers = h->d_layers_r;
    tp.po = po;
    tp.v_pre = h->d_vpre;
    tp.num_layers = (int)h->convs.size() + 1;
    tp.num_boards = B;
    tp.res_scratch = h->d_res_scratch;
    tp.dbg_backbone = h->d_dbg_backbone;
    const int max_pairs = h->sm_count / 2;
    {   // units of 8 boards (two chains that hide each other's epilogue) in rounds over the pairs; when the last round would
        // occupy at most half of the pairs, its units are cut into single chains of 4 boards on twice as many pairs
        const int units8 = (B + 7) / 8, rounds = units8 / max_pairs, last = units8 - rounds * max_pairs;
        if (last > 0 && last * 2 <= max_pairs) {
            tp.full_units = rounds * max_pairs;
            tp.num_units = tp.full_units + (B - tp.full_units * 8 + 3) / 4;
        } else {
            tp.full_units = units8;
            tp.num_units = units8;
        }
    }
    const int pairs = tp.num_units < max_pairs ? tp.num_units : max_pairs;
    CB_CUDA(h, cudaMemsetAsync(h->d_vpre, 0, (size_t)B * 64 * 4, h->stream));
    void (*kern)(const TowerR2Params) = h->dtype == CB_DTYPE_BF16 ? tower_r2_umma_kernel<true> : tower_r2_umma_kernel<false>;
    CB_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, TowerR2Cfg::kSmemBytes));
    cudaLaunchConfig_t lc = {};
    lc.gridDim = dim3(2 * pairs);
    lc.blockDim = dim3(TowerR2Cfg::kThreads);
    lc.dynamicSmemBytes = TowerR2Cfg::kSmemBytes;
    lc.stream = h->stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    lc.attrs = at;
    lc.numAttrs = 1;
    CB_CUDA(h, cudaLaunchKernelEx(&lc, kern, tp));
    const ValueParams vp = value_params(h, B, material_on, use_q, 0);
    value_head_kernel<<<(B + kValueBoardsPerCta - 1) / kValueBoardsPerCta, 256, 0, h->stream>>>(vp);
    CB_CUDA(h, cudaGetLastError());
    return CB_OK;
}

int launch_tower_r2_staged(cb_handle* h, int B) {
    PolicyOut po = {};
    if (h->staged_moves > 0) {
        po.move_idx = h->d_move_idx;
        po.move_off = h->d_move_off;
        po.priors = h->d_priors;
    }
    return launch_tower_r2(h, B, po, 1, true);
}

// The same launch with the outputs of the fast path (priors of the staged legal moves, material on).
int launch_tower_r_staged(cb_handle* h, int B) {
    PolicyOut po = {};
    if (h->staged_moves > 0) {
        po.move_idx = h->d_move_idx;
        po.move_off = h->d_move_off;
        po.priors = h->d_priors;
    }
    return launch_tower_r(h, B, po, 1, true);
}


template <int HID, bool BF16>
int launch_tower_t(cb_handle* h, int B) {
    using Cfg = ConvCfg<HID>;
    TowerParams tp;
    tp.maps = h->d_maps;
    tp.layers = h->d_layers;
    tp.num_layers = (int)h->convs.size();
    tp.num_tiles = (B + 1) / 2;
    const int units = (tp.num_tiles + 2 * Cfg::kTiles - 1) / (2 * Cfg::kTiles);
    const int grid = units < h->sm_count ? units : h->sm_count;
    auto kern = tower_umma_kernel<HID, BF16>;
    CB_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    kern<<<grid, Cfg::kThreads, Cfg::kSmemBytes, h->stream>>>(tp);
    CB_CUDA(h, cudaGetLastError());
    return CB_OK;
}

int launch_tower(cb_handle* h, int B) {
    const bool bf = h->dtype == CB_DTYPE_BF16;
    if (h->hid == 128) return bf ? launch_tower_t<128, true>(h, B) : launch_tower_t<128, false>(h, B);
    return bf ? launch_tower_t<256, true>(h, B) : launch_tower_t<256, false>(h, B);
}

int ensure_scratch(cb_handle* h, size_t floats) {
    if (floats <= h->scratch_floats) return CB_OK;
    free_dev(h->d_scratch);
    h->d_scratch = nullptr;
    h->scratch_floats = 0;
    CB_CUDA(h, cudaMalloc(&h->d_scratch, floats * 4));
    h->scratch_floats = floats;
    return CB_OK;
}

template <int HID, bool BF16>
int launch_conv_umma_t(cb_handle* h, const CUtensorMap& tin, const ConvLayer& L, const CUtensorMap& tout, int mode,
                       const ConvParams& p) {
    using Cfg = ConvCfg<HID>;
    const int rounds = (p.num_tiles + Cfg::kTiles - 1) / Cfg::kTiles;
    const int grid = rounds < h->sm_count ? rounds : h->sm_count;
    if (mode == MODE_RELU) {
        auto kern = conv3x3_umma_kernel<HID, BF16, MODE_RELU>;
        CB_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
        kern<<<grid, Cfg::kThreads, Cfg::kSmemBytes, h->stream>>>(tin, L.tm_w, tout, p);
    } else {
        auto kern = conv3x3_umma_kernel<HID, BF16, MODE_SE_RES>;
        CB_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
        kern<<<grid, Cfg::kThreads, Cfg::kSmemBytes, h->stream>>>(tin, L.tm_w, tout, p);
    }
    CB_CUDA(h, cudaGetLastError());
    return CB_OK;
}

int launch_conv_umma(cb_handle* h, const CUtensorMap& tin, const ConvLayer& L, const CUtensorMap& tout, int mode,
                     const ConvParams& p) {
    const bool bf = h->dtype == CB_DTYPE_BF16;
    if (h->hid == 128) return bf ? launch_conv_umma_t<128, true>(h, tin, L, tout, mode, p)
                                 : launch_conv_umma_t<128, false>(h, tin, L, tout, mode, p);
    return bf ? launch_conv_umma_t<256, true>(h, tin, L, tout, mode, p)
              : launch_conv_umma_t<256, false>(h, tin, L, tout, mode, p);
}

int launch_conv_fp32(cb_handle* h, const float* in, const ConvLayer& L, float* out, int B, int relu) {
    const size_t smem = (size_t)100 * L.cin * 4;
    if (h->hid == 128) {
        auto kern = conv3x3_fp32_kernel<128>;
        CB_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<B, 256, smem, h->stream>>>(in, L.cin, L.w32, L.bias, out, relu);
    } else {
        auto kern = conv3x3_fp32_kernel<256>;
        CB_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<B, 256, smem, h->stream>>>(in, L.cin, L.w32, L.bias, out, relu);
    }
    CB_CUDA(h, cudaGetLastError());
    return CB_OK;
}

template <typename T>
int launch_policy_t(cb_handle* h, const T* p, int B, const PolicyOut& po) {
    const size_t smem = ((size_t)64 * h->hid + (size_t)h->hid * 73 + CB_POLICY_SIZE) * 4;
    if (h->hid == 128) {
        auto kern = policy_head_kernel<T, 128>;
        CB_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<B, kPolicyThreads, smem, h->stream>>>(p, h->p_wt, h->p_bias, po, h->dtype != CB_DTYPE_FP32);
    } else {
        auto kern = policy_head_kernel<T, 256>;
        CB_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<B, kPolicyThreads, smem, h->stream>>>(p, h->p_wt, h->p_bias, po, h->dtype != CB_DTYPE_FP32);
    }
    CB_CUDA(h, cudaGetLastError());
    return CB_OK;
}

template <int HID, bool BF16>
int launch_policy_umma_t(cb_handle* h, int B, const PolicyOut& po) {
    using Cfg = PolicyCfg<HID>;
    const int tiles = (B + 1) / 2;
    const int grid = tiles < h->sm_count ? tiles : h->sm_count;
    auto kern = policy_umma_kernel<HID, BF16>;
    CB_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    kern<<<grid, Cfg::kThreads, Cfg::kSmemBytes, h->stream>>>(h->tm_act_st[1], h->tm_pw, h->p_bias, tiles, B, po);
    CB_CUDA(h, cudaGetLastError());
    return CB_OK;
}

int launch_policy(cb_handle* h, const void* p, int B, const PolicyOut& po) {
    if (h->dtype == CB_DTYPE_FP32) return launch_policy_t<float>(h, (const float*)p, B, po);
    const bool bf = h->dtype == CB_DTYPE_BF16;
    if (getenv("CB_SIMT_POLICY")) {
        if (bf) return launch_policy_t<__nv_bfloat16>(h, (const __nv_bfloat16*)p, B, po);
        return launch_policy_t<__half>(h, (const __half*)p, B, po);
    }
    if (h->hid == 128) return bf ? launch_policy_umma_t<128, true>(h, B, po) : launch_policy_umma_t<128, false>(h, B, po);
    return bf ? launch_policy_umma_t<256, true>(h, B, po) : launch_policy_umma_t<256, false>(h, B, po);
}

// Index of the buffer holding the backbone output after `blocks` residual blocks.
int backbone_index(int blocks) { return (blocks & 1) ? 2 : 0; }

// Enqueue the convolution tower only (start conv .. p_conv).  If ev is non-null it
// receives 2 events per conv launch (timing of individual layers).
int enqueue_tower(cb_handle* h, int B, std::vector<cudaEvent_t>* ev) {
    const int Bpad = (B + 1) & ~1;
    const int last = (int)h->convs.size() - 1;
    int cur = 0, e = 0, launched = 0;
    auto stop = [&]() { return h->layer_limit > 0 && launched++ >= h->layer_limit; };
    auto mark = [&]() -> int {
        if (ev) CB_CUDA(h, cudaEventRecord((*ev)[e++], h->stream));
        return CB_OK;
    };
    if (h->dtype == CB_DTYPE_FP32) {
        int rc;
        if (stop()) return CB_OK;
        if ((rc = mark())) return rc;
        if ((rc = launch_conv_fp32(h, (const float*)h->d_x0, h->convs[0], (float*)h->d_act[0], B, 1))) return rc;
        if ((rc = mark())) return rc;
        for (int i = 0; i < h->blocks; ++i) {
            const int nxt = cur == 0 ? 2 : 0;
            if (stop()) return CB_OK;
            if ((rc = mark())) return rc;
            if ((rc = launch_conv_fp32(h, (const float*)h->d_act[cur], h->convs[1 + 2 * i], (float*)h->d_act[1], B, 1))) return rc;
            if ((rc = mark())) return rc;
            if (stop()) return CB_OK;
            if ((rc = mark())) return rc;
            if ((rc = launch_conv_fp32(h, (const float*)h->d_act[1], h->convs[2 + 2 * i], h->d_tmp32, B, 0))) return rc;
            if ((rc = mark())) return rc;
            const bool lastblk = i == h->blocks - 1;
            if (h->hid == 128)
                se_res_fp32_kernel<128><<<B, 128, 0, h->stream>>>(h->d_tmp32, (const float*)h->d_act[cur], h->se_w1[i],
                                                                 h->se_w2[i], (float*)h->d_act[nxt],
                                                                 lastblk ? h->v_w : nullptr, lastblk ? h->d_vpre : nullptr);
            else
                se_res_fp32_kernel<256><<<B, 256, 0, h->stream>>>(h->d_tmp32, (const float*)h->d_act[cur], h->se_w1[i],
                                                                 h->se_w2[i], (float*)h->d_act[nxt],
                                                                 lastblk ? h->v_w : nullptr, lastblk ? h->d_vpre : nullptr);
            CB_CUDA(h, cudaGetLastError());
            cur = nxt;
        }
        if (stop()) return CB_OK;
        if ((rc = mark())) return rc;
        if ((rc = launch_conv_fp32(h, (const float*)h->d_act[cur], h->convs[last], (float*)h->d_act[1], B, 1))) return rc;
        if ((rc = mark())) return rc;
        return CB_OK;
    }
    ConvParams p = {};
    p.num_tiles = Bpad / 2;
    const char* dbg_env = getenv("CB_DBG");
    const int dbg = dbg_env ? atoi(dbg_env) : 0;
    p.dbg = dbg;
    int rc;
    // start conv: x0 (64-channel padded planes) -> act[0]
    p.kc_in = 1;
    p.bias = h->convs[0].bias;
    if (stop()) return CB_OK;
    if ((rc = mark())) return rc;
    if ((rc = launch_conv_umma(h, h->tm_x0, h->convs[0], h->tm_act_st[0], MODE_RELU, p))) return rc;
    if ((rc = mark())) return rc;
    for (int i = 0; i < h->blocks; ++i) {
        const int nxt = cur == 0 ? 2 : 0;
        ConvParams p1 = {};
        p1.dbg = dbg;
        p1.num_tiles = Bpad / 2;
        p1.kc_in = h->hid / 64;
        p1.bias = h->convs[1 + 2 * i].bias;
        if (stop()) return CB_OK;
        if ((rc = mark())) return rc;
        if ((rc = launch_conv_umma(h, h->tm_act_ld[cur], h->convs[1 + 2 * i], h->tm_act_st[1], MODE_RELU, p1))) return rc;
        if ((rc = mark())) return rc;
        ConvParams p2 = p1;
        p2.bias = h->convs[2 + 2 * i].bias;
        p2.se_w1 = h->se_w1[i];
        p2.se_w2 = h->se_w2[i];
        p2.residual = h->d_act[cur];
        if (i == h->blocks - 1) {
            p2.v_w = h->v_w;
            p2.v_pre = h->d_vpre;
        }
        if (stop()) return CB_OK;
        if ((rc = mark())) return rc;
        if ((rc = launch_conv_umma(h, h->tm_act_ld[1], h->convs[2 + 2 * i], h->tm_act_st[nxt], MODE_SE_RES, p2))) return rc;
        if ((rc = mark())) return rc;
        cur = nxt;
    }
    ConvParams pp = {};
    pp.dbg = dbg;
    pp.num_tiles = Bpad / 2;
    pp.kc_in = h->hid / 64;
    pp.bias = h->convs[last].bias;
    if (stop()) return CB_OK;
    if ((rc = mark())) return rc;
    if ((rc = launch_conv_umma(h, h->tm_act_ld[cur], h->convs[last], h->tm_act_st[1], MODE_RELU, pp))) return rc;
    if ((rc = mark())) return rc;
    return CB_OK;
}

// Enqueue one full forward on the handle's stream (inputs already in d_boards/d_q/CSR).
int enqueue_forward(cb_handle* h, int B, bool full_policy, bool csr, int material_on, bool use_q,
                    const uint16_t* move_cnt = nullptr) {
    const int kind = fused_kind(h);
    const int glog = 1;                                       // boards interleaved per activation group (pairs)
    const int Bpad = (B + (1 << glog) - 1) & ~((1 << glog) - 1);
    if (kind == 3 || kind == 4) {
        // the resident towers encode the planes themselves
    } else if (h->dtype == CB_DTYPE_FP32) {
        encode_nhwc_f32_kernel<<<(B * 64 + 255) / 256, 256, 0, h->stream>>>(h->d_boards, B, (float*)h->d_x0);
    } else if (h->dtype == CB_DTYPE_BF16) {
        encode_tiled16_kernel<true><<<(Bpad * 512 + 255) / 256, 256, 0, h->stream>>>(h->d_boards, B, Bpad, glog, (uint4*)h->d_x0);
    } else {
        encode_tiled16_kernel<false><<<(Bpad * 512 + 255) / 256, 256, 0, h->stream>>>(h->d_boards, B, Bpad, glog, (uint4*)h->d_x0);
    }
    CB_CUDA(h, cudaGetLastError());
    PolicyOut po = {};
    po.probs = full_policy ? h->d_probs : nullptr;
    if (csr) {
        po.move_idx = h->d_move_idx;
        po.move_off = h->d_move_off;
        po.priors = h->d_priors;
        po.move_cnt = move_cnt;    // device-side tree kernels: fixed-stride lists
    }
    int rc;
    if (kind == 3) return launch_tower_r(h, B, po, material_on, use_q);   // heads included
    if (kind == 4) return launch_tower_r2(h, B, po, material_on, use_q);  // policy head included, value FCs follow
    if (kind == 1)
        rc = launch_tower(h, B);
    else
        rc = enqueue_tower(h, B, nullptr);
    if (rc) return rc;
    // the two heads are independent: fork the value head onto the side stream
    CB_CUDA(h, cudaEventRecord(h->ev_fork, h->stream));
    CB_CUDA(h, cudaStreamWaitEvent(h->stream2, h->ev_fork, 0));
    const ValueParams vp = value_params(h, B, material_on, use_q, h->dtype == CB_DTYPE_FP32 ? 0 : (kind == 3 ? 0 : 1));
    value_head_kernel<<<(B + kValueBoardsPerCta - 1) / kValueBoardsPerCta, 256, 0, h->stream2>>>(vp);
    CB_CUDA(h, cudaGetLastError());
    CB_CUDA(h, cudaEventRecord(h->ev_join, h->stream2));
    if ((rc = launch_policy(h, h->d_act[1], B, po))) return rc;
    CB_CUDA(h, cudaStreamWaitEvent(h->stream, h->ev_join, 0));
    return CB_OK;
}

// Replay the forward graph for this shape.  A shape is launched eagerly the first two times it is seen (which also sets
// the kernels' attributes outside capture) and captured on the third; at most kMaxForwardGraphs shapes are kept — the
// host-driven pools submit a different batch size almost every step, and those ragged sizes simply stay eager.
constexpr size_t kMaxForwardGraphs = 24;
int run_forward(cb_handle* h, int B, bool full_policy, bool csr, int material_on, bool use_q) {
    if (!h->use_graphs) return enqueue_forward(h, B, full_policy, csr, material_on, use_q);
    const uint64_t key = (uint64_t)B | ((uint64_t)full_policy << 32) | ((uint64_t)csr << 33) |
                         ((uint64_t)(material_on != 0) << 34) | ((uint64_t)use_q << 35);
    auto it = h->graphs.find(key);
    if (it == h->graphs.end()) {
        if (h->graph_seen.size() > 4096) h->graph_seen.clear();
        if (++h->graph_seen[key] < 3 || h->graphs.size() >= kMaxForwardGraphs)
            return enqueue_forward(h, B, full_policy, csr, material_on, use_q);
        cudaGraph_t g = nullptr;
        CB_CUDA(h, cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
        const int rc = enqueue_forward(h, B, full_policy, csr, material_on, use_q);
        cudaError_t ce = cudaStreamEndCapture(h->stream, &g);
        if (rc) {
            if (g) cudaGraphDestroy(g);
            return rc;
        }
        CB_CUDA(h, ce);
        cudaGraphExec_t ge = nullptr;
        cudaError_t ie = cudaGraphInstantiate(&ge, g, 0);
        cudaGraphDestroy(g);
        CB_CUDA(h, ie);
        it = h->graphs.emplace(key, ge).first;
    }
    CB_CUDA(h, cudaGraphLaunch(it->second, h->stream));
    return CB_OK;
}

int upload_leaves(cb_handle* h, const cb_board* boards, const float* q, int B, const uint16_t* move_idx,
                  const uint32_t* move_off) {
    int rc = ensure_capacity(h, B);
    if (rc) return rc;
    h->staged_moves = 0;
    const bool csr = move_idx && move_off;
    const size_t n = csr ? move_off[B] : 0;
    if (n > h->csr_cap) return fail(h, CB_EINVAL, "legal-move CSR larger than 256 moves per position");
    memcpy(h->h_boards, boards, (size_t)B * sizeof(cb_board));
    if (q) memcpy(h->h_q, q, (size_t)B * 4);
    if (csr) {
        memcpy(h->h_move_off, move_off, ((size_t)B + 1) * 4);
        memcpy(h->h_move_idx, move_idx, n * 2);
        h->staged_moves = n;
    }
    if (q && csr && (size_t)B * 4 >= (size_t)h->cap * 3) {
        // nearly full batch: one transfer for the whole input block (gaps included)
        CB_CUDA(h, cudaMemcpyAsync(h->d_in, h->h_in, h->in_idx_off + n * 2, cudaMemcpyHostToDevice, h->stream));
    } else {
        CB_CUDA(h, cudaMemcpyAsync(h->d_boards, h->h_boards, (size_t)B * sizeof(cb_board), cudaMemcpyHostToDevice, h->stream));
        if (q) CB_CUDA(h, cudaMemcpyAsync(h->d_q, h->h_q, (size_t)B * 4, cudaMemcpyHostToDevice, h->stream));
        if (csr) {
            CB_CUDA(h, cudaMemcpyAsync(h->d_move_off, h->h_move_off, ((size_t)B + 1) * 4, cudaMemcpyHostToDevice, h->stream));
            if (n) CB_CUDA(h, cudaMemcpyAsync(h->d_move_idx, h->h_move_idx, n * 2, cudaMemcpyHostToDevice, h->stream));
        }
    }
    return CB_OK;
}

}  // namespace

// One move's training sample as logged by the device-side tree kernels.
struct DevSample {
    cb_board board;
    float material;
    bool completed = true;   // qsearch_completed of the material search
    std::vector<std::pair<uint16_t, float>> policy;
};

// host/selfplay.cpp: one 5763-f32 record (src/training_data.rs:20-60)
void cbsp_write_sample(FILE* f, const cb_board& b, float material, const std::pair<uint16_t, float>* policy, size_t n_policy,
                       int result_white, std::vector<float>& rec, bool qsearch_completed = true);

namespace {

// Copy the records appended since the last drain, rebuild the games and write the finished ones.
// Records [ring_read, cursor) of the device-side sample log -> finished games in the sample file.  `cursor` was read
// after the kernels that wrote those records completed; the copies run on `cs` (nullptr: the handle's stream), so the
// caller may already have the next graph replay running on the handle's stream (it appends beyond `cursor`).
int drain_samples(cb_handle* h, const MctsParams& mp, unsigned long long& ring_read, std::vector<uint8_t>& host,
                  std::vector<std::vector<DevSample>>& pending, FILE* f, unsigned long long cursor, cudaStream_t cs) {
    if (!cs) cs = h->stream;
    if (cursor == ring_read) return CB_OK;
    if (cursor - ring_read > mp.ring_slots) return fail(h, CB_ESTATE, "sample ring overflow");
    const size_t n = (size_t)(cursor - ring_read);
    host.resize(n * kSampleSlotBytes);
    const size_t first = (size_t)(ring_read % mp.ring_slots);
    const size_t run = n < mp.ring_slots - first ? n : mp.ring_slots - first;
    CB_CUDA(h, cudaMemcpyAsync(host.data(), mp.sample_ring + first * kSampleSlotBytes, run * kSampleSlotBytes,
                               cudaMemcpyDeviceToHost, cs));
    if (run < n)
        CB_CUDA(h, cudaMemcpyAsync(host.data() + run * kSampleSlotBytes, mp.sample_ring, (n - run) * kSampleSlotBytes,
                                   cudaMemcpyDeviceToHost, cs));
    CB_CUDA(h, cudaStreamSynchronize(cs));
    std::vector<float> rec;
    for (size_t i = 0; i < n; ++i) {
        const uint8_t* r = host.data() + i * kSampleSlotBytes;
        const uint32_t game = reinterpret_cast<const uint32_t*>(r)[0];
        const uint16_t cnt = reinterpret_cast<const uint16_t*>(r)[5];
        if (game >= pending.size()) return fail(h, CB_ESTATE, "corrupt sample record");
        if (cnt == 0xFFFFu) {
            const int result = reinterpret_cast<const int32_t*>(r)[3];
            for (const DevSample& s : pending[game])
                cbsp_write_sample(f, s.board, s.material, s.policy.data(), s.policy.size(), result, rec, s.completed);
            pending[game].clear();
            continue;
        }
        DevSample s;
        memcpy(&s.board, r + 16, sizeof(cb_board));
        s.material = reinterpret_cast<const float*>(r)[3];
        s.completed = (reinterpret_cast<const uint16_t*>(r)[4] & 0x8000u) == 0;
        const uint32_t* pairs = reinterpret_cast<const uint32_t*>(r + 144);
        uint32_t total = 0;
        for (uint16_t k = 0; k < cnt; ++k) total += pairs[k] >> 16;
        for (uint16_t k = 0; k < cnt && total > 0; ++k)
            s.policy.emplace_back((uint16_t)(pairs[k] & 0xFFFFu), (float)(pairs[k] >> 16) / (float)total);
        pending[game].push_back(std::move(s));
    }
    ring_read = cursor;
    return CB_OK;
}

}  // namespace

// ============================================================== C ABI
extern "C" {

const char* cb_version(void) { return "caissa_b200 0.1 (sm_100a)"; }

size_t cb_weight_count(int num_blocks, int hidden) {
    const size_t H = hidden, S = hidden / 16;
    size_t n = H * 17 * 9 + 4 * H;
    n += (size_t)num_blocks * (2 * (H * H * 9 + 4 * H) + 2 * S * H);
    n += H * H * 9 + 4 * H + 73 * H + 73;
    n += H + 1 + 4 + 256 * 65 + 256 + 256 + 1 + 1;
    return n;
}

int cb_create(cb_handle** out, int device, int num_blocks, int hidden, int dtype) {
    if (!out) return CB_EINVAL;
    *out = nullptr;
    if ((hidden != 128 && hidden != 256) || num_blocks < 1 || num_blocks > 64 || dtype < 0 || dtype > 2) return CB_EINVAL;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0 || device < 0 || device >= ndev) return CB_ENODEV;
    cb_handle* h = new (std::nothrow) cb_handle();
    if (!h) return CB_ENOMEM;
    h->device = device;
    h->blocks = num_blocks;
    h->hid = hidden;
    h->dtype = dtype;
    const char* ng = getenv("CB_NO_GRAPH");
    h->use_graphs = !(ng && ng[0] == '1');
    const char* nf = getenv("CB_NO_FUSE");
    h->use_fused = !(nf && nf[0] == '1');
    const char* tp = getenv("CB_TOWER_PAIR");   // 128-channel nets on the global-round-trip tower of the 256-channel ones (comparison)
    h->use_resident = !(tp && tp[0] == '1');
    // 2-CTA clusters sharing the weight stream (TMA multicast): measured, not the default — it saves power
    // (+5 % SM clock under the 1000 W cap) but the lock-step of the two CTAs on a 4-stage ring costs 10 % more
    // cycles (sustained 195.9 us vs 188.1 us per forward at batch 1024).  CB_CLUSTER=1 turns it on.
    const char* nc = getenv("CB_CLUSTER");
    h->use_cluster = nc && nc[0] == '1';
    const char* fe = getenv("CB_FP32_EXACT");   // CB_DTYPE_FP32 on CUDA cores (fp32 FMA: the exact mode) instead of fp16 pairs
    h->use_split = !(fe && fe[0] == '1');
    cudaDeviceProp prop;
    if (cudaSetDevice(device) != cudaSuccess || cudaGetDeviceProperties(&prop, device) != cudaSuccess) {
        delete h;
        return CB_ENODEV;
    }
    if (prop.major != 10) {  // sm_100a cubins only
        delete h;
        return CB_ENODEV;
    }
    h->sm_count = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&h->stream2, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreate(&h->ev0) != cudaSuccess || cudaEventCreate(&h->ev1) != cudaSuccess ||
        cudaEventCreate(&h->ev2) != cudaSuccess || cudaEventCreate(&h->ev3) != cudaSuccess) {
        cb_destroy(h);
        return CB_ECUDA;
    }
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn) {
        cb_destroy(h);
        return CB_ECUDA;
    }
    h->encode_tiled = (EncodeTiledFn)fn;
    *out = h;
    return CB_OK;
}

void cb_destroy(cb_handle* h) {
    if (!h) return;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    free_buffers(h);
    for (auto& L : h->convs) { free_dev(L.w16); free_dev(L.w16s); free_dev(L.w32); free_dev(L.bias); }
    for (auto p : h->se_w1) free_dev(p);
    for (auto p : h->se_w2) free_dev(p);
    free_dev(h->p_wt); free_dev(h->p_bias); free_dev(h->v_w); free_dev(h->fc_wt); free_dev(h->fc_b); free_dev(h->out_w);
    free_dev(h->p_w16); free_dev(h->p_w16r); free_dev(h->p_bias_r); free_dev(h->p_w16s);
    free_dev(h->d_maps); free_dev(h->d_layers);
    free_dev(h->d_maps_r); free_dev(h->d_layers_r); free_dev(h->d_res_scratch); free_dev(h->d_maps_r_half);
    free_dev(h->d_flush);
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    if (h->ev2) cudaEventDestroy(h->ev2);
    if (h->ev3) cudaEventDestroy(h->ev3);
    if (h->ev_fork) cudaEventDestroy(h->ev_fork);
    if (h->ev_join) cudaEventDestroy(h->ev_join);
    if (h->stream2) cudaStreamDestroy(h->stream2);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
}

const char* cb_last_error(const cb_handle* h) { return h ? h->err.c_str() : "null handle"; }
int cb_is_available(const cb_handle* h) { return h && h->loaded ? 1 : 0; }

int cb_launches_per_forward(const cb_handle* h) {
    if (!h) return 0;
    const int convs = 2 * h->blocks + 2;
    if (fused_kind(h) == 3) return 1;                          // resident tower: planes, convs and both heads in one launch
    if (fused_kind(h) == 4) return 2;                          // resident pair tower with the policy head, then the value FCs
    if (h->dtype != CB_DTYPE_FP32 && h->use_fused) return 4;  // encode, fused tower, policy, value
    return 1 + convs + (h->dtype == CB_DTYPE_FP32 ? h->blocks : 0) + 2;
}

int cb_get_timing(const cb_handle* h, cb_timing* out) {
    if (!h || !out) return CB_EINVAL;
    *out = h->timing;
    return CB_OK;
}
void cb_reset_timing(cb_handle* h) {
    if (h) h->timing = cb_timing{};
}

int cb_load_weights(cb_handle* h, const float* blob, size_t n_floats) {
    if (!h || !blob) return CB_EINVAL;
    if (n_floats != cb_weight_count(h->blocks, h->hid))
        return fail(h, CB_EINVAL, "weight blob has " + std::to_string(n_floats) + " floats, expected " +
                                      std::to_string(cb_weight_count(h->blocks, h->hid)));
    CB_CUDA(h, cudaSetDevice(h->device));
    CB_CUDA(h, cudaStreamSynchronize(h->stream));
    drop_graphs(h);
    for (auto& L : h->convs) { free_dev(L.w16); free_dev(L.w16s); free_dev(L.w32); free_dev(L.bias); }
    for (auto p : h->se_w1) free_dev(p);
    for (auto p : h->se_w2) free_dev(p);
    h->convs.clear(); h->se_w1.clear(); h->se_w2.clear();
    free_dev(h->p_wt); free_dev(h->p_bias); free_dev(h->v_w); free_dev(h->fc_wt); free_dev(h->fc_b); free_dev(h->out_w);
    free_dev(h->p_w16); free_dev(h->p_w16r); free_dev(h->p_bias_r); free_dev(h->p_w16s);
    h->p_w16 = nullptr; h->p_w16r = nullptr; h->p_bias_r = nullptr; h->p_w16s = nullptr;
    h->p_wt = h->p_bias = h->v_w = h->fc_wt = h->fc_b = h->out_w = nullptr;
    h->loaded = false;

    const int H = h->hid, S = H / 16;
    const float* cur = blob;
    auto take = [&](size_t n) { const float* p = cur; cur += n; return p; };
    auto upload = [&](const void* src, size_t bytes, void** dst) -> int {
        CB_CUDA(h, cudaMalloc(dst, bytes));
        CB_CUDA(h, cudaMemcpy(*dst, src, bytes, cudaMemcpyHostToDevice));
        return CB_OK;
    };
    // conv + BN -> folded layer.  w: [H][cin][3][3]; bn: gamma, beta, mean, var.
    auto add_conv = [&](int cin) -> int {
        const float* w = take((size_t)H * cin * 9);
        const float* gamma = take(H); const float* beta = take(H); const float* mean = take(H); const float* var = take(H);
        ConvLayer L;
        L.cin = cin;
        L.cin_pad = (cin + 63) / 64 * 64;
        std::vector<float> scale(H), bias(H);
        for (int co = 0; co < H; ++co) {
            scale[co] = gamma[co] / sqrtf(var[co] + kBnEps);
            bias[co] = beta[co] - mean[co] * scale[co];
        }
        int rc = upload(bias.data(), (size_t)H * 4, (void**)&L.bias);
        if (rc) return rc;
        if (h->dtype == CB_DTYPE_FP32) {
            std::vector<float> w32((size_t)9 * cin * H);
            for (int co = 0; co < H; ++co)
                for (int ci = 0; ci < cin; ++ci)
                    for (int t = 0; t < 9; ++t)
                        w32[((size_t)t * cin + ci) * H + co] = w[((size_t)co * cin + ci) * 9 + t] * scale[co];
            if ((rc = upload(w32.data(), w32.size() * 4, (void**)&L.w32))) return rc;
            if (H == 128) {
                // fp16 pairs: w = hi + lo, hi = fp16(w), lo = fp16(w - hi); [2][9][H][cin_pad]
                std::vector<uint16_t> ws((size_t)2 * 9 * H * L.cin_pad, 0);
                const size_t lo_off = (size_t)9 * H * L.cin_pad;
                for (int co = 0; co < H; ++co)
                    for (int ci = 0; ci < cin; ++ci)
                        for (int t = 0; t < 9; ++t) {
                            const float v = w[((size_t)co * cin + ci) * 9 + t] * scale[co];
                            const __half hi = __float2half_rn(v);
                            const __half lo = __float2half_rn(v - __half2float(hi));
                            const size_t at = ((size_t)t * H + co) * L.cin_pad + ci;
                            memcpy(&ws[at], &hi, 2);
                            memcpy(&ws[lo_off + at], &lo, 2);
                        }
                if ((rc = upload(ws.data(), ws.size() * 2, &L.w16s))) return rc;
            }
        } else {
            std::vector<uint16_t> w16((size_t)9 * H * L.cin_pad, 0);
            for (int co = 0; co < H; ++co)
                for (int ci = 0; ci < cin; ++ci)
                    for (int t = 0; t < 9; ++t)
                        w16[((size_t)t * H + co) * L.cin_pad + ci] = to16(w[((size_t)co * cin + ci) * 9 + t] * scale[co], h->dtype);
            if ((rc = upload(w16.data(), w16.size() * 2, &L.w16))) return rc;
            if ((rc = make_weight_map(h, L))) return rc;
        }
        h->convs.push_back(L);
        return CB_OK;
    };
    int rc;
    if ((rc = add_conv(17))) return rc;
    for (int i = 0; i < h->blocks; ++i) {
        if ((rc = add_conv(H))) return rc;
        if ((rc = add_conv(H))) return rc;
        float* d1 = nullptr; float* d2 = nullptr;
        if ((rc = upload(take((size_t)S * H), (size_t)S * H * 4, (void**)&d1))) return rc;
        h->se_w1.push_back(d1);
        if ((rc = upload(take((size_t)H * S), (size_t)H * S * 4, (void**)&d2))) return rc;
        h->se_w2.push_back(d2);
    }
    if ((rc = add_conv(H))) return rc;
    {   // p_head [73][H] -> transposed [H][73]
        const float* pw = take((size_t)73 * H);
        std::vector<float> wt((size_t)H * 73);
        for (int c = 0; c < 73; ++c)
            for (int k = 0; k < H; ++k) wt[(size_t)k * 73 + c] = pw[(size_t)c * H + k];
        if ((rc = upload(wt.data(), wt.size() * 4, (void**)&h->p_wt))) return rc;
        const float* pb = take(73);
        if ((rc = upload(pb, 73 * 4, (void**)&h->p_bias))) return rc;
        if (h->dtype == CB_DTYPE_FP32 && H == 128) {
            std::vector<uint16_t> ws((size_t)2 * H * H, 0);
            for (int c = 0; c < 73; ++c)
                for (int k = 0; k < H; ++k) {
                    const float v = pw[(size_t)c * H + k];
                    const __half hi = __float2half_rn(v);
                    const __half lo = __float2half_rn(v - __half2float(hi));
                    memcpy(&ws[(size_t)c * H + k], &hi, 2);
                    memcpy(&ws[(size_t)H * H + (size_t)c * H + k], &lo, 2);
                }
            if ((rc = upload(ws.data(), ws.size() * 2, &h->p_w16s))) return rc;
            std::vector<float> br(H, 0.0f);
            memcpy(br.data(), pb, 73 * 4);
            if ((rc = upload(br.data(), (size_t)H * 4, (void**)&h->p_bias_r))) return rc;
        }
        if (h->dtype != CB_DTYPE_FP32) {
            std::vector<uint16_t> w16((size_t)80 * H, 0);
            for (int c = 0; c < 73; ++c)
                for (int k = 0; k < H; ++k) w16[(size_t)c * H + k] = to16(pw[(size_t)c * H + k], h->dtype);
            if ((rc = upload(w16.data(), w16.size() * 2, &h->p_w16))) return rc;
            const cuuint64_t dims[2] = {(cuuint64_t)H, 80};
            const cuuint64_t strides[1] = {(cuuint64_t)H * 2};
            const cuuint32_t box[2] = {64, 80};
            const cuuint32_t estr[2] = {1, 1};
            const CUtensorMapDataType dt =
                h->dtype == CB_DTYPE_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
            if (h->encode_tiled(&h->tm_pw, dt, 2, h->p_w16, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
                return fail(h, CB_ECUDA, "cuTensorMapEncodeTiled(policy weights) failed");
            if (H == 128 || H == 256) {
                // resident towers: p_head as a one-tap "layer": weights [H][H] (73 real rows), bias padded to H
                std::vector<uint16_t> wr((size_t)H * H, 0);
                memcpy(wr.data(), w16.data(), (size_t)73 * H * 2);
                if ((rc = upload(wr.data(), wr.size() * 2, &h->p_w16r))) return rc;
                std::vector<float> br(H, 0.0f);
                memcpy(br.data(), pb, 73 * 4);
                if ((rc = upload(br.data(), (size_t)H * 4, (void**)&h->p_bias_r))) return rc;
                const cuuint64_t dims_r[2] = {(cuuint64_t)H, (cuuint64_t)H};
                const cuuint32_t box_r[2] = {64, 128};
                if (h->encode_tiled(&h->tm_pwr, dt, 2, h->p_w16r, dims_r, strides, box_r, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
                    return fail(h, CB_ECUDA, "cuTensorMapEncodeTiled(padded policy weights) failed");
                const cuuint32_t box_rh[2] = {64, 64};
                if (h->encode_tiled(&h->tm_pwr_half, dt, 2, h->p_w16r, dims_r, strides, box_rh, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
                    return fail(h, CB_ECUDA, "cuTensorMapEncodeTiled(padded policy weights, half box) failed");
            }
        }
    }
    if ((rc = upload(take(H), (size_t)H * 4, (void**)&h->v_w))) return rc;
    h->v_conv_bias = *take(1);
    {
        const float g = *take(1), b = *take(1), m = *take(1), v = *take(1);
        h->v_bn_scale = g / sqrtf(v + kBnEps);
        h->v_bn_bias = b - m * h->v_bn_scale;
    }
    {   // v_fc [256][65] -> transposed [65][256]
        const float* fw = take((size_t)256 * 65);
        std::vector<float> wt((size_t)65 * 256);
        for (int j = 0; j < 256; ++j)
            for (int e = 0; e < 65; ++e) wt[(size_t)e * 256 + j] = fw[(size_t)j * 65 + e];
        if ((rc = upload(wt.data(), wt.size() * 4, (void**)&h->fc_wt))) return rc;
        if ((rc = upload(take(256), 256 * 4, (void**)&h->fc_b))) return rc;
        if ((rc = upload(take(256), 256 * 4, (void**)&h->out_w))) return rc;
    }
    h->out_b = *take(1);
    {   // k = 0.47 * softplus(k_logit)  (python/model.py:109), evaluated like torch: log1p(exp(x)), x>20 -> x
        const float kl = *take(1);
        const float sp = kl > 20.0f ? kl : log1pf(expf(kl));
        h->k = 0.47f * sp;
    }
    if ((size_t)(cur - blob) != n_floats) return fail(h, CB_EINVAL, "internal: blob walk mismatch");
    h->loaded = true;
    return build_tower_tables(h);
}

int cb_encode(cb_handle* h, const cb_board* boards, int B, float* planes) {
    if (!h || !boards || !planes || B < 0) return CB_EINVAL;
    if (B == 0) return CB_OK;
    CB_CUDA(h, cudaSetDevice(h->device));
    int rc = upload_leaves(h, boards, nullptr, B, nullptr, nullptr);
    if (rc) return rc;
    if ((rc = ensure_scratch(h, (size_t)B * CB_PLANES))) return rc;
    encode_nchw_f32_kernel<<<(B * 64 + 255) / 256, 256, 0, h->stream>>>(h->d_boards, B, h->d_scratch);
    CB_CUDA(h, cudaGetLastError());
    CB_CUDA(h, cudaMemcpyAsync(planes, h->d_scratch, (size_t)B * CB_PLANES * 4, cudaMemcpyDeviceToHost, h->stream));
    CB_CUDA(h, cudaStreamSynchronize(h->stream));
    return CB_OK;
}

int cb_pst_eval_cp(cb_handle* h, const cb_board* boards, int B, int32_t* cp) {
    if (!h || !boards || !cp || B < 0) return CB_EINVAL;
    if (B == 0) return CB_OK;
    CB_CUDA(h, cudaSetDevice(h->device));
    int rc = upload_leaves(h, boards, nullptr, B, nullptr, nullptr);
    if (rc) return rc;
    if ((rc = ensure_scratch(h, (size_t)B))) return rc;
    pst_eval_kernel<<<(B + 127) / 128, 128, 0, h->stream>>>(h->d_boards, B, (int32_t*)h->d_scratch);
    CB_CUDA(h, cudaGetLastError());
    CB_CUDA(h, cudaMemcpyAsync(cp, h->d_scratch, (size_t)B * 4, cudaMemcpyDeviceToHost, h->stream));
    CB_CUDA(h, cudaStreamSynchronize(h->stream));
    return CB_OK;
}

// One warp per board: the tree kernels' legal-move generator (debug / test hook).
__global__ void __launch_bounds__(256) device_legal_moves_kernel(const cb_board* __restrict__ boards, int n, uint16_t* __restrict__ moves,
                                                                 int32_t* __restrict__ counts) {
    __shared__ cb_board s_b[8];
    __shared__ cbchess::Move s_pseudo[8][CB_MAX_MOVES], s_legal[8][CB_MAX_MOVES];
    const int wl = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int i = blockIdx.x * 8 + wl;
    if (i >= n) return;
    if (lane < 16) reinterpret_cast<uint64_t*>(&s_b[wl])[lane] = reinterpret_cast<const uint64_t*>(&boards[i])[lane];
    __syncwarp();
    const int cnt = mcts_gen_legal_warp(s_b[wl], s_pseudo[wl], s_legal[wl], lane);
    for (int j = lane; j < cnt; j += 32) moves[(size_t)i * CB_MAX_MOVES + j] = s_legal[wl][j];
    if (lane == 0) counts[i] = cnt;
}

int cb_debug_device_legal_moves(cb_handle* h, const cb_board* boards, int n, uint16_t* moves, int32_t* counts) {
    if (!h || n < 0 || (n > 0 && (!boards || !moves || !counts))) return CB_EINVAL;
    if (n == 0) return CB_OK;
    CB_CUDA(h, cudaSetDevice(h->device));
    if (!h->chess_tables_uploaded) {
        CB_CUDA(h, cudaMemcpyToSymbol(cbchess::d_chess_tables, &cbchess::host_tables(), sizeof(cbchess::Tables)));
        h->chess_tables_uploaded = true;
    }
    int rc = upload_leaves(h, boards, nullptr, n, nullptr, nullptr);
    if (rc) return rc;
    uint16_t* d_moves = nullptr;
    int32_t* d_counts = nullptr;
    CB_CUDA(h, cudaMalloc(&d_moves, (size_t)n * CB_MAX_MOVES * 2));
    cudaError_t ce = cudaMalloc(&d_counts, (size_t)n * 4);
    if (ce == cudaSuccess) {
        device_legal_moves_kernel<<<(n + 7) / 8, 256, 0, h->stream>>>(h->d_boards, n, d_moves, d_counts);
        ce = cudaGetLastError();
    }
    if (ce == cudaSuccess) ce = cudaMemcpyAsync(moves, d_moves, (size_t)n * CB_MAX_MOVES * 2, cudaMemcpyDeviceToHost, h->stream);
    if (ce == cudaSuccess) ce = cudaMemcpyAsync(counts, d_counts, (size_t)n * 4, cudaMemcpyDeviceToHost, h->stream);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(h->stream);
    cudaFree(d_moves);
    cudaFree(d_counts);
    CB_CUDA(h, ce);
    return CB_OK;
}

// One warp per board: the tree kernels' light gates (debug / test hook).
__global__ void __launch_bounds__(256) device_light_gates_kernel(const cb_board* __restrict__ boards, int n, int koth, int8_t* __restrict__ out) {
    __shared__ cb_board s_b[8], s_nb[8];
    __shared__ cbchess::Move s_pseudo[8][CB_MAX_MOVES], s_legal[8][CB_MAX_MOVES];
    const int wl = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int i = blockIdx.x * 8 + wl;
    if (i >= n) return;
    if (lane < 16) reinterpret_cast<uint64_t*>(&s_b[wl])[lane] = reinterpret_cast<const uint64_t*>(&boards[i])[lane];
    __syncwarp();
    const int cnt = mcts_gen_legal_warp(s_b[wl], s_pseudo[wl], s_legal[wl], lane);
    const int tv = light_gates_warp(s_b[wl], s_legal[wl], cnt, koth != 0, s_nb[wl], s_pseudo[wl], lane);
    if (lane == 0) out[i] = (int8_t)tv;
}

// One warp per board: the tree kernels' gate at full depth, run to completion (debug / test hook).
__global__ void __launch_bounds__(128) device_tier1_gates_kernel(const cb_board* __restrict__ boards, int n, int koth, GateCfg gc, int budget,
                                                                 GateJob* jobs, int8_t* __restrict__ out, uint8_t* __restrict__ dist) {
    __shared__ GateStack s_st[4];
    __shared__ Move s_pseudo[4][CB_MAX_MOVES];
    const int wl = threadIdx.x >> 5, lane = threadIdx.x & 31, i = blockIdx.x * 4 + wl;
    if (i >= n) return;
    GateJob& job = jobs[i];
    if (lane < 16) reinterpret_cast<uint64_t*>(&s_st[wl].boards[0])[lane] = reinterpret_cast<const uint64_t*>(&boards[i])[lane];
    __syncwarp();
    int v = 2, d = 0;
    const int us = s_st[wl].boards[0].w_to_move ? 0 : 1;
    if (koth && (s_st[wl].boards[0].pieces[(1 - us) * 6 + cbchess::KING] & cbchess::kKothCenter)) v = -1;
    else {
        // resumed in slices of `budget` positions, like the tree kernels do across steps (the shared stack is wiped in
        // between, as another game's job would do)
        bool fresh = true;
        while (!tier1_gates_run(job, s_st[wl], gc, GateBudget{budget, 0, nullptr, 0u}, fresh, s_pseudo[wl], lane, &v, &d)) {
            fresh = false;
            for (int k = lane; k < (int)(sizeof(GateStack) / 4); k += 32) reinterpret_cast<uint32_t*>(&s_st[wl])[k] = 0xDEADBEEFu;
            __syncwarp();
        }
    }
    if (lane == 0) { out[i] = (int8_t)v; dist[i] = (uint8_t)d; }
}

int cb_debug_device_tier1_gates(cb_handle* h, const cb_board* boards, int n, int koth, int mate_depth, int exhaustive_depth, int koth_depth,
                                int budget, int8_t* out, uint8_t* dist) {
    if (!h || n < 0 || (n > 0 && (!boards || !out || !dist))) return CB_EINVAL;
    if (mate_depth < 0 || mate_depth > 5 || koth_depth < 0 || koth_depth > 4 || exhaustive_depth < 0) return fail(h, CB_EINVAL, "tier1 depths");
    if (n == 0) return CB_OK;
    CB_CUDA(h, cudaSetDevice(h->device));
    int rc = ensure_capacity(h, n);
    if (rc) return rc;
    if (!h->chess_tables_uploaded) {
        CB_CUDA(h, cudaMemcpyToSymbol(cbchess::d_chess_tables, &cbchess::host_tables(), sizeof(cbchess::Tables)));
        h->chess_tables_uploaded = true;
    }
    GateJob* jobs = nullptr;
    int8_t* d_out = nullptr;
    CB_CUDA(h, cudaMalloc((void**)&jobs, (size_t)n * sizeof(GateJob)));
    cudaError_t e = cudaMalloc((void**)&d_out, (size_t)n * 2);
    if (e == cudaSuccess) e = cudaMemcpyAsync(h->d_boards, boards, (size_t)n * sizeof(cb_board), cudaMemcpyHostToDevice, h->stream);
    if (e == cudaSuccess) {
        const GateCfg gc = {koth ? koth_depth : 0, mate_depth, exhaustive_depth};
        device_tier1_gates_kernel<<<(n + 3) / 4, 128, 0, h->stream>>>(h->d_boards, n, koth, gc, budget > 0 ? budget : 1 << 30, jobs, d_out,
                                                                      (uint8_t*)d_out + n);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(out, d_out, (size_t)n, cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(dist, d_out + n, (size_t)n, cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    cudaFree(jobs);
    cudaFree(d_out);
    if (e != cudaSuccess) return fail(h, CB_ECUDA, std::string("device gates: ") + cudaGetErrorString(e));
    return CB_OK;
}

int cb_debug_device_light_gates(cb_handle* h, const cb_board* boards, int n, int koth, int8_t* out) {
    if (!h || n < 0 || (n > 0 && (!boards || !out))) return CB_EINVAL;
    if (n == 0) return CB_OK;
    CB_CUDA(h, cudaSetDevice(h->device));
    if (!h->chess_tables_uploaded) {
        CB_CUDA(h, cudaMemcpyToSymbol(cbchess::d_chess_tables, &cbchess::host_tables(), sizeof(cbchess::Tables)));
        h->chess_tables_uploaded = true;
    }
    int rc = upload_leaves(h, boards, nullptr, n, nullptr, nullptr);
    if (rc) return rc;
    if ((rc = ensure_scratch(h, (size_t)n))) return rc;
    device_light_gates_kernel<<<(n + 7) / 8, 256, 0, h->stream>>>(h->d_boards, n, koth, (int8_t*)h->d_scratch);
    CB_CUDA(h, cudaGetLastError());
    CB_CUDA(h, cudaMemcpyAsync(out, h->d_scratch, (size_t)n, cudaMemcpyDeviceToHost, h->stream));
    CB_CUDA(h, cudaStreamSynchronize(h->stream));
    return CB_OK;
}

// One warp per board: the principal-exchange line of host/qsearch.hpp.
__global__ void __launch_bounds__(256) principal_exchange_kernel(const cb_board* __restrict__ boards, int n, float* __restrict__ q) {
    __shared__ cb_board s_b[8], s_nb[8];
    __shared__ uint8_t s_mailbox[8][64];
    const int wl = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int i = blockIdx.x * 8 + wl;
    if (i >= n) return;
    if (lane < 16) reinterpret_cast<uint64_t*>(&s_b[wl])[lane] = reinterpret_cast<const uint64_t*>(&boards[i])[lane];
    __syncwarp();
    const int cp = cbchess::principal_exchange_cp_warp(s_b[wl], s_nb[wl], s_mailbox[wl], lane);
    if (lane == 0) q[i] = __fdiv_rn((float)cp, 100.0f);
}

// One warp per board: the cap = 1 quiescence search of mcts_device.cuh (cap1_cp_warp).
__global__ void __launch_bounds__(128) cap1_qsearch_kernel(const cb_board* __restrict__ boards, int n, float* __restrict__ q,
                                                           uint8_t* __restrict__ completed) {
    __shared__ Cap1Scratch s_S[4];
    __shared__ Move s_pseudo[4][CB_MAX_MOVES];
    __shared__ uint8_t s_mailbox[4][64];
    const int wl = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int i = blockIdx.x * 4 + wl;
    if (i >= n) return;
    if (lane < 16) reinterpret_cast<uint64_t*>(&s_S[wl].boards[0])[lane] = reinterpret_cast<const uint64_t*>(&boards[i])[lane];
    __syncwarp();
    bool c = true;
    const int cp = cap1_cp_warp(s_S[wl], s_pseudo[wl], s_mailbox[wl], lane, &c);
    if (lane == 0) { q[i] = __fdiv_rn((float)cp, 100.0f); completed[i] = c ? 1 : 0; }
}

int cb_cap1_qsearch_device(cb_handle* h, const cb_board* boards, int n, float* q, uint8_t* completed) {
    if (!h || n < 0 || (n > 0 && (!boards || !q || !completed))) return CB_EINVAL;
    if (n == 0) return CB_OK;
    CB_CUDA(h, cudaSetDevice(h->device));
    if (!h->chess_tables_uploaded) {
        CB_CUDA(h, cudaMemcpyToSymbol(cbchess::d_chess_tables, &cbchess::host_tables(), sizeof(cbchess::Tables)));
        h->chess_tables_uploaded = true;
    }
    int rc = upload_leaves(h, boards, nullptr, n, nullptr, nullptr);
    if (rc) return rc;
    if ((rc = ensure_scratch(h, (size_t)n * 2))) return rc;
    uint8_t* d_c = reinterpret_cast<uint8_t*>(h->d_scratch + n);
    cap1_qsearch_kernel<<<(n + 3) / 4, 128, 0, h->stream>>>(h->d_boards, n, h->d_scratch, d_c);
    CB_CUDA(h, cudaGetLastError());
    CB_CUDA(h, cudaMemcpyAsync(q, h->d_scratch, (size_t)n * 4, cudaMemcpyDeviceToHost, h->stream));
    CB_CUDA(h, cudaMemcpyAsync(completed, d_c, (size_t)n, cudaMemcpyDeviceToHost, h->stream));
    CB_CUDA(h, cudaStreamSynchronize(h->stream));
    return CB_OK;
}

int cb_principal_exchange_device(cb_handle* h, const cb_board* boards, int n, float* q) {
    if (!h || n < 0 || (n > 0 && (!boards || !q))) return CB_EINVAL;
    if (n == 0) return CB_OK;
    CB_CUDA(h, cudaSetDevice(h->device));
    if (!h->chess_tables_uploaded) {
        CB_CUDA(h, cudaMemcpyToSymbol(cbchess::d_chess_tables, &cbchess::host_tables(), sizeof(cbchess::Tables)));
        h->chess_tables_uploaded = true;
    }
    int rc = upload_leaves(h, boards, nullptr, n, nullptr, nullptr);
    if (rc) return rc;
    if ((rc = ensure_scratch(h, (size_t)n))) return rc;
    principal_exchange_kernel<<<(n + 7) / 8, 256, 0, h->stream>>>(h->d_boards, n, h->d_scratch);
    CB_CUDA(h, cudaGetLastError());
    CB_CUDA(h, cudaMemcpyAsync(q, h->d_scratch, (size_t)n * 4, cudaMemcpyDeviceToHost, h->stream));
    CB_CUDA(h, cudaStreamSynchronize(h->stream));
    return CB_OK;
}

// Device -> host read-back of the per-position scalars and (n_priors > 0) the gathered priors.
// Device -> pinned copies of a batch's results, enqueued behind the forward; ev3 marks their completion.
static int enqueue_results(cb_handle* h, int B, size_t n_priors) {
    float* s = h->h_small;
    if (n_priors && (size_t)B * 4 >= (size_t)h->cap * 3) {
        CB_CUDA(h, cudaMemcpyAsync(h->h_out, h->d_out, h->out_pri_off + n_priors * 4, cudaMemcpyDeviceToHost, h->stream));
    } else {
        CB_CUDA(h, cudaMemcpyAsync(s, h->d_vlogit, (size_t)B * 4, cudaMemcpyDeviceToHost, h->stream));
        CB_CUDA(h, cudaMemcpyAsync(s + h->cap, h->d_vfinal, (size_t)B * 4, cudaMemcpyDeviceToHost, h->stream));
        CB_CUDA(h, cudaMemcpyAsync(s + 2 * (size_t)h->cap, h->d_k, (size_t)B * 4, cudaMemcpyDeviceToHost, h->stream));
        if (n_priors)
            CB_CUDA(h, cudaMemcpyAsync(h->h_priors, h->d_priors, n_priors * 4, cudaMemcpyDeviceToHost, h->stream));
    }
    CB_CUDA(h, cudaEventRecord(h->ev3, h->stream));
    return CB_OK;
}

// Pinned -> caller buffers once ev3 has passed.
static int finish_results(cb_handle* h, int B, size_t n_priors, float* priors, float* v_logit, float* v_final, float* k) {
    float* s = h->h_small;
    CB_CUDA(h, cudaEventSynchronize(h->ev3));
    if (v_logit) memcpy(v_logit, s, (size_t)B * 4);
    if (v_final) memcpy(v_final, s + h->cap, (size_t)B * 4);
    if (k) memcpy(k, s + 2 * (size_t)h->cap, (size_t)B * 4);
    if (n_priors && priors) memcpy(priors, h->h_priors, n_priors * 4);
    return CB_OK;
}

static int read_results(cb_handle* h, int B, size_t n_priors, float* priors, float* v_logit, float* v_final, float* k) {
    const int rc = enqueue_results(h, B, n_priors);
    return rc ? rc : finish_results(h, B, n_priors, priors, v_logit, v_final, k);
}

static void account(cb_handle* h, int B) {
    float a = 0, b = 0, c = 0;
    cudaEventElapsedTime(&a, h->ev0, h->ev1);
    cudaEventElapsedTime(&b, h->ev1, h->ev2);
    cudaEventElapsedTime(&c, h->ev2, h->ev3);
    h->timing.calls++;
    h->timing.positions += (uint64_t)B;
    h->timing.h2d_ms += a;
    h->timing.forward_ms += b;
    h->timing.d2h_ms += c;
}

int cb_predict_batch(cb_handle* h, const cb_board* boards, const float* q, const uint8_t* qflag, int B,
                     float* policy_probs, float* v_logit, float* k) {
    (void)qflag;  // the network ignores scalars[:,1] (python/model.py:106)
    if (!h || !boards || B < 0) return CB_EINVAL;
    if (!h->loaded) return fail(h, CB_ESTATE, "weights not loaded");
    if (B == 0) return CB_OK;
    CB_CUDA(h, cudaSetDevice(h->device));
    CB_CUDA(h, cudaEventRecord(h->ev0, h->stream));
    int rc = upload_leaves(h, boards, q, B, nullptr, nullptr);
    if (rc) return rc;
    if (policy_probs && h->probs_cap < h->cap) {
        free_dev(h->d_probs);
        h->d_probs = nullptr;
        drop_graphs(h);
        CB_CUDA(h, cudaMalloc(&h->d_probs, (size_t)h->cap * CB_POLICY_SIZE * 4));
        h->probs_cap = h->cap;
    }
    CB_CUDA(h, cudaEventRecord(h->ev1, h->stream));
    if ((rc = run_forward(h, B, policy_probs != nullptr, false, 1, q != nullptr))) return rc;
    CB_CUDA(h, cudaEventRecord(h->ev2, h->stream));
    if (policy_probs)
        CB_CUDA(h, cudaMemcpyAsync(policy_probs, h->d_probs, (size_t)B * CB_POLICY_SIZE * 4, cudaMemcpyDeviceToHost, h->stream));
    if ((rc = read_results(h, B, 0, nullptr, v_logit, nullptr, k))) return rc;
    account(h, B);
    return CB_OK;
}

int cb_eval_leaves_submit(cb_handle* h, const cb_board* boards, const float* q, int B, int material_on,
                          const uint16_t* move_idx, const uint32_t* move_off) {
    if (!h || !boards || B < 0) return CB_EINVAL;
    if ((move_idx == nullptr) != (move_off == nullptr)) return CB_EINVAL;
    if (!h->loaded) return fail(h, CB_ESTATE, "weights not loaded");
    if (h->pending_B >= 0) return fail(h, CB_ESTATE, "a submitted batch is still waiting for cb_eval_leaves_wait");
    CB_CUDA(h, cudaSetDevice(h->device));
    h->pending_B = 0;
    h->pending_priors = 0;
    if (B == 0) return CB_OK;
    CB_CUDA(h, cudaEventRecord(h->ev0, h->stream));
    int rc = upload_leaves(h, boards, q, B, move_idx, move_off);   // the caller's buffers are free again on return
    if (!rc) {
        CB_CUDA(h, cudaEventRecord(h->ev1, h->stream));
        rc = run_forward(h, B, false, move_off != nullptr, material_on, q != nullptr);
    }
    if (!rc) {
        CB_CUDA(h, cudaEventRecord(h->ev2, h->stream));
        rc = enqueue_results(h, B, h->staged_moves);
    }
    if (rc) { h->pending_B = -1; return rc; }
    h->pending_B = B;
    h->pending_priors = h->staged_moves;
    return CB_OK;
}

int cb_eval_leaves_wait(cb_handle* h, float* priors, float* v_logit, float* v_final, float* k) {
    if (!h) return CB_EINVAL;
    if (h->pending_B < 0) return fail(h, CB_ESTATE, "no submitted batch to wait for");
    const int B = h->pending_B;
    const size_t n = h->pending_priors;
    h->pending_B = -1;
    if (B == 0) return CB_OK;
    if (n && !priors) return fail(h, CB_EINVAL, "the submitted batch has legal-move lists: priors must not be null");
    const int rc = finish_results(h, B, n, priors, v_logit, v_final, k);
    if (rc) return rc;
    account(h, B);
    return CB_OK;
}

int cb_eval_leaves(cb_handle* h, const cb_board* boards, const float* q, int B, int material_on,
                   const uint16_t* move_idx, const uint32_t* move_off, float* priors, float* v_logit,
                   float* v_final, float* k) {
    if (move_off && !priors) return CB_EINVAL;
    const int rc = cb_eval_leaves_submit(h, boards, q, B, material_on, move_idx, move_off);
    return rc ? rc : cb_eval_leaves_wait(h, priors, v_logit, v_final, k);
}

int cb_stage_leaves(cb_handle* h, const cb_board* boards, const float* q, int B, const uint16_t* move_idx,
                    const uint32_t* move_off) {
    if (!h || !boards || B <= 0) return CB_EINVAL;
    if (!h->loaded) return fail(h, CB_ESTATE, "weights not loaded");
    CB_CUDA(h, cudaSetDevice(h->device));
    int rc = upload_leaves(h, boards, q, B, move_idx, move_off);
    if (rc) return rc;
    if (!q) CB_CUDA(h, cudaMemsetAsync(h->d_q, 0, (size_t)B * 4, h->stream));
    CB_CUDA(h, cudaStreamSynchronize(h->stream));
    h->staged_B = B;
    return CB_OK;
}

int cb_time_resident(cb_handle* h, int material_on, int iters, int flush_l2, float* ms_each) {
    if (!h || iters <= 0 || !ms_each) return CB_EINVAL;
    if (!h->staged_B) return fail(h, CB_ESTATE, "nothing staged: call cb_stage_leaves first");
    CB_CUDA(h, cudaSetDevice(h->device));
    if (flush_l2 && !h->d_flush) CB_CUDA(h, cudaMalloc(&h->d_flush, kFlushBytes));
    std::vector<cudaEvent_t> ev(2 * (size_t)iters);
    for (auto& e : ev) CB_CUDA(h, cudaEventCreate(&e));
    int rc = CB_OK;
    for (int i = 0; i < iters && !rc; ++i) {
        if (flush_l2) CB_CUDA(h, cudaMemsetAsync(h->d_flush, i & 0xFF, kFlushBytes, h->stream));
        CB_CUDA(h, cudaEventRecord(ev[2 * i], h->stream));
        rc = run_forward(h, h->staged_B, false, h->staged_moves > 0, material_on, true);
        CB_CUDA(h, cudaEventRecord(ev[2 * i + 1], h->stream));
    }
    cudaError_t se = cudaStreamSynchronize(h->stream);
    if (!rc && se == cudaSuccess)
        for (int i = 0; i < iters; ++i) cudaEventElapsedTime(&ms_each[i], ev[2 * i], ev[2 * i + 1]);
    for (auto& e : ev) cudaEventDestroy(e);
    if (rc) return rc;
    CB_CUDA(h, se);
    return CB_OK;
}

int cb_time_sustained(cb_handle* h, int material_on, double seconds, float* ms_total, int* launches) {
    if (!h || seconds <= 0 || !ms_total || !launches) return CB_EINVAL;
    if (!h->staged_B) return fail(h, CB_ESTATE, "nothing staged: call cb_stage_leaves first");
    CB_CUDA(h, cudaSetDevice(h->device));
    constexpr int kChunk = 256;
    cudaEvent_t e0, e1;
    CB_CUDA(h, cudaEventCreate(&e0));
    CB_CUDA(h, cudaEventCreate(&e1));
    double total = 0.0;
    int n = 0, rc = CB_OK;
    cudaError_t se = cudaSuccess;
    const auto t0 = std::chrono::steady_clock::now();
    do {
        cudaEventRecord(e0, h->stream);
        for (int i = 0; i < kChunk && !rc; ++i) rc = run_forward(h, h->staged_B, false, h->staged_moves > 0, material_on, true);
        cudaEventRecord(e1, h->stream);
        se = cudaStreamSynchronize(h->stream);
        if (rc || se != cudaSuccess) break;
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        total += ms;
        n += kChunk;
    } while (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() < seconds);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (rc) return rc;
    CB_CUDA(h, se);
    *ms_total = (float)total;
    *launches = n;
    return CB_OK;
}

int cb_time_conv_layers(cb_handle* h, int iters, float* conv_ms_total, int* n_conv_launches) {
    return cb_time_conv_layers_detail(h, iters, conv_ms_total, n_conv_launches, nullptr);
}

int cb_time_conv_layers_detail(cb_handle* h, int iters, float* conv_ms_total, int* n_conv_launches,
                               float* per_layer_ms) {
    if (!h || iters <= 0) return CB_EINVAL;
    if (!h->staged_B) return fail(h, CB_ESTATE, "nothing staged: call cb_stage_leaves first");
    CB_CUDA(h, cudaSetDevice(h->device));
    const int nconv = 2 * h->blocks + 2;
    std::vector<cudaEvent_t> ev(2 * (size_t)nconv);
    for (auto& e : ev) CB_CUDA(h, cudaEventCreate(&e));
    double total = 0.0;
    int rc = CB_OK;
    for (int i = 0; i < iters && !rc; ++i) {
        rc = enqueue_tower(h, h->staged_B, &ev);
        if (rc) break;
        if (cudaStreamSynchronize(h->stream) != cudaSuccess) { rc = fail(h, CB_ECUDA, "sync failed in cb_time_conv_layers"); break; }
        // dominant kernel = the HID->HID layers: every conv launch except the start conv (index 0)
        for (int l = 0; l < nconv; ++l) {
            float ms = 0;
            cudaEventElapsedTime(&ms, ev[2 * l], ev[2 * l + 1]);
            if (l >= 1) total += ms;
            if (per_layer_ms) per_layer_ms[l] = (i == 0 ? 0.0f : per_layer_ms[l]) + ms / iters;
        }
    }
    for (auto& e : ev) cudaEventDestroy(e);
    if (rc) return rc;
    if (conv_ms_total) *conv_ms_total = (float)total;
    if (n_conv_launches) *n_conv_launches = nconv - 1;
    return CB_OK;
}

int cb_time_tower(cb_handle* h, int iters, float* ms_each, int* n_launches) {
    if (!h || iters <= 0 || !ms_each) return CB_EINVAL;
    if (!h->staged_B) return fail(h, CB_ESTATE, "nothing staged: call cb_stage_leaves first");
    CB_CUDA(h, cudaSetDevice(h->device));
    const int kind = fused_kind(h);
    const bool fused = kind != 0;
    std::vector<cudaEvent_t> ev(2 * (size_t)iters);
    for (auto& e : ev) CB_CUDA(h, cudaEventCreate(&e));
    int rc = CB_OK;
    for (int i = 0; i < iters && !rc; ++i) {
        CB_CUDA(h, cudaEventRecord(ev[2 * i], h->stream));
        rc = kind == 3   ? launch_tower_r_staged(h, h->staged_B)
             : kind == 4 ? launch_tower_r2_staged(h, h->staged_B)
             : kind == 1 ? launch_tower(h, h->staged_B)
                         : enqueue_tower(h, h->staged_B, nullptr);
        CB_CUDA(h, cudaEventRecord(ev[2 * i + 1], h->stream));
    }
    cudaError_t se = cudaStreamSynchronize(h->stream);
    if (!rc && se == cudaSuccess)
        for (int i = 0; i < iters; ++i) cudaEventElapsedTime(&ms_each[i], ev[2 * i], ev[2 * i + 1]);
    for (auto& e : ev) cudaEventDestroy(e);
    if (rc) return rc;
    CB_CUDA(h, se);
    if (n_launches) *n_launches = fused ? 1 : 2 * h->blocks + 2 + (h->dtype == CB_DTYPE_FP32 ? h->blocks : 0);
    return CB_OK;
}

int cb_read_resident(cb_handle* h, float* priors, float* v_logit, float* v_final, float* k) {
    if (!h) return CB_EINVAL;
    if (!h->staged_B) return fail(h, CB_ESTATE, "nothing staged");
    CB_CUDA(h, cudaSetDevice(h->device));
    return read_results(h, h->staged_B, priors ? h->staged_moves : 0, priors, v_logit, v_final, k);
}

int cb_debug_tower_trace(cb_handle* h, long long* out, int max_items) {
    if (!h || !out || max_items <= 0) return CB_EINVAL;
    const int kind = fused_kind(h);
    if (!h->staged_B || kind != 3) return fail(h, CB_ESTATE, "needs staged leaves and the resident tower");
    CB_CUDA(h, cudaSetDevice(h->device));
    const int items = 2 * ((int)h->convs.size() + 1);   // the resident tower's last "layer" is the policy head
    const int n = items < max_items ? items : max_items;
    drop_graphs(h);
    CB_CUDA(h, cudaMalloc(&h->d_trace, (size_t)items * 8 * sizeof(long long)));
    CB_CUDA(h, cudaMemsetAsync(h->d_trace, 0, (size_t)items * 8 * sizeof(long long), h->stream));
    int rc = launch_tower_r_staged(h, h->staged_B);
    cudaError_t e = cudaMemcpyAsync(out, h->d_trace, (size_t)n * 8 * sizeof(long long), cudaMemcpyDeviceToHost, h->stream);
    cudaError_t e2 = cudaStreamSynchronize(h->stream);
    cudaFree(h->d_trace);
    h->d_trace = nullptr;
    if (rc) return rc;
    CB_CUDA(h, e);
    CB_CUDA(h, e2);
    return n;
}

int cb_debug_cta_times(cb_handle* h, unsigned long long* out, int max_ctas) {
    if (!h || !out || max_ctas <= 0) return CB_EINVAL;
    if (!h->staged_B || fused_kind(h) != 3) return fail(h, CB_ESTATE, "needs staged leaves and the resident tower");
    CB_CUDA(h, cudaSetDevice(h->device));
    const int units = tower_r_units(h, h->staged_B);
    const int grid = units < h->sm_count ? units : h->sm_count;
    const int n = grid < max_ctas ? grid : max_ctas;
    CB_CUDA(h, cudaMalloc(&h->d_cta_times, (size_t)grid * 4 * sizeof(unsigned long long)));
    CB_CUDA(h, cudaMemsetAsync(h->d_cta_times, 0, (size_t)grid * 4 * sizeof(unsigned long long), h->stream));
    int rc = launch_tower_r_staged(h, h->staged_B);
    cudaError_t e = cudaMemcpyAsync(out, h->d_cta_times, (size_t)n * 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, h->stream);
    cudaError_t e2 = cudaStreamSynchronize(h->stream);
    cudaFree(h->d_cta_times);
    h->d_cta_times = nullptr;
    if (rc) return rc;
    CB_CUDA(h, e);
    CB_CUDA(h, e2);
    return n;
}

int cb_debug_set_layer_limit(cb_handle* h, int n_conv_launches) {
    if (!h) return CB_EINVAL;
    drop_graphs(h);
    h->layer_limit = n_conv_launches;
    return CB_OK;
}

// Self-play with the tree operations on the device (cb_selfplay_run, pipeline = 3): per simulation step one
// CUDA graph node each for select, the forward kernel and expand / play; the host only replays the graph
// and reads the counters.
// Simulation steps per CUDA-graph replay of the device self-play loop.  Between replays the host reads the counters
// and drains the sample log while the GPU idles (~30 us): 8 steps per replay cost 2 % of the run, 32 cost 0.5 %.
constexpr int kSelfplayStepsPerGraph = 32;

// Per-game tree state of the device drivers.
static int mcts_alloc(cb_handle* h, MctsParams& mp, int G, std::vector<void*>& owned) {
    auto dalloc = [&](void** ptr, size_t bytes) -> int {
        CB_CUDA(h, cudaMalloc(ptr, bytes));
        owned.push_back(*ptr);
        return CB_OK;
    };
    int rc;
    if ((rc = dalloc((void**)&mp.root, (size_t)G * sizeof(cb_board))) ||
        (rc = dalloc((void**)&mp.nodes, (size_t)G * 2 * mp.node_cap * sizeof(MNode))) ||
        (rc = dalloc((void**)&mp.n_nodes, (size_t)G * 4)) || (rc = dalloc((void**)&mp.arena, (size_t)G)) ||
        (rc = dalloc((void**)&mp.history, (size_t)G * kMctsHist * 8)) || (rc = dalloc((void**)&mp.hist_n, (size_t)G * 4)) ||
        (rc = dalloc((void**)&mp.rng, (size_t)G * 8)) || (rc = dalloc((void**)&mp.ply, (size_t)G * 4)) ||
        (rc = dalloc((void**)&mp.sims_done, (size_t)G * 4)) || (rc = dalloc((void**)&mp.samples_in_game, (size_t)G * 4)) ||
        (rc = dalloc((void**)&mp.path, (size_t)G * kMctsMaxDepth * 4)) || (rc = dalloc((void**)&mp.path_n, (size_t)G * 4)) ||
        (rc = dalloc((void**)&mp.wants_eval, (size_t)G)) || (rc = dalloc((void**)&mp.leaf_moves, (size_t)G * CB_MAX_MOVES * 2)) ||
        (rc = dalloc((void**)&mp.leaf_trank, (size_t)G * CB_MAX_MOVES)) || (rc = dalloc((void**)&mp.leaf_ntact, (size_t)G)) ||
        (rc = dalloc((void**)&mp.idle, (size_t)G)) ||
        (rc = dalloc((void**)&mp.eval_cnt, (size_t)G * 2)) || (rc = dalloc((void**)&mp.stats, MS_COUNT * 8)))
        return rc;
    CB_CUDA(h, cudaMemsetAsync(mp.idle, 0, (size_t)G, h->stream));
    if (mp.tier1_light >= 2 || (mp.qsearch == 2 && mp.material_on)) {   // the stacks of the games' gate / cap1 jobs
        if (mp.gate_warps < 1) mp.gate_warps = 1;
        if ((rc = dalloc((void**)&mp.gate_jobs, (size_t)G * mp.gate_warps * sizeof(GateJob)))) return rc;
        CB_CUDA(h, cudaMemsetAsync(mp.gate_jobs, 0, (size_t)G * mp.gate_warps * sizeof(GateJob), h->stream));
    }
    return CB_OK;
}

// The gate depths of tier1_light = 2 (the reference's self-play: mate-in-5 by checks, hill in 3) and the per-step budget.
static int mcts_gate_params(cb_handle* h, MctsParams& mp, int mate_depth, int exhaustive_depth, int koth_depth) {
    mp.t1_mate_depth = mate_depth > 0 ? mate_depth : 5;
    mp.t1_exh_depth = exhaustive_depth;
    mp.t1_koth_depth = koth_depth > 0 ? koth_depth : 3;
    if (mp.t1_mate_depth > 5 || mp.t1_koth_depth > 4 || exhaustive_depth < 0)
        return fail(h, CB_EINVAL, "tier1 depths: mate search up to 5, hill up to 4");
    const char* e = getenv("CB_GATE_BUDGET");
    mp.gate_budget = e && atoi(e) > 0 ? atoi(e) : 24;
    // a step whose jobs stop after `gate_budget` positions while most games are still inside long searches runs the forward for
    // a handful of leaves: the gate kernel lets the jobs go on (up to gate_stretch more positions) until gate_target games
    // of the launch are done with their step (CB_GATE_STRETCH = 0 turns it off; set where the pool size is known)
    const char* s2 = getenv("CB_GATE_STRETCH");
    mp.gate_stretch = s2 ? atoi(s2) : 400;
    // warps per game in the gate kernel: each takes every gate_warps-th move of the gate's root (1, 2, 4 or 8).  Measured
    // (profiles/r02_deep_gates_probe_warps.json): more than one does not pay - from ~4 warps per scheduler the kernel is bound
    // by instruction issue (1,400 instructions per position at a third of the lanes), and every share generates the root's
    // moves again - so the default is 1
    const char* w = getenv("CB_GATE_WARPS");
    const int gw = w ? atoi(w) : 1;
    mp.gate_warps = gw == 1 || gw == 2 || gw == 4 || gw == 8 ? gw : 1;
    return CB_OK;
}

// cb_debug_device_search: start positions and where the root children go.
struct DeviceSearchIO {
    const cb_board* boards;
    uint16_t* moves;
    uint32_t* visits;
    double* value_sums;
    int32_t* counts;
    uint32_t* root_visits;
};

static int selfplay_device_impl(cb_handle* h, const cb_selfplay_config* cfg, cb_selfplay_stats* out, const DeviceSearchIO* dbg) {
    if (!h || !cfg || !out || cfg->games <= 0 || cfg->sims_per_move <= 0) return CB_EINVAL;
    if (!h->loaded) return fail(h, CB_ESTATE, "weights not loaded");
    CB_CUDA(h, cudaSetDevice(h->device));
    const int G = cfg->games;
    const int pools = (!dbg && cfg->pipeline == 4) ? 2 : 1;
    if (pools == 2 && (G & 1)) return fail(h, CB_EINVAL, "two pools need an even number of games");
    int rc = ensure_capacity(h, G);
    if (rc) return rc;
    if (pools == 2 && fused_kind(h) != 3)
        return fail(h, CB_ESTATE, "two pools need the resident single-launch forward (128-channel nets, 16-bit mode)");
    memset(out, 0, sizeof(*out));
    if (!h->chess_tables_uploaded) {
        CB_CUDA(h, cudaMemcpyToSymbol(cbchess::d_chess_tables, &cbchess::host_tables(), sizeof(cbchess::Tables)));
        h->chess_tables_uploaded = true;
    }
    MctsParams mp = {};
    mp.G = G;
    mp.gn = G;
    mp.sims_per_move = cfg->sims_per_move;
    mp.max_plies = cfg->max_plies > 0 ? cfg->max_plies : 200;
    mp.material_on = cfg->material_on;
    mp.qsearch = cfg->qsearch;
    mp.tier1_light = cfg->tier1_light;
    if (cfg->tier1_light >= 2 || cfg->qsearch == 2) {   // (cap1 searches run under the same per-step budget)
        if ((rc = mcts_gate_params(h, mp, cfg->tier1_mate_depth, cfg->tier1_exhaustive_depth, cfg->tier1_koth_depth))) return rc;
    }
    mp.koth = cfg->koth;
    mp.c_puct = cfg->c_puct > 0 ? cfg->c_puct : 1.414;
    mp.explore_base = cfg->explore_base > 0 ? cfg->explore_base : 0.80;
    mp.shuffle = cfg->shuffle_children;
    mp.mock_eval = cfg->mock_eval;
    mp.search_only = dbg ? 1 : 0;
    {   // a move adds sims_per_move expansions of ~35 (at most 218) children to the subtree kept from the last move; when
        // the played child holds a share s of the visits the tree settles near new / (1 - s) nodes, so leave room for
        // s ~ 0.9 (32-byte nodes: 1024 games x 2 arenas x 65536 nodes = 4.3 GB).  A full arena is counted
        // (cb_selfplay_stats.overflow): the leaf then stays unexpanded and is evaluated again later.
        const long long want = (long long)cfg->sims_per_move * 400;
        mp.node_cap = (uint32_t)(want < 65536 ? 65536 : want > 524288 ? 524288 : want);
    }
    mp.seed = cfg->seed;
    std::vector<void*> owned;
    auto dalloc = [&](void** ptr, size_t bytes) -> int {
        CB_CUDA(h, cudaMalloc(ptr, bytes));
        owned.push_back(*ptr);
        return CB_OK;
    };
    auto release = [&] { for (void* q : owned) cudaFree(q); };
    if ((rc = mcts_alloc(h, mp, G, owned))) {
        release();
        return rc;
    }
    mp.eval_boards = h->d_boards;
    mp.eval_q = h->d_q;
    mp.eval_idx = h->d_move_idx;
    mp.priors = h->d_priors;
    mp.v_final = h->d_vfinal;
    // training samples: device-side log drained between graph replays, games assembled and written on the host
    FILE* sample_file = nullptr;
    std::vector<uint8_t> ring_host;
    std::vector<std::vector<DevSample>> pending;     // per game slot: samples of the game in progress
    unsigned long long ring_read = 0;
    if (cfg->sample_path) {
        sample_file = fopen(cfg->sample_path, "ab");
        if (!sample_file) { release(); return fail(h, CB_EINVAL, "cannot open the sample file"); }
        // a game appends at most one record per simulation step (+1 at its end): size the ring for 2 replays
        mp.ring_slots = (uint32_t)(4 * G * kSelfplayStepsPerGraph + 1024);
        if ((rc = dalloc((void**)&mp.sample_ring, (size_t)mp.ring_slots * kSampleSlotBytes)) ||
            (rc = dalloc((void**)&mp.ring_cursor, 8)) || (rc = dalloc((void**)&mp.game_serial, (size_t)G * 4))) {
            fclose(sample_file);
            release();
            return rc;
        }
        cudaMemsetAsync(mp.ring_cursor, 0, 8, h->stream);
        cudaMemsetAsync(mp.game_serial, 0, (size_t)G * 4, h->stream);
        pending.resize(G);
    }
    if (cfg->mock_eval && pools == 2) { if (sample_file) fclose(sample_file); release(); return fail(h, CB_EINVAL, "the mock evaluator runs with one pool"); }
    if (dbg) {
        cb_board* d_start = nullptr;
        if ((rc = dalloc((void**)&d_start, (size_t)G * sizeof(cb_board)))) { release(); return rc; }
        CB_CUDA(h, cudaMemcpyAsync(d_start, dbg->boards, (size_t)G * sizeof(cb_board), cudaMemcpyHostToDevice, h->stream));
        mp.start_boards = d_start;
    }
    cudaMemsetAsync(mp.stats, 0, MS_COUNT * 8, h->stream);
    cudaMemsetAsync(h->d_boards, 0, (size_t)G * sizeof(cb_board), h->stream);
    mcts_init_kernel<<<(G + 127) / 128, 128, 0, h->stream>>>(mp);
    // A simulation step = select, forward, expand.  In steady state the expand of step n and the select of
    // step n + 1 are one launch (mcts_step_kernel), so a step is two kernels; the run starts with a lone
    // select and ends with a lone expand, which completes the simulation in flight.
    //
    // Two pools (pipeline 4): the games are split in two halves that alternate — while the forward kernel of one
    // half runs, the tree kernel of the other half runs beside it.  The forward kernel takes a whole SM per CTA
    // (shared memory), so it is launched on ceil(Gp / 8) SMs (8 boards per CTA, its most efficient unit) and the
    // tree kernel is launched as its PROGRAMMATIC DEPENDENT: it is scheduled once every forward CTA is resident
    // and so lands on the SMs the forward leaves free.  One stream:
    //     fwd(1) , step(0)* , fwd(0) , step(1)* , ...          * = programmatic dependent of the kernel before it
    // step(p) reads the results of fwd(p), which completed before the forward kernel it runs beside started.
    const int Gp = G / pools;
    const int grid = (Gp + kMctsWarps - 1) / kMctsWarps;
    unsigned int* pdl_done = nullptr;          // [2] wrapping counters of finished warps (mcts_step_kernel)
    if ((rc = dalloc((void**)&pdl_done, 8))) { if (sample_file) fclose(sample_file); release(); return rc; }
    cudaMemsetAsync(pdl_done, 0, 8, h->stream);
    const int unit_cap = pools == 2 ? std::max(1, std::min(h->sm_count, (Gp + 7) / 8)) : 0;
    auto pool_params = [&](int pool, bool beside_forward) {
        MctsParams q = mp;
        q.g0 = pool * Gp;
        q.gn = Gp;
        q.pdl_done = beside_forward ? pdl_done + pool : nullptr;
        return q;
    };
    auto fwd = [&](int pool) -> int {
        if (cfg->mock_eval) {   // the reference's mock server in place of the network (tests)
            mcts_mock_eval_kernel<<<(G + 3) / 4, 128, 0, h->stream>>>(pool_params(0, false), cfg->mock_v_logit, cfg->mock_k, h->d_priors, h->d_vfinal);
            return CB_OK;
        }
        if (pools == 1) return enqueue_forward(h, G, false, true, cfg->material_on, true, mp.eval_cnt);
        PolicyOut po = {};
        po.move_idx = h->d_move_idx + (size_t)pool * Gp * CB_MAX_MOVES;
        po.move_off = h->d_move_off;
        po.priors = h->d_priors + (size_t)pool * Gp * CB_MAX_MOVES;
        po.move_cnt = mp.eval_cnt + pool * Gp;
        return launch_tower_r(h, Gp, po, cfg->material_on, true, nullptr, nullptr, pool * Gp, unit_cap);
    };
    // the tree kernel of `pool` beside the forward kernel launched just before it
    auto step_beside = [&](int pool) -> int {
        const MctsParams q = pool_params(pool, true);
        cudaLaunchConfig_t lc = {};
        lc.gridDim = dim3(grid);
        lc.blockDim = dim3(32 * kMctsWarps);
        lc.stream = h->stream;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        lc.attrs = at;
        lc.numAttrs = 1;
        CB_CUDA(h, cudaLaunchKernelEx(&lc, mcts_step_kernel<0>, q));
        return CB_OK;
    };
    // tier1_light = 2 with one pool: the gate jobs run in their own kernel after the tree kernel (mcts_select_body<MODE>)
    const bool split = mp.gate_jobs && pools == 1 && !getenv("CB_GATE_INLINE");
    if (split && mp.gate_stretch > 0 && (long long)G * std::max(1, mp.gate_warps) <= 4096) {   // (every block of the gate kernel must be resident for the count to make sense)
        if ((rc = dalloc((void**)&mp.gate_ready, 4))) { if (sample_file) fclose(sample_file); release(); return rc; }
        cudaMemsetAsync(mp.gate_ready, 0, 4, h->stream);
        const char* t = getenv("CB_GATE_TARGET_PCT");
        mp.gate_target = (unsigned int)((long long)G * (t && atoi(t) > 0 ? atoi(t) : 50) / 100);
    }
    const int gate_grid = (Gp * std::max(1, mp.gate_warps) + kMctsWarps - 1) / kMctsWarps;
    auto tree_step = [&] {   // expand + select of the one pool
        if (split) {
            mcts_step_kernel<1><<<grid, 32 * kMctsWarps, 0, h->stream>>>(pool_params(0, false));
            mcts_gate_kernel<<<gate_grid, 32 * kMctsWarps, 0, h->stream>>>(pool_params(0, false));
        } else {
            mcts_step_kernel<0><<<grid, 32 * kMctsWarps, 0, h->stream>>>(pool_params(0, false));
        }
    };
    auto first_select = [&](int pool) {
        if (split) {
            mcts_select_kernel<1><<<grid, 32 * kMctsWarps, 0, h->stream>>>(pool_params(pool, false));
            mcts_gate_kernel<<<gate_grid, 32 * kMctsWarps, 0, h->stream>>>(pool_params(pool, false));
        } else {
            mcts_select_kernel<0><<<grid, 32 * kMctsWarps, 0, h->stream>>>(pool_params(pool, false));
        }
    };
    // one step of every pool; leaves the run in the state it found it (pool 0 evaluated, the others selected)
    auto fused = [&]() -> int {
        int r;
        if (pools == 1) {
            if ((r = fwd(0))) return r;
            tree_step();
            return CB_OK;
        }
        if ((r = fwd(1)) || (r = step_beside(0)) || (r = fwd(0)) || (r = step_beside(1))) return r;
        return CB_OK;
    };
    if (dbg) {
        // one move's search for every position: a fresh root asks for its policy first (one step that is no simulation),
        // then sims_per_move simulations; games that are done go idle.  Eager launches, then the root children are read.
        first_select(0);
        for (int i = 0; i <= cfg->sims_per_move && !rc; ++i) rc = fused();
        cudaError_t e2 = cudaStreamSynchronize(h->stream);
        long long dbg_steps = cfg->sims_per_move + 1;
        if (mp.gate_jobs) {   // games whose gate / cap1 jobs were suspended need more steps: run until every game is idle
            std::vector<uint8_t> idle(G);
            // (a search whose leaves need ~2000 gate positions each, e.g. a king near the hill, takes sims x 2000 / budget steps)
            for (int round = 0; round < (1 << 16) && !rc && e2 == cudaSuccess; ++round) {
                e2 = cudaMemcpy(idle.data(), mp.idle, (size_t)G, cudaMemcpyDeviceToHost);
                if (e2 != cudaSuccess || std::all_of(idle.begin(), idle.end(), [](uint8_t x) { return x != 0; })) break;
                for (int i = 0; i < 64 && !rc; ++i) rc = fused();
                dbg_steps += 64;
                e2 = cudaStreamSynchronize(h->stream);
            }
            if (getenv("CB_GATE_DEBUG")) {
                std::vector<int> sd(G);
                cudaMemcpy(sd.data(), mp.sims_done, (size_t)G * 4, cudaMemcpyDeviceToHost);
                cudaMemcpy(idle.data(), mp.idle, (size_t)G, cudaMemcpyDeviceToHost);
                for (int g = 0; g < G; ++g) {
                    uint8_t js[16];
                    cudaMemcpy(js, reinterpret_cast<uint8_t*>(mp.gate_jobs) + (size_t)g * std::max(1, mp.gate_warps) * kGateJobBytes + kGateJobActiveOffset, 16, cudaMemcpyDeviceToHost);
                    if (!idle[g] || js[0]) fprintf(stderr, "game %d: idle %d sims_done %d job active %d phase %d it %d ply %d (steps %lld)\n", g, idle[g], sd[g], js[0], js[1], js[2], js[3], dbg_steps);
                }
                fprintf(stderr, "debug search: %lld steps\n", dbg_steps);
            }
        }
        std::vector<MNode> kids(CB_MAX_MOVES);
        std::vector<uint8_t> arena(G);
        unsigned long long st2[MS_COUNT] = {0};
        if (!rc && e2 == cudaSuccess) e2 = cudaMemcpy(arena.data(), mp.arena, (size_t)G, cudaMemcpyDeviceToHost);
        if (!rc && e2 == cudaSuccess) e2 = cudaMemcpy(st2, mp.stats, sizeof(st2), cudaMemcpyDeviceToHost);
        for (int g = 0; g < G && !rc && e2 == cudaSuccess; ++g) {
            const MNode* base = mp.nodes + ((size_t)g * 2 + arena[g]) * mp.node_cap;
            MNode root;
            e2 = cudaMemcpy(&root, base, sizeof(MNode), cudaMemcpyDeviceToHost);
            if (e2 != cudaSuccess) break;
            const int nc = root.state == 1 ? root.n_children : 0;
            dbg->counts[g] = nc;
            if (dbg->root_visits) dbg->root_visits[g] = root.visits;
            if (nc > 0) e2 = cudaMemcpy(kids.data(), base + root.first_child, (size_t)nc * sizeof(MNode), cudaMemcpyDeviceToHost);
            for (int c = 0; c < nc; ++c) {
                if (dbg->moves) dbg->moves[(size_t)g * CB_MAX_MOVES + c] = kids[c].move;
                if (dbg->visits) dbg->visits[(size_t)g * CB_MAX_MOVES + c] = kids[c].visits;
                if (dbg->value_sums) dbg->value_sums[(size_t)g * CB_MAX_MOVES + c] = kids[c].value_sum;
            }
        }
        release();
        if (rc) return rc;
        if (e2 != cudaSuccess) return fail(h, CB_ECUDA, std::string("device search: ") + cudaGetErrorString(e2));
        out->nn_evals = (long long)st2[MS_EVALS];
        out->terminal_hits = (long long)st2[MS_TERMINAL];
        out->overflow = (long long)st2[MS_OVERFLOW];
        out->gate_hits = (long long)st2[MS_GATE_HITS];
        out->gate_waits = (long long)st2[MS_GATE_WAITS] / std::max(1, mp.gate_warps);
        out->steps = dbg_steps;
        return CB_OK;
    }
    const long long target = cfg->max_sims > 0 ? (cfg->max_sims + G - 1) / G : 0;   // steps to run (0 = unbounded)
    long long done = 0;                        // completed steps; one more is always in flight
    for (int pool = 0; pool < pools; ++pool) first_select(pool);
    if (pools == 2) rc = fwd(0);
    // one eager fused step (sets the kernels' attributes outside capture)
    if (!rc && (target == 0 || target > 1)) {
        rc = fused();
        done = 1;
    }
    cudaError_t ce = cudaStreamSynchronize(h->stream);
    if (rc || ce != cudaSuccess) {
        if (sample_file) fclose(sample_file);
        release();
        if (rc) return rc;
        return fail(h, CB_ECUDA, std::string("device self-play step: ") + cudaGetErrorString(ce));
    }
    constexpr int kStepsPerGraph = kSelfplayStepsPerGraph;
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t gexec = nullptr;
    CB_CUDA(h, cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
    for (int i = 0; i < kStepsPerGraph && !rc; ++i) rc = fused();
    ce = cudaStreamEndCapture(h->stream, &graph);
    if (!rc && ce == cudaSuccess) ce = cudaGraphInstantiate(&gexec, graph, 0);
    if (graph) cudaGraphDestroy(graph);
    if (rc || ce != cudaSuccess) {
        if (sample_file) fclose(sample_file);
        release();
        if (rc) return rc;
        return fail(h, CB_ECUDA, std::string("device self-play graph: ") + cudaGetErrorString(ce));
    }
    unsigned long long st[MS_COUNT] = {0};
    float ms_fwd = 0, ms_tree = 0;             // split of one steady-state step (CUDA events, measured once mid-run)
    bool split_done = pools == 2;              // (two pools: the kernels overlap, there is no split to report)
    const auto t_start = std::chrono::steady_clock::now();
    auto elapsed = [&] { return std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count(); };
    // The sample log is drained while the NEXT graph replay runs: the cursor is read with the counters (those records
    // are complete), the replay is launched, then the ring is copied on a second stream and the finished games are
    // written.  The ring holds four replays' worth of records, the drain is at most one replay behind.
    cudaStream_t copy_stream = nullptr;
    if (sample_file) cudaStreamCreateWithFlags(&copy_stream, cudaStreamNonBlocking);
    unsigned long long cursor = 0;
    for (;;) {
        ce = cudaMemcpyAsync(st, mp.stats, sizeof(st), cudaMemcpyDeviceToHost, h->stream);
        if (ce == cudaSuccess && sample_file) ce = cudaMemcpyAsync(&cursor, mp.ring_cursor, 8, cudaMemcpyDeviceToHost, h->stream);
        if (ce == cudaSuccess) ce = cudaStreamSynchronize(h->stream);
        if (ce != cudaSuccess) break;
        // the step in flight is completed after the loop: stop once `done + 1` steps suffice
        if (target > 0 && done + 1 >= target) break;
        if (cfg->max_games > 0 && (long long)st[MS_GAMES] >= cfg->max_games) break;
        if (cfg->max_seconds > 0 && elapsed() >= cfg->max_seconds) break;
        if (!split_done && done >= 3 * kStepsPerGraph) {
            cudaEvent_t e[3];
            for (auto& x : e) cudaEventCreate(&x);
            cudaEventRecord(e[0], h->stream);
            rc = fwd(0);
            cudaEventRecord(e[1], h->stream);
            tree_step();
            cudaEventRecord(e[2], h->stream);
            ce = cudaStreamSynchronize(h->stream);
            cudaEventElapsedTime(&ms_fwd, e[0], e[1]);
            cudaEventElapsedTime(&ms_tree, e[1], e[2]);
            for (auto& x : e) cudaEventDestroy(x);
            if (rc || ce != cudaSuccess) break;
            done += 1;
            split_done = true;
            if (getenv("CB_MCTS_TIMES")) fprintf(stderr, "device self-play step: forward %.1f us, expand + select %.1f us\n", ms_fwd * 1e3, ms_tree * 1e3);
        } else if (target > 0 && target - 1 - done < kStepsPerGraph) {
            if ((rc = fused())) break;          // finish the simulation budget exactly with eager steps
            done += 1;
        } else {
            if ((ce = cudaGraphLaunch(gexec, h->stream)) != cudaSuccess) break;
            done += kStepsPerGraph;
        }
        if (sample_file && (rc = drain_samples(h, mp, ring_read, ring_host, pending, sample_file, cursor, copy_stream))) break;
    }
    if (rc || ce != cudaSuccess) cudaStreamSynchronize(h->stream);     // nothing in flight while the buffers go away
    // complete the simulations in flight
    if (!rc && ce == cudaSuccess) {
        if (pools == 2) {
            rc = fwd(1);
            for (int pool = 0; pool < 2 && !rc; ++pool) mcts_expand_kernel<<<grid, 32 * kMctsWarps, 0, h->stream>>>(pool_params(pool, false));
        } else {
            rc = fwd(0);
            if (!rc) mcts_expand_kernel<<<grid, 32 * kMctsWarps, 0, h->stream>>>(pool_params(0, false));
        }
        if (!rc) ce = cudaMemcpyAsync(st, mp.stats, sizeof(st), cudaMemcpyDeviceToHost, h->stream);
        if (!rc && ce == cudaSuccess && sample_file) ce = cudaMemcpyAsync(&cursor, mp.ring_cursor, 8, cudaMemcpyDeviceToHost, h->stream);
        if (!rc && ce == cudaSuccess) ce = cudaStreamSynchronize(h->stream);
        if (!rc && ce == cudaSuccess && sample_file) rc = drain_samples(h, mp, ring_read, ring_host, pending, sample_file, cursor, nullptr);
    }
    const long long steps = done + 1;
    const double secs = elapsed();
    if (copy_stream) cudaStreamDestroy(copy_stream);
    cudaGraphExecDestroy(gexec);
    if (sample_file) fclose(sample_file);
    release();
    if (rc) return rc;
    if (ce != cudaSuccess) return fail(h, CB_ECUDA, std::string("device self-play: ") + cudaGetErrorString(ce));
    // (with the deep gates a game whose gate job is suspended completes no simulation in that step: count them)
    out->sims = mp.gate_jobs ? (long long)st[MS_SIMS] : steps * G;
    out->nn_evals = (long long)st[MS_EVALS];
    out->terminal_hits = (long long)st[MS_TERMINAL];
    out->moves = (long long)st[MS_MOVES];
    out->games_finished = (long long)st[MS_GAMES];
    out->samples = (long long)st[MS_SAMPLES];
    out->steps = steps;
    out->white_wins = (long long)st[MS_WWINS];
    out->black_wins = (long long)st[MS_BWINS];
    out->draws = (long long)st[MS_DRAWS];
    out->overflow = (long long)st[MS_OVERFLOW];
    out->gate_hits = (long long)st[MS_GATE_HITS];
    out->gate_waits = (long long)st[MS_GATE_WAITS] / std::max(1, mp.gate_warps);   // (a share that is suspended counts 1 / gate_warps)
    out->seconds = secs;
    const double tot = (double)ms_fwd + ms_tree;
    out->eval_seconds = tot > 0 ? secs * ms_fwd / tot : 0.0;
    out->tree_seconds = tot > 0 ? secs * ms_tree / tot : 0.0;
    out->mean_batch = steps ? (double)out->nn_evals / (double)steps : 0.0;
    return CB_OK;
}

int cb_selfplay_device(cb_handle* h, const cb_selfplay_config* cfg, cb_selfplay_stats* out) {
    return selfplay_device_impl(h, cfg, out, nullptr);
}

int cb_debug_device_search(cb_handle* h, const cb_selfplay_config* cfg, const cb_board* boards, int n,
                           uint16_t* moves, uint32_t* visits, double* value_sums, int32_t* counts, uint32_t* root_visits) {
    if (!h || !cfg || !boards || n <= 0 || !counts) return CB_EINVAL;
    cb_selfplay_config c = *cfg;
    c.games = n;
    c.pipeline = 3;
    c.sample_path = nullptr;
    cb_selfplay_stats st;
    const DeviceSearchIO io = {boards, moves, visits, value_sums, counts, root_visits};
    return selfplay_device_impl(h, &c, &st, &io);
}

// ------------------------------------------------------------------ SPRT (src/mcts/sprt.rs)
int cb_sprt_llr(long long wins, long long losses, long long draws, double elo0, double elo1, double* llr) {
    const long long n = wins + losses + draws;
    if (!llr || wins < 0 || losses < 0 || draws < 0) return CB_EINVAL;
    if (n < 2) return 0;                                      // compute_llr: None for fewer than 2 games
    const double nf = (double)n, w = (double)wins, d = (double)draws;
    const double s0 = 1.0 / (1.0 + pow(10.0, -elo0 / 400.0));
    const double s1 = 1.0 / (1.0 + pow(10.0, -elo1 / 400.0));
    const double s_obs = (w + 0.5 * d) / nf;
    const double var = (w + d / 4.0) / nf - s_obs * s_obs;
    if (var <= 0.0) return 0;                                 // zero variance: None
    const double var_s = var / nf;
    *llr = (s1 - s0) * (2.0 * s_obs - s0 - s1) / (2.0 * var_s);
    return 1;
}

int cb_sprt_decision(long long wins, long long losses, long long draws, double elo0, double elo1, double alpha, double beta) {
    double llr = 0.0;
    if (cb_sprt_llr(wins, losses, draws, elo0, elo1, &llr) != 1) return 0;
    const double lower = log(beta / (1.0 - alpha)), upper = log((1.0 - beta) / alpha);
    return llr >= upper ? 1 : llr <= lower ? -1 : 0;
}

// Evaluation match between two models with the tree operations on the device: every simulation step runs
// select, the forward kernel of BOTH models on the same leaves, then expand / play (each move is a fresh
// search by the mover's model).  Results are consumed in finish order; with use_sprt the test runs after every
// game and the decision taken at the trigger point stands (src/bin/evaluate_models.rs:698-800).
int cb_match_run(cb_handle* cand, cb_handle* curr, const cb_match_config* cfg, cb_match_stats* out) {
    if (!cand || !curr || !cfg || !out || cand == curr || cfg->games <= 0 || cfg->total_games <= 0 || cfg->sims_per_move <= 0)
        return CB_EINVAL;
    cb_handle* h = cand;
    if (!cand->loaded || !curr->loaded) return fail(h, CB_ESTATE, "weights not loaded");
    if (cand->device != curr->device) return fail(h, CB_EINVAL, "both models must live on the same device");
    CB_CUDA(h, cudaSetDevice(h->device));
    const int G = cfg->games;
    int rc = ensure_capacity(cand, G);
    if (!rc) rc = ensure_capacity(curr, G);
    if (rc) return rc;
    if (fused_kind(cand) != 3 || fused_kind(curr) != 3)
        return fail(h, CB_ESTATE, "matches need the resident single-launch forward on both models (128-channel nets, 16-bit mode)");
    memset(out, 0, sizeof(*out));
    if (!h->chess_tables_uploaded) {
        CB_CUDA(h, cudaMemcpyToSymbol(cbchess::d_chess_tables, &cbchess::host_tables(), sizeof(cbchess::Tables)));
        h->chess_tables_uploaded = true;
    }
    CB_CUDA(h, cudaStreamSynchronize(curr->stream));
    MctsParams mp = {};
    mp.G = G;
    mp.gn = G;
    mp.sims_per_move = cfg->sims_per_move;
    mp.max_plies = cfg->max_plies > 0 ? cfg->max_plies : 200;
    mp.material_on = cfg->material_on;
    mp.qsearch = cfg->qsearch;
    mp.tier1_light = cfg->tier1_light;
    if ((cfg->tier1_light >= 2 || cfg->qsearch == 2) && (rc = mcts_gate_params(h, mp, 0, 0, 0))) return rc;   // 5 / 0 / 3: src/bin/evaluate_models.rs:251,270
    mp.koth = cfg->koth;
    mp.c_puct = cfg->c_puct > 0 ? cfg->c_puct : 1.414;
    mp.explore_base = cfg->explore_base > 0 ? cfg->explore_base : 0.90;   // DEFAULT_EXPLORE_BASE, src/bin/evaluate_models.rs:85
    {   // fresh tree every move: sims_per_move expansions of ~35 (at most 218) children
        const long long want = (long long)cfg->sims_per_move * 64;
        mp.node_cap = (uint32_t)(want < 8192 ? 8192 : want > 65536 ? 65536 : want);
    }
    mp.seed = cfg->seed;
    mp.match_mode = 1;
    mp.total_games = cfg->total_games;
    mp.result_cap = (uint32_t)(4 * G + 64);
    std::vector<void*> owned;
    auto dalloc = [&](void** ptr, size_t bytes) -> int {
        CB_CUDA(h, cudaMalloc(ptr, bytes));
        owned.push_back(*ptr);
        return CB_OK;
    };
    auto release = [&] { for (void* q : owned) cudaFree(q); };
    if ((rc = mcts_alloc(h, mp, G, owned)) ||
        (rc = dalloc((void**)&mp.cand_white, (size_t)G)) ||
        (rc = dalloc((void**)&mp.games_started, 8)) || (rc = dalloc((void**)&mp.result_cursor, 8)) ||
        (rc = dalloc((void**)&mp.result_log, (size_t)mp.result_cap * 4))) {
        release();
        return rc;
    }
    mp.eval_boards = cand->d_boards;
    mp.eval_q = cand->d_q;
    mp.eval_idx = cand->d_move_idx;
    mp.priors = cand->d_priors;
    mp.v_final = cand->d_vfinal;
    mp.priors_b = curr->d_priors;
    mp.v_final_b = curr->d_vfinal;
    const unsigned long long started0 = (unsigned long long)(cfg->total_games < G ? cfg->total_games : G);
    cudaMemsetAsync(mp.stats, 0, MS_COUNT * 8, h->stream);
    cudaMemsetAsync(mp.result_cursor, 0, 8, h->stream);
    cudaMemcpyAsync(mp.games_started, &started0, 8, cudaMemcpyHostToDevice, h->stream);
    cudaMemsetAsync(cand->d_boards, 0, (size_t)G * sizeof(cb_board), h->stream);
    mcts_init_kernel<<<(G + 127) / 128, 128, 0, h->stream>>>(mp);
    const int grid = (G + kMctsWarps - 1) / kMctsWarps;
    PolicyOut po_a = {}, po_b = {};
    po_a.move_idx = cand->d_move_idx; po_a.move_cnt = mp.eval_cnt; po_a.priors = cand->d_priors;
    po_b.move_idx = cand->d_move_idx; po_b.move_cnt = mp.eval_cnt; po_b.priors = curr->d_priors;
    auto forwards = [&]() -> int {
        int r = launch_tower_r(cand, G, po_a, cfg->material_on, true);
        if (!r) r = launch_tower_r(curr, G, po_b, cfg->material_on, true, cand, cand->stream);
        return r;
    };
    auto fused = [&]() -> int {
        int r = forwards();
        if (r) return r;
        mcts_step_kernel<0><<<grid, 32 * kMctsWarps, 0, h->stream>>>(mp);
        return CB_OK;
    };
    mcts_select_kernel<0><<<grid, 32 * kMctsWarps, 0, h->stream>>>(mp);
    rc = fused();
    cudaError_t ce = cudaStreamSynchronize(h->stream);
    constexpr int kStepsPerGraph = 8;
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t gexec = nullptr;
    if (!rc && ce == cudaSuccess) {
        ce = cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal);
        for (int i = 0; i < kStepsPerGraph && !rc && ce == cudaSuccess; ++i) rc = fused();
        cudaError_t ce2 = cudaStreamEndCapture(h->stream, &graph);
        if (ce == cudaSuccess) ce = ce2;
        if (!rc && ce == cudaSuccess) ce = cudaGraphInstantiate(&gexec, graph, 0);
        if (graph) cudaGraphDestroy(graph);
    }
    long long steps = 1, wins = 0, losses = 0, draws = 0;
    unsigned long long read = 0;
    int decision = 0;
    double llr_at_decision = 0.0;
    bool have_llr = false;
    std::vector<int> log_host(mp.result_cap);
    const auto t_start = std::chrono::steady_clock::now();
    auto elapsed = [&] { return std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count(); };
    while (!rc && ce == cudaSuccess) {
        unsigned long long cursor = 0;
        ce = cudaMemcpyAsync(&cursor, mp.result_cursor, 8, cudaMemcpyDeviceToHost, h->stream);
        if (ce == cudaSuccess) ce = cudaMemcpyAsync(log_host.data(), mp.result_log, (size_t)mp.result_cap * 4, cudaMemcpyDeviceToHost, h->stream);
        if (ce == cudaSuccess) ce = cudaStreamSynchronize(h->stream);
        if (ce != cudaSuccess) break;
        if (cursor - read > mp.result_cap) { rc = fail(h, CB_ESTATE, "match result log overflow"); break; }
        for (; read < cursor && decision == 0; ++read) {
            const int r = log_host[read % mp.result_cap];
            if (r > 0) ++wins; else if (r < 0) ++losses; else ++draws;
            if (cfg->use_sprt) {
                decision = cb_sprt_decision(wins, losses, draws, cfg->elo0, cfg->elo1, cfg->alpha, cfg->beta);
                if (decision != 0) have_llr = cb_sprt_llr(wins, losses, draws, cfg->elo0, cfg->elo1, &llr_at_decision) == 1;
            }
        }
        if (decision != 0) break;                                    // results after the trigger do not dilute it
        if (wins + losses + draws >= cfg->total_games) break;
        if (cfg->max_seconds > 0 && elapsed() >= cfg->max_seconds) break;
        if ((ce = cudaGraphLaunch(gexec, h->stream)) != cudaSuccess) break;
        steps += kStepsPerGraph;
    }
    cudaStreamSynchronize(h->stream);
    const double secs = elapsed();
    unsigned long long sims_counted = 0;   // (with the gates at full depth a suspended game completes no simulation in a step)
    if (mp.gate_jobs) cudaMemcpy(&sims_counted, mp.stats + MS_SIMS, 8, cudaMemcpyDeviceToHost);
    if (gexec) cudaGraphExecDestroy(gexec);
    release();
    if (rc) return rc;
    if (ce != cudaSuccess) return fail(h, CB_ECUDA, std::string("match: ") + cudaGetErrorString(ce));
    out->games = wins + losses + draws;
    out->candidate_wins = wins;
    out->current_wins = losses;
    out->draws = draws;
    out->decision = decision;
    if (decision != 0 && have_llr) { out->llr = llr_at_decision; out->llr_valid = 1; }
    else out->llr_valid = cb_sprt_llr(wins, losses, draws, cfg->elo0, cfg->elo1, &out->llr) == 1;
    out->sims = mp.gate_jobs ? (long long)sims_counted : steps * G;
    out->seconds = secs;
    return CB_OK;
}

int cb_debug_backbone(cb_handle* h, int B, float* out) {
    if (!h) return CB_EINVAL;
    if (fused_kind(h) == 3 || fused_kind(h) == 4) {
        // the resident tower never writes the backbone to global memory: rerun it on the boards of the
        // last call with the debug dump enabled
        if (!out || B <= 0 || B > h->cap) return CB_EINVAL;
        CB_CUDA(h, cudaSetDevice(h->device));
        const size_t total = (size_t)B * 64 * h->hid;
        int rc = ensure_scratch(h, total);
        if (rc) return rc;
        h->d_dbg_backbone = h->d_scratch;
        rc = fused_kind(h) == 4 ? launch_tower_r2_staged(h, B) : launch_tower_r_staged(h, B);
        h->d_dbg_backbone = nullptr;
        if (rc) return rc;
        CB_CUDA(h, cudaMemcpyAsync(out, h->d_scratch, total * 4, cudaMemcpyDeviceToHost, h->stream));
        CB_CUDA(h, cudaStreamSynchronize(h->stream));
        return CB_OK;
    }
    return cb_debug_act(h, B, backbone_index(h->blocks), out);
}

int cb_debug_act(cb_handle* h, int B, int buf, float* out) {
    if (!h || !out || B <= 0 || B > h->cap || buf < 0 || buf > 2) return CB_EINVAL;
    CB_CUDA(h, cudaSetDevice(h->device));
    const int total = B * 64 * h->hid;
    int rc = ensure_scratch(h, (size_t)total);
    if (rc) return rc;
    const void* src = h->d_act[buf];
    const int grid = (total + 255) / 256;
    // buffer 1 ends a forward holding p_conv's output, always in pair layout; the others follow the tower
    const int lay = 1;
    if (h->dtype == CB_DTYPE_FP32)
        act_to_nchw_f32_kernel<float><<<grid, 256, 0, h->stream>>>((const float*)src, h->hid, total, 0, h->d_scratch);
    else if (h->dtype == CB_DTYPE_BF16)
        act_to_nchw_f32_kernel<__nv_bfloat16><<<grid, 256, 0, h->stream>>>((const __nv_bfloat16*)src, h->hid, total, lay, h->d_scratch);
    else
        act_to_nchw_f32_kernel<__half><<<grid, 256, 0, h->stream>>>((const __half*)src, h->hid, total, lay, h->d_scratch);
    CB_CUDA(h, cudaGetLastError());
    CB_CUDA(h, cudaMemcpyAsync(out, h->d_scratch, (size_t)total * 4, cudaMemcpyDeviceToHost, h->stream));
    CB_CUDA(h, cudaStreamSynchronize(h->stream));
    return CB_OK;
}

}  // extern "C"
