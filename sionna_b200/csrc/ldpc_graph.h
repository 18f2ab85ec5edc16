// ldpc_graph.h -- decoding-graph handle shared by ldpc_bp.cu (generic kernel, graph construction) and
// ldpc_bp_qc.cu (quasi-cyclic fast path), plus small device helpers used by both kernels.
#pragma once
#include <stdint.h>
#include <vector>
#include "sb_common.h"

struct sb_ldpc_graph {
    int C = 0, N = 0, E = 0, Lc = 0, Lv = 0, n_in = 0, n_out = 0, n_sub = 1, n_active = 0;
    bool flooding = true;
    bool ref_order = false;   // node sums follow the reference's list orders (sb_ldpc_graph_create_ordered)
    std::vector<int> cn_off, cn_cnt, vn_off, vn_cnt, in_idx, out_pos, slot_of_edge, sched, cn_order, vn_order;
    std::vector<uint32_t> vn_slot;
    std::vector<int> h_cn, h_vn;   // the caller's edge list (reference VN order), kept for sb_ldpc_graph_set_qc
    // device copies (lazy)
    bool uploaded = false;
    int device = -1;
    int *d_cn_off = nullptr, *d_cn_cnt = nullptr, *d_vn_off = nullptr, *d_vn_cnt = nullptr, *d_in_idx = nullptr,
        *d_out_pos = nullptr, *d_slot_of_edge = nullptr, *d_sched = nullptr;
    uint16_t* d_vn_slot16 = nullptr;
    uint32_t* d_vn_slot32 = nullptr;
    int smem_optin = 0, num_sms = 0;
    // ---- quasi-cyclic description (optional, set by sb_ldpc_graph_set_qc; used by ldpc_bp_qc.cu) ----------
    bool qc = false;
    int qc_Z = 0, qc_rows = 0, qc_cols = 0, qc_nnz = 0, qc_max_row_deg = 0, qc_max_col_deg = 0;
    std::vector<int> qc_row_info;   // int4 per base row (processing order): {first base entry, deg, zrow, fused col | s << 16 or -1}
    std::vector<int> qc_col_info;   // int4 per base col (processing order): {first col-edge, deg, zcol, c*Z}
    std::vector<int> qc_col_edge;   // int2 per (col, entry), ascending base row: {be*Z*4, s*4 | (zrow*4) << 16}
    std::vector<int> qc_in_idx, qc_out_pos, qc_slot_of_edge;   // natural VN order / reference edge order
    std::vector<int> qc_row_edge;   // int2 per base entry (processing order): {column * Z, shift} for the syndrome pass
    std::vector<int> qc_row_cls_end, qc_col_cls_end;   // class boundaries in processing order (ldpc_bp_qc.cu)
    int *d_qc_row_info = nullptr, *d_qc_col_info = nullptr, *d_qc_col_edge = nullptr, *d_qc_in_idx = nullptr,
        *d_qc_out_pos = nullptr, *d_qc_slot_of_edge = nullptr, *d_qc_row_edge = nullptr;
    bool qc_uploaded = false;
};

// B200 (sm_100) opt-in shared memory per block; used for planning when no device is present.
static const int kSmemOptinB200 = 232448;

// ldpc_bp_qc.cu: runs the QC kernel if the graph / call qualifies; *handled tells the dispatcher.
int sb_qc_try_decode(sb_ldpc_graph* g, const float* d_llr, int64_t batch, int32_t num_iter, int32_t cn_rule,
                     int32_t vn_rule, float offset, float llr_max, int32_t hard_out, const float* d_state_in,
                     float* d_state_out, float* d_out, cudaStream_t stream, bool* handled, int32_t early = 0,
                     int32_t* d_iters = nullptr);
void sb_qc_free_device(sb_ldpc_graph* g);

#if defined(__CUDACC__)
__device__ __forceinline__ float clipf(float x, float c) { return fminf(fmaxf(x, -c), c); }

// ---- mbarrier / TMA bulk-copy helpers (cp.async.bulk, 1-D) ------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

#endif
