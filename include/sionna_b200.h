/* sionna_b200.h -- C-ABI of libsionna_b200.so (B200 / sm_100a kernels for the Sionna PHY hot path).
 *
 * The reference (NVlabs/sionna v1.2.1) has no FFI or plugin registry: its boundary is the Python
 * `Block.__call__ -> build(shapes) -> call()` protocol (/root/reference/src/sionna/phy/block.py:82-155)
 * and every op below is a chain of TensorFlow calls inside a `call()` method. Each entry point here
 * replaces the body of one such `call()`; the comment above it names the reference method
 * (file:line under /root/reference/src/sionna/phy/). The Python host layer (`sionna_b200/phy/...`)
 * keeps the reference's class names, constructor arguments, shapes and sign conventions and calls
 * these functions through ctypes on the current torch CUDA stream.
 *
 * Conventions
 *   - plain C types only: device pointers, sizes, a `void* stream` (cudaStream_t; NULL = default stream).
 *   - every function returns 0 on success or a negative SB_E* code; `sb_last_error()` returns a
 *     thread-local human-readable message for the last failure on the calling thread.
 *   - outputs and workspaces are allocated by the caller; no function synchronises the stream or
 *     allocates device memory, except `*_create` (device copies of index tables owned by the handle,
 *     released by `*_destroy`).
 *   - all pointers named `d_*` are device pointers, `h_*` host pointers.
 *   - real tensors are fp32, complex tensors interleaved (re, im) fp32 ("single" precision).
 */
#ifndef SIONNA_B200_H
#define SIONNA_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SB_OK 0
#define SB_EINVAL (-1)   /* bad argument */
#define SB_ECUDA (-2)    /* CUDA runtime error (message holds cudaGetErrorString) */
#define SB_ENOMEM (-3)   /* workspace too small / allocation failed */
#define SB_EUNSUPPORTED (-4)

const char* sb_last_error(void);
/* Library / device sanity: returns SB_OK and fills sm count, compute capability major/minor, and the
 * opt-in shared memory per block of the CURRENT device. */
int sb_device_info(int* sm_count, int* cc_major, int* cc_minor, int* smem_optin_bytes);
int sb_version(void);

/* ------------------------------------------------------------------------------------------------
 * LDPC belief propagation
 * replaces LDPCBPDecoder.__init__ graph set-up   fec/ldpc/decoding.py:277-345
 *          LDPCBPDecoder.call / _bp_iter          fec/ldpc/decoding.py:416-637
 *          vn_update_sum / cn_update_*            fec/ldpc/decoding.py:681-1166
 *          LDPC5GDecoder.call rate recovery       fec/ldpc/decoding.py:1427-1536
 * ---------------------------------------------------------------------------------------------- */
typedef struct sb_ldpc_graph sb_ldpc_graph;

enum { SB_CN_BOXPLUS_PHI = 0, SB_CN_BOXPLUS = 1, SB_CN_MINSUM = 2, SB_CN_OFFSET_MINSUM = 3, SB_CN_IDENTITY = 4 };
enum { SB_VN_SUM = 0, SB_VN_IDENTITY = 1 };

/* Build a decoding graph.
 *   h_cn_of_edge, h_vn_of_edge [num_edges]: the reference's edge list in its VN order, i.e.
 *       `self._cn_idx`, `self._vn_idx` after `idx = np.argsort(vn_idx)` (decoding.py:282-288). Edge e of
 *       this list is row e of the `msg_v2c` decoder state ([num_edges, batch], decoding.py:575-579).
 *   h_in_map [num_vn] or NULL: where VN v takes its channel logit from: >= 0 index into the caller's
 *       input row (length n_in); -1 punctured (logit 0, decoding.py:1444-1458); -2 filler / shortened
 *       (logit -llr_max, decoding.py:1472-1475). NULL = identity (n_in must equal num_vn).
 *   h_out_vn [n_out] or NULL: VN whose estimate is written to output column j (systematic slice,
 *       filler / 2Z removal and output interleaver of decoding.py:1486-1536 folded into one gather).
 *       NULL = identity (n_out must equal num_vn).
 *   h_schedule [n_sub * n_active] or NULL: CN indices updated in each sub-iteration
 *       (`cn_schedule`, decoding.py:253-271, 464-497). NULL = flooding.
 */
int sb_ldpc_graph_create(sb_ldpc_graph** out, int32_t num_cn, int32_t num_vn, int32_t num_edges,
                         const int32_t* h_cn_of_edge, const int32_t* h_vn_of_edge,
                         const int32_t* h_in_map, int32_t n_in,
                         const int32_t* h_out_vn, int32_t n_out,
                         const int32_t* h_schedule, int32_t n_sub, int32_t n_active);
void sb_ldpc_graph_destroy(sb_ldpc_graph* g);
/* 1 if one codeword's messages + channel LLRs fit in one SM's shared memory (the on-chip path),
 * 0 if the decoder will keep messages in an L2-resident global workspace. */
int sb_ldpc_graph_on_chip(const sb_ldpc_graph* g);
/* Bytes of device workspace `sb_ldpc_decode` needs for this graph (0 on the on-chip path). */
size_t sb_ldpc_workspace_bytes(const sb_ldpc_graph* g);

/* Decode `batch` codewords.
 *   d_llr    [batch, n_in]   channel logits log p(1)/p(0) (decoding.py:159-164); clipped to +-llr_max and
 *                            negated internally (decoding.py:552-565).
 *   d_out    [batch, n_out]  hard_out != 0: 1.0f where the internal LLR <= 0 else 0.0f (decoding.py:622-624);
 *                            hard_out == 0: soft logits (decoding.py:626).
 *   d_state_in / d_state_out [batch, num_edges] or NULL: the `msg_v2c` decoder state, one row per
 *       codeword, columns in the reference's edge order (the host layer transposes to/from the
 *       reference's [num_edges, batch] layout, decoding.py:569-573, 633-637).
 *   offset: only for SB_CN_OFFSET_MINSUM (decoding.py:755).
 */
int sb_ldpc_decode(const sb_ldpc_graph* g, const float* d_llr, int64_t batch, int32_t num_iter,
                   int32_t cn_rule, int32_t vn_rule, float offset, float llr_max, int32_t hard_out,
                   const float* d_state_in, float* d_state_out, float* d_out,
                   void* d_workspace, size_t workspace_bytes, void* stream);
/* Number of kernels the last sb_ldpc_decode on this thread launched (for bench.py's gpu_launches). */
int sb_ldpc_last_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* SIONNA_B200_H */
