/* graingnn_b200.h — C ABI of libgraingnn_b200.so (hand-written CUDA for sm_100a).
 *
 * Drop-in boundary for the GrainGNN rollout message-passing path.  Every entry point
 * replaces a piece of the reference's Python hot path; the citation on each prototype is the reference
 * file:line whose arithmetic it takes over (paths relative to the reference repository root).
 *
 * Conventions
 *  - All pointers are DEVICE pointers unless the name ends in `_host`.  No entry point allocates, frees or
 *    retains memory; scratch is passed in by the caller (sizes from the *_workspace_bytes helpers).
 *  - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).  Launches are asynchronous.
 *  - Return value: 0 on success, a negative GG_E* code for a rejected argument, or a positive cudaError_t.
 *    gg_error_string() renders either.  Nothing falls back to the CPU.
 *  - Row-major everywhere; `ld*` are row strides in elements.  Node features are fp32, the reference's edge
 *    index is int64 [2,E] (row 0 = source, row 1 = target), CSR arrays are int32.
 *  - Hidden width C must be a multiple of 32 with C <= 128 (the reference uses 96, parameters.py:18-21).
 *  - "Gate block": G*C contiguous columns, gate-major (gate g occupies columns [g*C,(g+1)*C)).
 */
#ifndef GRAINGNN_B200_H
#define GRAINGNN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GG_EINVAL   (-1)  /* bad argument (null pointer, negative size, unsupported width)              */
#define GG_ERANGE   (-2)  /* an edge endpoint outside [0, N) was found (reported by gg_csr_build)      */
#define GG_EALIGN   (-3)  /* a pointer / leading dimension violates a documented 16-byte alignment     */
#define GG_ENOSPC   (-4)  /* workspace too small                                                       */
#define GG_EARCH    (-5)  /* device is not sm_100 (tcgen05 paths only)                                 */

const char* gg_error_string(int code);
int gg_version(void);                 /* 100 * major + minor                                        */
int gg_device_is_sm100(void);         /* 1 when the current device has compute capability 10.x      */

/* ------------------------------------------------------------------------------------------------
 * (a) dst-sorted CSR builder.  Replaces the per-call COO handling of PyG `MessagePassing.propagate`
 *     (called at periodGATconv.py:174) and torch-scatter's atomics: edges are stably sorted by target,
 *     so row i of the CSR lists the in-edges of node i in ORIGINAL edge order (bit-exact vs
 *     numpy.argsort(kind="stable")).
 *       rowptr[n_dst+1], col[E] = source of the e-th sorted edge, perm[E] = its original edge id.
 *     status (device int32[1]) receives 0, or GG_ERANGE if any endpoint is out of range.
 * ---------------------------------------------------------------------------------------------- */
size_t gg_csr_workspace_bytes(int64_t n_edges, int32_t n_dst);
int gg_csr_build(const int64_t* edge_index, int64_t n_edges, int32_t n_src, int32_t n_dst,
                 int32_t* rowptr, int32_t* col, int32_t* perm, int32_t* status,
                 void* workspace, size_t workspace_bytes, void* stream);

/* out[i] = src[perm[i]] (fp32) — puts a per-edge attribute into CSR order. */
int gg_permute_f32(const float* src, const int32_t* perm, float* out, int64_t n, void* stream);

/* ------------------------------------------------------------------------------------------------
 * (a11) wrapped 2-D edge length, test.py:562-575:  d = x_src[s,:2] - x_dst[t,:2];  d += (d<-.5) - (d>.5);
 *       out = sqrt(dx^2+dy^2).  `out` is in original edge order; if out_csr != NULL it also receives the
 *       same values in CSR order (out_csr[i] = out[perm[i]]), so kernel (b) needs no indirection.
 * ---------------------------------------------------------------------------------------------- */
int gg_edge_length(const float* x_src, int32_t ld_src, const float* x_dst, int32_t ld_dst,
                   const int64_t* edge_index, int64_t n_edges,
                   const int32_t* perm /* nullable */, float* out, float* out_csr /* nullable */, void* stream);

/* ------------------------------------------------------------------------------------------------
 * (c) node projections.  out[m, n] = sum_k [A1|A2][m,k] * W[n,k] + bias[n]
 *     Takes over the per-EDGE lin_key / lin_value / lin_query GEMMs of periodGATconv.py:216-218 by
 *     re-associating them per NODE: W is the row-concatenation of the PyG [out,in] weights
 *     of every role x gate a node type plays, A = cat([X, h]) (heteropgclstm.py:112) given as two pieces so
 *     the concat is never materialised (A2 == NULL means h == 0, the encoder case of models.py:237-238).
 *     mode 0: fp32 SIMT FMA.   mode 1: tcgen05 3xTF32 tensor-core path (needs gg_tc_* packing, see below).
 * ---------------------------------------------------------------------------------------------- */
int gg_node_proj(const float* A1, int32_t lda1, int32_t K1,
                 const float* A2, int32_t lda2, int32_t K2,
                 const float* W, int32_t ldw, const float* bias /* nullable */,
                 float* out, int32_t ldo, int32_t M, int32_t N, void* stream);

/* Tensor-core variant of gg_node_proj: tcgen05.mma kind::tf32 with TMEM accumulators, TMA-staged operands.
 * fp32-grade accuracy through the 3xTF32 split  A W^T = A_lo W_hi^T + A_hi W_lo^T + A_hi W_hi^T.
 *   gg_tc_supported : 1 when the device is sm_100 and the driver exposes cuTensorMapEncodeTiled.
 *   gg_split_tf32   : builds A_hi, A_lo [M, Kp] (Kp % 32 == 0) from X (K1 columns, placed in [0, K1p32)) and h (K2 columns,
 *                     placed from K1p32); every value is rounded to TF32 with cvt.rna, a = hi + lo up to 2^-22 |a|.
 *   gg_node_proj_tc : out[M, N] = A W^T + bias; W_hi / W_lo are [N, Kp] with the same K layout. n_sms <= 0: all SMs.
 *                     k_first (1..32): columns [k_first, 32) of A are zero (feature padding) - their MMAs are not issued.
 */
int gg_tc_supported(void);
int gg_split_tf32(const float* X, int32_t ldx, int32_t K1, const float* H /* nullable */, int32_t ldh, int32_t K2,
                  int32_t M, float* A_hi, float* A_lo, int32_t Kp, int32_t K1p32, void* stream);
int gg_node_proj_tc(const float* A_hi, const float* A_lo, int32_t Kp, int32_t k_first, const float* W_hi, const float* W_lo,
                    int32_t N, const float* bias /* nullable */, float* out, int32_t ldo, int32_t M, int32_t n_sms,
                    void* stream);
/* The same projection with the TF32 split fused into the kernel (no gg_split_tf32 pass, no A_hi / A_lo in HBM): the fp32 chunk of
 * [X zero-extended to 32 columns | H] is the hi operand as it stands (the tensor core reads the upper 19 bits), converter warps
 * write a - trunc_tf32(a) next to it.  X [M, K1] (4 <= K1 <= 32, K1 % 4 == 0, row stride ldx), H [M, K2] (K2 % 32 == 0, may be
 * NULL with K2 == 0); W_hi / W_lo [N, 32 + K2].  Replaces lin_key / lin_value / lin_query on the gathered rows
 * (periodGATconv.py:216-218) like gg_node_proj. */
int gg_node_proj_fused(const float* X, int32_t ldx, int32_t K1, const float* H /* nullable */, int32_t ldh, int32_t K2,
                       const float* W_hi, const float* W_lo, int32_t N, const float* bias /* nullable */, float* out, int32_t ldo,
                       int32_t M, int32_t n_sms, void* stream);

/* ------------------------------------------------------------------------------------------------
 * (b) fused periodic-attention gather.  One launch = one edge type, all G gates of one cell.
 *     Per target node i and gate g (PeriodConv.message, periodGATconv.py:204-236, applied to per-node
 *     projections):
 *        w_e   = (r<-.5) - (r>.5),  r = p_j - p_i                       (:209-210, periodic wrap)
 *        s_e   = (Q_i . K_j + QX_i[0:3] . w_e + QX_i[3] * a_e) / sqrt(C) (:216-226; the p_i term and every
 *                 per-target constant cancel in the softmax)
 *        al_e  = exp(s_e - max_i) / (sum_i + 1e-16)                      (:227, PyG utils.softmax)
 *        agg_i = sum_e al_e * relu(V_j + Wv3 (w_e - p_i))                (:211,:218 — lin_l2 is applied after
 *                 aggregation by gg_gate_update)          ea_i = sum_e al_e * a_e     (:222,:233)
 *     With weighted == 0 (periodconv.py:235) al_e = 1.
 *     Layout: P_src row = [... K block @k_off (G*C) ... V block @v_off (G*C) ...];
 *             P_dst row = [... Q block @q_off (G*C) ... QX block @qx_off (G*4: Wk3^T q (3), We . q (1)) ...];
 *             Wv3 is [G*C][4] (x,y,z weights of lin_value, 4th = 0); pos_* point at column 0 (x,y,z) of the node
 *             features; eattr_csr is a_e in CSR order.  Outputs: agg [n_dst, ld_agg] gate block, ea [n_dst, G].
 *     One warp per target node (8 lanes x C/8 channels per gate, 128-bit loads), no atomics; edges of a row
 *     are accumulated in CSR (= original) order, matching the sequential index_add_ of the CPU reference.
 *     Fast path (sm_100, G <= 4): pass the flat work list of gg_csr_items (items, item_ptr) and the per-edge wrap codes of
 *     gg_edge_wrap (wrap_csr, computed from the CURRENT positions); the kernel then stages K|V rows with cp.async.bulk and
 *     never reads source positions.  With items == NULL or wrap_csr == NULL the 128-bit-load kernel runs and derives the
 *     wraps from pos_src / pos_dst itself.
 *     raw_k = 16 (fast path only, cells without hidden state, weighted): raw-score mode.  q_i . (Wk x_j) = x_j . (Wk^T q_i), so
 *     no key rows exist: P_src holds [raw features of the source (16 floats, zero padded) @k_off | V block @v_off = k_off+16]
 *     and P_dst holds Q'_i = [Wk[:, :F]^T q_i (15) | We . q_i] per gate (G*16 floats @q_off; qx_off unused).  raw_k = 0: classic.
 *     All row pointers / leading dimensions / offsets must be multiples of 4 floats (GG_EALIGN otherwise).
 * ---------------------------------------------------------------------------------------------- */
int gg_pgat_gather(const float* P_src, int32_t ld_src, int32_t k_off, int32_t v_off,
                   const float* P_dst, int32_t ld_dst, int32_t q_off, int32_t qx_off,
                   const float* pos_src, int32_t ld_pos_src, const float* pos_dst, int32_t ld_pos_dst,
                   const int32_t* rowptr, const int32_t* col, const float* eattr_csr,
                   const int32_t* items /* nullable */, const int32_t* item_ptr /* nullable */, const int32_t* wrap_csr /* nullable */,
                   int32_t raw_k,
                   const float* Wv3, int32_t n_dst, int32_t G, int32_t C, int32_t weighted,
                   float* agg, int32_t ld_agg, float* ea, void* stream);

/* Work list of the fast gather path: node i contributes max(1, ceil(deg_i / dcap)) items of <= dcap consecutive in-edges,
 * items[k] = {node, first CSR edge, count | first << 8 | last << 9, 0} (int32 x 4, 16-byte aligned), item_ptr[n_dst + 1] =
 * exclusive scan of the per-node item counts (item_ptr[n_dst] = number of items <= n_dst + E / dcap).
 * dcap must be gg_gather_dcap().  workspace: gg_csr_workspace_bytes(0, n_dst). */
int gg_gather_dcap(void);
int gg_csr_items(const int32_t* rowptr, int32_t n_dst, int32_t dcap, int32_t* item_ptr, int32_t* items,
                 void* workspace, size_t workspace_bytes, void* stream);

/* Warp-specialised form of (b) (gather_tiled.cu; sm_100, weighted, G <= 4): same arithmetic and outputs as gg_pgat_gather on
 * these layouts: raw_k = 0: P_src row holds K|V adjacent at k_off, P_dst row Q|QX|x,y,z,0 adjacent at q_off; raw_k = 16
 * (cells without hidden state) or 32 + C (with hidden state): P_src row holds [raw input (raw_k floats) | V] at k_off and
 * P_dst row Q' (raw_k floats per gate: Wk^T q laid out like the raw input, We . q in slot 15 / 31, and the target's x, y, z in
 * the three slots in front of it of the FIRST gate — the source input is zero there) at q_off.  Driven by the TILE
 * index instead of the item list: the CSR edge array is cut into tiles of ecap = gg_gather_tile_ecap(G, C, raw_k) consecutive
 * in-edges (0 = this shape is not supported), producer warps stage each tile with cp.async.bulk, consumer warps compute.
 *   gg_csr_compact: nz[NZ] = targets with in-edges (ascending), nzptr[NZ + 1] = where their rows start (nzptr[NZ] = E),
 *                   nz_count[0] = NZ;  nz / nzptr sized n_dst / n_dst + 1, scratch int32[n_dst + 1],
 *                   workspace gg_csr_workspace_bytes(0, n_dst).
 *   gg_csr_tiles:   targets are grouped into units by the CSR position of their first in-edge (unit k: [4 ecap k, 4 ecap (k+1)))
 *                   and each unit's edges are cut into tiles of <= ecap edges; tiles[f] = {first CSR edge, edges, index in nz of
 *                   the target owning the first edge, ... the last edge} (int32 x 4, 16-byte aligned), ordered by CTA (unit k
 *                   belongs to CTA k mod n_ctas), cta_ptr[n_ctas + 1] = where each CTA's tiles start.  tiles holds
 *                   gg_csr_tiles_capacity(E, ecap, n_ctas) entries, scratch gg_csr_tiles_scratch_ints(...) int32.
 *                   n_ctas must be gg_gather_ctas() (one persistent CTA per SM) and is the grid gg_pgat_gather_tiled launches.
 * Replaces PyG propagate + utils.softmax + scatter-add (periodGATconv.py:174, :204-236) like gg_pgat_gather. */
int gg_gather_tile_ecap(int32_t G, int32_t C, int32_t raw_k);
int gg_csr_compact(const int32_t* rowptr, int32_t n_dst, int32_t* nz, int32_t* nzptr, int32_t* nz_count,
                   int32_t* scratch, void* workspace, size_t workspace_bytes, void* stream);
int gg_gather_ctas(void);
int64_t gg_csr_tiles_capacity(int64_t n_edges, int32_t ecap, int32_t n_ctas);
size_t gg_csr_tiles_scratch_ints(int64_t n_edges, int32_t ecap, int32_t n_ctas);
int gg_csr_tiles(const int32_t* nzptr, const int32_t* nz_count, int64_t n_edges, int32_t ecap, int32_t n_ctas,
                 int32_t* tiles, int32_t* cta_ptr, int32_t* scratch, void* stream);
int gg_pgat_gather_tiled(const float* P_src, int32_t ld_src, int32_t k_off,
                         const float* P_dst, int32_t ld_dst, int32_t q_off,
                         const int32_t* rowptr, const int32_t* col, const float* eattr_csr, const int32_t* wrap_csr,
                         const int32_t* nz, const int32_t* nzptr, const int32_t* tiles, const int32_t* cta_ptr,
                         int32_t n_ctas, int32_t ecap, int64_t n_edges,
                         int32_t raw_k, const float* Wv3, int32_t n_dst, int32_t G, int32_t C,
                         float* agg, int32_t ld_agg, float* ea, void* stream);

/* The same kernel for up to 3 edge types of ONE cell in one launch (same C, G, raw_k, hence one stage layout): every persistent
 * CTA walks its tiles of segment 0, then segment 1, ... without draining the stage ring in between — one launch ramp and one
 * tail per cell instead of one per edge type (the 12 PeriodConv calls of a HeteroPGCLSTM cell, heteropgclstm.py:111-142, in a
 * single launch).  Fields as the arguments of gg_pgat_gather_tiled; segments with n_dst == 0 are skipped. */
typedef struct gg_gather_segment {
    const float* P_src; int32_t ld_src, k_off;
    const float* P_dst; int32_t ld_dst, q_off;
    const int32_t* rowptr; const int32_t* col; const float* eattr_csr; const int32_t* wrap_csr;
    const int32_t* nz; const int32_t* nzptr; const int32_t* tiles; const int32_t* cta_ptr;
    int64_t n_edges;
    const float* Wv3;
    int32_t n_dst;
    float* agg; int32_t ld_agg;
    float* ea;
} gg_gather_segment;
int gg_pgat_gather_tiled_multi(const gg_gather_segment* segs, int32_t n_seg, int32_t n_ctas, int32_t ecap,
                               int32_t raw_k, int32_t G, int32_t C, void* stream);

/* Periodic wrap of every edge (periodGATconv.py:209-210), CSR order: r = p_src - p_dst per coordinate, code 1 where
 * r < -0.5 (+1), 2 where r > 0.5 (-1), else 0; wrap_csr[e] = cx | cy << 2 | cz << 4.  pos_* point at column 0 (x,y,z). */
int gg_edge_wrap(const float* pos_src, int32_t ld_pos_src, const float* pos_dst, int32_t ld_pos_dst,
                 const int32_t* rowptr, const int32_t* col, int32_t n_dst, int32_t* wrap_csr, void* stream);

/* gg_edge_wrap and gg_edge_length in ONE pass over the CSR rows of an edge type (the rollout step's per-step rebuild,
 * test.py:562-575 + periodGATconv.py:209-210): wrap_csr and eattr_csr in CSR order, eattr[perm[k]] in original edge order.
 * Lengths are bit-identical to gg_edge_length. */
int gg_edge_refresh(const float* pos_src, int32_t ld_pos_src, const float* pos_dst, int32_t ld_pos_dst,
                    const int32_t* rowptr, const int32_t* col, const int32_t* perm, int32_t n_dst,
                    int32_t* wrap_csr, float* eattr_csr, float* eattr, void* stream);

/* ------------------------------------------------------------------------------------------------
 * (c) post-aggregation gate GEMM fused with the LSTM update.  Per node m of one node type, gate g:
 *        pre_g = sum_t [ W2_{t,g} agg_t[m,g] + We_{t,g} ea_t[m,g] + b2_{t,g} * cnt_t(m) ]      (lin_l2, lin_edge;
 *                 periodGATconv.py:218,:222,:231-235 moved after the sum; cnt = [deg>0] weighted, deg unweighted)
 *              + Wskip_g [X|h][m] + btot_g                     (lin_skip :186,192 summed over edge types
 *                 + HeteroConv sum + gate bias b_g, heteropgclstm.py:113-116)
 *     mode GG_GATE_RAW : out_h = pre (G*C wide)                      — a bare PeriodConv.forward
 *     mode GG_GATE_RELU: out_h = relu(pre), G == 1                   — HeteroPGC, heteropgclstm.py:243-251
 *     mode GG_GATE_LSTM: gates (i,f,c,o): c' = s(f) c + s(i) tanh(c~); h' = s(o) tanh(c')   (:133,:145)
 *     mode GG_GATE_LSTM0: gates (i,c,o) with c == 0 (encoder, models.py:237-238): c' = s(i) tanh(c~)
 * ---------------------------------------------------------------------------------------------- */
enum { GG_GATE_RAW = 0, GG_GATE_RELU = 1, GG_GATE_LSTM = 2, GG_GATE_LSTM0 = 3 };

typedef struct gg_agg_input {
    const float*   agg;      /* [M, ld_agg] gate block written by gg_pgat_gather                 */
    int32_t        ld_agg;
    const float*   ea;       /* [M, G]                                                            */
    const int32_t* rowptr;   /* CSR rowptr of that edge type (for cnt)                            */
    const float*   W2;       /* [G][C][C]  lin_l2.weight per gate, PyG [out,in]                   */
    const float*   We;       /* [G][C]     lin_edge.weight[:,0]                                   */
    const float*   b2;       /* [G][C]     lin_l2.bias                                            */
    int32_t        weighted; /* 1: attention (cnt = deg>0), 0: periodconv sum variant (cnt = deg) */
} gg_agg_input;

int gg_gate_update(const gg_agg_input* inputs_host, int32_t n_inputs,
                   const float* X, int32_t ldx, int32_t K1, const float* H /* nullable */, int32_t ldh,
                   const float* Wskip, int32_t ldw, const float* btot,
                   const float* c_in /* nullable */, float* out_h, float* out_c /* nullable */,
                   int32_t M, int32_t G, int32_t C, int32_t mode, void* stream);

/* Tensor-core variant of gg_gate_update (tcgen05 kind::tf32 in the A-from-TMEM form, 3xTF32, accumulators in TMEM).
 *   Every A operand is plain fp32 and is split into TF32 hi / lo INSIDE the kernel (converter warps -> tensor memory):
 *   inputs[e].agg [M, G*C] as written by gg_pgat_gather (agg_lo == NULL), X [M, K1] with 4 <= K1 <= 32 - (2G+3) (the TMA
 *   box zero-extends it to 32 columns), H [M, C] or NULL (encoder).  Modes: LSTM (G=4), LSTM0 (G=3), RAW / RELU (G=1).
 *   W_hi/W_lo: TF32 split of Wall [G*C, Ktot], row g*C+n, Ktot = n_inputs*C + 32 (+ C with H), K layout
 *     [lin_l2 of input 0 (C) | lin_l2 of input 1 (C) | feature chunk (32) | summed lin_skip on h (C)].
 *   Feature chunk of gate g: summed lin_skip on X in columns [0, K1); the rank-1 terms ride in its tail, RB = 32-(2G+3):
 *     col RB+g: lin_edge of input 0 (gate g only) | RB+G: lin_l2 bias of input 0 | RB+G+1+g: lin_edge of input 1 |
 *     RB+2G+1: lin_l2 bias of input 1 | 31: summed lin_skip bias + gate bias;
 *   the kernel supplies [ea_0[0..G) | cnt_0 | ea_1[0..G) | cnt_1 | 1] there per node (cnt = [deg>0] weighted, deg otherwise),
 *   so inputs[e].W2 / We / b2 are not read (they live in Wall); ea, rowptr, weighted are.  1 <= n_inputs <= 2.
 *   X, H, agg, out rows must be 16-byte aligned with leading dimensions that are multiples of 4 floats; out_h is [M, C]
 *   (LSTM modes; out_c [M, C] nullable) or [M, G*C] (RAW / RELU), both dense.
 */
int gg_gate_update_tc(const gg_agg_input* inputs_host, int32_t n_inputs,
                      const float* X, int32_t ldx, int32_t K1, const float* H /* nullable */, int32_t ldh,
                      const float* W_hi, const float* W_lo, int32_t Ktot,
                      const float* c_in /* nullable */, float* out_h, float* out_c /* nullable */,
                      int32_t M, int32_t G, int32_t C, int32_t mode, int32_t n_sms, void* stream);

/* ------------------------------------------------------------------------------------------------
 * (a7) SAGEConv mean aggregation for HeteroGCLSTM / HeteroGC (heterogclstm.py:52-54, PyG SAGEConv aggr='mean'):
 *      out[i, 0:width] = mean_{e -> i} src[col[e], 0:width], 0 for rows without in-edges.  width % 4 == 0.
 *      The lin_l / lin_r GEMMs and the gate math that follow run in gg_gate_update on [mean_e0 | mean_e1 | X] and h.
 * ---------------------------------------------------------------------------------------------- */
int gg_segment_mean(const float* src, int32_t ld_src, int32_t width, const int32_t* rowptr, const int32_t* col,
                    int32_t n_dst, float* out, int32_t ld_out, void* stream);

/* ------------------------------------------------------------------------------------------------
 * (d) heads.
 *  gg_node_head: y = W h + b with n_out <= 4 rows (models.py:433), act[j]: 0 none, 1 tanh, 2 relu (:443,:450-452);
 *      if area_out: area_out[m] = tanh(raw y[m,0]) / area_scale + area_in[m*ld_area]   (:445)
 *  gg_edge_head: pair = [h[src], h[dst], a]  (:602) -> edge_event = lin2(pair) (:607), edge = tanh(lin1(pair)) (:609)
 *      fused with the joint-pair gather; outputs in ORIGINAL edge order (Cmodel.update indexes them by edge id,
 *      models.py:626-628).  W1 [2, 2C+1], W2 [1, 2C+1] are the PyG/torch [out,in] weights.
 * ---------------------------------------------------------------------------------------------- */
int gg_node_head(const float* h, int32_t ldh, int32_t C, const float* W, const float* b, int32_t n_out,
                 const int32_t* act_host, float* y, int32_t ldy,
                 const float* area_in, int32_t ld_area, float area_scale, float* area_out,
                 int32_t M, void* stream);
int gg_edge_head(const float* h, int32_t ldh, int32_t C, const int64_t* edge_index, int64_t n_edges,
                 const float* eattr, const float* W1, const float* b1, const float* W2, const float* b2,
                 float* edge_event, float* edge /* [E,2] nullable */, void* stream);

/* ------------------------------------------------------------------------------------------------
 * (a12) in-place feature update, models.py:503-516 + test.py:401-407.
 *   x_j[:, :2] += y_j/5; x_j[:,6:8] = y_j; x_g[:,3] += y_g0/20; x_g[:,4] = y_g1; x_g[:,n_grain_feat-1] = y_g0;
 *   z += dz on both, then if x_g[0,2] > z_max: z = z_max everywhere.   scratch: device int32[1].
 * ---------------------------------------------------------------------------------------------- */
int gg_feature_update(float* x_joint, int32_t ld_j, int32_t n_joint, const float* y_joint,
                      float* x_grain, int32_t ld_g, int32_t n_grain, int32_t n_grain_feat, const float* y_grain,
                      float dz, float z_max, int32_t* scratch, void* stream);
/* Ensemble form (BASELINE config 5: a block-diagonal batch of independent rollouts, each with its own span): dz_joint /
 * dz_grain hold span_of_graph / 121 per node; z = min(z + dz, z_max) per node (all nodes of a graph carry one z, so this
 * is test.py:405-407 per graph). */
int gg_feature_update_batched(float* x_joint, int32_t ld_j, int32_t n_joint, const float* y_joint,
                              float* x_grain, int32_t ld_g, int32_t n_grain, int32_t n_grain_feat, const float* y_grain,
                              const float* dz_joint, const float* dz_grain, float z_max, void* stream);

/* ------------------------------------------------------------------------------------------------
 * (f2) geometry feedback: grain centres from the joint positions, periodic boundary.  Replaces the host loop of
 *      graph.update (graph_datastruct.py:672-708) that graph_trajectory.GNN_update (graph_trajectory.py:1010-1098) runs
 *      after every NN step, and the write-back of test.py:556-559.  Bit-exact against the reference's numpy arithmetic:
 *      per grain, its joints in the order of the reference's joint2vertex dict, each unwrapped against the PREVIOUS
 *      moved joint (periodic_move :55-72; the first vertex stays float32, later ones become float64), the whole grain
 *      shifted by +1 along an axis where a vertex is <= -1e-12, centre = np.mean (numpy's pairwise float64 sum).
 *  gg_joint_rank: rank[j] = index of the first grain->joint edge whose target is j (dict order of
 *      graph_trajectory.py:1062-1080); gj_dst = row 1 of the grain->joint edge_index.
 *  gg_region_key: key[k] = rank[col[k]] for the CSR of the grain->joint edges taken BY GRAIN (gg_csr_build on the edge
 *      list with its two rows swapped: rowptr over grains, col = joints).
 *  gg_region_sort: col_sorted = the joints of every grain in dict order (increasing key) — once per topology, so that the
 *      per-step kernel walks each grain's joints without keys.
 *  gg_region_center: one thread per grain.  key == NULL: `col` is already in dict order (gg_region_sort); otherwise the
 *      joints are visited in increasing key order by repeated minimum search.  x_joint rows hold (x, y) in columns 0..1 (ld_j even, 8-byte aligned);
 *      domain_factor > 1: global = (x + joint_offset[j]) / domain_factor in fp32 (test.py:472-474), joint_offset [Nj,2];
 *      centers (nullable): float64 [n_grain, 2], NaN for grains with <= 1 joint (skipped by :684, features untouched);
 *      x_grain (nullable): columns 0..1 <- fp32(centre), `(c * domain_factor) % 1` on scaled patches (test.py:558-559).
 *      Preconditions (the reference's own, graph_trajectory.py:1066-1067): (grain, joint) pairs are unique, and no two
 *      joints touch the same three grains (the reference's dict drops one of such a pair).
 * ---------------------------------------------------------------------------------------------- */
int gg_joint_rank(const int64_t* gj_dst, int64_t n_edges, int32_t n_joint, int32_t* rank, void* stream);
int gg_region_key(const int32_t* col, const int32_t* rank, int64_t n_edges, int32_t* key, void* stream);
int gg_region_sort(const int32_t* rowptr, const int32_t* col, const int32_t* key, int32_t n_grain,
                   int32_t* col_sorted, void* stream);
int gg_region_center(const float* x_joint, int32_t ld_j, const float* joint_offset /* nullable */, float domain_factor,
                     const int32_t* rowptr, const int32_t* col, const int32_t* key /* nullable */, int32_t n_grain,
                     double* centers /* nullable */, float* x_grain /* nullable */, int32_t ld_g, void* stream);

/* QoI bookkeeping of GNN_update (graph_trajectory.py:1041-1051 `area_counts`, `extraV_traj`; :1100-1103 `vertex_area`) from the
 * resident grain rows (column 3 = area, 4 = extra volume):
 *   area_counts[g] = area_g s^2 / area_sum for live grains (mask > 0; NaN otherwise), area_sum = sum(area mask) / (lxd / 40)^2 (given);
 *   extra_v[g] = mask_g extraV_g / v_scale s^3 (v_scale = targets_scaling['grain'] = 20);
 *   vertex_area[j] = mesh2 * sum over the grains g of joint j of area_counts[g] / #joints(g)   (rowptr_j / col_g: grain->joint CSR by
 *   joint, rowptr_g: the same edges by grain; vertex_area may be NULL). */
int gg_area_bookkeeping(const float* x_grain, int32_t ld_g, const float* mask_grain /* nullable */, int32_t ld_m, int32_t n_grain,
                        double s, double area_sum, double v_scale, double* area_counts, double* extra_v,
                        const int32_t* rowptr_j, const int32_t* col_g, const int32_t* rowptr_g, int32_t n_joint,
                        double mesh2, double* vertex_area /* nullable */, void* stream);

/* ------------------------------------------------------------------------------------------------
 * (f1) topology update on the device.  Replaces GrainNN_classifier.update (models.py:614-768: grain elimination in area order,
 *      neighbour switching in probability order, the two-sided sweep; without the optional nucleation branch :771-835),
 *      switching_edge_index (:899-1053) and delete_grain_index (:864-896); `cleanup` (:846-862) is the caller's stable compaction of
 *      the columns whose first row is not -1.
 *  Edge arrays: int64 [2, cap] row-major (row r at edges + r * cap), n used columns; pp = joint->joint, pq = joint->grain.  The
 *      update edits them in place (overwrites, appended columns, -1 = deleted) exactly as the reference does, so the compacted
 *      result equals the reference's position for position.
 *  gg_topology_lists: ascending position lists per (row, value): list_r[v * cap_r + k], cnt_r[v]; caps from gg_topology_caps
 *      (joint rows: cap_joint, the grain row of pq: cap_grain); status[0] != 0 on overflow / out-of-range ids.
 *  gg_topology_update: the update of one step in one CTA (eliminations one after the other, the plain switches concurrently in
 *      conflict-free rounds; results are those of the reference's sequential order).  Candidates are the device buffers of gg_select_events: grains
 *      (ge_ids, ge_vals = predicted area; sorted here by area) and joint-joint columns with src < dst (l1_ids, l1_vals = the
 *      candidate's PROBABILITY, i.e. sigmoid of its logit as torch computes it; sorted here by probability descending, equal
 *      probabilities - saturated or colliding in fp32 - in ascending column order like the reference's sort, models.py:730-731), each with its count (clamped to *_cap).
 *      x_joint rows hold (x, y) in columns 0..1 and the predicted (dx, dy) in columns col_dxy, col_dxy + 1; joint_row (nullable)
 *      maps a joint id to its row of x_joint.  y_joint [Nj, 2] and mask_grain / mask_joint (fp32 [N], 1 = live) are updated in
 *      place; act_* are scratch (uint8 [N]).  ahead_cnt int32 [Nj] and ahead_flag uint8 [cap_pp] must be zero on entry (they are
 *      zero again on exit); dirty_flag uint8 [Ng] zero on entry; dirty_list int32 [Ng]; scratch int32 [Ng + 2 (l1_cap + ge_cap) + 128];
 *      ge_sorted int32 [ge_cap], l1_work int32 [l1_cap], l1_logit_work fp32 [l1_cap]; work int32 [gg_topology_work_ints(...)].
 *      result int64 [8]: {n_pp, n_pq, switches, grain events out, error (GGTopoError, topology_core.h), ge_count, l1_count};
 *      switching_list int64 [l1_cap, 2] (models.py:741); grain_event_out int32 [ge_cap + Ng]: the candidates followed by the forced and
 *      two-sided eliminations (models.py:757-759).
 * ---------------------------------------------------------------------------------------------- */
int gg_topology_caps(int32_t* cap_joint, int32_t* cap_grain);
int64_t gg_topology_work_ints(int64_t n_l1, int64_t n_ge, int64_t n_grain);
int gg_topology_lists(const int64_t* edges, int64_t cap, int64_t n, int32_t* list0, int32_t* cnt0, int32_t cap0, int64_t n0,
                      int32_t* list1, int32_t* cnt1, int32_t cap1, int64_t n1, int32_t* status, void* stream);
int gg_topology_update(int64_t* pp, int64_t cap_pp, int64_t n_pp, int64_t* pq, int64_t cap_pq, int64_t n_pq,
                       int32_t* pp_list0, int32_t* pp_cnt0, int32_t* pp_list1, int32_t* pp_cnt1,
                       int32_t* pq_list0, int32_t* pq_cnt0, int32_t* pq_list1, int32_t* pq_cnt1,
                       int32_t* ahead_cnt, uint8_t* ahead_flag,
                       float* x_joint, int32_t ld_xj, const int32_t* joint_row /* nullable */, int32_t col_dxy,
                       float* y_joint, const float* y_grain, int32_t ld_yg,
                       float* mask_grain, float* mask_joint, uint8_t* act_grain, uint8_t* act_joint,
                       int32_t n_joint, int32_t n_grain,
                       const int32_t* ge_count, const int32_t* ge_ids, const float* ge_vals, int32_t ge_cap,
                       const int32_t* l1_count, const int32_t* l1_ids, const float* l1_vals, int32_t l1_cap,
                       uint8_t* dirty_flag, int32_t* dirty_list, int32_t* scratch,
                       int32_t* ge_sorted, int32_t* l1_work, float* l1_logit_work,
                       int64_t* switching_list, int32_t* grain_event_out, int32_t* work, int64_t* result, void* stream);

/* ------------------------------------------------------------------------------------------------
 * (f4) polygon raster + layer error.  Replaces graph.plot_polygons (graph_datastruct.py:553-610, periodic branch: PIL
 *      ImageDraw.polygon per grain into a 2s x 2s image in `region_coors` order, grain id as colour, the four quadrants folded with
 *      max) and graph.compute_error_layer (:346-348).
 *  gg_raster_polygons: polygon p has the integer vertices verts[2 * poly_ptr[p] .. 2 * poly_ptr[p + 1]) (x, y pairs, `int(coor * s)`
 *      as the reference truncates them, :585) and the grain id ids[p]; polygons are given in DRAW order (a pixel keeps the last
 *      polygon that covers it).  scratch: int32 [2s x 2s]; alpha: int32 [s x s] (Image convention [ny, nx]), 0 = never drawn.
 *      Polygons with <= 1 or > 32 vertices are skipped (:588).
 *  gg_count_mismatch: count[0] = number of i with a[i] != b[i] (error_layer = count / n).
 * ---------------------------------------------------------------------------------------------- */
int gg_raster_polygons(const int32_t* poly_ptr, const int32_t* verts, const int32_t* ids, int32_t n_poly, int32_t s,
                       int32_t* scratch, int32_t* alpha, void* stream);
int gg_count_mismatch(const int32_t* a, const int32_t* b, int64_t n, unsigned long long* count, void* stream);

/* ------------------------------------------------------------------------------------------------
 * (f1, first stage) event candidates.  Replaces the host scans of the full prediction arrays,
 *      L1 = ((sigmoid(edge_event) > threshold) & (src < dst)).nonzero()            models.py:627-629
 *      grain_event = ((mask_grain > 0) & (grain_area < threshold)).nonzero()       test.py:414
 *  by one streaming pass that leaves only the candidates on the device:
 *      mode 0 keeps values[i*ld] >= threshold, mode 1 keeps values[i*ld] < threshold;
 *      src/dst (both or neither; int64 [n]): additionally src[i] < dst[i];  mask (nullable, fp32, stride ld_mask): mask > 0.
 *  For mode 0 the caller passes the smallest fp32 logit whose sigmoid exceeds the probability threshold (found on the
 *  host with the reference's sigmoid), so the decision equals the reference's bit for bit.
 *  count (device int32[1]) receives the number of candidates — it may exceed cap, in which case only cap of them were
 *  stored and the caller repeats with larger buffers; ids / vals [cap] hold them in NO particular order (the caller sorts
 *  the few survivors: by id for `nonzero` order, then as test.py:416 / models.py:730-731 prescribe).
 * ---------------------------------------------------------------------------------------------- */
int gg_select_events(const float* values, int64_t n, int32_t ld, float threshold, int32_t mode,
                     const int64_t* src /* nullable */, const int64_t* dst /* nullable */,
                     const float* mask /* nullable */, int32_t ld_mask,
                     int32_t cap, int32_t* count, int32_t* ids, float* vals, void* stream);

/* ------------------------------------------------------------------------------------------------
 * (e) halo pack / unpack for the slab-partitioned domain: out[i, :] = src[idx[i], :] and the inverse.
 *     width must be a multiple of 4 floats and rows 16-byte aligned.
 * ---------------------------------------------------------------------------------------------- */
int gg_gather_rows(const float* src, int32_t ld_src, const int32_t* idx, int32_t n, int32_t width,
                   float* out, int32_t ld_out, void* stream);
int gg_scatter_rows(const float* src, int32_t ld_src, const int32_t* idx, int32_t n, int32_t width,
                    float* out, int32_t ld_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GRAINGNN_B200_H */
