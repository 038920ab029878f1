/*
 * i3d.h — C ABI of lib3dinfomax_b200.so: the sm_100a kernels behind the 3DInfomax hot path
 * (PNA 2-D encoder, Net3D 3-D encoder, NTXent losses, optimizer step).
 *
 * The reference (HannesStark/3DInfomax) is pure Python: every entry point below replaces a
 * *library call site* of the reference (DGL message passing / PyTorch ATen ops), cited as
 * file:line relative to the reference root.  The Python host layer (3dinfomax_b200/*.py) binds these
 * with ctypes and mirrors the reference's model_type / model3d_type / loss_func plugin surface.
 *
 * Conventions
 *   - Every pointer is a DEVICE pointer owned by the caller unless the comment says "host".
 *     The library allocates nothing and keeps no mutable global state (error string: thread local).
 *   - All matrices are row-major fp32 with an explicit leading dimension (elements).
 *   - `stream` is a cudaStream_t passed as void*; every kernel is enqueued on it (CUDA-graph safe).
 *   - Return value: 0 = ok, <0 = error (I3D_ERR_*); i3d_last_error_string() describes the last error
 *     of the calling thread.  Nothing throws or aborts.
 */
#ifndef I3D_H_
#define I3D_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define I3D_OK 0
#define I3D_ERR_INVALID (-1)     /* bad argument (null pointer, negative size, unsupported width ...) */
#define I3D_ERR_UNSUPPORTED (-2) /* valid request the library has no kernel for */
#define I3D_ERR_CUDA (-3)        /* a CUDA runtime call / launch failed */

#define I3D_ACT_NONE 0
#define I3D_ACT_RELU 1
#define I3D_ACT_SILU 2
#define I3D_ACT_LEAKY_RELU 3 /* nn.LeakyReLU() default slope 0.01 (models/pna_original.py:302) */

#define I3D_GEMM_NT 0 /* C[m,n] = sum_k A[m,k]   * B[n,k]   (y = x W^T: nn.Linear forward) */
#define I3D_GEMM_NN 1 /* C[m,n] = sum_k A[m,k]   * B[k,n]   (dx = dy W)                    */
#define I3D_GEMM_TN 2 /* C[m,n] = sum_k A[k,m]   * B[k,n]   (dW = dy^T x)                  */

#define I3D_RO_SUM 0
#define I3D_RO_MEAN 1
#define I3D_RO_MAX 2
#define I3D_RO_MIN 3

int i3d_version(void);
const char* i3d_last_error_string(void);
/* number of kernels this library has launched from the calling process (bench.py's gpu_launches) */
int64_t i3d_launch_count(void);
/* Programmatic dependent launch (griddepcontrol) between consecutive kernels of the library: on by default
 * (environment I3D_PDL=0 disables); returns the previous setting.  Results are identical either way. */
int i3d_set_pdl(int enabled);

/* ------------------------------------------------------------------------------------------------
 * Graph structure.  Replaces DGL's per-call degree bucketing (host numpy sort + sync) behind
 * g.update_all(message_func, reduce_func)  [models/pna.py:206]  and  fn.mean/fn.sum [models/net3d.py:94-96,109].
 * ---------------------------------------------------------------------------------------------- */

/* Stable counting sort of the E edges by `key` (dst for the in-CSR, src for the out-CSR).
 *   rowptr[N+1], eid[E] (edge ids, ascending inside every row == argsort(key, stable) -- BIT EXACT),
 *   col[E] = other[eid[k]], rowid[E] = key[eid[k]].  cursor_ws: N int32 scratch.                    */
int i3d_csr_build(const int64_t* key, const int64_t* other, int64_t E, int64_t N, int32_t* rowptr,
                  int32_t* col, int32_t* rowid, int32_t* eid, int32_t* cursor_ws, void* stream);
/* same with int32 keys (used to sort the CSR-ordered edge list by source: the out-CSR position map) */
int i3d_csr_build_i32(const int32_t* key, const int32_t* other, int64_t E, int64_t N, int32_t* rowptr,
                      int32_t* col, int32_t* rowid, int32_t* eid, int32_t* cursor_ws, void* stream);

/* ptr[0]=0, ptr[i+1]=ptr[i]+counts[i]  (graph_ptr from g.batch_num_nodes(), self_supervised_trainer.py:28) */
int i3d_segment_ptr(const int64_t* counts, int64_t B, int32_t* ptr, void* stream);

/* amp[v]=(float)ln(D+1), att[v]=(float)(1/ln(D+1)) with D=in-degree; both 0 for D=0.
 * scale_amplification / scale_attenuation with avg_d["log"]=1.0  [models/pna.py:61-68,153]          */
int i3d_degree_scalers(const int32_t* rowptr, int64_t N, float* amp, float* att, void* stream);
/* same with the SCALAR avg_d of the tower PNA: amp = (float)(ln(D+1)/avg_d), att = (float)(avg_d/ln(D+1))
 * [models/pna_original.py:28-35]                                                                        */
int i3d_degree_scalers_avg(const int32_t* rowptr, int64_t N, double avg_d, float* amp, float* att, void* stream);
/* y[m, :] = x[m, :] * s[m]  (graph_norm: h * snorm_n, models/pna_original.py:258-259; its backward is the same op) */
int i3d_scale_rows(const float* x, int ldx, const float* s, int64_t M, int F, float* y, int ldy, void* stream);

/* Degree plan: the reference evaluates the posttrans FC on cat[h, A, A*amp_D, A*att_D] (13F columns,
 * models/pna.py:207,232); amp_D / att_D depend on the in-degree D only (models/pna.py:57-68), so nodes grouped by D
 * can share one merged weight W_id + amp_D W_amp + att_D W_att and the GEMM runs with K = 5F.  (DGL itself buckets
 * nodes by degree inside update_all; here the buckets feed the tensor-core tiles instead of ~25 launches each.)
 * Bucket b = nodes of in-degree b, b < n_buckets <= 16; every bucket owns whole 128-row tiles of a "virtual row"
 * order, node ids ascending inside a bucket.  With tiles = ceil(N/128):
 *   perm[128 * (tiles + n_buckets)]       virtual row -> node id, -1 for padding rows
 *   tile_bucket[tiles + n_buckets]        bucket of each row tile, -1 for unused tail tiles
 *   chunk_tab[3 * (ceil(tiles / chunk_tiles) + n_buckets)]   split-K chunks for the weight gradient:
 *                                         (first virtual row, rows, bucket); rows = 0 for unused chunks
 *   overflow[1]                           1 if some node has D >= n_buckets (clamped into the last bucket: the
 *                                         caller must not use the results)                                      */
int i3d_degree_plan(const int32_t* rowptr, int64_t N, int n_buckets, int chunk_tiles, int32_t* perm,
                    int32_t* tile_bucket, int32_t* chunk_tab, int32_t* overflow, void* stream);

/* Device-side batch construction from a packed molecule store in HBM (replaces B x QM9Dataset.__getitem__
 * [datasets/qm9_dataset.py:189-244: get_graph :219-229, get_complete_graph :231-244, get_pairwise :207-217] +
 * dgl.batch [datasets/custom_collate.py:105-114] + the host->device copy of the collated graphs).
 * Store (device, layout of the reference's processed file, qm9_dataset.py:454-467): atom_slices / edge_slices [M+1]
 * with leading 0, edge_indices [2, Etot] molecule-local ids, atom_features [Ntot, n_atom_feat] int64,
 * edge_features [Etot, n_edge_feat] int64, coordinates [Ntot, 3] fp32.
 * Per batch (device): idx[B] molecule ids; node_ptr / edge_ptr / edge3_ptr [B+1] exclusive prefix sums of the
 * batch's atom counts n_k, bond-edge counts and n_k (n_k - 1).
 *   i3d_collate_2d: src/dst [E] = molecule-local ids + node_ptr[k]; x_atom [N, n_atom_feat]; e_attr [E, n_edge_feat]
 *   i3d_collate_3d: complete digraph without self loops in the reference's order (src = repeat_interleave(arange(n),
 *                   n-1), dst ascending), d3 [E3] = ||x_src - x_dst||_2 as torch.norm evaluates it in fp32
 *                   (x*x, fma(y,y,.), fma(z,z,.), sqrt: bit-equal to the CPU reference on the pinned vectors)   */
int i3d_collate_2d(const int64_t* idx, int64_t B, const int64_t* atom_slices, const int64_t* edge_slices,
                   const int64_t* edge_indices, int64_t Etot, const int64_t* atom_features, int n_atom_feat,
                   const int64_t* edge_features, int n_edge_feat, const int64_t* node_ptr, const int64_t* edge_ptr,
                   int64_t N, int64_t E, int64_t* src, int64_t* dst, int64_t* x_atom, int64_t* e_attr, void* stream);
int i3d_collate_3d(const int64_t* idx, int64_t B, const int64_t* atom_slices, const float* coordinates,
                   const int64_t* node_ptr, const int64_t* edge3_ptr, int64_t E3, int64_t* src3, int64_t* dst3,
                   float* d3, void* stream);

/* Shape-bucketed batch construction: the same batches PADDED to a bucket's capacities (n_cap nodes, e_cap edges) and
 * emitted together with their CSR structure, so that one captured CUDA graph (static shapes) serves every batch of
 * the bucket although each batch of an epoch has a different (N, E, E3) [train.py:595-598,
 * datasets/custom_collate.py:105-114].  Valid sizes are read on the DEVICE (node_ptr[B], edge_ptr[B], edge3_ptr[B]).
 * Padding convention (all consumers in this library honour it): nodes [N, n_cap) have no edges and feature index 0;
 * edges [E, e_cap) have src = dst = -1 in both id orders (a negative gather index reads a zero row), eid / out_pos =
 * own position.  rowptr[n_cap] = E and graph_ptr[B] = N are the device scalars the *_v entry points take as m_valid.
 *   2-D: dgl.batch keeps each molecule's node / edge order, so the destination-sorted edge-id-stable CSR of the batch
 *        is the concatenation of per-molecule CSRs.  The store carries them (molecule-local, int32, computed once):
 *        in_rowptr_l[Ntot] (in-edges of earlier atoms of the molecule), in_eid_l[Etot] (local edge id at local CSR
 *        position), out_rowptr_l[Ntot], out_pos_l[Etot] (local CSR position of the molecule's edges sorted by source).
 *        Outputs: src/dst [e_cap] int64 (edge-id order), x_atom [n_cap, n_atom_feat], e_attr [e_cap, n_edge_feat],
 *        rowptr / out_rowptr [n_cap+1], src_csr / dst_csr / eid / out_pos [e_cap], graph_ptr [B+1]  — bit-equal to
 *        i3d_csr_build on the collated edge list.  code_csr [e_cap] (optional): sum_c e_attr[eid[k], c] * code_mult[c],
 *        the mixed-radix index of the CSR-ordered edge's categorical feature row (the table index of
 *        i3d_edge_gather_add); 0 for padding edges.
 *   3-D: C conformer graphs per molecule, molecule-major [datasets/qmugs_dataset.py:149-166]; coords [Ntot, ld_coords]
 *        fp32 with conformer c in columns [3c, 3c+3).  Complete digraphs in the reference's order have a closed-form
 *        CSR; the out-CSR row pointer equals rowptr.  Outputs: src3/dst3 [e_cap] int64, d3 [e_cap] (edge-id order),
 *        rowptr [n_cap+1], src_csr / dst_csr / eid / out_pos [e_cap], graph_ptr [B*C+1], num_nodes3 [B*C] (optional). */
int i3d_collate_2d_struct(const int64_t* idx, int64_t B, const int64_t* atom_slices, const int64_t* edge_slices,
                          const int64_t* edge_indices, int64_t Etot, const int64_t* atom_features, int n_atom_feat,
                          const int64_t* edge_features, int n_edge_feat, const int32_t* in_rowptr_l,
                          const int32_t* in_eid_l, const int32_t* out_rowptr_l, const int32_t* out_pos_l,
                          const int64_t* node_ptr, const int64_t* edge_ptr, int64_t n_cap, int64_t e_cap, int64_t* src,
                          int64_t* dst, int64_t* x_atom, int64_t* e_attr, int32_t* rowptr, int32_t* src_csr,
                          int32_t* dst_csr, int32_t* eid, int32_t* out_rowptr, int32_t* out_pos, int32_t* graph_ptr,
                          const int64_t* code_mult, int64_t* code_csr, void* stream);
int i3d_collate_3d_struct(const int64_t* idx, int64_t B, int C, const int64_t* atom_slices, const float* coords,
                          int ld_coords, const int64_t* node_ptr, const int64_t* edge3_ptr, int64_t n_cap, int64_t e_cap,
                          int64_t* src3, int64_t* dst3, float* d3, int32_t* rowptr, int32_t* src_csr, int32_t* dst_csr,
                          int32_t* eid, int32_t* out_pos, int32_t* graph_ptr, int64_t* num_nodes3, void* stream);

/* ------------------------------------------------------------------------------------------------
 * AtomEncoder / BondEncoder  [commons/mol_encoder.py:34-42,65-73; models/pna.py:162-163]
 *   out[r,:] = sum_c table[col_off[c] + idx[perm ? perm[r] : r, c], :]
 * `table` is the row-concatenation of the C embedding tables, col_off[c] the first row of table c.
 * ---------------------------------------------------------------------------------------------- */
int i3d_embed_sum_fwd(const int64_t* idx, int64_t R, int C, const int32_t* col_off, const int32_t* perm,
                      const float* table, int F, float* out, void* stream);
/* gtable (caller-zeroed, table_rows x F) += scatter of gout.  max_dim = largest per-column vocabulary: when the
 * column slice fits in shared memory the sums are formed per CTA first (few global atomics), else plain fp32 atomics;
 * pass table_rows = max_dim = 0 to force the latter. */
int i3d_embed_sum_bwd(const int64_t* idx, int64_t R, int C, const int32_t* col_off, const int32_t* perm,
                      const float* gout, int F, float* gtable, int table_rows, int max_dim, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Dense stage.  Replaces cat + nn.Linear (cuBLAS SGEMM) at models/pna.py:249-252 (pretrans over
 * cat[h_src,h_dst,e]), models/pna.py:207-209 (posttrans over cat[h, agg x scalers]),
 * models/net3d.py:112-115,122-124 and the MLP heads [models/base_layers.py:100-101].
 * The K dimension is a list of segments so that torch.cat / index_select / degree scalers never
 * materialise:   C = sum_s  op(diag(scale_s) * gather(A_s, a_idx_s)) * op(gather(B_s, b_idx_s))
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  const float* A;       /* NT/NN: [*, K] rows indexed by m;  TN: [K, M] rows indexed by k           */
  const float* B;       /* NT: [N, K] rows indexed by n;  NN/TN: [K, N] rows indexed by k            */
  const int32_t* a_idx; /* optional row gather for A (NT/NN: per m; TN: per k)                        */
  const int32_t* b_idx; /* optional row gather for B (TN only: per k)                                 */
  const float* scale;   /* optional multiplier (NT/NN: per output row m; TN: per k)                   */
  int32_t K;            /* depth of this segment                                                      */
  int32_t lda, ldb;     /* leading dimensions (elements)                                              */
} i3d_gemm_seg;

/* segs: HOST array of n_seg (<=4) descriptors.  bias[N] optional (added once).  accumulate!=0: C +=
 * Backend: tcgen05.mma kind::tf32 with 3xTF32 operand splitting (fp32-level accuracy, accumulator in TMEM) for
 * the shapes it covers, the fp32 SIMT kernel otherwise.  i3d_gemm_backend(1) forces SIMT,
 * (2) keeps tensor cores and routes TN through the MN-major descriptor kernel; returns the old value. */
int i3d_gemm_backend(int backend);
/* out[c, r] = in[r, c]; used to hand W^T to the NT kernel so that dx = dy W runs on the same tensor-core path */
int i3d_transpose(const float* in, int64_t rows, int cols, int ld_in, float* out, int ld_out, void* stream);
int i3d_gemm(int mode, int64_t M, int N, int n_seg, const i3d_gemm_seg* segs, float* C, int ldc,
             const float* bias, int accumulate, void* stream);
/* Same, with caller-owned device scratch: when ws_bytes >= i3d_gemm_ws_bytes(...) the NT kernel splits the dense B
 * operand (weights) into tf32 hi/lo copies once and streams them by TMA instead of staging them through threads.
 * i3d_gemm_ws_bytes returns 0 for problems that do not use scratch.
 */
size_t i3d_gemm_ws_bytes(int mode, int64_t M, int N, int n_seg, const i3d_gemm_seg* segs);
/* OR into `stats_act` (GEMM entry points) or `act` (i3d_bn_bwd_reduce*) when the fp64 statistics buffer handed in is
 * ALREADY ZERO: the library then skips its own cudaMemsetAsync.  A caller that owns an arena zeroed once per step saves
 * ~50 memset nodes per training step this way (each one also breaks programmatic dependent launch between its
 * neighbours). */
#define I3D_STATS_PREZEROED 0x100
/* Layout of every fp64 column-statistics buffer of this library (col_stats, sums, sums2): statistic j of column c lives
 * at element (j * F + c) * I3D_STATS_STRIDE, i.e. one accumulator per 128-byte line; a buffer for F columns holds
 * 2 * F * I3D_STATS_STRIDE doubles.  Hundreds of CTAs add their partial sums into these with atomics; with 4 doubles
 * per 32-byte sector the L2 serialised them (measured on B200, 296 CTAs x 400 columns: 27.5 us packed -> 16.3 us
 * spread for the same kernel). */
#define I3D_STATS_STRIDE 16
/* col_stats (optional, NT without accumulate): fp64 [2N * I3D_STATS_STRIDE] = column sums of act(C) and act(C)^2, i.e. the train-mode
 * BatchNorm statistics of the FCLayer tail, produced by the GEMM epilogue instead of a separate pass over C. */
int i3d_gemm_ex(int mode, int64_t M, int N, int n_seg, const i3d_gemm_seg* segs, float* C, int ldc,
                const float* bias, int accumulate, void* ws, size_t ws_bytes, double* col_stats, int stats_act,
                void* stream);

/* *_v variants (here and in the FCLayer tail below): m_valid is an optional DEVICE int32 scalar, the number of valid
 * leading rows of a shape-bucketed (padded) batch.  Rows >= *m_valid are excluded from BatchNorm statistics and counts,
 * written as zeros by i3d_bn_apply_v / i3d_bn_bwd_apply_v, and never read by the reductions.  NULL = all M rows. */
int i3d_gemm_ex_v(int mode, int64_t M, int N, int n_seg, const i3d_gemm_seg* segs, float* C, int ldc,
                  const float* bias, int accumulate, void* ws, size_t ws_bytes, double* col_stats, int stats_act,
                  const int32_t* m_valid, void* stream);

/* Weights change once per optimizer step, not once per GEMM: a caller that owns persistent scratch can prepare the
 * hi/lo copies of MANY GEMMs' B operands with ONE launch (after the optimizer step) and then run each GEMM with
 * i3d_gemm_nt_prepared, which skips the per-call split (and, for `transposed` operands, the weight transpose that
 * dx = dy W [models/base_layers.py:100 backward] otherwise needs).
 *   i3d_gemm_prep_describe  host-only: fills items_out[0..n_seg) for one GEMM whose scratch is `ws`
 *                           (i3d_gemm_ws_bytes of the same N / segments).  transposed != 0: segs[s].B points at the
 *                           operand stored as [K, N] (ld = ldb), e.g. a column block W[:, o:o+k] of a weight used as
 *                           the B of dx = dy W.  tile0: running tile offset; *tiles_out: tiles this GEMM adds.
 *   i3d_gemm_prep_run       dev_items: the described items copied to DEVICE memory by the caller. */
typedef struct i3d_prep_item {
  const float* B;
  float* hi;
  float* lo;
  int32_t ldb, N, K, kpad, ldo, col0, transposed, tile0;
} i3d_prep_item;
int i3d_gemm_prep_describe(int N, int n_seg, const i3d_gemm_seg* segs, int transposed, void* ws, int tile0,
                           i3d_prep_item* items_out, int* tiles_out);
int i3d_gemm_prep_run(const i3d_prep_item* dev_items, int n_items, int total_tiles, void* stream);
/* NT GEMM on the tensor-core path with B taken from prepared scratch (segs[s].B / ldb are ignored).  Fails with
 * I3D_ERR_INVALID when the shape is not eligible for that path (use i3d_gemm_nt_prepared_ok to ask first). */
int i3d_gemm_nt_prepared_ok(int64_t M, int N, int n_seg, const i3d_gemm_seg* segs);
int i3d_gemm_nt_prepared(int64_t M, int N, int n_seg, const i3d_gemm_seg* segs, float* C, int ldc, const float* bias,
                         int accumulate, const void* ws, double* col_stats, int stats_act, void* stream);

int i3d_gemm_nt_prepared_v(int64_t M, int N, int n_seg, const i3d_gemm_seg* segs, float* C, int ldc, const float* bias,
                           int accumulate, const void* ws, double* col_stats, int stats_act, const int32_t* m_valid,
                           void* stream);

/* Degree-bucketed GEMMs over an i3d_degree_plan (posttrans FC of PNALayer, models/pna.py:207-211, and its backward).
 *   i3d_posttrans_merge    W [Fout, 13F] (ld = ldw) -> tf32 hi/lo operands of the merged weights
 *                          Wm_b = [Wh | W_id + a_b W_amp + t_b W_att], a_b = (float)ln(b+1), t_b = (float)(1/ln(b+1)):
 *                          fwd_hi/lo [n_buckets*Fout, kpad(F)+kpad(4F)]  (B of y = [h|A] Wm_b^T; kpad = round up to 32)
 *                          bwd_hi/lo [n_buckets*5F, kpad(Fout)]          (B of d[h|A] = dy Wm_b)
 *                          rows are fwd_pitch / bwd_pitch floats apart (>= the padded K extent; a pitch that is not
 *                          a power of two keeps the TMA row fetches spread over the L2 slices); pad columns are
 *                          not written: zero-fill the buffers once when allocating them.
 *   i3d_gemm_nt_bucketed   C[row_map[m], n] = bias[n] + sum_s sum_k A_s[a_idx_s[m] or m, k] * B[bucket(m)*N + n, k]
 *                          for the Mv = 128*(tiles+n_buckets) virtual rows; rows with row_map[m] < 0 are skipped (and
 *                          excluded from col_stats).  Segments: unscaled, K % 4 == 0; a_idx entries < 0 read zeros.
 *   i3d_gemm_tn_chunked    C[bucket] (+)= sum over the chunk's virtual rows k of A[a_idx[k], :]^T B[b_idx[k], :]
 *                          (C + bucket * c_bucket_stride; vector atomics: zero C before the call); seg->K = Mv.
 *   i3d_posttrans_unmerge  dW[:, F:5F] += sum_b dWb[b]; dW[:, 5F:9F] += sum_b a_b dWb[b]; dW[:, 9F:13F] += sum_b t_b dWb[b]
 *                          with dWb [n_buckets, Fout, 4F] dense.                                                  */
int i3d_posttrans_merge(const float* W, int ldw, int Fout, int F, int n_buckets, float* fwd_hi, float* fwd_lo,
                        int fwd_pitch, float* bwd_hi, float* bwd_lo, int bwd_pitch, void* stream);
int i3d_gemm_nt_bucketed(int64_t Mv, int N, int n_seg, const i3d_gemm_seg* segs, float* C, int ldc, const float* bias,
                         const float* b_hi, const float* b_lo, int b_pitch, int n_buckets, const int32_t* tile_bucket,
                         const int32_t* row_map, double* col_stats, int stats_act, void* stream);
/* m_valid: output rows (row_map values) >= *m_valid are stored but excluded from col_stats */
int i3d_gemm_nt_bucketed_v(int64_t Mv, int N, int n_seg, const i3d_gemm_seg* segs, float* C, int ldc, const float* bias,
                           const float* b_hi, const float* b_lo, int b_pitch, int n_buckets, const int32_t* tile_bucket,
                           const int32_t* row_map, double* col_stats, int stats_act, const int32_t* m_valid,
                           void* stream);
int i3d_gemm_tn_chunked(int64_t M, int N, const i3d_gemm_seg* seg, float* C, int ldc, int64_t c_bucket_stride,
                        const int32_t* chunk_tab, int n_chunks, void* stream);
/* Tuning aid: per-role blocked-cycle counters of the warp-specialised NT kernel (16 HOST uint64; read-and-clear).
 * All zero unless the library was built with -DI3D_WS_DEBUG (see i3d_gemm_tc_ws.cu). */
int i3d_gemm_debug_counters(unsigned long long* out16);
int i3d_posttrans_unmerge(const float* dWb, int n_buckets, int Fout, int F, float* dW, int ldw, void* stream);

/* ------------------------------------------------------------------------------------------------
 * FCLayer tail: activation -> BatchNorm1d (train: batch statistics)  [models/base_layers.py:102-110]
 *   h = act(Y);  O = gamma*(h-mean)*rstd + beta (+ residual)
 * ---------------------------------------------------------------------------------------------- */
/* sums[(0:F) * S]=sum_rows act(Y), sums[(F:2F) * S]=sum_rows act(Y)^2, S = I3D_STATS_STRIDE (fp64; zeroed inside) */
int i3d_act_colstats(const float* Y, int64_t M, int F, int ldy, int act, double* sums, void* stream);
/* training!=0: statistics from `sums` (biased var for normalisation), running stats updated with
 * `momentum` (unbiased var), *num_batches_tracked += 1;  else running stats are used.
 * save_mean_rstd[2F] receives the (mean, rstd) used.  residual (optional, ld = ldo) is added to O.   */
int i3d_bn_apply(const float* Y, int64_t M, int F, int ldy, int act, const double* sums, float* running_mean,
                 float* running_var, int64_t* num_batches_tracked, const float* gamma, const float* beta,
                 float momentum, float eps, int training, float* save_mean_rstd, const float* residual,
                 float* O, int ldo, void* stream);
/* sums2[(0:F) * S]=sum dO, sums2[(F:2F) * S]=sum dO*xhat, S = I3D_STATS_STRIDE (fp64; zeroed inside) */
int i3d_bn_bwd_reduce(const float* dO, int ldd, const float* Y, int ldy, int64_t M, int F, int act,
                      const float* save_mean_rstd, double* sums2, void* stream);
/* same; additionally sets zero_buf[0:zero_n] = 0 (fp32) — the bias-gradient accumulator the following
 * i3d_bn_bwd_apply adds into, so that no separate fill kernel sits between the two */
int i3d_bn_bwd_reduce_ex(const float* dO, int ldd, const float* Y, int ldy, int64_t M, int F, int act,
                         const float* save_mean_rstd, double* sums2, float* zero_buf, int zero_n, void* stream);
/* dY = act'(Y) * gamma*rstd*(dO - mean(dO) - xhat*mean(dO*xhat))   (training)  |  act'(Y)*gamma*rstd*dO (eval)
 * has_bn==0: dY = act'(Y)*dO.  dbias[F] (caller-zeroed) += sum_rows dY; dgamma/dbeta[F] written.      */
int i3d_bn_bwd_apply(const float* dO, int ldd, const float* Y, int ldy, int64_t M, int F, int act, int has_bn,
                     int training, const float* save_mean_rstd, const float* gamma, const double* sums2,
                     float* dY, int lddy, float* dbias, float* dgamma, float* dbeta, void* stream);
/* Workspace of the two-stage column reductions (optional argument `rws` below; NULL = fp atomics into the output).
 * Same-address atomics serialise in L2: with 300-450 CTAs the atomic tail was most of a column-statistics kernel.  With a
 * workspace every CTA stores its partial sums in its own slot, takes a ticket on `counter`, and the last CTA sums the
 * slots in slot order and STORES the result (the output then needs no pre-zeroing and is bit-reproducible).
 *   slots       caller-owned device scratch, slot_bytes >= CTAs x columns x element size (else: atomics);
 *   counter     device uint32, 0 on entry, left 0 on exit; one counter per kernel that may run concurrently. */
typedef struct {
  void* slots;
  int64_t slot_bytes;
  unsigned int* counter;
} i3d_reduce_ws;

/* valid-row variants of the four entry points above (see i3d_gemm_ex_v) */
int i3d_act_colstats_v(const float* Y, int64_t M, int F, int ldy, int act, double* sums, const int32_t* m_valid,
                       const i3d_reduce_ws* rws, void* stream);
int i3d_bn_apply_v(const float* Y, int64_t M, int F, int ldy, int act, const double* sums, float* running_mean,
                   float* running_var, int64_t* num_batches_tracked, const float* gamma, const float* beta,
                   float momentum, float eps, int training, float* save_mean_rstd, const float* residual, float* O,
                   int ldo, const int32_t* m_valid, void* stream);
int i3d_bn_bwd_reduce_v(const float* dO, int ldd, const float* Y, int ldy, int64_t M, int F, int act,
                        const float* save_mean_rstd, double* sums2, float* zero_buf, int zero_n,
                        const int32_t* m_valid, const i3d_reduce_ws* rws, void* stream);
/* dbias_stride: element stride of dbias (column c accumulates into dbias[c * dbias_stride]); 8 puts every accumulator in
 * its own 32-byte sector, which the L2 needs to run the ~300 CTAs' atomics in parallel.  With rws, dbias is STORED. */
int i3d_bn_bwd_apply_v(const float* dO, int ldd, const float* Y, int ldy, int64_t M, int F, int act, int has_bn,
                       int training, const float* save_mean_rstd, const float* gamma, const double* sums2, float* dY,
                       int lddy, float* dbias, int dbias_stride, float* dgamma, float* dbeta, const int32_t* m_valid,
                       const i3d_reduce_ws* rws, void* stream);
/* i3d_bn_bwd_reduce_v followed by i3d_bn_bwd_apply_v (has_bn = 1) as ONE call — the backward of activation ->
 * BatchNorm1d [models/base_layers.py:102-110] — and, when the problem fits, as ONE launch: the two phases are separated
 * by a grid-wide barrier on `barrier` (DEVICE counter, zero on entry, not reset) instead of a kernel boundary, every
 * thread keeping its rows of dO in shared memory — opt-in with I3D_BN_BWD=fused (measured 0.4 % of the step, within
 * noise: the default is the two-kernel path).  barrier == NULL or a problem that does not fit (more than 16 rows per
 * thread at full occupancy) always take the two-kernel path.  The barrier spins: callers run at most one
 * such launch at a time per device (3dinfomax_b200.kernels.bn_bwd keeps it to one stream).  act may carry
 * I3D_STATS_PREZEROED for sums2; zero_buf / zero_n as in i3d_bn_bwd_reduce_v. */
int i3d_bn_bwd_fused_v(const float* dO, int ldd, const float* Y, int ldy, int64_t M, int F, int act, int training,
                       const float* save_mean_rstd, const float* gamma, double* sums2, float* dY, int lddy,
                       float* dbias, int dbias_stride, float* dgamma, float* dbeta, float* zero_buf, int zero_n,
                       const int32_t* m_valid, unsigned int* barrier, void* stream);
/* Factored first layer of the edge MLP [models/pna.py:237-252]: cat[h[src], h[dst], e] W^T =
 * (h W_s^T)[src] + (h W_d^T)[dst] + e W_e^T.  P [N, >=2F] = h [W_s; W_d]^T comes from ONE node-level GEMM; the bond
 * features take prod(5,6,2) = 60 distinct values [commons/mol_encoder.py:4-7], so e W_e^T is a table T [n_codes, F]
 * indexed by the edge's feature code.  Y[m,:] = P[src[m], 0:F] + P[dst[m], F:2F] + T[code[m],:] + bias[:]; negative
 * src / dst read zeros; T / code / bias optional.  col_stats (optional, fp64 [2F]): column sums of act(Y), act(Y)^2 over
 * the rows < *m_valid (all M when m_valid is NULL) — the FCLayer's train-mode BatchNorm statistics; stats_act takes
 * I3D_STATS_PREZEROED like the GEMM entry points. */
int i3d_edge_gather_add(const float* P, int ldp, const int32_t* src, const int32_t* dst, const float* T, int ldt,
                        const int32_t* code, const float* bias, int64_t M, int F, float* Y, int ldy, double* col_stats,
                        int stats_act, const int32_t* m_valid, const i3d_reduce_ws* rws, void* stream);
/* The tables of i3d_edge_gather_add for ALL message-passing layers in one launch, and their gradients:
 *   fwd   T_l[c, n]             = sum_k combo[c, k] W_l[n, col0 + k]          l < L, c < n_codes <= 64, n < Fout
 *   bwd   dW_l[n, col0 + k]    += sum_c dT_l[c, n] combo[c, k]                 (layers with dT[l] or dW[l] NULL skipped)
 *         dcombo[c, k]         += sum_l sum_n dT_l[c, n] W_l[n, col0 + k]      (caller-zeroed; NULL: not computed)
 * combo [n_codes, F] = BondEncoder applied to every categorical combination [commons/mol_encoder.py:45-73];
 * W / T / dT / dW are HOST arrays of L (<= 16) device pointers; W_l is the first pretrans weight of layer l
 * [models/pna.py:249-252], leading dimension ldw, `e` segment at columns [col0, col0 + F); T_l, dT_l [n_codes, Fout]. */
int i3d_bond_tables_fwd(const float* combo, int n_codes, int F, int L, const float* const* W, int ldw, int col0,
                        int Fout, float* const* T, void* stream);
int i3d_bond_tables_bwd(const float* combo, int n_codes, int F, int L, const float* const* W, int ldw, int col0,
                        int Fout, const float* const* dT, float* const* dW, float* dcombo, void* stream);
/* y = act(x) elementwise (used where there is no BN, and for Net3D's second SiLU, models/net3d.py:81) */
int i3d_act_fwd(const float* x, int64_t n, int act, float* y, void* stream);
int i3d_act_bwd(const float* gy, const float* x, int64_t n, int act, float* gx, void* stream);

/* ------------------------------------------------------------------------------------------------
 * PNA aggregation  [models/pna.py:17-37 aggregators, :221-235 reduce_func, :206 update_all]
 * msg[E,F] is in CSR (dst-sorted, edge-id-stable) order, so node v's mailbox is the contiguous row
 * block [rowptr[v], rowptr[v+1]).  out[v, 0:4F] = [mean | max | min | std], std = sqrt(relu(E[x^2]-E[x]^2)+1e-5);
 * zero in-degree -> zeros (DGL zero-fill).  The three degree scalers are NOT materialised: they are
 * per-row scalars folded into the consuming GEMM (i3d_gemm scale segments).
 * ---------------------------------------------------------------------------------------------- */
int i3d_pna_aggregate_fwd(const float* msg, const int32_t* rowptr, int64_t N, int F, float* out, int ldo,
                          void* stream);
/* dmsg_k = g_mean/D + g_max*[k==first argmax] + g_min*[k==first argmin] + g_std*[var>0]*(x_k-mean)/(D*std) */
int i3d_pna_aggregate_bwd(const float* g, int ldg, const float* msg, const float* out, int ldo,
                          const int32_t* rowptr, int64_t N, int F, float* dmsg, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Segment ops over contiguous row blocks.
 *   readout: dgl.readout_nodes(graph,'feat',op) for op in readout_aggregators  [models/pna.py:133; net3d.py:73]
 *   segment_sum: builtin fn.sum / fn.mean reduce  [models/net3d.py:94-96,109]
 * ---------------------------------------------------------------------------------------------- */
/* out[b, i*F:(i+1)*F] = op_i over rows [ptr[b], ptr[b+1]); ops: n_ops (<=4) I3D_RO_* codes (host array) */
int i3d_segment_readout_fwd(const float* x, int ldx, const int32_t* ptr, int64_t B, int F, int n_ops,
                            const int32_t* ops, float* out, void* stream);
/* max/min route the whole gradient to the FIRST row attaining the extremum (DGL segment_reduce arg) */
int i3d_segment_readout_bwd(const float* g, const float* x, int ldx, const float* out, const int32_t* ptr,
                            int64_t B, int F, int n_ops, const int32_t* ops, float* dx, int lddx, void* stream);
/* same; additionally zero-fills dx rows [ptr[B], n_rows): the padding nodes of a shape-bucketed batch */
int i3d_segment_readout_bwd_v(const float* g, const float* x, int ldx, const float* out, const int32_t* ptr,
                              int64_t B, int F, int n_ops, const int32_t* ops, float* dx, int lddx, int64_t n_rows,
                              void* stream);
/* out[v,:] = (mean ? 1/max(D,1) : 1) * sum_{k in [rowptr[v],rowptr[v+1])} x[idx ? idx[k] : k, :]  (+ addend[v,:])
 * idx==NULL: rows of x are already in CSR order.  Also the backward of the h[src] / h[dst] gathers feeding the
 * edge MLP (models/pna.py:249): in-CSR rowptr without idx, out-CSR rowptr with idx = position map.          */
int i3d_segment_sum_fwd(const float* x, int ldx, const int32_t* rowptr, const int32_t* idx, int64_t N, int F,
                        int mean, const float* addend, int lda, float* out, int ldo, void* stream);
/* gx[k,:] = g[rowid[k],:] * (mean ? 1/D : 1) */
int i3d_segment_sum_bwd(const float* g, const int32_t* rowptr, const int32_t* rowid, int64_t E, int F, int mean,
                        float* gx, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Net3D element-wise pieces  [commons/utils.py:103-110; models/net3d.py:61,112-118]
 * ---------------------------------------------------------------------------------------------- */
/* out[r,:] = [sin(d/2^0..2^(k-1)), cos(...), d] with d = dist[perm ? perm[r] : r] */
int i3d_fourier_encode(const float* dist, const int32_t* perm, int64_t E, int k, float* out, void* stream);
/* w = sigmoid(msg . ws + bs);  m = msg * w   (soft_edge_network, net3d.py:117-118) */
int i3d_soft_gate_fwd(const float* msg, int64_t E, int H, const float* ws, const float* bs, float* m, float* w,
                      void* stream);
/* gmsg = gm*w + (sum_h gm*msg) w(1-w) ws ; gws[H], gbs[1] (caller-zeroed) accumulate */
int i3d_soft_gate_bwd(const float* gm, const float* msg, const float* w, int64_t E, int H, const float* ws,
                      float* gmsg, float* gws, float* gbs, void* stream);
/* out[r,:] = vec[:]   /   out[c] = sum_rows x[r,c] */
int i3d_broadcast_rows(const float* vec, int64_t M, int F, float* out, void* stream);
int i3d_colsum(const float* x, int ldx, int64_t M, int F, float* out, void* stream);
/* y = a + b (n elements) */
int i3d_add(const float* a, const float* b, int64_t n, float* y, void* stream);
/* y[m, :F] = a[m, :F] + b[m, :F], every matrix with its own leading dimension: the gradient of a layer input that arrives
 * through two paths of ONE backward node — the residual `+ h` and the first K-segment of the posttrans GEMM, a column
 * slice of the [N, 5F] dx buffer (models/pna.py:207-211) — summed without a strided framework kernel */
int i3d_add_rows(const float* a, int lda, const float* b, int ldb, int64_t M, int F, float* y, int ldy, void* stream);

/* ------------------------------------------------------------------------------------------------
 * NTXent / NTXentMultiplePositives  [commons/losses.py:143-155, 225-246]
 *   dot[B, Bc*C] = z1 z2^T (i3d_gemm NT);  s = dot/(n1 n2 + eps) (norm) ; p = exp(s/tau);
 *   Q_ij = sum_u p[i, j*C+u];  l_i = -log(Q_ii / (sum_j Q_ij - Q_ii));  loss = sum_i l_i / B_total.
 * Rows are the local 2-D embeddings, columns the (all-gathered) 3-D embeddings; the positive of local row i
 * is molecule (row_offset + i) in column space.  eps = 1e-8 for NTXent, 0 for MultiplePositives.
 * ---------------------------------------------------------------------------------------------- */
int i3d_row_norms(const float* z, int64_t R, int D, float* norms, void* stream);
/* P[B, BcC] in: dot, out: p = exp(s/tau).  rowstats[B,2] = (pos, rowsum-pos).  loss_rows[B] = l_i   */
int i3d_ntxent_rows_fwd(float* P, int64_t B, int64_t Bc, int C, const float* n1, const float* n2, int norm,
                        float eps, float tau, int64_t row_offset, float* rowstats, float* loss_rows, void* stream);
/* Contrastive metrics logged by the pre-training configs (configs_clean/pre-train_QM9.yml:15-20): PositiveSimilarity,
 * NegativeSimilarity, TruePositiveRate, TrueNegativeRate, ContrastiveAccuracy [trainer/metrics.py:240-334,444-463],
 * global-vs-global case (pos_mask == None).  dot[B,B] = z1 z2^T (i3d_gemm NT), n1/n2 = row norms, S = dot/(n1 n2),
 * pred = ((S+1)/2 > threshold).  part: [B,4] scratch.  out5 = (positive_similarity, negative_similarity,
 * true_positive_rate, true_negative_rate, contrastive_accuracy).  One pass instead of five [B,B] einsums.      */
int i3d_contrastive_metrics(const float* dot, int64_t B, const float* n1, const float* n2, float threshold,
                            float* part, float* out5, void* stream);
/* The other four metrics of configs_clean/pre-train_QM9.yml [trainer/metrics.py:161-176,212-230 with cov_loss /
 * uniformity_loss of commons/losses.py:946-964], x1 [B1, D], x2 [B2 >= B1, D] (extra rows of x2 are the appended noisy
 * samples: Alignment uses x2[:B1], the others all rows, as the reference does).  ws: >= 8 + 2 D doubles of scratch.
 * out4 = (dimension_covariance = cov_loss(x1) + cov_loss(x2),  batch_variance = x1.std(0).mean() + x2.std(0).mean(),
 *         alignment = mean_i ||x1_i - x2_i||^alpha,  uniformity = (log mean_{i<j} exp(-t d_ij^2) over x1 + same over x2) / 2;
 *         the reference's Uniformity always uses t = 2).  One pass over covariance / pair tiles instead of materialising
 *         [D, D] and B(B-1)/2 tensors per metric. */
int i3d_embedding_metrics(const float* x1, int64_t B1, const float* x2, int64_t B2, int D, float alpha, float t,
                          double* ws, float* out4, void* stream);
/* out[0] = scale * sum_i x[i]  (deterministic single-block reduction) */
int i3d_sum_scaled(const float* x, int64_t n, float scale, float* out, void* stream);
/* in: P (= p), dot recomputed as tau*log(p)*(n1n2+eps); out: P <- d loss / d dot; dn1[B], dn2[BcC] (caller-zeroed)
 * gscale = d total / d loss_i  is read from the device scalar gout[0] times inv_B                           */
int i3d_ntxent_rows_bwd(float* P, int64_t B, int64_t Bc, int C, const float* n1, const float* n2, int norm,
                        float eps, float tau, int64_t row_offset, const float* rowstats, const float* gout,
                        float inv_B, float* dn1, float* dn2, void* stream);
/* dz[r,:] += dn[r] * z[r,:] / norms[r] */
int i3d_norm_bwd_accum(const float* z, const float* norms, const float* dn, int64_t R, int D, float* dz,
                       void* stream);

/* ------------------------------------------------------------------------------------------------
 * Optimizer  [trainer/trainer.py:119-123 optim.step(); torch.optim.Adam semantics, L2 weight decay]
 * ---------------------------------------------------------------------------------------------- */
/* hyper_dev (optional, DEVICE double[6] = lr, beta1, beta2, eps, weight_decay, grad_scale) and step_dev
 * (optional, DEVICE int64) override the by-value arguments so that a captured CUDA graph can be replayed
 * while the schedule advances (WarmUpWrapper changes lr every step, trainer/lr_schedulers.py:30-52).     */
int i3d_adam_step(float* p, const float* g, float* m, float* v, int64_t n, double lr, double beta1, double beta2,
                  double eps, double weight_decay, double grad_scale, int64_t step, const double* hyper_dev,
                  const int64_t* step_dev, void* stream);
/* Data-parallel step fused over NVLink/NVSwitch multicast memory: gradient all-reduce + Adam + parameter broadcast in
 * ONE kernel (replaces NCCL all-reduce + i3d_adam_step; the DDP form of trainer/trainer.py:119-123).
 * local_base: this rank's symmetric allocation holding [p | g | m | v], n floats each (n % 4 == 0); mc_base: the
 * multicast address bound to the same allocation on all `world` ranks.  Rank r reduces slice r of g through the switch
 * (multimem.ld_reduce.add), applies Adam to that slice and multicast-stores the new p, m, v into every rank's buffers.
 * The caller orders the ranks: a cross-rank barrier on the stream before (all gradients written) and after (all
 * stores landed) this call.  Other arguments as i3d_adam_step. */
int i3d_adam_step_nvls(float* mc_base, float* local_base, int64_t n, int rank, int world, double lr, double beta1,
                       double beta2, double eps, double weight_decay, double grad_scale, int64_t step,
                       const double* hyper_dev, const int64_t* step_dev, void* stream);
int i3d_add_i64(int64_t* x, int64_t delta, void* stream);
/* flat[off[t] : off[t]+len[t]] = src_t (to_flat!=0) or the reverse.  ptrs/off/len: DEVICE arrays of T entries */
int i3d_multi_copy(const uint64_t* ptrs, const int64_t* off, const int64_t* len, int T, float* flat, int to_flat,
                   void* stream);
/* same with a per-tensor element stride on the tensor side (stride[t] >= 1; NULL = all 1): strided bias gradients */
int i3d_multi_copy_strided(const uint64_t* ptrs, const int64_t* off, const int64_t* len, const int64_t* stride, int T,
                           float* flat, int to_flat, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* I3D_H_ */
