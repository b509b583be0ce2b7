/*
 * mmdfn_b200 -- C ABI of the B200 (sm_100a) implementation of MM-DFN's per-dialogue
 * forward/backward hot path.
 *
 * The reference (zerohd4869/MM-DFN) has no FFI / plugin layer: its boundary is the Python
 * nn.Module API.  Each entry point below therefore replaces a group
 * of ATen library calls at the cited reference lines (paths relative to the reference
 * root); the Python drop-in modules under mm-dfn_b200/dropin/ bind them with ctypes
 * (INTEGRATION.md shows the stub).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer to fp32 row-major contiguous data unless stated
 *     (int = int32, long long = int64, unsigned char = keep-mask bytes);
 *   - the caller (PyTorch) owns every input, output, saved-for-backward and workspace
 *     buffer; nothing is allocated or freed here.  The only process-global state is a handful of
 *     host-side knobs read at enqueue time -- mmdfn_gru_set_tile, mmdfn_adj_spmm_set_variant,
 *     mmdfn_gemm_tc_set_variant, the *_set_debug profiling hooks -- and the launch counter; they
 *     are not thread-safe (the reference's trainer is single-threaded Python);
 *   - `stream` is a cudaStream_t; calls only enqueue work on it (asynchronous, re-entrant,
 *     CUDA-graph capturable);
 *   - return 0 on success, a positive cudaError_t value on a CUDA failure, a negative
 *     MMDFN_E* value on an argument error.  There is no CPU fallback.
 *   - node order of every (3N, .) array is the reference's stack [a; v; l]
 *     (code/model_mm.py:98); dialogue i owns rows dia_off[i] .. dia_off[i+1]-1 of each third.
 */
#ifndef MMDFN_B200_H_
#define MMDFN_B200_H_

#ifdef __cplusplus
extern "C" {
#endif

#define MMDFN_EINVAL (-1)
#define MMDFN_ENULL (-2)
#define MMDFN_ERANGE (-3)

/* Library identification: returns the ABI version (1). */
int mmdfn_abi_version(void);
/* Number of kernel launches this library has issued in this process (monotonic; for bench accounting). */
long long mmdfn_launch_count(void);

/* ---- k1 + every dense contraction --------------------------------------------------------
 * C[M,N] = act(alpha * op(A) op(B) + beta * C + bias[N]); op(A)(m,k) = transA ? A[k*lda+m] : A[m*lda+k];
 * op(B)(k,n) = transB ? B[n*ldb+k] : B[k*ldb+n]; act: 0 none, 1 ReLU.  transA && transB unsupported.
 * Replaces nn.Linear / torch.mm: code/model.py:1065,1094,1129,1337 ; code/model_GCN.py:186,454. */
int mmdfn_gemm(int transA, int transB, int M, int N, int K, float alpha, const float* A, long long lda,
               const float* B, long long ldb, float beta, float* C, long long ldc, const float* bias, int act,
               void* stream);
/* Same contract on the tcgen05 tensor cores (kind::tf32, accumulators in TMEM) with the 3-term TF32 split
 * (a_hi*b_hi + a_lo*b_hi + a_hi*b_lo): fp32-level accuracy at tensor-core rate.  mmdfn_gemm dispatches here
 * for large problems; exposed for tests and profiling. */
int mmdfn_gemm_tc(int transA, int transB, int M, int N, int K, float alpha, const float* A, long long lda,
                  const float* B, long long ldb, float beta, float* C, long long ldc, const float* bias, int act,
                  void* stream);
/* profiling aid: 128 x int64 device buffer receiving clock64() phase stamps of CTA 0 of the next mmdfn_gemm_tc launches (NULL = off) */
int mmdfn_gemm_tc_set_debug(long long* device_buf);
int mmdfn_gemm_tc_set_variant(int v);
/* debug aid: reads back the shared-memory word the tensor core uses for each operand element (see umma_probe.cu) */
int mmdfn_umma_probe(float* out, int N, int lbo, int sbo, int mn_major, int probe_a, void* stream);
/* debug aid: one product with the A operand in tensor memory; out (128 x 16) must read back 16 m + n */
int mmdfn_umma_probe_ta(float* out, int a_col, void* stream);
/* debug aid: cycles for `reps` and 2*reps back-to-back 128 x N MMAs (kind 0 = tf32, 1 = bf16; A from smem or tensor
 * memory; `issuers` threads of different warps issue concurrently into separate accumulators) */
int mmdfn_umma_rate(long long* out, int kind, int N, int a_tmem, int reps, int issuers, void* stream);
/* zero-fill `bytes` bytes at p with cudaMemsetAsync on `stream` (gradient buffers: a memset node instead of a fill kernel) */
int mmdfn_memset_zero(void* p, long long bytes, void* stream);
/* out[n] = beta*out[n] + sum_m A[m*lda+n]   (bias gradients) */
int mmdfn_colsum(int M, int N, const float* A, long long lda, float beta, float* out, void* stream);

/* Sequences per CTA of the recurrence kernels launched by the NEXT mmdfn_bigru2_fwd / _bwd calls: 2, 3, 4 or 8;
 * 0 (default) = chosen from the sequence count.  Lets the caller size two encoders that run concurrently on two
 * streams so that both fit in one wave of CTAs.  Host-side setting read at enqueue time (not stream-ordered). */
int mmdfn_gru_set_tile(int nb);
/* ---- k2: nn.GRU(200,100,num_layers=2,bidirectional=True), no packing, h0 = 0 -------------------
 * code/model.py:866 (lstm_l), :868 (rnn_parties); forward calls :1132, :1082, :1113, :1146.
 * x: (rows, 200) row table.  rowmap == NULL: rows == T*nseq and slot (t,s) reads row t*nseq+s.
 * rowmap (T, nseq) int32: slot reads row rowmap[t,s], -1 = all-zero input (speaker-party gather,
 * fused).  w: 16 pointers in nn.GRU state_dict order {weight_ih, weight_hh, bias_ih, bias_hh} x
 * {l0, l0_reverse, l1, l1_reverse}.  mask (T,nseq,200): inter-layer dropout keep mask or NULL.
 * y2: (T, nseq, 200).  ws: mmdfn_bigru2_ws_floats() floats, kept for the backward. */
long long mmdfn_bigru2_ws_floats(int T, int nseq, long long rows);
int mmdfn_bigru2_fwd(int T, int nseq, long long rows, const float* x, const int* rowmap, const float* const* w,
                     const unsigned char* mask, float mask_scale, float* y2, float* ws, void* stream);
long long mmdfn_bigru2_bwd_ws_floats(int T, int nseq, long long rows);
/* dx (rows,200): "=" or "+=" (accumulate_dx), may be NULL; dw: 16 gradient pointers.  dw_zeroed != 0: the caller
 * zero-filled them (one memset of a shared buffer) and gradients are accumulated, which saves the zero-init
 * launches of the split-K GEMMs; when weight_ih_l{k} and weight_ih_l{k}_reverse gradients are adjacent in memory
 * both directions are computed by one GEMM. */
int mmdfn_bigru2_bwd(int T, int nseq, long long rows, const float* x, const int* rowmap, const float* const* w,
                     const unsigned char* mask, float mask_scale, const float* y2, const float* dy2,
                     const float* ws_fwd, float* dx, int accumulate_dx, float* const* dw, int dw_zeroed, float* ws,
                     void* stream);
/* the same encoder with a layer-0 input width other than 200 (x: (rows, in_dim), weight_ih_l0[_reverse]: (300, in_dim),
   dx: (rows, in_dim)): the text-only configuration's `lstm` = nn.GRU(hidden_, D_e, 2, bidirectional) with
   hidden_ in {100, 150, 250} (code/model.py:836-849, 1035-1036).  Workspace sizes as for the 200-wide form. */
int mmdfn_bigru2_fwd_in(int in_dim, int T, int nseq, long long rows, const float* x, const int* rowmap,
                        const float* const* w, const unsigned char* mask, float mask_scale, float* y2, float* ws,
                        void* stream);
int mmdfn_bigru2_bwd_in(int in_dim, int T, int nseq, long long rows, const float* x, const int* rowmap,
                        const float* const* w, const unsigned char* mask, float mask_scale, const float* y2,
                        const float* dy2, const float* ws_fwd, float* dx, int accumulate_dx, float* const* dw,
                        int dw_zeroed, float* ws, void* stream);
/* the same backward in two calls that may run on different streams: _data = both recurrences, the layer-1 input gradient,
   the scatter and dx (the step's dependency chain; also accumulates the bias gradients); _wgrad = the weight-gradient
   contractions, reading the workspace the data part filled (the caller orders the two with an event). */
int mmdfn_bigru2_bwd_data(int in_dim, int T, int nseq, long long rows, const float* x, const int* rowmap,
                          const float* const* w, const unsigned char* mask, float mask_scale, const float* y2,
                          const float* dy2, const float* ws_fwd, float* dx, int accumulate_dx, float* const* dw,
                          int dw_zeroed, float* ws, void* stream);
int mmdfn_bigru2_bwd_wgrad(int in_dim, int T, int nseq, long long rows, const float* x, const int* rowmap,
                           const unsigned char* mask, const float* y2, const float* ws_fwd, float* const* dw,
                           int dw_zeroed, float* ws, void* stream);
/* the two calls by layer (parts: bit 1 = layer 1, bit 0 = layer 0; 3 = the calls above): the layer-1 weight gradients need
   only the layer-1 recurrence, so data(2), wgrad(2) [other stream], data(1), wgrad(1) [other stream] lets them run in the
   shadow of the layer-0 recurrence, which occupies a fraction of the SMs. */
int mmdfn_bigru2_bwd_data_part(int parts, int in_dim, int T, int nseq, long long rows, const float* x, const int* rowmap,
                               const float* const* w, const unsigned char* mask, float mask_scale, const float* y2,
                               const float* dy2, const float* ws_fwd, float* dx, int accumulate_dx, float* const* dw,
                               int dw_zeroed, float* ws, void* stream);
int mmdfn_bigru2_bwd_wgrad_part(int parts, int in_dim, int T, int nseq, long long rows, const float* x, const int* rowmap,
                                const unsigned char* mask, const float* y2, const float* ws_fwd, float* const* dw,
                                int dw_zeroed, float* ws, void* stream);

/* ---- k3/k4: speaker-party partition + fused scatter/combine/ragged pack ------------------------
 * code/model.py:1070-1090, 1101-1121, 1134-1154 and simple_batch_graphify :553-565.
 * qmask (T,B,S).  pos (T,B,S): rank of t among the non-zero entries of qmask[:,b,p] or -1;
 * cnt (B,S); sel (T,B): last p with qmask != 0 or -1; rowmap (T, 3*B*S) or NULL: row of the
 * stacked projection table (3,T,B,200) feeding slot (k, (m*B+b)*S+p).  Integer outputs are
 * bit-exact w.r.t. torch.nonzero ordering. */
int mmdfn_spk_partition(int T, int B, int S, const float* qmask, int* pos, int* cnt, int* sel, int* rowmap,
                        void* stream);
/* X[(m*N + dia_off[b] + t), :] = base_m[t,b,:] + w_m * Q[pos[t,b,sel], (m*B+b)*S+sel, :]  for t < L_b.
 * Q (T, 3*B*S, 200) may be NULL (use_crn_speaker off). */
int mmdfn_party_pack_fwd(int T, int B, int S, int N, const int* dia_off, const int* sel, const int* pos,
                         const float* base_a, const float* base_v, const float* base_l, const float* Q, float wa,
                         float wv, float wl, float* X, void* stream);
int mmdfn_party_pack_bwd(int T, int B, int S, int N, const int* dia_off, const int* sel, const int* pos,
                         const float* dX, float wa, float wv, float wl, float* dbase_a, float* dbase_v,
                         float* dbase_l, float* dQ, void* stream);

/* padded (T,B,D) view of one modality for the relation path's edge attention: rows t < L_b from the packed graph
 * input, the padded tail from pad_src (the padded encoder output); and its adjoint ("=" on both outputs). */
int mmdfn_unpack_pad_fwd(int T, int B, int D, const int* dia_off, const float* packed, const float* pad_src, float* out,
                         void* stream);
int mmdfn_unpack_pad_bwd(int T, int B, int D, const int* dia_off, const float* dM, float* d_packed, float* d_pad,
                         void* stream);

/* ---- k5: MM_GCN.create_big_adj in block-compact form (code/model_mm.py:122-180) ---------------
 * blk_off (B+1) int64: float offset of dialogue b's 3 blocks (3*L_b^2 floats each dialogue).
 * adj_blk: sum_b 3 L_b^2; adj_diag (3,N) pairs (a,v),(a,l),(v,l); dinv, rinv (3N); cos_blk, cos_diag
 * saved for the backward; deg_ws (3N) scratch. */
int mmdfn_adj_fwd(int B, int N, int Lmax, const int* dia_off, const long long* blk_off, const float* X,
                  float modal_weight, float* adj_blk, float* adj_diag, float* dinv, float* rinv, float* cos_blk,
                  float* cos_diag, float* deg_ws, void* stream);
/* d_blk / d_diag: gradient w.r.t. the stored entries (d_blk is clobbered).  dX = add + grad (add may be NULL). */
int mmdfn_adj_bwd(int B, int N, int Lmax, const int* dia_off, const long long* blk_off, const float* X,
                  float modal_weight, const float* adj_blk, const float* adj_diag, const float* dinv,
                  const float* rinv, const float* cos_blk, const float* cos_diag, float* d_blk, const float* d_diag,
                  const float* add, float* dX, float* dd_ws, void* stream);
/* dense (3N,3N) materialisation for callers that want the reference's tensor (not on the hot path) */
int mmdfn_adj_densify(int B, int N, int Lmax, const int* dia_off, const long long* blk_off, const float* adj_blk,
                      const float* adj_diag, float* dense, void* stream);

/* timing / cross-check aid: 0 (default) = for G == 100 the aggregate runs on the any-length tcgen05 kernel (128-row
 * tiles, streamed contraction; spmm_tc_long.cu), 1 = FFMA kernels only, 2 = the whole-block tcgen05 kernel
 * (spmm_tc.cu) when every dialogue has <= 128 utterances */
int mmdfn_adj_spmm_set_variant(int variant);
/* profiling aid: 64 x int64 device buffer receiving clock64() phase stamps of CTA 0 of the tcgen05 aggregate (NULL = off) */
int mmdfn_adj_spmm_set_debug(long long* device_buf);
/* ---- k6: message aggregate y = A_hat x, x,y (3N,G)  (torch.spmm, code/model_GCN.py:178) -------- */
int mmdfn_adj_spmm(int B, int N, int Lmax, const int* dia_off, const long long* blk_off, const float* adj_blk,
                   const float* adj_diag, const float* x, int G, float* y, void* stream);
/* d_blk[r,j] (+)= dhi_r . z_j ; d_diag[p][n] (+)= both orientations */
int mmdfn_adj_grad(int B, int N, int Lmax, const int* dia_off, const long long* blk_off, const float* dhi,
                   const float* z, int G, float* d_blk, float* d_diag, int accumulate, void* stream);

/* ---- k6 fused: one launch per GraphConvolution layer (code/model_GCN.py:176-189 + the ReLU / dropout / "+= q" of the
 * GCNII_lyc loop :469-472) for G = 100, any dialogue length -- tcgen05 aggregate chained with the weight product and
 * the epilogue on chip (gcn_layer.cu).  With theta_l = ln(lamda/l + 1):
 *   Mtop_l = theta_l W_l[0:100] + (1-theta_l)(1-alpha) I,   Mbot_l = theta_l W_l[100:200] + (1-theta_l) alpha I
 *   u = (A_hat zin) Mtop_l + h0 Mbot_l   ==   theta [hi|h0] W + (1-theta)((1-alpha) hi + alpha h0)
 * mmdfn_gcn_layer_prep builds, for K layers, mtop_all / mbot_all ((100, 100 K) row-major, column block l) and the
 * pre-split tensor-core operand images img_f / img_b (K x mmdfn_gcn_layer_img_floats() floats: forward / transposed).
 * The caller computes r = h0 Mbot_l (for all layers at once: R_all = h0 mbot_all, one mmdfn_gemm; ldr = 100 K).
 * fwd : out = dropout(relu((A_hat zin) Mtop + r)) (+ q), flags (3N,100) bytes = [relu and keep]; zin / q row stride 100,
 *       out row stride ldo, mask (3N,100) keep bytes or NULL.
 * bwd : t_out = A_hat du (row stride ldt), out = t_out Mtop^T (+ add); du row stride ldu.  (dMtop = zin^T t_out.) */
long long mmdfn_gcn_layer_img_floats(void);
int mmdfn_gcn_layer_prep(int K, const float* const* convW, double lamda, double alpha, float* mtop_all, float* mbot_all,
                         float* img_f, float* img_b, void* stream);
int mmdfn_gcn_layer_fwd(int B, int N, int Lmax, const int* dia_off, const long long* blk_off, const float* adj_blk,
                        const float* adj_diag, const float* zin, const float* wimg, const float* r, long long ldr,
                        const float* q, const unsigned char* mask, float mask_scale, unsigned char* flags, float* out,
                        long long ldo, void* stream);
int mmdfn_gcn_layer_bwd(int B, int N, int Lmax, const int* dia_off, const long long* blk_off, const float* adj_blk,
                        const float* adj_diag, const float* du, long long ldu, const float* wimg_t, float* t_out,
                        long long ldt, const float* add, float* out, void* stream);
/* timing aid (process-global, like mmdfn_adj_spmm_set_variant): 0 = default (second-generation persistent kernel with
 * tensor-memory A operands for batches whose dialogues have <= 128 utterances, first generation with 16-wide K chunks
 * otherwise), 1 = first generation, 8-wide chunks / 3 stages, 2 = first generation, 16-wide chunks / 2 stages.  Weight
 * images must be (re)built by mmdfn_gcn_layer_prep under the same setting. */
int mmdfn_gcn_layer_set_variant(int v);
/* profiling aid: 256 x int64 device buffer receiving clock64() stamps of CTA (0,0) (NULL switches it off) */
int mmdfn_gcn_layer_set_debug(long long* device_buf);

/* ---- k6/k7/k8: GCNII_lyc stack (code/model_GCN.py:444-488, GraphConvolution :176-189, LSTM :466)
 * X (3N,200) -> F (3N,300) = [dropout(X) | z_K].  convW: K pointers to (200,100).  masks: keep bytes
 * (3N,200), (3N,100), (K,3N,100) or NULL. */
long long mmdfn_gcn_stack_ws_floats(int n3, int K);
int mmdfn_gcn_stack_fwd(int B, int N, int Lmax, const int* dia_off, const long long* blk_off, const float* adj_blk,
                        const float* adj_diag, const float* X, int K, int reason_flag, double lamda, double alpha,
                        const float* W0, const float* b0, const float* const* convW, const float* w_ih,
                        const float* w_hh, const float* b_ih, const float* b_hh, const unsigned char* mask_x,
                        const unsigned char* mask_h0, const unsigned char* mask_layers, float mask_scale, float* F,
                        float* ws, void* stream);
long long mmdfn_gcn_stack_bwd_ws_floats(int n3, int K);
/* grads_zeroed != 0 (here and in mmdfn_head_bwd): every weight/bias gradient buffer was zero-filled by the caller
 * (one memset of a shared buffer) and is accumulated into -- no per-GEMM zero-init launches. */
int mmdfn_gcn_stack_bwd(int B, int N, int Lmax, const int* dia_off, const long long* blk_off, const float* adj_blk,
                        const float* adj_diag, int K, int reason_flag, double lamda, double alpha, const float* W0,
                        const float* const* convW, const float* w_ih, const float* w_hh,
                        const unsigned char* mask_x, const unsigned char* mask_h0, const unsigned char* mask_layers,
                        float mask_scale, const float* F, const float* ws_fwd, const float* dF, float* dX,
                        float* d_adj_blk, float* d_adj_diag, float* dW0, float* db0, float* const* dconvW,
                        float* dw_ih, float* dw_hh, float* db_ih, float* db_hh, int grads_zeroed, float* ws,
                        void* stream);

/* ---- k9: head + loss (code/model.py:1328-1337 ; code/loss.py:14-34) ---------------------------
 * F (3N,300); mask (N,900) keep bytes or NULL; Wc (C,900); R (3N,300) saved; log_prob (N,C). C <= 16.
 * relu = 1 on the GDF path (:1329), 0 on the relation path (:1241-1242: dropout -> smax_fc, no ReLU). */
int mmdfn_head_fwd(int N, int C, const float* F, const unsigned char* mask, float mask_scale, int relu,
                   const float* Wc, const float* bc, float* R, float* log_prob, void* stream);
int mmdfn_head_bwd(int N, int C, const unsigned char* mask, float mask_scale, int relu, const float* Wc, const float* R,
                   const float* log_prob, const float* dlog_prob, float* dF, float* dWc, float* dbc,
                   int grads_zeroed, float* dlogits_ws, void* stream);
int mmdfn_focal_loss_fwd(int N, int C, const float* log_prob, const long long* target, const float* alpha,
                         float gamma, int size_average, float* loss, void* stream);
int mmdfn_focal_loss_bwd(int N, int C, const float* log_prob, const long long* target, const float* alpha,
                         float gamma, int size_average, const float* dloss, float* dlog_prob, void* stream);

/* device-side evaluation metrics (replaces the per-batch argmax -> .cpu() -> sklearn of code/run_train_erc.py:202-203,
 * 228-235): pred (N) int64 (nullable) = argmax over classes (first maximum), conf (C*C) uint64 (nullable) accumulates
 * conf[target*C + pred] += 1 over as many batches as the caller likes. */
int mmdfn_confusion_accumulate(int N, int C, const float* log_prob, const long long* target, long long* pred,
                               unsigned long long* conf, void* stream);

/* ---- a13 (star row): Memory Fusion Network block -- MFN.forward (code/model_fusion.py:62-120) and its backward --------
 * x (T, n, 900) = [l | a | v] -> out (T, n, 400) = [h_l | h_a | h_v | mem].  params: the first 28 tensors of the module's
 * state_dict in order (lstm_{l,a,v}.{weight_ih, weight_hh, bias_ih, bias_hh}, att1_fc1/2, att2_fc1/2, gamma1_fc1/2,
 * gamma2_fc1/2: weight, bias; out_fc1 / out_fc2 are constructed by the reference but never used).  masks: NULL (eval) or
 * 4 uint8 keep masks (T n, 100) for the block's four Dropout(0.2) layers (after att1_fc1, att2_fc1, gamma1_fc1,
 * gamma2_fc1), mask_scale = 1 / (1 - 0.2).  ws: mmdfn_mfn_ws_floats floats, kept for the backward.  dparams receive "=". */
long long mmdfn_mfn_ws_floats(int T, int n);
long long mmdfn_mfn_bwd_ws_floats(int T, int n);
int mmdfn_mfn_fwd(int T, int n, const float* x, const float* const* params, const unsigned char* const* masks,
                  float mask_scale, float* out, float* ws, void* stream);
int mmdfn_mfn_bwd(int T, int n, const float* x, const float* const* params, const unsigned char* const* masks,
                  float mask_scale, const float* out, const float* ws_fwd, const float* dout, float* dx,
                  float* const* dparams, float* ws, void* stream);
/* glue of the 'mfn' head (code/model.py:1263-1291, 1303-1330): stacked node features F (3N, 300) -> padded time-major
 * window x (T, B, 900) with x[.., 300 j ..] taken from modality block p_j; MFN output (T, B, 400) -> node rows (N, 400);
 * y = relu(x) * keep * scale (Dropout then ReLU) and its backward dx = y != 0 ? dy * ind_scale : 0. */
int mmdfn_mfn_pack_fwd(int T, int B, int N, const int* dia_off, int p0, int p1, int p2, const float* F, float* x, void* stream);
int mmdfn_mfn_pack_bwd(int T, int B, int N, const int* dia_off, int p0, int p1, int p2, const float* dx, float* dF, void* stream);
int mmdfn_mfn_unpad_fwd(int T, int B, const int* dia_off, const float* out, float* feat, void* stream);
int mmdfn_mfn_unpad_bwd(int T, int B, const int* dia_off, const float* dfeat, float* dout, void* stream);
int mmdfn_relu_mask_fwd(long long n, const float* x, const unsigned char* mask, float scale, float* y, void* stream);
int mmdfn_relu_mask_bwd(long long n, const float* dy, const float* y, float ind_scale, float* dx, void* stream);

/* ---- k10/k11 (relation graph type): edges and masked edge attention ----------------------------
 * edge_perms + batch_graphify (code/model.py:532-550, 568-611) and MaskedEdgeAttention 'attn1' (:449-471).
 * Canonical edge order: dialogue, source j, target i ascending; edge_off (B+1) int64 = per-dialogue edge
 * offsets (host-computed from the lengths), E = edge_off[B].  edge_index (2,E) int64 [row 0 = source j,
 * row 1 = target i, both with the dialogue's node offset], edge_type (E) int64 = 2*(S*spk_j+spk_i)+(j>=i),
 * row_ptr (N+1) int64: first edge of each source node, node_dia (N) int32: dialogue of each node. */
int mmdfn_edges_build(int T, int B, int S, int N, int window_past, int window_future, const int* dia_off,
                      const long long* edge_off, long long E, const float* qmask, long long* edge_index,
                      long long* edge_type, long long* row_ptr, int* node_dia, void* stream);
/* s_all (T*B, ncol) = M W_att[:ncol]^T (mmdfn_gemm), ncol >= Lmax.  edge_norm (E) in edge order;
 * stat (N,2) = {row max, denominator} saved for the backward. */
int mmdfn_edge_attn_fwd(int T, int B, int Lmax, int ncol, int window_past, int window_future, const int* dia_off,
                        const long long* row_ptr, const float* s_all, float* edge_norm, float* stat, void* stream);
int mmdfn_edge_attn_bwd(int T, int B, int Lmax, int ncol, int window_past, int window_future, const int* dia_off,
                        const long long* row_ptr, const float* s_all, const float* edge_norm, const float* stat,
                        const float* g_norm, float* ds_all, void* stream);
/* dense (B, msl, T) <-> compact (E) edge scores: gather = 0 scatters (dense zero-filled first), 1 gathers */
int mmdfn_edge_scores_dense(long long E, int B, int msl, int T, const long long* edge_index, const int* node_dia,
                            const int* dia_off, float* edge_norm, float* dense, int gather, void* stream);

/* ---- k12 (relation graph type): RGCNConv -> GraphConv of GraphNetwork (code/model.py:675-715; arithmetic of
 * torch-geometric 1.4.3, restated -- parity unpinned).  Windowed edges make every node's in- and out-edges contiguous
 * ranges, so aggregation is a warp-per-node gather-reduce without atomics.  node_spk (N) int32 = speaker of each node
 * (first p with qmask == 1).  xw (N, R, G) = x W_r for all relations (one mmdfn_gemm); norm (E) edge weights. */
int mmdfn_node_speakers(int B, int S, const int* dia_off, const float* qmask, int* node_spk, void* stream);
/* out[i,:] += sum_{j->i} norm[e] * xw[j, type(e), :] */
int mmdfn_rgcn_aggregate_fwd(int N, int G, int R, int S, int window_past, int window_future, const int* dia_off,
                             const int* node_dia, const int* node_spk, const long long* row_ptr, const float* xw,
                             const float* norm, float* out, void* stream);
/* dxw (N,R,G) "=" (zero-filled here), dnorm (E) "=" */
int mmdfn_rgcn_aggregate_bwd(int N, int G, int R, int S, int window_past, int window_future, const int* dia_off,
                             const int* node_dia, const int* node_spk, const long long* row_ptr, const float* xw,
                             const float* norm, const float* dout, float* dxw, float* dnorm, void* stream);
/* out[i,:] (+)= sum of h[j,:] over j in [i-reach_back, i+reach_fwd] inside i's dialogue (-1 = unbounded):
 * GraphConv forward uses (window_future, window_past), its backward (window_past, window_future). */
int mmdfn_window_sum(int N, int G, int reach_back, int reach_fwd, const int* dia_off, const int* node_dia,
                     const float* h, float* out, int accumulate, void* stream);
/* out[b][c][r] = in[b][r][c] */
int mmdfn_transpose_batched(int batch, int rows, int cols, const float* in, float* out, void* stream);

/* ---- ☆ f4: low-rank multimodal fusion (LMF, code/model_fusion.py:214-310; att_type = 'lmf_only') -------------------
 * out = sum_r w_r prod_m ([1, h_m] . factor_m[r]) + bias with h: three (N, H) activations {audio, video, text} (the
 * sub-network Linears run on mmdfn_gemm), factor: three (R, H + 1, O) tensors (row 0 = the weight of the appended 1),
 * w (R) = fusion_weights, bias (O); R <= 8.  fz / dfz: 3 R N O floats (mmdfn_lmf_ws_floats); fz is saved for backward.
 * _bwd overwrites dh[m], dfactor[m], dbias and ACCUMULATES into dw (zero it first). */
long long mmdfn_lmf_ws_floats(int N, int R, int O);
int mmdfn_lmf_fuse_fwd(int N, int H, int O, int R, const float* const* h, const float* const* factor, const float* w,
                       const float* bias, float* fz, float* out, void* stream);
int mmdfn_lmf_fuse_bwd(int N, int H, int O, int R, const float* const* h, const float* const* factor, const float* w,
                       const float* fz, const float* dout, float* dfz, float* const* dh, float* const* dfactor, float* dw,
                       float* dbias, void* stream);

/* ---- ☆ f4: tensor fusion network (TFN, code/model_fusion.py:123-211; att_type = 'tfn_only') ------------------------
 * y1 = dropout([1,h_a] (x) [1,h_v] (x) [1,h_t]) W1^T + b1 with h_m (N, 100) from the three sub-network Linears and
 * W1 (300, 101^3 = 1030301).  The 4 MB-per-row fusion tensor exists only for one chunk of mmdfn_tfn_chunk_rows() rows at a
 * time (ws: mmdfn_tfn_ws_floats(N) floats); its dropout keep bits are a function of (seed, offset + n * 1030301 + e), the
 * same in both passes (p = 0: none).  _bwd takes dy1 = d/d(pre-activation), overwrites dha / dhv / dht (N, 100) and db1
 * (300), overwrites or (accumulate != 0) accumulates into dW1; d1: scratch of 3 N 101 floats. */
int mmdfn_tfn_chunk_rows();
long long mmdfn_tfn_ws_floats(int N);
int mmdfn_tfn_fuse_fwd(int N, const float* ha, const float* hv, const float* ht, const float* W1, const float* b1, float p,
                       unsigned long long seed, unsigned long long offset, float* y1, float* ws, void* stream);
int mmdfn_tfn_fuse_bwd(int N, const float* ha, const float* hv, const float* ht, const float* W1, float p,
                       unsigned long long seed, unsigned long long offset, const float* dy1, float* dha, float* dhv,
                       float* dht, float* dW1, float* db1, int accumulate, float* d1, float* ws, void* stream);

/* ---- k13 (☆ SURVEY 8f rank 3): nodal attention of the relation path's classifier head ---------------------------
 * Replaces attentive_node_features + MatchingAttention('general2') (code/model.py:614-645, 66-76) on the ragged node
 * rows: per dialogue b (rows dia_off[b]..dia_off[b+1]-1 of E, Q, O: (N, D), D <= 512)
 *     S = tanh(Q_b E_b^T), P = softmax_rows(S), O_b = P E_b          with Q = E W^T + bias (mmdfn_gemm, by the caller).
 * P, S, dA: sum_b L_b^2 floats, block b row-major at sq_off[b]; row_dia (N): dialogue index of every row.
 * _bwd overwrites dQ = dA E and dE = P^T dO + dA^T Q (the caller adds dQ W and forms dW = dQ^T E, db = colsum dQ). */
int mmdfn_nodal_attn_fwd(int B, int N, int D, int Lmax, const int* dia_off, const long long* sq_off, const int* row_dia,
                         const float* E, const float* Q, float* P, float* S, float* O, void* stream);
int mmdfn_nodal_attn_bwd(int B, int N, int D, int Lmax, const int* dia_off, const long long* sq_off, const int* row_dia,
                         const float* E, const float* Q, const float* P, const float* S, const float* dO, float* dA,
                         float* dQ, float* dE, void* stream);

/* ---- a13: MMGatedAttention 'general' (code/model.py:757-781) -------------------------------------
 * h_m = tanh(P_m), P_m = W_m x_m + b_m (computed by the caller with mmdfn_gemm); z_mn = sigmoid(w_mn.[x_m, x_n, x_m*x_n]
 * + b_mn) for the pairs av, al, vl; out (N, 3C) = [z_av h_a + (1-z_av) h_v | z_al h_a + (1-z_al) h_l | z_vl h_v +
 * (1-z_vl) h_l].  x* (N, D) are the (dropped-out) inputs, w (3, 3D) = transform_av/al/vl.weight stacked, b (3) their
 * biases, z (N, 3) is saved for the backward.  Backward: dP* (N, C) and dx* (N, D) (the gate path only; the GEMMs'
 * backward adds the projection path), dw (3, 3D), db (3) are overwritten; dzpre_ws is an (N, 3) scratch. */
int mmdfn_gated_fuse_fwd(int N, int D, int C, const float* xa, const float* xv, const float* xl, const float* Pa,
                         const float* Pv, const float* Pl, const float* w, const float* b, float* out, float* z,
                         void* stream);
int mmdfn_gated_fuse_bwd(int N, int D, int C, const float* dout, const float* xa, const float* xv, const float* xl,
                         const float* Pa, const float* Pv, const float* Pl, const float* w, const float* b,
                         const float* z, float* dPa, float* dPv, float* dPl, float* dxa, float* dxv, float* dxl,
                         float* dw, float* db, float* dzpre_ws, void* stream);
/* log_softmax over the class axis and its backward (the relation path's 'gated' head, code/model.py:1238-1239) */
int mmdfn_log_softmax_fwd(int N, int C, const float* logits, float* log_prob, void* stream);
int mmdfn_log_softmax_bwd(int N, int C, const float* log_prob, const float* dlog_prob, float* dlogits, void* stream);
/* y = mask ? x * scale : 0 (nn.Dropout with an explicit keep mask, code/model.py:742-744; its own backward) */
int mmdfn_mask_scale(long long n, const float* x, const unsigned char* mask, float scale, float* y, void* stream);

/* ---- support: dropout keep masks, fused flat-buffer Adam(+L2) (code/run_train_erc.py:512) ----- */
int mmdfn_dropout_mask(long long n, float p, unsigned long long seed, unsigned long long offset,
                       unsigned char* mask, void* stream);
int mmdfn_adam_step(long long n, float* param, const float* grad, float* exp_avg, float* exp_avg_sq, float lr,
                    float beta1, float beta2, float eps, float weight_decay, int step, float grad_scale,
                    void* stream);
/* CUDA-graph friendly variants: everything that changes from step to step is read from a device-resident state
 * (2 x uint64: [0] = optimizer steps taken, [1] = two floats {1 - beta1^t, sqrt(1 - beta2^t)} of the step in flight).
 * mmdfn_step_advance increments [0] and refreshes [1] (call it once per step, before mmdfn_adam_step_dev);
 * mmdfn_dropout_mask_dev draws with counter base offset + state[0] * per_step. */
int mmdfn_step_advance(unsigned long long* state, float beta1, float beta2, void* stream);
int mmdfn_dropout_mask_dev(long long n, float p, unsigned long long seed, const unsigned long long* state,
                           unsigned long long per_step, unsigned long long offset, unsigned char* mask,
                           void* stream);
int mmdfn_adam_step_dev(long long n, float* param, const float* grad, float* exp_avg, float* exp_avg_sq, float lr,
                        float beta1, float beta2, float eps, float weight_decay, const unsigned long long* state,
                        float grad_scale, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MMDFN_B200_H_ */
