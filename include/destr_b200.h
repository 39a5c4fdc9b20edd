/* destr_b200.h -- C ABI of libdestr_b200.so: the B200 (sm_100a) kernels behind the DESTR
 * transformer-half hot path.
 *
 * The reference (mio0115/object_detection_destr) is pure Python/PyTorch and has no FFI; the
 * "plugin interface" of this path is its nn.Module / function API.  Each entry point below names
 * the reference code (file:line, relative to the reference root) whose arithmetic it replaces.
 * The Python host layer (object_detection_destr_b200/) binds these with ctypes and exposes
 * drop-in modules with the reference's own names and signatures.
 *
 * Conventions
 *  - plain pointers and sizes only; every pointer is a DEVICE pointer unless named host_*.
 *  - bf16 tensors are passed as `const void*` / `void*` (16-bit storage), fp32 as float*.
 *  - the caller owns every buffer (inputs, outputs, workspaces); kernels never allocate.
 *  - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, nothing syncs.
 *  - return value: 0 = ok; nonzero = error, message via destr_last_error() (thread-local).
 *  - token-major activations: row r = b*N + n (encoder tokens) or b*Q + q (object queries).
 */
#ifndef DESTR_B200_H_
#define DESTR_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

int destr_version(void);
const char* destr_last_error(void);
/* bring-up only: overrides UMMA descriptor fields (tools/gpu_check.py); not part of the product API */
int destr_debug_knob(int idx, int value);

/* ---------------- masks and positional embeddings ---------------- */

/* Pack a key-padding mask (1 = padded; encoder_block.py:31, self_attention.py:34-37) into
 * bit words for the attention kernels: bits[b][w] bit i = key (32w+i) is masked OR >= n_keys.
 * kpm: uint8 [B, n_keys] (may be NULL = nothing masked); bits: uint32 [B, words_per_row],
 * words_per_row >= ceil(n_keys/128)*4. */
int destr_pack_key_mask(const uint8_t* kpm, uint32_t* bits, int B, int n_keys, int words_per_row, void* stream);

/* PositionEmbeddingSine(num_pos_feats=128, normalize=True) (position_encoding_cdetr.py:39-63).
 * mask uint8 [B,H,W] (1 = padded) -> pos token-major [B, H*W, 256] (fp32 and/or bf16; either
 * pointer may be NULL). */
int destr_sine_pos2d(const uint8_t* mask, float* pos_f32, void* pos_bf16, int B, int H, int W, void* stream);

/* gen_sineembed_for_position (positional_embedding.py:6-39), d_model = 256.
 * centers fp32 [M,2] (x,y) -> out [M,256] (fp32 and/or bf16). */
int destr_query_sine_embed(const float* centers, float* out_f32, void* out_bf16, int M, void* stream);

/* ---------------- encoder elementwise / normalisation ---------------- */

/* y = x + pos * s   (encoder_block.py:38,95;  to_q_k = inputs + pos_embed*scale), bf16 [M,256]. */
int destr_pos_mul_add_fwd(const void* x, const void* pos, const void* s, void* y, int64_t n_elem, void* stream);
/* backward: dx = dy (aliasing is the caller's business), ds = dy * pos */
int destr_pos_mul_add_bwd(const void* dy, const void* pos, void* ds, int64_t n_elem, void* stream);
/* dx_out = dx_in + dy ; ds = dy*pos  (pos_mul_add backward with the residual gradient folded in) */
int destr_pos_mul_add_bwd_acc(const void* dy, const void* pos, const void* dx_in, void* ds, void* dx_out,
                              int64_t n_elem, void* stream);
/* ReLU backward fused with the bias gradient of the Linear in front of it (encoder_block.py:107,
 * decoder_block.py:255): dpre = dy * (h > 0), dbias[c] += sum_rows dpre[:,c] (fp32, accumulated).
 * With h = dpre = NULL it is a plain column sum of dy (bias gradient of a Linear without activation).
 * bf16 [M,C] operands with row pitches lddy/ldh/ldo. */
int destr_relu_bwd_colsum(const void* dy, int lddy, const void* h, int ldh, void* dpre, int ldo, float* dbias, int M,
                          int C, float scale, void* stream);
/*   scale multiplies dpre: 1/(1-p) when a dropout followed the ReLU (h is then the DROPPED activation, whose zeros
 *   already mask the dropped elements), else 1. */
/* x <- dropout(x) in place, bf16 [M,C] with row pitch ld: the dropout that follows a fused GEMM+ReLU
 * (encoder_block.py:108 dropout2; decoder_block.py:255). */
int destr_dropout_inplace(void* x, int ld, int M, int C, const uint32_t* drop_seed, uint32_t drop_thr16,
                          uint32_t drop_site, void* stream);
/* y = a * b, bf16 (fine_pos = pos * pos_scale(enc_out), model.py:89-92; decoder_block.py:49) */
int destr_mul_fwd(const void* a, const void* b, void* y, int64_t n_elem, void* stream);

/* y = LayerNorm(a + b) * gamma + beta over the last dim D (256 or 512), eps 1e-5
 * (encoder_block.py:104-110, :40; decoder_block.py:65, 253-258).  a, b, y bf16 rows of D channels
 * with row pitches lda/ldb/ldy elements (so the 256-wide cls/reg halves of a [M,512] tensor can be
 * normalised in place, decoder_block.py:185-187,218); b may be NULL; gamma/beta fp32 [D].
 * Saves mean/rstd (fp32 [M]) for the backward when the pointers are non-NULL. */
int destr_add_layernorm_fwd(const void* a, int lda, const void* b, int ldb, const float* gamma, const float* beta,
                            void* y, int ldy, float* mean, float* rstd, int M, int D, const uint32_t* drop_seed,
                            uint32_t drop_thr16, uint32_t drop_site, void* stream);
/* DROPOUT ARGUMENTS (every entry point that has them): the reference's dropout sites (encoder_block.py:67-69,
 * 104-109; decoder_block.py:132,182-184,234,253-256; self_attention.py:40) are applied INSIDE the kernels from a
 * counter-based mask keep(seed, site, row, col): drop_seed = DEVICE pointer to a 32-bit seed (may change between
 * CUDA-graph replays), drop_thr16 = round(p * 65536) (0 = no dropout), drop_site = id of the dropout call (the
 * forward and the backward of one site must pass the same).  Kept values are scaled by 65536 / (65536 - thr16).
 * Here: y = LN(a + dropout(b)). */
/* Two chained LayerNorms in one pass (D = 256): y1 = LN1(a + dropout(b)), y2 = LN2(c + y1) with both sets of row
 * statistics -- `norm2(x + dropout3(fc2 ..))` followed by the Encoder's shared `norm(x + block(x))`
 * (encoder_block.py:108-110, :40).  y1 enters the second LayerNorm as stored (bf16). */
int destr_add_layernorm2_fwd(const void* a, int lda, const void* b, int ldb, const float* gamma1, const float* beta1,
                             void* y1, int ldy1, float* mean1, float* rstd1, const void* c, int ldc,
                             const float* gamma2, const float* beta2, void* y2, int ldy2, float* mean2, float* rstd2,
                             int M, int D, const uint32_t* drop_seed, uint32_t drop_thr16, uint32_t drop_site,
                             void* stream);

/* Backward of destr_add_layernorm2_fwd in one pass (dense [M,256] bf16 operands): d3 = gradient w.r.t. (c + y1) (also
 * the residual-stream gradient of c), dxb = gradient w.r.t. b through its dropout mask, dsum (may be NULL) = the
 * un-masked gradient w.r.t. (a + dropout(b)); dgamma2/dbeta2 (outer LayerNorm), dgamma1/dbeta1 (inner) and dbias
 * (column sums of dxb, may be NULL) are ACCUMULATED into. */
int destr_add_layernorm2_bwd(const void* dy, const void* c, const void* y1, const float* gamma2, const float* mean2,
                             const float* rstd2, const void* a, const void* b, const float* gamma1, const float* mean1,
                             const float* rstd1, void* d3, void* dxb, void* dsum, float* dgamma2, float* dbeta2,
                             float* dgamma1, float* dbeta1, float* dbias, int M, int D, const uint32_t* drop_seed,
                             uint32_t drop_thr16, uint32_t drop_site, void* stream);

/* backward of y = LN(a+b): dx (bf16, pitch lddx) = d(a+b); dgamma/dbeta fp32 [D] are ACCUMULATED into
 * (caller zeroes).  a+b is recomputed from a and b. */
int destr_add_layernorm_bwd(const void* dy, int lddy, const void* a, int lda, const void* b, int ldb,
                            const float* gamma, const float* mean, const float* rstd, void* dx, int lddx,
                            float* dgamma, float* dbeta, float* dbias, const void* res_in, int ldri, void* res_out,
                            int ldro, int M, int D, const uint32_t* drop_seed, uint32_t drop_thr16,
                            uint32_t drop_site, void* stream);
/*   dx = gradient w.r.t. b (through b's dropout mask when one is active).  Optional fusions: dbias (fp32 [D],
 *   accumulated) = column sum of dx = gradient of the bias of the Linear that produced b;  res_out = [res_in +]
 *   d(a+b) = gradient flowing on into the residual stream (NOT masked; res_in may be NULL). */

/* out = lam*LN1(x+o1) + (1-lam)*LN2(x+o2eff)  (decoder_block.py:182-184), D = 512, fused with the
 * head-group slot masking of PairSelfAttention (pair_self_attention.py:101-105):
 *   o2eff[c] = [pairs[row,0]==i]*o2[row,c] + [pairs[row,1]==i]*o2[row,512+c],  i = row % Q.
 * x, o1, out bf16 [M,512]; o2 bf16 [M,1024] (head-major pair-attention output); pairs int32 [M,2];
 * stats fp32 [M,4] = mean1,rstd1,mean2,rstd2 (saved for the backward, may be NULL). */
int destr_dual_ln_mix_fwd(const void* x, const void* o1, const void* o2, const int32_t* pairs, const float* g1,
                          const float* b1, const float* g2, const float* b2, float lam, void* out, float* stats,
                          int M, int Q, const uint32_t* drop_seed, uint32_t drop_thr16, uint32_t drop_site1,
                          uint32_t drop_site2, void* stream);
/*   dropout1(o1) uses drop_site1, dropout1(o2eff) drop_site2 (two independent masks, decoder_block.py:182-184) */
/* backward: dx bf16 [M,512]; do1 bf16 [M,512] and do2 bf16 [M,1024] token-major, or (head_major = 1)
 * do1 [B,8,Q,64] and do2 [B,8,Q,128] as the attention backward wants them; dg1,db1,dg2,db2 fp32 [512] are
 * ACCUMULATED into.  With head_major, delta1/delta2 (fp32 [B,8,Q], may be NULL) receive rowsum(dO o O) per
 * head for the self / pair attention backward. */
int destr_dual_ln_mix_bwd(const void* dout, const void* x, const void* o1, const void* o2, const int32_t* pairs,
                          const float* g1, const float* g2, const float* stats, float lam, void* dx, void* do1,
                          void* do2, float* dg1, float* db1, float* dg2, float* db2, int M, int Q, int head_major,
                          float* delta1, float* delta2, const uint32_t* drop_seed, uint32_t drop_thr16,
                          uint32_t drop_site1, uint32_t drop_site2, void* stream);

/* ---------------- encoder multi-head self-attention (tcgen05) ---------------- */

/* Fused QK^T / softmax / PV of nn.MultiheadAttention as used at encoder_block.py:97-103
 * (torch functional.py multi_head_attention_forward: q scaled by 1/sqrt(d_head), bool
 * key-padding mask -> -inf, softmax, PV), flash style: the N x N score matrix is never written.
 *   q, k, v : bf16 token-major, row r = b*N + n, head h occupies columns [h*32, h*32+32);
 *             ld_* = row pitch in elements (q and k may alias one [M,512] projection output).
 *   mask_bits : from destr_pack_key_mask (words_per_row as given there).
 *   out     : bf16 [B*N, heads*32], heads merged (the layout out_proj consumes).
 *   lse     : fp32 [B, heads, N], log2-domain log-sum-exp of scale*log2e*scores (for backward).
 *   scale   : softmax scale (1/sqrt(32) for the encoder).
 * d_head is fixed at 32, heads*32 <= 256.  Needs sm_100a (tcgen05/TMEM/TMA). */
int destr_enc_attn_fwd(const void* q, const void* k, const void* v, int ld_q, int ld_k, int ld_v,
                       const uint32_t* mask_bits, int words_per_row, void* out, float* lse, int B, int N,
                       int heads, float scale, const uint32_t* drop_rowbits, int drop_words, uint32_t drop_thr16,
                       void* stream);
/*   dropout acts on the attention probabilities, as nn.MultiheadAttention(dropout=p) does in training
 *   (encoder_block.py:58-60): mask row = (b*heads + h)*N + query, column = key, the same counter-based mask as
 *   destr_add_layernorm_fwd's; the softmax denominator is not dropped.  The two flash kernels read the mask as bit
 *   matrices (1 = dropped) written once per layer by destr_attn_dropout_bits -- evaluating the hash per score inside
 *   their issue-bound loops cost 40-80% of the kernel time:
 *     rowbits [n_sites][B*heads][words][Np]  bit i of word (w, q) <-> key   32w+i  (forward;  Np = 128*ceil(N/128))
 *     colbits [n_sites][B*heads][words][Np]  bit i of word (w, k) <-> query 32w+i  (backward)
 *   words >= max(Np/32, 3*ceil(N/96)); either output may be NULL; only the ceil(N/32)^2 blocks of real (query, key)
 *   pairs are written.  One call can fill the matrices of n_sites dropout sites (site = site0 + i*site_stride, e.g.
 *   all encoder layers of a step); the attention calls take one [B*heads][words][Np] slice.  drop_thr16 = 0 in the
 *   attention calls disables dropout. */
int destr_attn_dropout_bits(const uint32_t* drop_seed, uint32_t drop_thr16, uint32_t drop_site0,
                            uint32_t drop_site_stride, int n_sites, int BH, int N, int words, uint32_t* rowbits,
                            uint32_t* colbits, void* stream);
/* backward: dq, dk, dv bf16 with the same row pitches as q, k, v (ld_dq, ld_dk, ld_dv).
 * Caller-provided workspaces: stats (fp32, destr_enc_attn_bwd_stats_floats(B,N,heads) elements: -lse and
 * -delta = -rowsum(dO o O), padded to 128-query tiles) and dq_acc (fp32 [B*N, heads*32]). */
int destr_enc_attn_bwd_stats_floats(int B, int N, int heads);
int destr_enc_attn_bwd(const void* q, const void* k, const void* v, int ld_q, int ld_k, int ld_v,
                       const uint32_t* mask_bits, int words_per_row, const void* out, const void* dout,
                       const float* lse, float* stats, float* dq_acc, void* dq, void* dk, void* dv, int ld_dq,
                       int ld_dk, int ld_dv, int B, int N, int heads, float scale, const uint32_t* drop_colbits,
                       int drop_words, uint32_t drop_thr16, void* stream);

/* ---------------- decoder: pairing, self + pair attention, split cross-attention ---------------- */

/* _get_pairs (pair_self_attention.py:110-171).  coords fp32 [B,Q,4] (cx,cy,h,w) -> pairs
 * int32 [B,Q,2].  fp32 arithmetic in the reference's operation order (index parity). */
int destr_pair_indices(const float* coords, int32_t* pairs, int B, int Q, void* stream);

/* sigmoid(delta + [logit(cx), logit(cy), 0, 0]) (decoder_block.py:41,51-54; model.py:123-129;
 * inverse_sigmoid misc.py:59-62).  delta fp32 [M,4], centers fp32 [M,2] -> boxes fp32 [M,4]. */
int destr_box_refine(const float* delta, const float* centers, float* boxes, int M, void* stream);

/* The box head's output layer fused with the refinement: boxes = sigmoid(hidden W2^T + b2 + [logit(cx), logit(cy), 0, 0])
 * (decoder_block.py:51-54 with bbox_embed[2] = Linear(256, 4), fp32 weights): hidden bf16 [M,256] (pitch ldh) is the
 * post-ReLU output of bbox_embed[0]; W2 fp32 [4,256], b2 fp32 [4]; boxes fp32 [M,4]. */
int destr_box_head_refine(const void* hidden, int ldh, const float* W2, const float* b2, const float* centers,
                          float* boxes, int M, void* stream);

/* SA operand preparation (decoder_block.py:167-177) + the left/right gathers of pair attention
 * (pair_self_attention.py:47-89) in one pass.
 *   qkv_obj bf16 [B*Q,1536] = [W_q x | W_k x | W_v x];  qk_pos bf16 [B*Q,512] (row pitch ld_pos) = [W_qp p | W_kp p]
 *   -> qkv bf16 [3][B,8,Q,64]  head-major: x_0 = q_obj+[qp|qp], x_1 = k_obj+[kp|kp], x_2 = v
 *   -> cat bf16 [3][B,8,Q,128] head-major: cat[w][b,h,i, 0..64) = x_w[b,h,L_i], [64..128) = x_w[b,h,R_i]
 *      with (L_i, R_i) = pairs[b,i] (indices within the image). */
int destr_dec_qkv_prep(const void* qkv_obj, const void* qk_pos, int ld_pos, const int32_t* pairs, void* qkv,
                       void* cat, int B, int Q, void* stream);

/* Decoder self-attention (self_attention.py:26-45, 8 heads x 64, scale 1/8) and pair self-attention
 * (pair_self_attention.py:91-99: softmax(Ql.Kl^T + Qr.Kr^T)/sqrt(128) . [Vl|Vr]) in ONE launch on
 * tcgen05.  qkv / cat head-major as written by destr_dec_qkv_prep.  o1 bf16 [B*Q,512]; o2 bf16 [B*Q,1024]
 * (head-major, before the slot masking that destr_dual_ln_mix_fwd applies).  lse1/lse2 fp32 [B,8,Q]
 * (log2 domain, for the backward; may be NULL).  Q <= 384. */
int destr_dec_self_pair_attn_fwd(const void* qkv, const void* cat, void* o1, void* o2, float* lse1, float* lse2,
                                 int B, int Q, const uint32_t* drop_seed, uint32_t drop_thr16,
    uint32_t drop_site, void* stream);

/* Backward of destr_dec_self_pair_attn_fwd, stage 1 (one launch, tcgen05): recomputes S and dP = dO.v^T and
 * applies the softmax backward.  qkv / cat / do1 [B,8,Q,64] / do2 [B,8,Q,128] head-major; lse*, delta*
 * fp32 [B,8,Q] (delta = rowsum(dO o O), from destr_dual_ln_mix_bwd).  Outputs bf16 [B,8,Q,Qp],
 * Qp = ceil(Q/128)*128 (keys >= Q written as 0): P1,dS1 (self) and P2,dS2 (pair).  The caller finishes with
 * batched GEMMs: dV = P^T dO, dQ = dS K, dK = dS^T Q (see ops.dec_self_pair_attn_bwd). */
int destr_dec_self_pair_attn_bwd_ds(const void* qkv, const void* cat, const void* do1, const void* do2,
                                    const float* lse1, const float* lse2, const float* delta1, const float* delta2,
                                    void* P1, void* dS1, void* P2, void* dS2, int B, int Q, const uint32_t* drop_seed, uint32_t drop_thr16,
    uint32_t drop_site, void* stream);
/* FUSED backward of destr_dec_self_pair_attn_fwd for Q <= 128 (the training shape, 100 queries): S and dP are
 * recomputed on the tensor cores, the softmax backward runs in registers, P and dS stay in shared memory as bf16 tiles
 * and dQ = dS K, dK = dS^T Q, dV = P^T dO are tcgen05 products in the same kernel -- nothing but the gradients is
 * written (autograd of self_attention.py:26-45 with 8 x 64 heads and of pair_self_attention.py:91-99).
 *   qkv [3][B,8,Q,64], cat [3][B,8,Q,128], do1 [B,8,Q,64], do2 [B,8,Q,128] head-major bf16; lse / delta fp32 [B,8,Q]
 *   d_qkv [3][B,8,Q,64], d_cat [3][B,8,Q,128] bf16 out.  Q > 128: destr_dec_self_pair_attn_bwd_ds + batched GEMMs. */
int destr_dec_self_pair_attn_bwd(const void* qkv, const void* cat, const void* do1, const void* do2, const float* lse1,
                                 const float* lse2, const float* delta1, const float* delta2, void* d_qkv, void* d_cat,
                                 int B, int Q, const uint32_t* drop_seed, uint32_t drop_thr16, uint32_t drop_site,
                                 void* stream);

/* Backward of destr_dec_qkv_prep (gather formulation of the scatter-add, deterministic): head-major
 * d_qkv [3][B,8,Q,64], d_cat [3][B,8,Q,128] -> d_qkv_obj bf16 [B*Q,1536], d_qk_pos bf16 [B*Q,512] (pitch ld_pos). */
int destr_dec_qkv_prep_bwd(const void* d_qkv, const void* d_cat, const int32_t* pairs, void* d_qkv_obj,
                           void* d_qk_pos, int ld_pos, int B, int Q, void* stream);

/* Split cross-attention of both ClsRegBranch'es (decoder_block.py:212-217, 246-251 ->
 * self_attention.py:26-45 with one head, d_qk = 512, d_v = 256, scale 1/sqrt(512)).
 * Uses score = <q_obj, k_enc> + <q_pos, k_pos> (the per-head interleave of decoder_block.py:
 * 195-210 is one permutation applied to both q and k, so it cancels in the dot product).
 *   q_obj : bf16 [B*Q, 512]   cls half = cols [0,256), reg half = cols [256,512)
 *   q_pos : bf16 [B*Q, 256]   shared by both branches
 *   k_enc, k_pos, v : bf16 [B*N, 256] (row pitch ld_kv elements; may be slices of one wide GEMM output)
 *   out   : bf16 [B*Q, 512]   cls result in cols [0,256), reg result in cols [256,512)
 *   lse   : fp32 [B, 2, Q] (log2 domain)
 *   ws_partial : fp32 workspace of destr_split_cross_attn_ws_floats(B,Q,N) floats (split-KV partials) */
int64_t destr_split_cross_attn_ws_floats(int B, int Q, int N);
int destr_split_cross_attn_fwd(const void* q_obj, const void* q_pos, const void* k_enc, const void* k_pos,
                               const void* v, int ld_kenc, int ld_kpos, int ld_v, const uint32_t* mask_bits,
                               int words_per_row, void* out, float* lse, float* ws_partial, int B, int Q, int N,
                               float scale, const uint32_t* drop_seed, uint32_t drop_thr16,
    uint32_t drop_site, void* stream);

/* Backward of destr_split_cross_attn_fwd, stage 1: recomputes S and dP = dO.V^T on tcgen05 and applies
 * the softmax backward.  Rows are ordered (2q + br):
 *   P_all, dS_all : bf16 [B, 2Q, Np]   (Np = ceil(N/128)*128; padded / masked keys are written as 0)
 *   dS_sum        : bf16 [B,  Q, Np]   = dS_cls + dS_reg
 *   delta         : fp32 [B, 2, Q] workspace (rowsum(dO o O), computed here)
 * The caller finishes with five batched GEMMs on plain views (see ops.split_cross_attn_bwd):
 *   dV = P^T dO, dK_enc = dS^T q_obj, dK_pos = dSsum^T q_pos, dq_obj = dS k_enc, dq_pos = dSsum k_pos. */
int destr_split_cross_attn_bwd_ds(const void* q_obj, const void* q_pos, const void* k_enc, const void* k_pos,
                                  const void* v, int ld_kenc, int ld_kpos, int ld_v, const uint32_t* mask_bits,
                                  int words_per_row, const void* out, const void* dout, const float* lse,
                                  float* delta, void* P_all, void* dS_all, void* dS_sum, int B, int Q, int N,
                                  float scale, const uint32_t* drop_seed, uint32_t drop_thr16,
    uint32_t drop_site, void* stream);

/* FUSED backward of destr_split_cross_attn_fwd for Q <= 128: as the _ds entry point, but P never leaves the chip and
 * the key-side gradients are finished inside the kernel -- dv = sum_br P_br^T dO_br, dk_enc = sum_br dS_br^T q_obj_br,
 * dk_pos = sum_br dS_br^T q_pos, complete per key tile, stored bf16 into dk_enc / dk_pos / dv ([B*N,256] views with
 * row pitches ld_*: e.g. column slices of the packed d_kv_all / d_kpos_all buffers).  dS_all [B,2Q,Np] (rows 2q+br) and
 * dS_sum [B,Q,Np] (bf16, Np = 128*ceil(N/128), pad columns zero) are written for the query-side products
 * dq_obj = dS k_enc, dq_pos = dS_sum k_pos (destr_gemm_bf16_batched). */
int destr_split_cross_attn_bwd_fused(const void* q_obj, const void* q_pos, const void* k_enc, const void* k_pos,
                                     const void* v, int ld_kenc, int ld_kpos, int ld_v, const uint32_t* mask_bits,
                                     int words_per_row, const void* out, const void* dout, const float* lse,
                                     float* delta, void* dS_all, void* dS_sum, void* dk_enc, int ld_dke, void* dk_pos,
                                     int ld_dkp, void* dv, int ld_dv, int B, int Q, int N, float scale,
                                     const uint32_t* drop_seed, uint32_t drop_thr16, uint32_t drop_site, void* stream);

/* ---------------- set-prediction cost matrix ---------------- */

/* Matching cost of HungarianMatcher / HungarianMatcherWoL1 (matcher.py:72-107, 158-184;
 * complete_iou bbox_utils.py:160-198), computed ONLY for the per-image diagonal blocks that
 * linear_sum_assignment consumes (matcher.py:109-112), fp32, in the reference's operation order.
 *   logits fp32 [B,Q,C]; boxes fp32 [B,Q,4] cxcyhw; tgt_ids int32 [T]; tgt_boxes fp32 [T,4] xyxy;
 *   tgt_offsets int32 [B+1] (prefix sums of T_i); cost fp32, image b's block is row-major
 *   [Q, T_b] starting at element Q*tgt_offsets[b].
 *   with_l1 = 1 adds w_bbox * L1(pred cxcyhw, tgt xyxy) (matcher.py:96). */
int destr_match_cost_blockdiag(const float* logits, const float* boxes, const int32_t* tgt_ids,
                               const float* tgt_boxes, const int32_t* tgt_offsets, float* cost, int B, int Q, int C,
                               float w_class, float w_bbox, float w_ciou, int with_l1, void* stream);

/* Per-image linear sum assignment on the block-diagonal cost buffer of destr_match_cost_blockdiag, bit-identical
 * to scipy.optimize.linear_sum_assignment (matcher.py:109-112, 186-189), one warp per image.
 *   cost / tgt_offsets: as above; max_targets >= max_b T_b.
 *   pred_idx, tgt_idx int64 [B, n_slots], valid uint8 [B, n_slots], n_slots >= min(Q, max_targets): image b's
 *   min(Q,T_b) pairs in ascending query order (scipy's row order), padded slots = (Q, 0, 0).
 *   status int32 [B]: 0 ok, 1 = NaN/-inf in the block (scipy raises ValueError), 2 = infeasible. */
int destr_lsap_blockdiag(const float* cost, const int32_t* tgt_offsets, int B, int Q, int max_targets, int n_slots,
                         int64_t* pred_idx, int64_t* tgt_idx, uint8_t* valid, int32_t* status, void* stream);

/* ---------------- set-prediction loss (after matching) ---------------- */

/* SetCriterion.forward after the matcher (criterion.py:29-79; sigmoid_focal_loss misc.py:99-128, L1Loss,
 * CompleteIOULoss criterion.py:82-89 = mean of the FULL n x n complete_iou matrix), forward and backward fused:
 *   logits fp32 [B,Q,C]; boxes fp32 [B,Q,4] cxcyhw; tgt_labels int64 [B,t_max] and tgt_boxes fp32 [B,t_max,4]
 *   xyxy (per-image targets, padded); pred_idx / tgt_idx int64 [B,n] and valid uint8 [B,n]: the matching
 *   (linear_sum_assignment rows / columns, padded slots have valid = 0).
 *   losses fp32 [4] = {class, bbox, ciou, w_class*class + w_bbox*bbox + w_ciou*ciou};
 *   dlogits fp32 [B,Q,C], dboxes fp32 [B,Q,4] = gradient of losses[3] (torch autograd conventions).
 *   workspace: 3*B + 1 floats, zero-initialised ONCE by the caller (the kernel leaves it reusable). */
int destr_set_loss_fwd_bwd(const float* logits, const float* boxes, const int64_t* tgt_labels,
                           const float* tgt_boxes, const int64_t* pred_idx, const int64_t* tgt_idx,
                           const uint8_t* valid, int B, int Q, int C, int t_max, int n, float w_class,
                           float w_bbox, float w_ciou, float* losses, float* dlogits, float* dboxes,
                           float* workspace, void* stream);

/* ---------------- FFN first layer: GEMM + bias + ReLU + dropout (tcgen05) ---------------- */

/* out = dropout(relu(a w^T + bias)) in bf16: `dropout2(relu(fc1(x)))` of encoder_block.py:107-108 and the branch FFNs
 * of decoder_block.py:255 as one tcgen05 GEMM whose epilogue does the whole tail (first member).
 *   a bf16 [M,K], row pitch lda;  w bf16 [N,K] (nn.Linear layout);  bias fp32 [N] or NULL;  out bf16 [M,N], pitch ldo
 *   K == 256 (the weight block of a CTA stays resident in shared memory), N % 256 == 0, pitches multiples of 8
 *   elements;  relu = 0 skips the ReLU
 *   dropout: the mask of destr_dropout_inplace on the [M,N] output (row = output row, column = output column). */
int destr_linear_bias_relu_dropout(const void* a, int lda, const void* w, const float* bias, void* out, int ldo,
                                   int M, int N, int K, int relu, const uint32_t* drop_seed, uint32_t drop_thr16,
                                   uint32_t drop_site, void* stream);

/* ---------------- tcgen05 GEMM family with fused epilogues (csrc/gemm_tc.cu) ---------------- */

/* C[M,N] = A[M,K] . op(B) with the tail of the surrounding layer in the epilogue -- replaces nn.Linear / its dX product
 * plus the elementwise kernels around it (encoder_block.py:24-44, 88-112 and their autograd).
 *   a bf16 [M,K] row-major, pitch lda.  b_kn = 0: b bf16 [N,K] row-major (nn.Linear weight; forward y = x W^T)
 *                                       b_kn = 1: b bf16 [K,N] row-major (the same weight used for dX = dY W)
 *   x    = act(acc + bias) [relu = 1], then the dropout mask of destr_dropout_inplace on the [M,N] output
 *   out  = add + mul * x     (mul, add: bf16 [M,N] or NULL)       e.g. x + pos * pos_scale(x), dX + residual gradient
 *   out2 = add2 + x          (out2 NULL: not written)             e.g. ds = dxq * pos (out) together with dx += dxq (out2)
 *   N % 32 == 0, K % 8 == 0, every pitch a multiple of 8 elements; out may alias add (in-place accumulation). */
int destr_gemm_bf16(const void* a, int lda, const void* b, int ldb, int b_kn, int M, int N, int K, const float* bias,
                    int relu, const uint32_t* drop_seed, uint32_t drop_thr16, uint32_t drop_site, const void* mul,
                    int ldmul, const void* add, int ldadd, void* out, int ldo, const void* add2, int ldadd2, void* out2,
                    int ldo2, void* stream);

/* Backward of `dropout(relu(fc1 x))` seen from fc2 (encoder_block.py:107-109): dpre = scale * (dy w) * (h > 0) with
 * w bf16 [K,N] row-major (fc2.weight: [256, 2048]), h the saved (dropped) activation bf16 [M,N] whose zeros carry both
 * the ReLU and the dropout mask, scale = 1/(1-p); dbias[n] += column sums of dpre (fc1.bias gradient; NULL to skip). */
int destr_gemm_relu_bwd(const void* dy, int lddy, const void* w, int ldw, int M, int N, int K, const void* h, int ldh,
                        float scale, void* dpre, int ldo, float* dbias, void* stream);

/* Linear + dropout + residual + LayerNorm in one kernel, N = 256 (encoder_block.py:104-106 `norm1(x + dropout1(attn))`,
 * :108-110 `norm2(x + dropout3(fc2 ..))` and, through res2, the encoder's shared `norm(x + blk(x))` of :40):
 *   z = res + dropout(a w^T + bias);  y = LN(z; gamma, beta);  mean/rstd fp32 [M]
 *   res2 != NULL:  y2 = LN(res2 + y; gamma2, beta2), mean2/rstd2
 *   w bf16 [256,K]; z (bf16, may be NULL) is what destr_add_layernorm_bwd(a = z, b = NULL, drop) needs. */
int destr_gemm_res_ln(const void* a, int lda, const void* w, int ldw, int M, int K, const float* bias,
                      const uint32_t* drop_seed, uint32_t drop_thr16, uint32_t drop_site, const void* res, int ldres,
                      const float* gamma, const float* beta, void* z, int ldz, void* y, int ldy, float* mean,
                      float* rstd, const void* res2, int ldres2, const float* gamma2, const float* beta2, void* y2,
                      int ldy2, float* mean2, float* rstd2, void* stream);

/* `batch` independent products out_i[M,N] = a_i[M,K] . op(b_i) in one launch (plain store epilogue).  a is ONE
 * row-major matrix holding the a_i stacked every a_batch_rows rows, b likewise every b_batch_rows rows (b_kn as above),
 * out the stack of the [M,N] results.  Used for the query side of the split cross-attention backward:
 * dq_obj_b = dS_b k_enc_b, dq_pos_b = (dS_cls + dS_reg)_b k_pos_b per image b (decoder_block.py:189-217 autograd). */
int destr_gemm_bf16_batched(const void* a, int lda, int a_batch_rows, const void* b, int ldb, int b_batch_rows,
                            int b_kn, int batch, int M, int N, int K, void* out, int ldo, void* stream);

/* Two batched products with the same (N, K, batch, b_kn) -- e.g. dq_obj and dq_pos of the split cross-attention
 * backward -- in ONE launch (the second problem rides on gridDim.z): they are independent, latency-bound and small. */
int destr_gemm_bf16_batched2(const void* a1, int lda1, int a1_batch_rows, const void* b1, int ldb1, int b1_batch_rows,
                             int M1, void* out1, int ldo1, const void* a2, int lda2, int a2_batch_rows, const void* b2,
                             int ldb2, int b2_batch_rows, int M2, void* out2, int ldo2, int b_kn, int batch, int N,
                             int K, void* stream);

/* Weight gradient dw[Nout,Kin] += dy[M,Nout]^T x[M,Kin] (fp32 accumulation INTO dw with red.global.add: zero it first),
 * split over the M rows so small weight matrices still fill the GPU.  dy, x bf16 row-major; dw fp32, pitch lddw. */
int destr_gemm_dw(const void* dy, int lddy, const void* x, int ldx, int M, int Nout, int Kin, float* dw, int lddw,
                  void* stream);

/* ---------------- mini-detector query selection ---------------- */

/* MiniDetector.get_topk_index + the gathers of MiniDetector.forward (mini_detector.py:70-104, 142-170), two small launches:
 *   scores fp32 [B,N,C]   class scores as passed at :156 (sigmoid'ed, padded rows zeroed); key = max over classes
 *   mask   u8 [B,N] (1 = padded) or NULL;   cls_feat, reg_feat fp32 [B,N,D] (D % 4 == 0);   coords fp32 [B,N,4]
 *   k      = min(top_k, N, valid positions of image 0), computed by the caller as the reference does (:153-154)
 *   topk_idx int64 [B,k]: positions by decreasing key (equal keys: lower position first -- the reference's order
 *            inside a tie is torch.topk's unspecified one); an image with valid < k un-padded positions keeps its
 *            top `valid` and fills slot s >= valid with idx[valid-1-(s % valid)] (:86-98)
 *   sel_f32 [B,k,2D] and/or sel_bf16 (either may be NULL) = [cls_feat | reg_feat] rows; centers fp32 [B,k,2]
 *   status int32 [B]: 1 = image without valid positions (the reference raises ZeroDivisionError).
 *   key_ws: scratch of B*N + B 32-bit words. */
int destr_select_queries(const float* scores, const uint8_t* mask, const float* cls_feat, const float* reg_feat,
                         const float* coords, int B, int N, int C, int D, int k, int64_t* topk_idx, float* sel_f32,
                         void* sel_bf16, float* centers, int32_t* status, uint32_t* key_ws, void* stream);

/* ---------------- prediction heads ---------------- */

/* Class and box heads of the detector (model.py:120-131) on the decoder output, fp32 weights and arithmetic:
 *   logits = dec[:, :256] Wc^T + bc;   boxes = sigmoid(W2 relu(W1 dec[:, 256:] + b1) + b2 + [inverse_sigmoid(centers), 0, 0])
 *   dec bf16 [M,512]; centers fp32 [M,2]; Wc [C,256], W1 [256,256], W2 [4,256] row-major as nn.Linear stores them
 *   (C <= 128); logits fp32 [M,C], boxes fp32 [M,4] (cx,cy,h,w in (0,1)); hidden fp32 [M,256] is kept for backward.
 * Backward: d_dec bf16 [M,512] and the six parameter gradients (overwritten, not accumulated) from dlogits, dboxes;
 * dh_ws (M*256 floats) and dz_ws (M*4 floats) are scratch. */
int destr_heads_fwd(const void* dec, const float* centers, const float* Wc, const float* bc, int C, const float* W1,
                    const float* b1, const float* W2, const float* b2, float* logits, float* boxes, float* hidden,
                    int M, void* stream);
/* which: 1 = row gradients only (d_dec + the dh / dz workspaces), 2 = parameter gradients only (reads the workspaces a
 * which = 1 call filled: lets the caller run them off the critical path, on another stream), 3 = both. */
int destr_heads_bwd(const void* dec, const float* hidden, const float* boxes, const float* dlogits,
                    const float* dboxes, const float* Wc, const float* W1, const float* W2, int C, void* d_dec,
                    float* dh_ws, float* dz_ws, float* dWc, float* dbc, float* dW1, float* db1, float* dW2,
                    float* db2, int M, int which, void* stream);

/* ---------------- optimizer step on the flat parameter buffer ---------------- */

/* torch.optim.AdamW arithmetic (decoupled weight decay, bias correction; amsgrad off) on flat fp32 buffers of
 * n elements (n % 4 == 0, 16-byte aligned), plus the refresh of the bf16 weight shadow the GEMMs read.
 * `step` points to a device float holding the 1-based step count of THIS update.
 * The gradients of parameters [bf16_begin, n_bf16) (both % 4 == 0; n_bf16 = 0: none) are read from `grad_bf16` instead of
 * `grad` -- weight-matrix gradients left in bf16 by library dW GEMMs or by a bf16 gradient exchange; destr_gemm_dw
 * accumulates in fp32 straight into `grad` -- and every gradient is multiplied by grad_scale (1/world for the
 * data-parallel mean, else 1); the fp32 value actually used is written back to `grad`, so that buffer (the
 * parameters' .grad) is complete after the call. */
int destr_flat_adamw(float* master, float* grad, float* exp_avg, float* exp_avg_sq, void* shadow_bf16, int64_t n,
                     float lr, float beta1, float beta2, float eps, float weight_decay, const float* step,
                     const void* grad_bf16, int64_t bf16_begin, int64_t n_bf16, float grad_scale, void* stream);

/* Batch hand-over into the static inputs of a captured training step (engine.py: load_batch / step_prefetched; the
 * reference feeds a fresh batch per iteration through its DataLoader, src/train/train.py:164-170): n <= 16
 * device-to-device copies in ONE launch.  `src`, `dst`, `bytes` are HOST arrays of n device pointers / byte counts;
 * buffers may overlap in nothing; any alignment (16-byte aligned pairs take the vector path). */
int destr_copy_many(const void* const* src, void* const* dst, const int64_t* bytes, int n, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DESTR_B200_H_ */
