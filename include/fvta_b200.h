/*
 * fvta_b200.h -- C ABI of the B200-native FVTA hot path (libfvta_b200.so).
 *
 * The reference (JunweiLiang/FVTA_MemexQA) has no FFI: its hot path sits behind Python call
 * signatures executed by TensorFlow.  Each entry point below replaces the TensorFlow ops of
 * one reference function; the citation is the reference file:line it stands in for.  The
 * Python host layer (fvta_memexqa_b200/ops.py) binds these with ctypes and re-exposes the
 * reference's own signatures (attention_3d, attention, Trainer.step, ...).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host; all tensors are
 *     contiguous row-major fp32 (masks: uint8, 0/1; indices: int32);
 *   - the caller owns every buffer; nothing here allocates, frees or synchronises;
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*);
 *   - return value: FVTA_OK (0) or a negative FVTA_E_* code; never throws;
 *   - re-entrant across streams and devices; no global state (fvta_last_error is per thread,
 *     fvta_launch_count a monotonic diagnostic counter).
 */
#ifndef FVTA_B200_H
#define FVTA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FVTA_OK 0
#define FVTA_E_BADARG (-1)      /* null pointer / non-positive dimension */
#define FVTA_E_UNSUPPORTED (-2) /* unknown simiMatrix / warp_type (reference: sys.exit / raise) */
#define FVTA_E_WORKSPACE (-3)   /* workspace too small */
#define FVTA_E_LAUNCH (-4)      /* cudaGetLastError() != cudaSuccess after a launch */
#define FVTA_E_SHAPE (-5)       /* shape outside what the kernels were built for */

int fvta_version(void);
/* Last CUDA error string seen by this library on the calling thread ("" if none). */
const char* fvta_last_error(void);
/* Kernels this library has launched in this process so far (diagnostic; bench.py's gpu_launches). */
unsigned long long fvta_launch_count(void);

/* ------------------------------------------------------------------------------------------
 * Focal visual-text attention: time-warp gate + similarity + masks + two-level softmax +
 * weighted sum.  Replaces model_v2.py:953-1009 (time warp), :301-341 (time_indication_func),
 * :210-298 (attention_3d), :23-48 (softmax/softsel), utils.py:210-213 (exp_mask).
 * With K=1, use_time_warp=0 it is model_v2.py:125-201 `attention` (bidirect=False).
 * ---------------------------------------------------------------------------------------- */
typedef struct {
  int32_t N;             /* QA samples */
  int32_t K;             /* modalities ("sequences"): 6, or 5 with --no_photo; 1 for attention() */
  int32_t T;             /* M * JMAX context positions per modality */
  int32_t D;             /* 2 * hidden_size */
  int32_t JQ;            /* question positions */
  int32_t simi;          /* --simiMatrix 1..4 (model_v2.py:241-257) */
  int32_t add_tanh;      /* --add_tanh */
  int32_t use_time_warp; /* --use_time_warp */
  int32_t warp_type;     /* --warp_type 1..5 (model_v2.py:301-341) */
  int32_t window;        /* ceil(window_t) for warp_type 5 (model_v2.py:333-337); 3 */
  int32_t time_warp_att; /* --use_time_warp_att (model_v2.py:269-275) */
  int32_t has_mask;      /* 0: hinfo_mask/hq_mask are None (exp_mask skipped, model_v2.py:263) */
  int32_t ws_from_fwd;   /* 1 (2: incl. the tensor-core question images, i.e. the forward ran FVTA_PATH_TC) =
                            `workspace` still holds what an earlier fvta_attn_fwd call left there for
                            the same hq / hq_mask / lq / parameters, so the operand folding is not redone
                            (fvta_attn_bwd after its forward; fvta_attn_fwd over a new context with the
                            same question, e.g. repeated inference) */
  int32_t path;          /* forward kernel selection: FVTA_PATH_AUTO, or FVTA_PATH_FFMA to pin the
                            CUDA-core tile kernel (A/B parity checks).  Per call, not process state */
} FvtaAttnDesc;

#define FVTA_PATH_AUTO 0
#define FVTA_PATH_SMALL 1 /* one fused CTA per sample (question attention: K = 1, JQ = 1, no time warp) */
#define FVTA_PATH_FFMA 2  /* CUDA-core tile kernel + combine kernels */
#define FVTA_PATH_TC 3    /* tcgen05 / TMEM / TMA tile kernel with the combine fused in */

typedef struct {
  const float* att_W; /* att_logits/W [rows,1], rows = 3D,2D,4D,0 for simi 1..4 */
  const float* att_b; /* att_logits/b [1] */
  const float* WH_W;  /* time_warp/WH/W [2D,D]  (model_v2.py:986) */
  const float* WH_b;  /* time_warp/WH/b [D] */
  const float* WC_W;  /* time_warp/WC/W [D,1]   (model_v2.py:989) */
  const float* WC_b;  /* time_warp/WC/b [1] */
} FvtaAttnParams;

typedef struct {      /* gradients, same shapes as FvtaAttnParams; each is OVERWRITTEN */
  float* att_W;
  float* att_b;
  float* WH_W;
  float* WH_b;
  float* WC_W;
  float* WC_b;
} FvtaAttnGrads;

/* Which kernel fvta_attn_fwd runs for this descriptor (a pure function of its arguments): one of
 * FVTA_PATH_SMALL / FVTA_PATH_FFMA / FVTA_PATH_TC, or a negative FVTA_E_* code.  want_c_col /
 * want_warp_h: whether those optional outputs will be requested. */
int fvta_attn_fwd_path(const FvtaAttnDesc* desc, int want_c_col, int want_warp_h);

/* Bytes of scratch the fwd/bwd calls need, and floats of `stats` fwd saves for bwd. */
size_t fvta_attn_workspace_bytes(const FvtaAttnDesc* desc);
size_t fvta_attn_stats_floats(const FvtaAttnDesc* desc);

/*
 * hall      [N,K,T,D]   un-warped context tensor (model_v2.py:905-912 reshaped :974)
 * hall_mask [N,K,T]     uint8
 * hq        [N,JQ,D]    question encoder outputs; hq_mask [N,JQ] uint8
 * lq        [N,D]       question final state (only read when use_time_warp)
 * g1_all    [N,D]       out: attended vector            (attention_3d's h_a)
 * a_logits  [N,K,T,JQ]  out, nullable: masked logits    (attention_3d's a_logits)
 * c_col     [N,T]       out, nullable: tanh(z) = any column of model.C (model_v2.py:995)
 * warp_h    [N,K,T,D]   out, nullable: gated context    (model.warp_h, model_v2.py:1009)
 * stats     fvta_attn_stats_floats() floats, out: softmax statistics for the backward
 */
int fvta_attn_fwd(const FvtaAttnDesc* desc, const float* hall, const uint8_t* hall_mask,
                  const float* hq, const uint8_t* hq_mask, const float* lq,
                  const FvtaAttnParams* params, float* g1_all, float* a_logits, float* c_col,
                  float* warp_h, float* stats, void* workspace, size_t workspace_bytes,
                  void* stream);

/*
 * Backward of fvta_attn_fwd w.r.t. hall, hq, lq and the parameters, given d_g1_all [N,D].
 * (The reference differentiates only through g1_all: a_logits is returned for visualisation.)
 * d_hall [N,K,T,D], d_hq [N,JQ,D], d_lq [N,D] and every non-null member of `grads` are
 * overwritten.  d_lq may be null when use_time_warp == 0.
 */
int fvta_attn_bwd(const FvtaAttnDesc* desc, const float* hall, const uint8_t* hall_mask,
                  const float* hq, const uint8_t* hq_mask, const float* lq,
                  const FvtaAttnParams* params, const float* stats, const float* d_g1_all,
                  float* d_hall, float* d_hq, float* d_lq, const FvtaAttnGrads* grads,
                  void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * Bidirectional LSTM reader over packed variable-length sequences.
 * Replaces tf.nn.bidirectional_dynamic_rnn(BasicLSTMCell(d)) at the eight call sites
 * model_v2.py:694-823 (gate order i,j,f,o; forget_bias 1; zero outputs past the length;
 * final state = state at the last valid step; backward direction reverses the valid prefix).
 * ---------------------------------------------------------------------------------------- */
#define FVTA_LSTM_AUTO 0   /* the faster measured recurrence for the shape (today: CUDA cores at d = 50) */
#define FVTA_LSTM_FFMA 1   /* packed-fp32 CUDA-core recurrence, W_h in registers */
#define FVTA_LSTM_TC 2     /* forward recurrence on tcgen05 (d <= 64): W_h^T . h^T as split-fp16 UMMAs */

typedef struct {
  int32_t B;     /* sequences */
  int32_t J;     /* padded length */
  int32_t In;    /* input width */
  int32_t d;     /* hidden size */
  int32_t path;  /* FVTA_LSTM_AUTO / _FFMA / _TC: forward recurrence kernel (per call, not process state) */
} FvtaLstmDesc;

size_t fvta_bilstm_workspace_bytes(const FvtaLstmDesc* desc);
size_t fvta_bilstm_reserve_floats(const FvtaLstmDesc* desc);

/*
 * x       [B,J,In]; lens [B] int32
 * kernel_fw/bw [In+d,4d], bias_fw/bw [4d]   (the reference shares one cell: pass the same
 *                                            pointers for both directions)
 * out     [B,J,2d]  fw in [:d], bw in [d:], zeros past the length
 * last    [B,2d]    final h of each direction
 * reserve fvta_bilstm_reserve_floats() floats kept for the backward (gate activations, c)
 */
int fvta_bilstm_fwd(const FvtaLstmDesc* desc, const float* x, const int32_t* lens,
                    const float* kernel_fw, const float* bias_fw, const float* kernel_bw,
                    const float* bias_bw, float* out, float* last, float* reserve,
                    void* workspace, size_t workspace_bytes, void* stream);

/*
 * d_out [B,J,2d], d_last [B,2d] (either nullable = zero) -> d_x [B,J,In] (overwritten),
 * d_kernel_fw/bw [In+d,4d], d_bias_fw/bw [4d] ACCUMULATED (+=) so that the shared-weight case
 * and the seven text call sites can add into one gradient slot.
 */
int fvta_bilstm_bwd(const FvtaLstmDesc* desc, const float* x, const int32_t* lens,
                    const float* kernel_fw, const float* kernel_bw, const float* out,
                    const float* reserve, const float* d_out, const float* d_last, float* d_x,
                    float* d_kernel_fw, float* d_bias_fw, float* d_kernel_bw, float* d_bias_bw,
                    void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * Packed reader: the same BiLSTM over sequences whose valid tokens are PACKED (no padding rows)
 * and whose outputs are written straight into an output arena at per-sequence row offsets --
 * which is how all seven text call sites of model_v2.py:694-812 (one shared cell) run as ONE
 * launch and land directly in the context tensor hall[N,K,M,JMAX,2d] of model_v2.py:862-914
 * (the tf.pad / tf.stack copies disappear).  fw and bw share `kernel`/`bias`.
 * ---------------------------------------------------------------------------------------- */
typedef struct {
  int32_t B;     /* sequences */
  int32_t ntok;  /* packed token rows (>= sum of lengths; extra rows are ignored) */
  int32_t In;    /* input width */
  int32_t d;     /* hidden size */
  int32_t Jmax;  /* upper bound on any sequence length */
  int32_t path;  /* FVTA_LSTM_AUTO / _FFMA / _TC */
} FvtaReaderDesc;

size_t fvta_reader_workspace_bytes(const FvtaReaderDesc* desc);
size_t fvta_reader_reserve_floats(const FvtaReaderDesc* desc);

/*
 * x_tok        [ntok,In]   inputs of the valid tokens, sequence after sequence
 * seq_tok_off  [B] int32   first token row of each sequence
 * seq_len      [B] int32   valid length (0 allowed)
 * seq_out_row  [B] int64   arena row receiving step 0 of the sequence (step t -> row+t);
 *                          negative: outputs are not stored (only `last` is wanted)
 * out_arena    [rows,2d]   caller-zeroed; fw in [:d], bw in [d:]
 * last         [B,2d]      final h of each direction
 */
int fvta_reader_fwd(const FvtaReaderDesc* desc, const float* x_tok, const int32_t* seq_tok_off,
                    const int32_t* seq_len, const int64_t* seq_out_row, const float* kernel,
                    const float* bias, float* out_arena, float* last, float* reserve, void* workspace,
                    size_t workspace_bytes, void* stream);

/* d_out_arena [rows,2d], d_last [B,2d] (nullable) -> d_x_tok [ntok,In] (overwritten),
 * d_kernel [In+d,4d], d_bias [4d] ACCUMULATED (+=).
 * The three products behind the recurrence (d kernel[In:], d x, d kernel[:In] + d bias) are
 * independent: this call and fvta_bilstm_bwd fork them over two auxiliary streams (a pool of four
 * per device, created on the first call -- make that call outside a stream capture) and join on
 * `stream` before returning, so the caller sees ordinary stream semantics. */
int fvta_reader_bwd(const FvtaReaderDesc* desc, const float* x_tok, const int32_t* seq_tok_off,
                    const int32_t* seq_len, const int64_t* seq_out_row, const float* kernel,
                    const float* reserve, const float* d_out_arena, const float* d_last, float* d_x_tok,
                    float* d_kernel, float* d_bias, void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * fp32 GEMM with fused epilogue: C[M,N] = act(A[M,K] @ B[K,N] + bias[N]).
 * Replaces `linear` (model_v2.py:75-100: tf.matmul + bias + optional tanh).
 * act: 0 none, 1 tanh, 2 relu.  trans_a/trans_b: operand stored transposed.
 * accumulate != 0: C += (no bias/act) -- used for weight gradients.
 * ---------------------------------------------------------------------------------------- */
int fvta_gemm(int32_t M, int32_t N, int32_t K, const float* A, int32_t lda, int32_t trans_a,
              const float* B, int32_t ldb, int32_t trans_b, const float* bias, int32_t act,
              float* C, int32_t ldc, int32_t accumulate, void* stream);

/* The same product with the rows of op(A) GATHERED from a larger table: storage row r of A is
 * A_table[a_rows[r]] (trans_a == 0: a_rows has M entries; trans_a != 0: K entries).  Replaces
 * tf.nn.embedding_lookup(image_emb_mat, pis) followed by `linear` (model_v2.py:634-645) and its
 * weight gradient without materialising the gathered [tokens, image_feat_dim] matrix. */
int fvta_gemm_gather(int32_t M, int32_t N, int32_t K, const float* A_table, int32_t lda,
                     const int64_t* a_rows, int32_t trans_a, const float* B, int32_t ldb, int32_t trans_b,
                     const float* bias, int32_t act, float* C, int32_t ldc, int32_t accumulate, void* stream);

/* ------------------------------------------------------------------------------------------
 * Character CNN (replaces conv1d, model_v2.py:52-70, when its input dropout is off): the conv
 * output of a window depends only on its `height` character ids, so with
 * table[:,h,:] = char_emb @ filter[h] (caller, [VC,height,cw]) the layer is
 *   y[w,:] = relu( max_p ( bias + sum_h table[ids[w,p+h]][h][:] ) ),  p in [0, W-height].
 * argpos [ntok,cw] u8 records the arg-max position for the backward, which accumulates
 * d_table / d_bias (+=, caller zeroes them).  cw <= 128, W <= 64 (float4 path when cw % 4 == 0, the row
 * stride is a multiple of 4 and the bases are 16-byte aligned; one channel per lane otherwise).
 * y and dy are addressed with row strides ldy / lddy (floats) so the features can live inside the
 * [char | word] token matrix of model_v2.py:617.
 * ---------------------------------------------------------------------------------------- */
int fvta_charcnn_fwd(int32_t ntok, int32_t W, int32_t VC, int32_t cw, int32_t height, const int32_t* ids,
                     const float* table, const float* bias, float* y, int32_t ldy, uint8_t* argpos,
                     void* stream);
int fvta_charcnn_bwd(int32_t ntok, int32_t W, int32_t VC, int32_t cw, int32_t height, const int32_t* ids,
                     const float* y, int32_t ldy, const uint8_t* argpos, const float* dy, int32_t lddy,
                     float* d_table, float* d_bias, void* stream);

/* ------------------------------------------------------------------------------------------
 * Word vectors of the packed tokens (replaces the tf.concat of the trainable and the fed
 * embedding table + tf.nn.embedding_lookup, model_v2.py:590-600):
 *   out[i, 0:dim] = ids[i] < VW ? table_new[ids[i]] : table_fixed[ids[i] - VW]   (row stride ld_out)
 * and its backward, d_table_new[ids[i]] += dy[i, 0:dim] for ids[i] < VW (caller zeroes; the fed
 * table is not trained).  Ids are clamped into [0, VW + VF).
 * ---------------------------------------------------------------------------------------- */
int fvta_embed_rows(int32_t ntok, int32_t dim, int32_t VW, int32_t VF, const int64_t* ids,
                    const float* table_new, const float* table_fixed, float* out, int32_t ld_out, void* stream);
int fvta_embed_rows_bwd(int32_t ntok, int32_t dim, int32_t VW, const int64_t* ids, const float* dy,
                        int32_t ld_dy, float* d_table_new, void* stream);

/* ------------------------------------------------------------------------------------------
 * Bias / activation part of the backward of `linear` (model_v2.py:75-100):
 *   dz[M,N] = dy * act'(y)   (act 1: tanh, y = tanh output; 2: relu; 0: dz is not written, dz == dy)
 *   db[N]  += column sums of dz   (nullable)
 * ---------------------------------------------------------------------------------------- */
int fvta_bias_act_grad(int32_t M, int32_t N, const float* dy, int32_t ld_dy, const float* y, int32_t ld_y,
                       int32_t act, float* dz, int32_t ld_dz, float* db, void* stream);

/* ------------------------------------------------------------------------------------------
 * Output layer + softmax + mean cross-entropy, forward and backward in ONE launch (replaces
 * model_v2.py:1061-1083, the non-`use_eu_output` branch, and :1088-1090).
 *   gq [N,D], g1 [N,D], gc [N,C,D] (choice encodings), W [5D,1] = blocks for
 *   [gq, g1, gc, g1*gc, gq*gc], b [1], y [N,C] u8 (nullable at test time)
 *   -> logits [N,C], yp [N,C], *loss += mean CE (caller zeroes it; nullable),
 *      and, when all five gradient pointers are non-null: d_gq, d_g1, d_gc (overwritten),
 *      dW [5D], db [1] (ACCUMULATED, caller zeroes) for an upstream gradient of 1.
 * ---------------------------------------------------------------------------------------- */
int fvta_output_loss(int32_t N, int32_t C, int32_t D, const float* gq, const float* g1, const float* gc,
                     const float* W, const float* b, const uint8_t* y, float* logits, float* yp, float* loss,
                     float* d_gq, float* d_g1, float* d_gc, float* dW, float* db, void* stream);

/* ------------------------------------------------------------------------------------------
 * Adadelta over one flat buffer.  Replaces tf.train.AdadeltaOptimizer(lr).apply_gradients
 * (trainer.py:16-27; TF defaults rho=0.95, eps=1e-8).  `update_mask` (nullable) is 1.0 where
 * the slot is trainable-and-has-a-gradient, 0.0 otherwise (window_t is never updated).
 * grad_scale multiplies the gradient first (1/world_size after a sum all-reduce).
 * touch_index [n] int32 / touch (both nullable): sparse rows.  The reference reads its embedding
 * tables through tf.nn.embedding_lookup (model_v2.py:534, :596-600), whose gradient is IndexedSlices,
 * so TF runs sparse_apply_adadelta on them: a row no id of the batch looked up keeps var AND both
 * accumulators.  touch_index[i] >= 0 selects the entry of `touch` that is > 0 iff the row holding
 * element i was looked up this step; -1 marks a dense variable.
 * ---------------------------------------------------------------------------------------- */
int fvta_adadelta(int64_t n, float* var, const float* grad, float* accum, float* accum_update,
                  const float* update_mask, const int32_t* touch_index, const float* touch,
                  float lr, float rho, float eps, float grad_scale, void* stream);

/* ------------------------------------------------------------------------------------------
 * Data parallelism (north_star d; the reference is single-device, model_v2.py:14): the sum over the
 * W ranks of the flat gradient buffer FUSED with the Adadelta update of trainer.py:16-27, in one
 * kernel per rank over NVLink peer memory (no NCCL call on the path).
 *   grad_local      [n + n_tail]  this rank's gradient, then the sparse tables' "row looked up" flags
 *   stage_ptrs_dev  device array [world] of peer-mapped staging buffers, each 2 * (n + n_tail) floats
 *                   (the halves alternate by step); entry `rank` is the local one
 *   flag_ptrs_dev   device array [world] of peer-mapped flag arrays (>= world uint32, zero-initialised)
 *   state           local, 4 x uint32, zero-initialised once: step counter and block counters
 * Every rank must call it once per step, in step order, with the same n / n_tail.  The sum runs in rank
 * order on every rank (bitwise identical replicas); the gradient is scaled by 1/world (batch mean over
 * the global batch, model_v2.py:1090).  touch_index: see fvta_adadelta (flags come from the summed tail).
 * ---------------------------------------------------------------------------------------- */
int fvta_p2p_allreduce_adadelta(int32_t world, int32_t rank, int64_t n, int64_t n_tail,
                                const float* grad_local, float* const* stage_ptrs_dev,
                                uint32_t* const* flag_ptrs_dev, uint32_t* state, float* var, float* accum,
                                float* accum_update, const float* update_mask, const int32_t* touch_index,
                                float lr, float rho, float eps, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FVTA_B200_H */
