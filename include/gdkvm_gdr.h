/*
 * gdkvm_gdr.h -- C ABI of the B200-native GDKVM memory op (LKVA readout + Gated Delta Rule).
 *
 * This is the drop-in boundary for the ONE hot path of wangrui2025/GDKVM named by BASELINE.json
 * north_star.  The mounted reference exposes no code for it (reference README.md:1 "Code:
 * github.com/wangrui2025/gdkvm_code", README.md:36-38 "Quick Start TBD"), so there is no reference
 * FFI to bind; every entry point below instead replaces the reference *concept* cited next to it
 * (README.md:20, website/src/content/homepage/en.json:20) with the call surface north_star fixes:
 *     (q, k, v, gate, beta, initial_state) -> (readout, final_state)
 * which is argument-compatible with flash-linear-attention's chunk_gated_delta_rule
 * (fla/ops/gated_delta_rule/chunk.py:365-377), the op the upstream model most plausibly calls.
 *
 * Conventions: plain C, no torch/C++ types, caller-owned buffers, errors as negative ints (no
 * exceptions cross the ABI), asynchronous on the caller's CUDA stream, CUDA-graph capturable, re-entrant.
 * No host synchronisation and no user-visible allocation, ever.  Internal memory: a launch whose chains are cut into
 * time segments (and every packed variable-length launch) takes a per-launch scratch with a stream-ordered
 * allocation from a memory pool the LIBRARY owns (created on first use per device; the process-wide default pool and
 * its release policy are never touched); under stream capture that scratch becomes alloc / free nodes of the graph.
 */
#ifndef GDKVM_GDR_H_
#define GDKVM_GDR_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GDKVM_ABI_VERSION 1

/* element types of q/k/v/o (io_dtype) and of g/beta (gate_dtype) */
enum GdkvmDtype { GDKVM_F32 = 0, GDKVM_BF16 = 1 };

/* error codes (0 = success).  gdkvm_strerror() gives the text. */
enum GdkvmStatus {
    GDKVM_OK = 0,
    GDKVM_ERR_NULL = -1,        /* a required pointer is NULL                               */
    GDKVM_ERR_ABI = -2,         /* params->struct_size does not match this library          */
    GDKVM_ERR_SHAPE = -3,       /* B/T/H/K/V/frame_tokens out of the supported set           */
    GDKVM_ERR_DTYPE = -4,       /* unknown dtype enum                                        */
    GDKVM_ERR_ALIGN = -5,       /* pointer/stride alignment not met                          */
    GDKVM_ERR_ARCH = -6,        /* device is not sm_100 (no fallback is provided on purpose)  */
    GDKVM_ERR_CUDA = -7,        /* a CUDA runtime/driver call failed (see gdkvm_last_cuda_error) */
    GDKVM_ERR_UNSUPPORTED = -8  /* a forced path (flags) cannot run this problem             */
};

/* flags */
#define GDKVM_FLAG_FORCE_RECURRENT 0x1u /* token-recurrent fp32 CUDA-core kernel (exact fp32 math)   */
#define GDKVM_FLAG_FORCE_CHUNKED   0x2u /* tcgen05 WY/UT chunk kernel; error if the shape is not
                                           covered instead of falling back                           */
#define GDKVM_FLAG_FLAT_CHUNKS     0x4u /* chunked kernel: tile the flat token stream in 64-token
                                           chunks instead of frame-aligned chunks (same results,
                                           no 49->64 padding work)                                    */
#define GDKVM_FLAG_FRAME_CHUNKS    0x8u /* chunked kernel: always one chunk per frame (sub-chunks of
                                           64 for longer frames).  Default: frame-aligned when
                                           frame_tokens is a multiple of 64, flat tiling otherwise --
                                           token-causal semantics make the two exactly equivalent     */

#define GDKVM_FLAG_SEGMENTS(n) (((uint32_t)(n) & 0xfu) << 8)
                                        /* chunked kernel: cut every chain into n (1..15) time
                                           segments scheduled as separate work units (fills the
                                           ragged last wave of chains over SMs; bit-identical
                                           results).  0 = let the library choose.  Under stream
                                           capture the hand-off scratch becomes alloc / free nodes of
                                           the graph                                                  */

/*
 * One forward call of the memory module over a batch of clips.
 *   replaces: "Linear Key-Value Association ... state transition matrix; Gated Delta Rule ...
 *   managing memory" (reference website/src/content/homepage/en.json:20; README.md:20).
 *
 * Per clip b and head h, tokens i = 0..T-1 in frame-major raster order, fp32 state S[K][V]:
 *     S <- exp(g_i) S ;  r = v_i - S^T k_i ;  S <- S + k_i (beta_i r)^T ;  o_i = scale S^T q_i
 *
 * Tensors (strides in ELEMENTS; innermost K or V dimension contiguous):
 *     q,k [B,T,H,K]   v,o [B,T,H,V]   g,beta [B,T,H]   initial/final state [B,H,K,V] fp32 contiguous
 * stride arrays are {batch, token, head}.  frame_tokens = C > 0 declares T = F*C with every frame
 * (C pixel tokens of one key/value feature map) forming one chunk; 0 = no frame structure.
 */
typedef struct GdkvmGdrParams {
    uint32_t struct_size;        /* = sizeof(GdkvmGdrParams), ABI guard                       */
    uint32_t flags;              /* GDKVM_FLAG_*                                              */
    const void* q;
    const void* k;
    const void* v;
    const void* g;               /* log-space gate (<= 0), one per token-head                 */
    const void* beta;            /* write strength in (0,1)                                   */
    const float* initial_state;  /* may be NULL (zero state)                                  */
    void* o;                     /* readout, io_dtype                                         */
    float* final_state;          /* may be NULL (not written)                                 */
    int64_t q_stride[3];
    int64_t k_stride[3];
    int64_t v_stride[3];
    int64_t o_stride[3];
    int64_t g_stride[3];
    int64_t beta_stride[3];
    int32_t B, T, H, K, V;
    int32_t frame_tokens;
    int32_t io_dtype;            /* GdkvmDtype of q,k,v,o                                     */
    int32_t gate_dtype;          /* GdkvmDtype of g,beta                                      */
    float scale;                 /* readout scale; the Python op defaults it to K^-0.5        */
    int32_t reserved;
} GdkvmGdrParams;

/* ABI version of the loaded library (compare with GDKVM_ABI_VERSION). */
int gdkvm_abi_version(void);

/* Static, never-NULL description of a GdkvmStatus. */
const char* gdkvm_strerror(int status);

/* cudaError_t of the most recent failing CUDA call on this thread (0 if none). */
int gdkvm_last_cuda_error(void);

/*
 * Launch the forward op on `cuda_stream` (a cudaStream_t; NULL = legacy default stream) of the
 * current device.  Returns immediately after enqueueing; 0 or a negative GdkvmStatus.
 */
int gdkvm_gdr_fwd(const GdkvmGdrParams* params, void* cuda_stream);

/*
 * The same op over PACKED variable-length clips (videos of different lengths in one batch):
 *   replaces: the `cu_seqlens` argument of fla's chunk_gated_delta_rule (fla/ops/gated_delta_rule/chunk.py:375);
 *   SURVEY.md section 8f rank 4.
 * params->B must be 1 and params->T the total number of tokens: q,k [1,T,H,K], v,o [1,T,H,V], g,beta [1,T,H].  Clip n
 * is rows cu_seqlens[n] .. cu_seqlens[n+1]-1 (n = 0..n_seqs-1; non-decreasing, 0 <= cu_seqlens[0], cu_seqlens[n_seqs]
 * <= T; rows of o outside every clip are never written, rows of q/k/v there may be read and must hold finite
 * numbers), an array of n_seqs+1 offsets in DEVICE memory of cu_seqlens_bytes (4 or 8) bytes each -- the
 * library never reads it on the host, so the call stays asynchronous.  initial_state / final_state are
 * [n_seqs, H, K, V]; a clip without tokens passes its initial state through.  frame_tokens is ignored (flat 64-token
 * tiling; token-causal semantics make that exactly equivalent).  Uses a stream-ordered per-launch workspace
 * (cudaMallocAsync / cudaFreeAsync on cuda_stream); under stream capture these become allocation / free nodes of the graph.
 */
int gdkvm_gdr_fwd_varlen(const GdkvmGdrParams* params, const void* cu_seqlens, int32_t cu_seqlens_bytes, int32_t n_seqs,
                         void* cuda_stream);

/*
 * Which kernel gdkvm_gdr_fwd would pick for `params` without launching anything:
 * 0 = recurrent fp32 CUDA-core kernel, 1 = tcgen05 chunked kernel, negative = GdkvmStatus.
 * Needs no GPU.
 */
int gdkvm_gdr_plan(const GdkvmGdrParams* params);

/*
 * Why gdkvm_gdr_fwd would NOT take the tcgen05 chunk kernel for `params` (static string; "" when it would): a call that
 * drops to the fp32 CUDA-core kernel is several times slower, and this is how a caller finds out why.  Needs no GPU.
 */
const char* gdkvm_gdr_plan_reason(const GdkvmGdrParams* params);

/*
 * Time segments per chain gdkvm_gdr_fwd would cut the problem into on a device with `sm_count` SMs (<= 0: 148, a B200):
 * 1 = one work unit per (clip, head) chain; n > 1 = n units per chain, scheduled in order, so that the last wave of
 * units over the SMs is not ragged (GDKVM_FLAG_SEGMENTS overrides; launches under stream capture are cut the same
 * way when the device supports memory pools -- the scratch lives in the graph -- and stay uncut otherwise).
 * Pure host arithmetic, needs no GPU.  Returns 1 for the recurrent path, negative = GdkvmStatus.
 */
int gdkvm_gdr_plan_segments(const GdkvmGdrParams* params, int sm_count);

/*
 * The work units of gdkvm_gdr_fwd for this problem: out = {units, uncut clips, cut clips, segments of a cut clip}.  A batch of
 * equal-length clips that fills more than one wave of SMs is not cut uniformly: whole waves of chains stay uncut (a unit
 * boundary costs ~3 chunk periods) and only the clips left over for the last wave are cut, into about one unit per SM
 * ("mixed plan": flat 64-token tiling or frames of whole 64-token chunks, no GDKVM_FLAG_SEGMENTS, a batch that is contiguous in memory; results are bit-identical
 * to any other plan).  Returns 1 for the mixed plan, 0 for the uniform one (out[1] = 0, out[3] = gdkvm_gdr_plan_segments), negative
 * = GdkvmStatus.  Pure host arithmetic.
 */
int gdkvm_gdr_plan_units(const GdkvmGdrParams* params, int sm_count, int32_t out[4]);

/*
 * Row-wise L2 normalisation  y[r][:] = x[r][:] * rsqrt(sum(x[r][:]^2) + eps)  of `rows` rows of D elements (dtype:
 * GdkvmDtype; D in {32, 64, 128, 256}; row strides in ELEMENTS, multiples of 16 bytes; 16-byte aligned bases; y may
 * alias x).  The step immediately before the memory op when the caller asks for normalised q / k:
 *   replaces: fla's `use_qk_l2norm_in_kernel=True` (fla/ops/gated_delta_rule/chunk.py:374); SURVEY.md section 8f rank 3.
 * Asynchronous on `cuda_stream`; 0 or a negative GdkvmStatus.
 */
int gdkvm_l2norm_fwd(const void* x, void* y, int64_t rows, int32_t D, int64_t x_row_stride, int64_t y_row_stride,
                     int32_t dtype, float eps, void* cuda_stream);

/*
 * ---- training: forward that keeps what the backward pass needs, and the backward pass ----
 *   replaces: the autograd formula of the reference memory module (the upstream model is trained: reference
 *   website/src/pages/[lang]/reprod/index.astro:238-252; shape of the work: fla/ops/gated_delta_rule/chunk.py:117).
 *
 * gdkvm_gdr_fwd_train = gdkvm_gdr_fwd on the tcgen05 chunk kernel (bf16 I/O, K = 64, V in {64, 128, 256}; anything else returns
 * GDKVM_ERR_UNSUPPORTED -- there is no slow training path) which ALSO writes the bf16 state at the start of every 64-token
 * chunk into `chunk_states`, a caller-owned device buffer of gdkvm_gdr_chunk_states_bytes(B, T, H, K, V) bytes laid out
 * [B*H][ceil(T/64)][V][K], 32-byte aligned (GDKVM_ERR_ALIGN otherwise: it is written with 256-bit stores).  The token stream is
 * tiled flat in 64-token chunks (frame_tokens is ignored; results equal within the op's tolerance).
 */
int gdkvm_gdr_fwd_train(const GdkvmGdrParams* params, void* chunk_states, void* cuda_stream);
int64_t gdkvm_gdr_chunk_states_bytes(int32_t B, int32_t T, int32_t H, int32_t K, int32_t V);
/*
 * The same for packed variable-length clips (gdkvm_gdr_fwd_varlen's arguments): chunk_states holds
 * gdkvm_gdr_chunk_states_bytes_varlen(T, n_seqs, H, K, V) bytes laid out [T / 64 + n_seqs + 1 slots][H][V][K]; chunk c of clip n
 * uses slot cu_seqlens[n] / 64 + n + c (distinct for distinct chunks, computable without a scan over the offsets).
 */
int gdkvm_gdr_fwd_train_varlen(const GdkvmGdrParams* params, const void* cu_seqlens, int32_t cu_seqlens_bytes, int32_t n_seqs,
                               void* chunk_states, void* cuda_stream);
int64_t gdkvm_gdr_chunk_states_bytes_varlen(int32_t T, int32_t n_seqs, int32_t H, int32_t K, int32_t V);

/*
 * Gradients of (readout, final_state) with respect to (q, k, v, g, beta, initial_state), given the cotangents d_o [B,T,H,V]
 * (io dtype) and d_final_state [B,H,K,V] (fp32, may be NULL = zero).  dq, dk, dv: io dtype, strides in elements, innermost
 * dimension contiguous; dg, dbeta: fp32, CONTIGUOUS [B,T,H]; d_initial_state: fp32 [B,H,K,V], may be NULL (not written).
 * One kernel: per (clip, head) chain a reverse-time scan over 64-token chunks with the state cotangent in registers.
 */
typedef struct GdkvmGdrBwdParams {
    uint32_t struct_size;        /* = sizeof(GdkvmGdrBwdParams)                                */
    uint32_t flags;              /* GDKVM_FLAG_SEGMENTS(n): cut every chain into n time segments (separate work units, the state
                                    cotangent handed over in fp32: bit-identical results); 0 = the library chooses           */
    const void* q;
    const void* k;
    const void* v;
    const void* g;
    const void* beta;
    const void* d_o;             /* cotangent of the readout                                   */
    const float* d_final_state;  /* cotangent of the final state, may be NULL                  */
    const void* chunk_states;    /* written by gdkvm_gdr_fwd_train on the same inputs          */
    void* dq;
    void* dk;
    void* dv;
    float* dg;
    float* dbeta;
    float* d_initial_state;      /* may be NULL                                                */
    int64_t q_stride[3];
    int64_t k_stride[3];
    int64_t v_stride[3];
    int64_t do_stride[3];
    int64_t g_stride[3];
    int64_t beta_stride[3];
    int64_t dq_stride[3];
    int64_t dk_stride[3];
    int64_t dv_stride[3];
    int32_t B, T, H, K, V;
    int32_t io_dtype;
    int32_t gate_dtype;
    float scale;
    /* packed variable-length clips (the backward of gdkvm_gdr_fwd_train_varlen): B = 1, T = total tokens, states and their
       cotangents [n_seqs, H, K, V]; NULL / 0 / 0 for the batched call */
    const void* cu_seqlens;      /* device, n_seqs + 1 offsets                                 */
    int32_t cu_seqlens_bytes;    /* 4 or 8                                                     */
    int32_t n_seqs;
} GdkvmGdrBwdParams;

int gdkvm_gdr_bwd(const GdkvmGdrBwdParams* params, void* cuda_stream);

/*
 * ---- fused projection prologue: features -> the op's operands in ONE kernel ----
 *   replaces: the projections that turn the fused key/pixel feature into q, k, v, gate, beta ("Key-Pixel Feature Fusion fuses
 *   the local key feature, the global key feature with the pixel feature", reference website/src/content/homepage/en.json:20)
 *   plus the q/k L2 normalisation (fla/ops/gated_delta_rule/chunk.py:374) and the gate / beta activations.
 *
 *   y = x w^T (+ bias),  x [R, D] bf16 (row stride in elements),  w [N, D] bf16 row-major (the nn.Linear layout),
 *   N = H (2 K + V) + 2 H with the rows of w ordered   q (H x K) | k (H x K) | v (H x V) | g (H) | beta (H);
 *   q, k <- y_q, y_k L2-normalised per head (x rsqrt(sum x^2 + eps)) -> bf16 [R, H, K];   v <- y_v -> bf16 [R, H, V];
 *   g <- logsigmoid(y_g), beta <- sigmoid(y_beta) -> fp32 [R, H].   K = 64, H even (2..32), V % 64 == 0, D % 64 == 0.
 * A tcgen05 GEMM (TMA-staged operands, TMEM accumulators) whose epilogue writes the op-ready tensors: y never touches HBM.
 */
#define GDKVM_PROJ_FLAG_TILE_ROWS_128 1u   /* force 128-token row blocks (128 x 256 output tiles)                       */
#define GDKVM_PROJ_FLAG_TILE_ROWS_256 2u   /* force 256-token row blocks (256 x 128 output tiles; needs D <= 256)       */
typedef struct GdkvmProjParams {
    uint32_t struct_size;        /* = sizeof(GdkvmProjParams)                                  */
    uint32_t flags;              /* 0 = the library picks the tile shape; GDKVM_PROJ_FLAG_*    */
    const void* x;
    const void* w;
    const float* bias;           /* [N] fp32, may be NULL                                      */
    void* q;
    void* k;
    void* v;
    float* g;
    float* beta;
    int64_t R;                   /* rows = tokens (B * T)                                      */
    int64_t x_row_stride;        /* elements                                                   */
    int32_t D, H, K, V;
    float eps;
    int32_t reserved;
} GdkvmProjParams;

int gdkvm_qkvgb_project_fwd(const GdkvmProjParams* params, void* cuda_stream);

/* Number of kernels this library has launched in the calling process (bench "gpu_launches"). */
uint64_t gdkvm_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* GDKVM_GDR_H_ */
