/*
 * dedark_b200.h -- C-ABI of libdedark_b200.so: the B200 (sm_100a) implementation of the
 * Dedark-YOLO low-light hot path.
 *
 * The reference (cvYouTian/Dedark-YOLO) is pure Python: it has no FFI for this path.  The seam this
 * header defines is the one its host code would bind: plain pointers,
 * sizes and a CUDA stream; no torch types.  Each entry point names the reference code it replaces
 * (paths relative to the reference checkout).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless its comment says "host";
 *   - the caller owns every buffer (the library never allocates or frees device memory, creates no
 *     streams, keeps no mutable global state apart from the launch counter and the last-error
 *     string, which is thread-local);
 *   - all work is enqueued on `stream` (a cudaStream_t passed as void*), nothing synchronises;
 *   - images are contiguous NCHW fp32 with C == 3;
 *   - return value: DD_OK or a DD_ERR_* code; `dd_last_error()` then describes the failure.  The
 *     codes mirror the reference's own failure modes (reflect pad needs H,W > 12; the rgb2lum quirk
 *     needs W >= 3).  The library never calls exit/abort.
 */
#ifndef DEDARK_B200_H
#define DEDARK_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DD_OK 0
#define DD_ERR_INVALID 1      /* null pointer, B/H/W <= 0, bad enum                                */
#define DD_ERR_REFLECT_PAD 2  /* H <= 12 or W <= 12: F.pad(mode='reflect', 12) fails (filtersB.py:167) */
#define DD_ERR_WIDTH_LT3 3    /* W < 3: rgb2lum indexes columns 0..2 (util_filters.py:270-273)     */
#define DD_ERR_WORKSPACE 4    /* workspace too small (see dd_workspace_bytes)                     */
#define DD_ERR_CUDA 5         /* a CUDA runtime/driver call failed                                */

#define DD_NUM_FEATURES 15    /* filter_cfg.py:17 */
#define DD_RESIZE 256         /* llie.py:43       */

/* workspace kinds for dd_workspace_bytes */
#define DD_WS_SYNTH 0          /* dd_synth_fwd / dd_synth_resize_fwd partial sums (depends on B)   */
#define DD_WS_PREDICTOR_ACTS 1 /* activations + prepared tensor-core weights kept from dd_predictor_fwd for dd_predictor_bwd */
#define DD_WS_PREDICTOR_BWD 2  /* scratch of dd_predictor_bwd                                      */
#define DD_WS_RECOVERY_BWD 3   /* partial sums of dd_recovery_bwd                                  */
#define DD_WS_DARK_PRIOR 4     /* histograms of dd_dark_prior (depends on B)                       */

/* element types of the *_ex entry points (bf16 I/O mode) */
#define DD_F32 1
#define DD_BF16 2

/* source dtypes for dd_synth_fwd */
#define DD_SRC_U8 0
#define DD_SRC_F32 1

int dd_version(void);
const char* dd_last_error(void);             /* host string, valid until the next failing call on this thread */
unsigned long long dd_launch_count(void);    /* kernels launched by this library since load (monotonic)       */
size_t dd_workspace_bytes(int kind, int B, int H, int W);

/* ---- a1 + a2: low-light synthesis and the recovery-loss scalar ---------------------------------
 * Replaces models/yolo/detect/train.py:72,79,103,108-109 (`clean = u8.float()/255`,
 * `dark = torch.pow(clean, p)`, `rec = F.mse_loss(dark, clean)`) and the offline writer
 * utils/lowlight_process.py:57,68,74 (`(dark*255).astype(uint8)`, truncation).
 *   src        n elements, uint8 (DD_SRC_U8) or fp32 in [0,1] (DD_SRC_F32)
 *   lut256     optional 256-entry table dark = lut256[u8] (u8 source only).  NULL: the kernel fills
 *              the table itself with powf(clean[k], p) -- bit-identical to torch.pow on CUDA.  A host-
 *              computed table makes the result bit-identical to the reference's CPU path.
 *   clean_lut256  optional 256-entry table clean = clean_lut256[u8].  NULL: k * (1.f/255.f), which is
 *              what `u8.float() / 255` evaluates to on CUDA (ATen multiplies by the reciprocal); the
 *              CPU kernel performs a true division, so a host table reproduces those bits.
 *   clean_out  optional fp32 clean image (u8 source only; ignored for fp32 source)
 *   dark_out   optional fp32 darkened image
 *   dark_u8    optional truncated uint8 darkened image (offline writer)
 *   rec_out    optional: one float, mean((dark-clean)^2) over the n elements
 *   ws         DD_WS_SYNTH bytes (only needed when rec_out != NULL)
 */
int dd_synth_fwd(const void* src, int src_dtype, float p, const float* lut256,
                 const float* clean_lut256, float* clean_out, float* dark_out, uint8_t* dark_u8, float* rec_out, long long n, void* ws,
                 size_t ws_bytes, void* stream);

/* ---- a1 + a2 + a4 in one pass: synthesis fused with the 256x256 bilinear resize the module applies to the darkened
 * batch (train.py:72,103,108 + llie.py:43).  Same outputs as dd_synth_fwd followed by dd_resize256 (r bit-identical), but
 * the dark batch is not read back from HBM for the resize.  Needs W % 4 == 0 and a band of source rows that fits in shared
 * memory: dd_synth_resize_supported(H, W) == 1; ws: DD_WS_SYNTH bytes for this B. */
int dd_synth_resize_supported(int H, int W);
int dd_synth_resize_fwd(const void* src, int src_dtype, float p, const float* lut256, const float* clean_lut256,
                        float* clean_out, float* dark_out, float* r_out, float* rec_out, int B, int H, int W, void* ws,
                        size_t ws_bytes, void* stream);

/* ---- uint8 batch -> filter chain in one pass, no darkened fp32 batch in HBM ---------------------
 * (models/yolo/detect/train.py:72-109: the dataloader hands uint8 NCHW, data/dataset.py:172-188.)  dd_dark_table writes the 256
 * darkened values dark[k] = pow(k/255, p) -- the same bits dd_synth_fwd produces; lut256 overrides as there -- and
 * dd_recovery_fwd_u8 / dd_recovery_bwd_u8 are dd_recovery_fwd / dd_recovery_bwd reading `src` (uint8 [B,3,H,W]) through that
 * table in their stage phase: y and dfeat are bit-identical to the fp32-source calls on the materialised batch.  Needs
 * W % 4 == 0.  dd_synth_resize_fwd with dark_out == NULL then only produces r and the recovery loss.  No dx (x is data). */
int dd_dark_table(float p, const float* lut256, float* table_out, void* stream);
int dd_recovery_fwd_u8(const uint8_t* src, const float* dark_table, const float* A, const float* IcA, const float* feat, float* y,
                       int B, int H, int W, void* stream);
int dd_recovery_bwd_u8(const uint8_t* src, const float* dark_table, const float* A, const float* IcA, const float* feat, const float* g,
                       float* dfeat, int B, int H, int W, void* ws, size_t ws_bytes, void* stream);

/* ---- the dark-channel prior on the GPU ------------------------------------------
 * Replaces the per-batch D2H + numpy loop of DetectionTrainer.preprocess_batch (models/yolo/detect/train.py:81-97:
 * DarkChannel :42-45, AtmLight :47-61, DarkIcA :63-67) that produces batch['dedark_A'] / batch['IcA'].  Works on the
 * uint8 source batch directly: the darkened image is quantised exactly as train.py:84 does
 * (`(pow(u8/255, p) * 255).astype(uint8)`, a 256-entry table), then
 *   dc[h,w]  = min_c dark_u8[c,h,w]                                            (DarkChannel)
 *   A_u8[c]  = (sum of dark_u8[c] over the numpx - 1 pixels of largest dc) / numpx,  numpx = max(HW / 1000, 1)
 *              (AtmLight, including its `range(1, numpx)` off-by-one; pixels tied at the selection threshold contribute
 *              in equal shares -- numpy's argsort picks an arbitrary subset of them)
 *   IcA[h,w] = min_c dark_u8[c,h,w] / max(A_u8[c], 1)                          (DarkIcA as intended: the reference
 *              indexes rows instead of channels of an uninitialised uint8 array, so its own IcA is not reproducible)
 * Outputs: A_out [B,3] = A_u8 / 255 (the image's [0,1] units, what DeDarkFilter expects), IcA_out [B,1,H,W] fp32.
 * lut256: optional darkening table as for dd_synth_fwd.  ws: DD_WS_DARK_PRIOR bytes.  Integer histograms: bit-reproducible. */
int dd_dark_prior(const uint8_t* src, float p, const float* lut256, float* A_out, float* IcA_out, int B, int H, int W, void* ws,
                  size_t ws_bytes, void* stream);

/* ---- a4: bilinear resize to 256x256 (llie.py:43; align_corners=False, no antialias) ----------- */
int dd_resize256(const float* x, float* r, int B, int H, int W, void* stream);
/* adjoint of the above: dx += resize^T(dr)  (only used when the caller wants dL/dx) */
int dd_resize256_bwd(const float* dr, float* dx, int B, int H, int W, void* stream);

/* ---- a5: the parameter predictor (nn/modules/common.py:9-23,52-78) -----------------------------
 * Weights in the reference's state-dict layout: conv_w[i] is [Cout,Cin,3,3], fc1_w [64,2048],
 * fc2_w [15,64].  The same struct with writable pointers receives the gradients (overwritten). */
typedef struct dd_predictor_tensors {
    float* conv_w[5];
    float* conv_b[5];
    float* fc1_w;
    float* fc1_b;
    float* fc2_w;
    float* fc2_b;
} dd_predictor_tensors;

/* r [B,3,256,256] -> feat [B,15]; acts: DD_WS_PREDICTOR_ACTS bytes, kept by the caller for bwd.  conv2..conv5 run on the
 * tensor cores (tcgen05, split-TF32 operands, fp32 accumulation); behind the activations `acts` also receives the
 * weights of those layers re-laid-out for the tensor-core operands, which dd_predictor_bwd of the SAME weights reuses. */
int dd_predictor_fwd(const float* r, const dd_predictor_tensors* w, float* acts, float* feat, int B,
                     void* stream);
/* dfeat [B,15] -> grads (all 14 tensors overwritten) and, if dr != NULL, dL/dr [B,3,256,256].  `acts` must come from
 * dd_predictor_fwd called with the same `w` (it carries the prepared weights).  Deterministic (fixed-order sums). */
int dd_predictor_bwd(const float* r, const dd_predictor_tensors* w, const float* acts,
                     const float* dfeat, const dd_predictor_tensors* grads, float* dr, int B,
                     void* ws, size_t ws_bytes, void* stream);

/* The same backward in two calls, for callers that overlap the reduction of the gradients over the ranks with the rest of
 * the backward (DDP does the same with its buckets, engine/trainer.py:223): part 1 = the fully connected layers (one launch;
 * fc1 / fc2 gradients -- 132 111 of the 164 943 floats, the tail of the flat state-dict order -- are final afterwards),
 * part 2 = the five convolutions, part 0 = everything (== dd_predictor_bwd). */
int dd_predictor_bwd_part(const float* r, const dd_predictor_tensors* w, const float* acts,
                          const float* dfeat, const dd_predictor_tensors* grads, float* dr, int B, void* ws,
                          size_t ws_bytes, int part, void* stream);

/* ---- 8(e): predictor backward fused with the exchange of its gradients over peer memory ----------------------
 * One process per GPU on one NVLink/NVSwitch node.  Every rank owns an exchange buffer of dd_exchange_bytes() bytes
 * (zero-filled once by the caller, then owned by the library) that is mapped into every peer (CUDA IPC / symmetric
 * memory: done by the host layer); buf[r] is rank r's buffer as addressed from THIS process.
 * dd_predictor_bwd_allreduce == dd_predictor_bwd followed by all-reduce(sum) of the 14 gradient tensors, as ONE
 * stream-ordered sequence with no collective library call: the kernels that produce the gradients (the FC backward at
 * the start of the backward, 80 % of the bytes; the deferred weight-gradient slice reduction at its end) store every
 * value as one 8-byte {value, tag} word straight into slot [rank] of every peer's buffer over NVLink, and a final kernel
 * polls the world's words element by element and sums them in rank order (deterministic, identical on every rank).
 * Replaces DDP's bucketed all-reduce of these parameters (engine/trainer.py:223) for the stand-alone pipeline. */
#define DD_MAX_PEERS 8
typedef struct dd_peer_exchange {
    int rank, world;
    void* buf[DD_MAX_PEERS];
} dd_peer_exchange;
size_t dd_exchange_bytes(void);
int dd_predictor_bwd_allreduce(const float* r, const dd_predictor_tensors* w, const float* acts,
                               const float* dfeat, const dd_predictor_tensors* grads, int B, void* ws,
                               size_t ws_bytes, const dd_peer_exchange* px, void* stream);

/* ---- a6..a12: regressors + DeDark -> WB -> Gamma -> Contrast -> USM, one fused pass -------------
 * Replaces the body of lowlight_recovery.forward after the predictor (llie.py:34-40,49-52;
 * filtersB.py:32-37,144-259,289-303; util_filters.py:270-273,295-304,316-317).
 *   x    [B,3,H,W]      A  [B,3] or NULL (=> 0.8)      IcA [B,1,H,W] or NULL (=> 0.5)
 *   feat [B,15] raw fc2 output      y [B,3,H,W] (must not alias x)                                 */
int dd_recovery_fwd(const float* x, const float* A, const float* IcA, const float* feat, float* y,
                    int B, int H, int W, void* stream);

/* ---- a14: backward of the fused pass ------------------------------------------------------------
 * g = dL/dy [B,3,H,W]  ->  dfeat [B,15] (columns 1 and 5..12 are written as exact zeros) and, if
 * dx != NULL, dL/dx through the filter chain (overwritten).  Nothing is saved by the forward: the
 * chain is recomputed from x.  ws: DD_WS_RECOVERY_BWD bytes.  Reductions are fixed-order
 * (deterministic; no float atomics). */
int dd_recovery_bwd(const float* x, const float* A, const float* IcA, const float* feat,
                    const float* g, float* dfeat, float* dx, int B, int H, int W, void* ws,
                    size_t ws_bytes, void* stream);

/* ---- bf16 I/O mode ---------------------------------------------------------------------------------
 * The reference trains under autocast (engine/trainer.py:330); SURVEY.md section 8(d): x and the cotangent are read as bf16
 * (6 bytes per pixel instead of 12), y is fp32 like the reference's output unless the caller opts into bf16.  Same
 * semantics as the fp32 entry points with element types DD_F32 / DD_BF16 for x (and dx), y and g; A, IcA, feat, dfeat and r
 * stay fp32.  When any of them is bf16 the blur runs as plain TF32 tcgen05 GEMMs (bf16 values are exact in TF32; gate
 * 2e-2 relative to the reference on the same bf16-rounded input) and the shape must satisfy W % 4 == 0 (backward: also H,
 * W >= 14); with all-fp32 types the calls forward to dd_recovery_fwd / dd_recovery_bwd / dd_resize256. */
int dd_synth_fwd_ex(const void* src, int src_dtype, float p, const float* lut256, const float* clean_lut256,
                    float* clean_out, void* dark_out, int dark_dtype, uint8_t* dark_u8, float* rec_out, long long n,
                    void* ws, size_t ws_bytes, void* stream);   /* dark_out fp32 or bf16 (rounded to nearest even from the exact fp32 value; the loss uses the fp32 value) */
int dd_resize256_ex(const void* x, int x_dtype, float* r, int B, int H, int W, void* stream);
int dd_recovery_fwd_ex(const void* x, int x_dtype, const float* A, const float* IcA, const float* feat, void* y,
                       int y_dtype, int B, int H, int W, void* stream);
int dd_recovery_bwd_ex(const void* x, int x_dtype, const float* A, const float* IcA, const float* feat,
                       const void* g, int g_dtype, float* dfeat, void* dx, int B, int H, int W, void* ws,
                       size_t ws_bytes, void* stream);

/* ---- unit-test hook of the tensor-core blur engine -------------------------------------------------
 * y = the reflect-padded 25x25 Gaussian of x (filtersB.py:154-175: F.pad(..., 'reflect') + conv2d per channel), computed by
 * the same tcgen05 engine the fused kernels use (x3 != 0: 3xTF32 operand split, the fp32 mode; 0: plain TF32, the bf16
 * mode).  Needs W % 4 == 0 and a 16-byte aligned x.  Selecting the engine of the fp32 entry points: the CUDA-core (FFMA2)
 * kernels by default (still the faster ones at fp32 precision), the tensor-core kernels (3xTF32) with the environment
 * variable DEDARK_BLUR=tc when the rows are 16-byte aligned; the bf16 entry points always use the tensor cores. */
int dd_debug_blur_tc(const float* x, float* y, int B, int H, int W, int x3, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DEDARK_B200_H */
