/*
 * escb200.h — C ABI of the B200-native ESC encode/decode hot path.
 *
 * The reference (yzGuu830/efficient-speech-codec) is pure Python on stock PyTorch ops and has no FFI of
 * its own, so each entry point below cites the reference *Python* interface it replaces
 * (paths relative to the reference root).  INTEGRATION.md shows the ctypes binding a maintainer adds.
 *
 * Conventions
 *   - plain pointers and sizes only; no torch / C++ types cross this boundary;
 *   - every function returns 0 on success or a negative ESCB_E* code; the message for the calling thread's
 *     last failure is returned by escb_last_error(); nothing throws across the ABI;
 *   - pointers named *_dev are device pointers on the handle's CUDA device, owned by the caller;
 *     the library owns only the repacked weights inside the handle;
 *   - all work is enqueued on the `stream` argument (a cudaStream_t passed as void*; NULL = legacy
 *     default stream) with no implicit synchronisation unless stated;
 *   - the handle is immutable after escb_finalize(): concurrent calls on different streams with
 *     different workspaces are safe.
 *   - there is NO CPU fallback: every compute entry point fails with ESCB_ENODEV when no CUDA device
 *     is usable.
 */
#ifndef ESCB200_H
#define ESCB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define ESCB_API __attribute__((visibility("default")))
#else
#define ESCB_API
#endif

#define ESCB_ABI_VERSION 2
#define ESCB_MAX_LEVELS 8
#define ESCB_MAX_DEPTH 8
#define ESCB_NUM_OPS 20

enum {
    ESCB_OK = 0,
    ESCB_EINVAL = -1,     /* bad argument / unsupported geometry (reference raises ValueError / AssertionError) */
    ESCB_ENODEV = -2,     /* no usable CUDA device */
    ESCB_ECUDA = -3,      /* a CUDA runtime call or kernel launch failed */
    ESCB_ESTATE = -4,     /* call order violated (e.g. compute before escb_finalize, missing weight) */
    ESCB_ENOMEM = -5,     /* workspace too small / allocation failed */
    ESCB_EKEY = -6        /* unknown weight name */
};

/* Mirrors the keyword arguments of ESC.__init__  (esc/models/codecs.py:11-18) and the STFT parameters derived
 * from them in BaseAudioCodec.__init__ (esc/models/base.py:22-27). */
typedef struct escb_config {
    int32_t in_freq;                        /* 192: one-sided STFT bins; n_fft = 2*(in_freq-1)                 */
    int32_t win_length;                     /* int(win_len*sr*1e-3) = 320 samples                               */
    int32_t hop_length;                     /* int(hop_len*sr*1e-3) = 80 samples                                */
    int32_t patch_freq, patch_time;         /* patch_size = [3, 2]                                              */
    int32_t num_levels;                     /* len(h_dims) == max_streams (6)                                   */
    int32_t h_dims[ESCB_MAX_LEVELS];        /* encoder channel widths, fine to coarse [45,72,96,144,192,384]    */
    int32_t swin_heads[ESCB_MAX_LEVELS];    /* num_levels-1 entries [3,6,12,24,24]                              */
    int32_t swin_depth;                     /* blocks per TransformerLayer: 2 (Base) / 4 (Large)                */
    int32_t window_size;                    /* 4 (only value supported)                                         */
    int32_t mlp_hidden_mult;                /* int(mlp_ratio): hidden = mult * C (4)                            */
    int32_t overlap;                        /* frames quantized together (2)                                    */
    int32_t group_size;                     /* product-VQ groups (3, only value supported)                      */
    int32_t codebook_size;                  /* 1024                                                             */
    int32_t codebook_dims[ESCB_MAX_LEVELS]; /* per stream                                                       */
    int32_t l2norm;                         /* 1: cosine-style argmin (only value supported)                    */
    int32_t num_rvqs;                       /* 0: ESC (cross-scale product VQ, codecs.py:9-94).  > 0: RVQCodecs
                                             * (codecs.py:96-181): ONE ProductResidualVectorQuantize at the bottleneck
                                             * with num_rvqs residual codebooks per group of dimension codebook_dims[0]
                                             * (quantization.py:139-378) and the plain Decoder (base.py:161-203)      */
} escb_config;

typedef struct escb_handle escb_handle;

/* ------------------------------------------------------------------------------------------- lifecycle */

/* Library/ABI version (ESCB_ABI_VERSION the library was built with). */
ESCB_API int escb_abi_version(void);

/* Message of the calling thread's last failure ("" if none). Never NULL. */
ESCB_API const char* escb_last_error(void);

/* Replaces ESC.__init__ (esc/models/codecs.py:11-28): validates the geometry and allocates weight storage on
 * the CURRENT CUDA device. */
ESCB_API int escb_create(const escb_config* cfg, escb_handle** out);
ESCB_API void escb_destroy(escb_handle* h);

/* The checkpoint tensors the handle expects, named exactly like the reference's state_dict keys
 * (e.g. "encoder.blocks.0.swint_blocks.1.attn.qkv.weight").  Buffers that are pure functions of
 * the config (ft.window / ift.window / relative_position_index) are not listed: they are regenerated. */
ESCB_API int escb_num_weights(const escb_handle* h);
ESCB_API const char* escb_weight_name(const escb_handle* h, int index);
ESCB_API int64_t escb_weight_numel(const escb_handle* h, int index);

/* Replaces nn.Module.load_state_dict for one tensor (scripts/compress.py:23-25): copies `numel` contiguous
 * fp32 values in the reference's own layout.  `is_device` selects a device or host source pointer. */
ESCB_API int escb_set_weight(escb_handle* h, const char* name, const float* data, int64_t numel, int is_device);

/* Repack all weights into kernel layouts (transposed/padded GEMM operands, gathered relative-position bias,
 * normalised codebooks + |c|^2, windowed DFT bases).  Synchronous.  Must be called after all weights are set
 * and again after any weight changes. */
ESCB_API int escb_finalize(escb_handle* h);

/* ------------------------------------------------------------------------------------------- geometry */

/* Time patches W for a clip of num_samples (Encoder.forward, esc/models/base.py:149): (1 + L/hop) / patch_time.
 * Fails with ESCB_EINVAL if W is not a multiple of `overlap` (assert at esc/modules/vq/quantization.py:407). */
ESCB_API int escb_time_patches(const escb_handle* h, int64_t num_samples, int32_t* W);
/* Decoded samples for W time patches: hop * (patch_time*W - 1). */
ESCB_API int64_t escb_decoded_samples(const escb_handle* h, int32_t W);
/* Scratch bytes needed by encode/decode/forward for `batch` clips of W time patches. */
ESCB_API int escb_workspace_bytes(const escb_handle* h, int32_t batch, int32_t W, size_t* bytes);

/* ------------------------------------------------------------------------------------------- the hot path */

/* Replaces ESC.encode (esc/models/codecs.py:68-81 -> base.py:29-37,143-158 -> csrvq.py:131-158).
 *   audio_dev  [batch, num_samples] fp32
 *   codes_dev  [batch, num_streams, group_size, W/overlap] int64 (written)
 * feat_shape of the reference's return value is (bottom_freq, W) with W from escb_time_patches(). */
ESCB_API int escb_encode(escb_handle* h, const float* audio_dev, int32_t batch, int64_t num_samples, int32_t num_streams,
                int64_t* codes_dev, void* workspace_dev, size_t workspace_bytes, void* stream);

/* Replaces ESC.decode (esc/models/codecs.py:83-94 -> csrvq.py:160-182 -> base.py:39-47).
 *   codes_dev      [batch, num_streams, group_size, W/overlap] int64
 *   audio_dev      [batch, escb_decoded_samples(W)] fp32 (written)
 *   recon_feat_dev optional [batch, 2, in_freq, patch_time*W] fp32 (written when non-NULL) */
ESCB_API int escb_decode(escb_handle* h, const int64_t* codes_dev, int32_t batch, int32_t num_streams, int32_t W,
                float* audio_dev, float* recon_feat_dev, void* workspace_dev, size_t workspace_bytes, void* stream);

/* Replaces ESC.forward in eval mode (esc/models/codecs.py:30-66 -> csrvq.py:97-129): one fused
 * encode+decode pass.  Optional outputs may be NULL.
 *   raw_feat_dev   [batch, 2, in_freq, 1 + num_samples/hop]
 *   recon_feat_dev [batch, 2, in_freq, patch_time*W]
 *   vq_loss_dev    [batch]  (cm_loss == cb_loss in eval mode, esc/modules/vq/codebook.py:71-73) */
ESCB_API int escb_forward(escb_handle* h, const float* audio_dev, int32_t batch, int64_t num_samples, int32_t num_streams,
                 int64_t* codes_dev, float* audio_out_dev, float* raw_feat_dev, float* recon_feat_dev,
                 float* vq_loss_dev, void* workspace_dev, size_t workspace_bytes, void* stream);

/* The same pass from a precomputed spectrum: ESC.forward(x, x_feat, ...) with x_feat given (esc/models/codecs.py:33-34,
 * after its rearrange to [batch, 2, in_freq, frames]); the STFT is skipped, W = (frames - patch_time) / patch_time + 1. */
ESCB_API int escb_forward_feat(escb_handle* h, const float* planes_dev, int32_t batch, int32_t frames, int32_t num_streams,
                      int64_t* codes_dev, float* audio_out_dev, float* recon_feat_dev, float* vq_loss_dev,
                      void* workspace_dev, size_t workspace_bytes, void* stream);

/* Host-buffer variants: the same calls for callers that hold host memory (scripts/compress.py:19-35 runs the
 * codec on a wav it just read).  They allocate device scratch internally, copy in, run, copy out and
 * synchronise `stream` before returning.  Host buffers should be pinned for full PCIe bandwidth. */
ESCB_API int escb_encode_host(escb_handle* h, const float* audio_host, int32_t batch, int64_t num_samples,
                     int32_t num_streams, int64_t* codes_host, void* stream);
ESCB_API int escb_decode_host(escb_handle* h, const int64_t* codes_host, int32_t batch, int32_t num_streams, int32_t W,
                     float* audio_host, void* stream);

/* ------------------------------------------------------------------------------------------- unit entry points
 * One per reference module on the path, so the parity tests read like tests of the reference's modules. */

/* BaseAudioCodec.spec_transform (esc/models/base.py:29-37): audio [B,L] -> planes [B,2,in_freq,1+L/hop]. */
ESCB_API int escb_stft(escb_handle* h, const float* audio_dev, int32_t batch, int64_t num_samples, float* planes_dev,
              void* workspace_dev, size_t workspace_bytes, void* stream);
/* BaseAudioCodec.audio_reconstruct (esc/models/base.py:39-47): planes [B,2,in_freq,T] -> audio [B,hop*(T-1)]. */
ESCB_API int escb_istft(escb_handle* h, const float* planes_dev, int32_t batch, int32_t frames, float* audio_dev,
               void* workspace_dev, size_t workspace_bytes, void* stream);
/* PatchEmbed.forward (esc/modules/transformer/scale.py:42-50): planes [B,2,in_freq,T] -> tokens [B,H*W,C0]. */
ESCB_API int escb_patch_embed(escb_handle* h, const float* planes_dev, int32_t batch, int32_t frames, float* tokens_dev,
                     void* workspace_dev, size_t workspace_bytes, void* stream);
/* PatchDeEmbed.forward (esc/modules/transformer/scale.py:73-81): tokens [B,H*W,C0] -> planes [B,2,in_freq,2W]. */
ESCB_API int escb_patch_deembed(escb_handle* h, const float* tokens_dev, int32_t batch, int32_t W, float* planes_dev,
                       void* workspace_dev, size_t workspace_bytes, void* stream);
/* TransformerLayer.forward (esc/modules/transformer/attention.py:48-91) for the layer at `layer_index`:
 *   0 = encoder.pre_nn, 1..L-1 = encoder.blocks[i-1], L..2L-2 = decoder.blocks[i-L], 2L-1 = decoder.post_nn
 *   (L = num_levels).  x [B,H*W,C] -> y [B,H'*W,C'] dense row-major, H given, H' = H/2, 2H or H. */
ESCB_API int escb_swin_layer(escb_handle* h, int32_t layer_index, const float* x_dev, int32_t batch, int32_t H, int32_t W,
                    float* y_dev, void* workspace_dev, size_t workspace_bytes, void* stream);
/* ProductVectorQuantize.encode (esc/modules/vq/quantization.py:74-91) of stream `q` applied to enc - dec
 * (CrossScaleRVQ.csrvq_encode, esc/models/csrvq.py:50-54); dec_dev may be NULL.
 *   enc_dev/dec_dev [B, in_freq_q*W, in_dim_q]; codes_dev [B, group_size, W/overlap] int64. */
ESCB_API int escb_pvq_encode(escb_handle* h, int32_t q, const float* enc_dev, const float* dec_dev, int32_t batch, int32_t W,
                    int64_t* codes_dev, void* workspace_dev, size_t workspace_bytes, void* stream);
/* ProductVectorQuantize.decode + post_fuse (quantization.py:93-108, csrvq.py:56-60): out = vq.decode(codes) + dec
 * (dec_dev may be NULL).  out_dev [B, in_freq_q*W, in_dim_q]. */
ESCB_API int escb_pvq_decode(escb_handle* h, int32_t q, const int64_t* codes_dev, const float* dec_dev, int32_t batch,
                    int32_t W, float* out_dev, void* workspace_dev, size_t workspace_bytes, void* stream);
/* One cross-scale RVQ stream step, CrossScaleRVQ.csrvq in eval mode (esc/models/csrvq.py:23-48 = csrvq_encode +
 * csrvq_decode, :50-60): codes = vq.encode(enc - dec), out = vq.decode(codes) + dec, in ONE kernel launch.  dec_dev may be
 * NULL (stream 0), out_dev may be NULL (codes only: the last transmitted stream of ESC.encode) and may alias dec_dev.
 *   enc_dev/dec_dev/out_dev [B, in_freq_q*W, in_dim_q]; codes_dev [B, group_size, W/overlap] int64. */
ESCB_API int escb_pvq_stream(escb_handle* h, int32_t q, const float* enc_dev, const float* dec_dev, int32_t batch, int32_t W,
                    int64_t* codes_dev, float* out_dev, void* workspace_dev, size_t workspace_bytes, void* stream);
/* Codebook.quantize_to_code (esc/modules/vq/codebook.py:20-43), THE RVQ argmin, for group `g` of stream `q`:
 *   z_dev [rows, codebook_dim_q] (already down-projected) -> idx_dev [rows] int64. */
ESCB_API int escb_codebook_argmin(escb_handle* h, int32_t q, int32_t g, const float* z_dev, int64_t rows, int64_t* idx_dev,
                         void* stream);

/* EntropyCounter.update of the evaluation sweep (scripts/metrics.py:37-51, fed by scripts/test.py:44); needs no
 * handle (the counter is model-agnostic in the reference too):
 *   codes_dev [B, S, G, T] int64 -> counts_dev [S*G, codebook_size] fp32, counts += occurrences (the caller zeroes
 *   counts at EntropyCounter.reset_stats).  Indices outside [0, codebook_size) are not counted. */
ESCB_API int escb_code_histogram(const int64_t* codes_dev, int32_t batch, int32_t num_streams, int32_t group_size,
                        int32_t frames, int32_t codebook_size, float* counts_dev, void* stream);

/* Per-kernel-class timing (bench.py's roofline figures).  Between escb_profile_begin and escb_profile_end every
 * kernel launched for this handle is bracketed by CUDA events on its stream; escb_profile_end synchronises the
 * device and returns, per class, the launch count, the summed device time and the summed ALGORITHMIC flops /
 * bytes (true dims, no padding).  Debug facility: not thread-safe, not for use inside a timed throughput run. */
typedef struct escb_op_stat {
    const char* name;
    int64_t launches;
    double ms;
    double flops;
    double bytes;
} escb_op_stat;
ESCB_API int escb_profile_begin(escb_handle* h);
ESCB_API int escb_profile_end(escb_handle* h, escb_op_stat* stats /* [ESCB_NUM_OPS] */, int32_t* n);

/* Code indices handed to escb_decode / escb_pvq_decode are caller data (saved .pth files).  An index outside
 * [0, codebook_size) - where the reference's F.embedding raises IndexError (esc/modules/vq/codebook.py:53) - is
 * decoded as index 0 (no out-of-bounds read) and latched; this call returns ESCB_EINVAL once if a kernel that has
 * COMPLETED latched one since the last poll, else ESCB_OK.  escb_decode_host polls after its synchronisation. */
ESCB_API int escb_poll_error(escb_handle* h);

/* Host-only diagnostic (no device, no handle): how the tcgen05 engine would tile an nn.Linear weight [N, K] and split
 * its accumulators (csrc/tc_gemm.cuh choose_tiling / acc_policy; DESIGN.md section 2 "accumulator split").  role: 0, 1, 2
 * = 8 + 16, 16 + 8, 12 + 12 epilogue + producer warps.  out[8] = { BN (UMMA N), sub-tiles per output tile, n-tiles,
 * 32-wide K blocks, weights resident in shared memory (0/1), main accumulators per sub-tile, corrections in their own
 * accumulator (0/1), TMEM columns of one output tile }.  Returns ESCB_EINVAL for shapes the engine cannot tile. */
ESCB_API int escb_tiling_info(int32_t N, int32_t K, int32_t role, int32_t* out);

#ifdef ESCB_TC_TRACE
/* Debug builds only (make trace): per-role cycle counters of the tcgen05 engine, 16 x 1024 u64 (tools/trace_step.py).
 * First call allocates and arms the buffer; later calls synchronise and copy it to out_host. */
ESCB_API int escb_debug_trace(escb_handle* h, unsigned long long* out_host);
#endif

/* Number of kernels the library has launched on behalf of this handle since creation (bench.py's gpu_launches). */
ESCB_API int64_t escb_launch_count(const escb_handle* h);

#ifdef __cplusplus
}
#endif
#endif /* ESCB200_H */
