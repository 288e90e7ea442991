// Internal (non-ABI) declarations shared by the translation units of libescb200.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <vector>

#include "../../include/escb200.h"
#include "gemm.cuh"
#include "loaders.cuh"
#include "tc_gemm.cuh"
#include "mlp_fused.h"

namespace escb {

inline int ldc(int C) { return (C + 3) & ~3; }
// head width of the qkv buffer [rows][3][heads][hdp]: heads are padded to a multiple of 4 floats (15 -> 16) so the
// attention core reads q / k / v rows with aligned vector loads; 6 stays 6 (8-byte vectors)
inline int head_pad(int hd) { return hd == 6 ? 6 : (hd + 3) & ~3; }

struct LnW { const float* g; const float* b; };     // padded to a multiple of 4 with zeros

struct BlockW {
    LnW n1, n2;
    GemmWeight qkv, proj, fc1, fc2;
    GemmWeight qkvh_p, fc1_p;                        // gamma-folded images for the post-GEMM LayerNorm mode (tc.img null: not built)
    GemmWeight qkvh;                                 // qkv in head-major columns for the fused attention epilogue (tc.img null: not fusable)
    const float* relbias;                            // [heads][16][16], gathered from the (49, heads) table
    mf::Weights mlpf;                                // fused LN2 -> fc1 -> GELU -> fc2 -> +x kernel (plan.ok == 0: not built, C > 96)
};

struct LayerW {
    int C, heads, hd, hdp, depth, scale, out_dim;    // scale: 0 none, 1 down (PatchMerge), 2 up (PatchSplit); hdp = head_pad(hd)
    BlockW blk[ESCB_MAX_DEPTH];
    LnW sn;
    GemmWeight sub;
    GemmWeight sub_p;                                // PatchSplit weight, post-GEMM LayerNorm variant
};

struct QuantW {
    int in_dim, in_freq, d, frame_dim, ncodes, l2norm;
    GemmWeight down;                                 // K = frame_dim in (h,o,c) order, N = 3d (block structured)
    GemmWeight down_g[3];                            // per group: K = frame_dim / 3 (its (o,c) third of every h run), N = d
    int run;                                         // 2C/3 when the groups are equal (o,c) thirds (else 0: stacked path only)
    GemmWeight up;                                   // K = 3d, N = frame_dim in (h,o,c) order
    const float* raw;                                // [3][ncodes][d]
    const float* cbt;                                // [3][d][ncodes] L2-normalised, transposed
    const float* cnorm;                              // [3][ncodes] squared norms of cbn rows
};

// RVQCodecs: ProductResidualVectorQuantize at the bottleneck (quantization.py:276-378).  The projections reuse the
// product-VQ layouts of QuantW (per-group down-projections, block-structured up-projection); every group owns `S`
// residual codebooks.
struct RvqW {
    QuantW q;                                        // geometry + down_g / up; its own raw / cbt / cnorm are unused
    int S;                                           // num_rvqs
    const float* raw;                                // [3][S][ncodes][d]
    const float* cbt;                                // [3][S][d][ncodes] L2-normalised, transposed
    const float* cnorm;                              // [3][S][ncodes]
};

struct Conv3Weights { float w[9 * 64 * 2]; };        // [tap][c][2], c < kEmbedMaxC
// patch-embedding parameters of the shipped geometry (C0 = 45, 3 x 2 patches of 2 planes) as a by-value kernel parameter
struct EmbedWeights { float w[45 * 12]; float b[45], g[45], be[45]; };

struct FrontW {
    int F, win, hop, nov, C0, pf, pt;
    GemmWeight dft;                                  // [win][2F] windowed DFT basis
    GemmWeight idft;                                 // [nov*2F][hop] windowed inverse basis
    const float* wsq;                                // [win] squared synthesis window
    const float* embed_w; const float* embed_b; LnW embed_ln;
    EmbedWeights embed_k;                            // the same values for patch_embed45_kernel (valid when embed_k_ok)
    int embed_k_ok;
    GemmWeight de1;                                  // conv5x5 as implicit GEMM, K = 25*ldc(C0)
    const float* de2_w; const float* de2_b;          // [9][C0][2], [2]
    float de2_bias[2];
    Conv3Weights de2_k;                              // the same taps as a by-value kernel parameter (constant bank)
};

// Kernel classes for launch accounting / per-op timing (escb_profile_begin/end).
enum OpId {
    OP_STFT, OP_EMBED, OP_QKV, OP_ATTN, OP_PROJ, OP_MLP1, OP_MLP2, OP_MERGE, OP_SPLIT, OP_PVQ_DOWN, OP_ARGMIN,
    OP_PVQ_UP, OP_VQLOSS, OP_DEEMBED1, OP_DEEMBED2, OP_ISTFT, OP_LAYOUT, OP_QKV_ATTN, OP_MLP_FUSED, OP_PVQ_STREAM, OP_COUNT
};
static_assert(OP_COUNT == ESCB_NUM_OPS, "escb200.h ESCB_NUM_OPS out of date");

struct ProfRec { cudaEvent_t a, b; int op; double flops, bytes; };
struct Profiler { std::vector<ProfRec> recs; };

// Post-GEMM LayerNorm (tc_gemm.cuh LNP) measured per class at 36 clips: PatchSplit 0.77 -> 0.65 ms, but the two
// 8-producer-warp kernels are slower with it (fused attention 6.05 -> 6.47 ms, mlp1 5.36 -> 5.64 ms), so bits 4 and 8 are on (PatchMerge 0.75 -> 0.52 ms).
constexpr int kLnPostDefault = 12;

struct Launcher {                                    // stream + launch accounting + first-error latch
    cudaStream_t st = nullptr;
    long long launches = 0;
    cudaError_t err = cudaSuccess;
    bool tc = true;                                  // dense layers on the tcgen05 engine (false: fp32 SIMT engine)
    bool pvq_tc = true;                              // product-VQ projections on the tcgen05 engine
    bool emit_stats = true;                          // the fused MLP emits the next LayerNorm's statistics (no pre-kernel for it)
    bool fuse_pvq = true;                            // one launch per RVQ stream step (ESCB_FUSE_PVQ=0: down GEMM + argmin + up GEMM)
    bool fuse_mlp = true;                            // LN2 -> fc1 -> GELU -> fc2 -> +x in one launch where a plan exists (ESCB_FUSE_MLP=0: the unfused pair)
    int ln_post = kLnPostDefault;                    // bit mask (ESCB_LN_POST): LayerNorm applied after the GEMM in 1 fused qkv+attention, 2 mlp1, 4 PatchSplit, 8 PatchMerge
    Profiler* prof = nullptr;                        // non-null: bracket every launch with CUDA events
    float2* ln_stats = nullptr;                      // [max rows] LayerNorm statistics scratch (tc engine)
    int* code_err = nullptr;                         // device view of the handle's host-mapped bad-code latch (ACodes)
    unsigned long long* trace = nullptr;             // ESCB_TC_TRACE builds: 16 counters per GEMM launch
    int trace_n = 0;
    unsigned long long* next_trace() { return trace ? trace + 16 * (size_t)(trace_n++ % 1024) : nullptr; }
    bool open = false;
    // flops / bytes: ALGORITHMIC work of the launch that follows (true dims, no padding, no 3x anything)
    void begin(int op, double flops, double bytes) {
        if (!prof) return;
        ProfRec r{nullptr, nullptr, op, flops, bytes};
        cudaEventCreate(&r.a);
        cudaEventCreate(&r.b);
        cudaEventRecord(r.a, st);
        prof->recs.push_back(r);
        open = true;
    }
    void note(cudaError_t e) {
        ++launches;
        if (err == cudaSuccess && e != cudaSuccess) err = e;
        if (prof && open) { cudaEventRecord(prof->recs.back().b, st); open = false; }
    }
};

constexpr float kLnEps = 1e-5f;
#ifndef ESCB_ROLE_PVQUP
#define ESCB_ROLE_PVQUP 1
#endif
constexpr int kPvqUpWide = ESCB_ROLE_PVQUP;      // product-VQ up-projection (K <= 96, N up to 4608: all epilogue): 16 epilogue warps
#ifndef ESCB_ROLE_SPLIT
#define ESCB_ROLE_SPLIT 1
#endif
constexpr int kSplitWide = ESCB_ROLE_SPLIT;      // PatchSplit GEMM role split
#ifndef ESCB_ROLE_PROJ
#define ESCB_ROLE_PROJ 1
#endif
constexpr int kProjWide = ESCB_ROLE_PROJ;       // attention output projection role split
#ifndef ESCB_ROLE_MLP2
#define ESCB_ROLE_MLP2 1
#endif
// mlp2 (K = 4N).  Round 1 (every level, one accumulator): 8 + 16 measured 4.65 ms, 16 + 8 5.01 ms.  Since the fused MLP
// kernel took C <= 96 and the accumulator split left the deeper levels ONE accumulator region (MMAs and epilogue of a
// tile alternate), the epilogue is on the critical path: 16 + 8 measured 2.50 ms, 12 + 12 2.79, 8 + 16 3.06.
constexpr int kMlp2Wide = ESCB_ROLE_MLP2;
#ifndef ESCB_ROLE_MLP1
#define ESCB_ROLE_MLP1 2
#endif
constexpr int kMlp1Wide = ESCB_ROLE_MLP1;       // role split of the GELU GEMM (codes: tc::role_e): 2 = 12 + 12 (5.14 ms), 1 = 16 epilogue + 8 producer warps (5.37), 0 = 8 + 16 (7.1)
constexpr int kEmbedMaxC = 64;     // patch_embed_kernel register budget: h_dims[0] <= 64
constexpr int kEmbedMaxK = 16;     // 2 * patch_freq * patch_time <= 16

// ---- swin.cu : one reference SwinBlock = qkv -> attention -> proj -> mlp1 -> mlp2
void op_qkv(Launcher& L, const BlockW& w, const float* x, int ld, const WindowGeom& g, long long M, float* qkv, int ldq);
void op_attention(Launcher& L, const float* qkv, int ldq, float* att, int ldo, const float* relbias, int heads,
                  int hd, int hdp, int C, long long nwin, bool masked, const WindowGeom& g);
void op_qkv_attn(Launcher& L, const BlockW& w, int heads, int hd, const float* x, int ld, const WindowGeom& g, long long M,
                 float* att, int ldo, bool masked, bool stats_ready = false);
bool qkv_attn_takes_stats(const Launcher& L, const BlockW& w);
bool attention_fusable(int hd);
void op_proj(Launcher& L, const BlockW& w, const float* att, int lda, const float* resid, float* y, int ld,
             const WindowGeom& g, long long M);
void op_mlp1(Launcher& L, const BlockW& w, const float* x, int ld, long long M, float* hid, int ldh);
void op_mlp2(Launcher& L, const BlockW& w, const float* hid, int ldh, long long M, float* x, int ld);
void op_mlp_fused(Launcher& L, const BlockW& w, float* x, int ld, long long M, const mf::StatsOut& so);
void op_merge(Launcher& L, const LayerW& w, const float* x, int ld, int B, int H, int W, float* y, int ldy);
void op_split(Launcher& L, const LayerW& w, const float* x, int ld, int B, int H, int W, float* y, int ldy, bool stats_ready = false);
cudaError_t swin_init();
cudaError_t frontend_init();

// ---- pvq.cu : product VQ of one stream
void op_pvq_down(Launcher& L, const QuantW& q, const float* enc, const float* dec, int B, int W, float* ze, int ldz);
void op_argmin(Launcher& L, const QuantW& q, int g_first, int groups, const float* ze, int ldz, long long rows,
               long long* out, int T, long long bstride);
void op_pvq_up(Launcher& L, const QuantW& q, const long long* codes, int S, int s, const float* dec, int B, int W,
               float* out);
void op_code_histogram(Launcher& L, const long long* codes, int B, int S, int G, int T, int ncodes, float* counts);
// fused stream step (kernels.cuh pvq_stream_kernel); returns false when this quantizer's geometry has no fused kernel
bool op_pvq_stream(Launcher& L, const QuantW& q, const float* enc, const float* dec, int B, int W, long long* codes, int S,
                   int s, float* out, float* ze, int ldz);
cudaError_t pvq_init();
// RVQCodecs (pvq.cu): residual chain on the projected vectors, code gather-sum for decode, eval-mode loss reduce
void op_rvq_chain(Launcher& L, const RvqW& w, const float* ze, int ldz, long long rows, int S, long long* codes, int T,
                  float* zq, float* se);
void op_rvq_gather(Launcher& L, const RvqW& w, const long long* codes, int S, long long rows, int T, float* zq, int ldz);
void op_rvq_up(Launcher& L, const RvqW& w, const float* zq, int ldz, int B, int W, float* out);
void op_rvq_loss(Launcher& L, const float* se, int B, int T, int d, float* loss);
void op_vq_loss(Launcher& L, const QuantW& q, const float* ze, int ldz, const long long* codes, int S, int s, int B,
                int T, float* loss);

// ---- frontend.cu : STFT / patch embed / patch de-embed / inverse STFT / layout helpers
void op_stft(Launcher& L, const FrontW& f, const float* audio, int B, long long Ls, int T, float* Sf);
void op_patch_embed(Launcher& L, const FrontW& f, const float* Sf, int B, int T, int H, int W, float* tok, int ld);
void op_deembed(Launcher& L, const FrontW& f, const float* tok, int ld, int B, int H, int W, float* Y1, float* Xf);
void op_istft(Launcher& L, const FrontW& f, const float* Xf, int B, int T, float* audio);
void op_transpose(Launcher& L, const float* in, float* out, int B, int R, int C);
void op_repitch(Launcher& L, const float* src, int lds, float* dst, int ldd, int C, long long rows);

}  // namespace escb
