// Swin block ops: host launchers over gemm.cuh / kernels.cuh.  Reference: attention.py:129-178 (SwinBlock),
// :215-244 (WindowAttention), :258-272 (FeedForward); scale.py:83-145 (PatchMerge / PatchSplit).
#include <stdlib.h>

#include <algorithm>

#include "internal.h"
#include "kernels.cuh"

namespace escb {

static inline LnParams lnp(Launcher& L, const LnW& w) { return LnParams{w.g, w.b, kLnEps, L.ln_stats, L.next_trace(), nullptr, nullptr}; }
static inline LnParams noln(Launcher& L) { return LnParams{nullptr, nullptr, 0.f, nullptr, L.next_trace(), nullptr, nullptr}; }
// post-GEMM LayerNorm: statistics from the pre-kernel, column vectors of the gamma-folded weight
static inline LnParams lnpost(Launcher& L, const GemmWeight& gw) { return LnParams{nullptr, nullptr, kLnEps, L.ln_stats, L.next_trace(), gw.cs, gw.bw}; }

void op_qkv(Launcher& L, const BlockW& w, const float* x, int ld, const WindowGeom& g, long long M, float* qkv, int ldq) {
    L.begin(OP_QKV, 2.0 * M * 3.0 * w.qkv.K * w.qkv.K, 4.0 * (1.0 * M * w.qkv.K + 3.0 * M * w.qkv.K));   // true dims (3C x C)
    AWindow al{x, ld, g};
    EpiRows<false, false> ep{qkv, w.qkv.bias, nullptr, ldq, 0};
    if (L.tc) ++L.launches, L.note(tc::launch<true, AWindow, EpiRows<false, false>>(L.st, al, lnp(L, w.n1), w.qkv, M, ep));
    else L.note(GemmLauncher<true, AWindow, EpiRows<false, false>, 7, 9>::launch(L.st, al, lnp(L, w.n1), w.qkv, M, ep));
}

// qkv projection + window attention core in one launch: LN1 -> window gather -> GEMM -> attention in the epilogue
#define ESCB_FUSED_HDS(X) X(6) X(8) X(12) X(15) X(16) X(24)
bool attention_fusable(int hd) {
    switch (hd) {
#define X(n) case n: return true;
        ESCB_FUSED_HDS(X)
#undef X
        default: return false;
    }
}

// The fused kernel can take LayerNorm statistics written by the producer of x (window-row indexed, mlp_fused.cu) when it
// normalises in its A producers: the post-GEMM form reads the statistics of padded rows too, which no producer writes.
bool qkv_attn_takes_stats(const Launcher& L, const BlockW& w) { return !((L.ln_post & 1) && w.qkvh_p.tc.img); }

void op_qkv_attn(Launcher& L, const BlockW& w, int heads, int hd, const float* x, int ld, const WindowGeom& g, long long M,
                 float* att, int ldo, bool masked, bool stats_ready) {
    const int C = w.qkvh.K;
    L.begin(OP_QKV_ATTN, 2.0 * M * 3.0 * C * C + 64.0 * M * C, 4.0 * 2.0 * M * C);    // qkv + (q k^T, p v) ; x in, attention out
    const float scale = (float)(1.0 / sqrt((double)hd));
    cudaError_t e = cudaErrorInvalidValue;
    AWindow al{x, ld, g};
    switch (hd) {
#define X(n)                                                                                                        \
    case n: {                                                                                                       \
        using EP = EpiAttn<n, (n == 6 ? 6 : (n + 3) & ~3)>;                                                         \
        EP ep{att, ldo, w.qkvh.bias, w.relbias, heads, scale, masked ? 1 : 0, g.nW, g.nWw, g.Hp, g.Wp, g.dW, g.dWw};             \
        if ((L.ln_post & 1) && w.qkvh_p.tc.img) {                                                                         \
            ep.bias = w.qkvh_p.bias;                                                                                \
            e = tc::launch<false, AWindow, EP, 0, true>(L.st, al, lnpost(L, w.qkvh_p), w.qkvh_p, M, ep);        \
        } else                                                                                                      \
            e = tc::launch<true, AWindow, EP>(L.st, al, lnp(L, w.n1), w.qkvh, M, ep, stats_ready);                  \
    } break;
        ESCB_FUSED_HDS(X)
#undef X
        default: break;
    }
    if (!stats_ready) ++L.launches;          // the LayerNorm statistics pre-kernel
    L.note(e);
}

void op_proj(Launcher& L, const BlockW& w, const float* att, int lda, const float* resid, float* y, int ld,
             const WindowGeom& g, long long M) {
    L.begin(OP_PROJ, 2.0 * M * w.proj.N * w.proj.K, 4.0 * 3.0 * M * w.proj.K);
    ARows al{att, lda};
    EpiWindow ep{y, resid, w.proj.bias, ld, g};
    if (L.tc) L.note(tc::launch<false, ARows, EpiWindow, kProjWide>(L.st, al, noln(L), w.proj, M, ep));
    else L.note(GemmLauncher<false, ARows, EpiWindow, 3, 5, 6, 8, 9>::launch(L.st, al, noln(L), w.proj, M, ep));
}

void op_mlp1(Launcher& L, const BlockW& w, const float* x, int ld, long long M, float* hid, int ldh) {
    L.begin(OP_MLP1, 2.0 * M * w.fc1.N * w.fc1.K, 4.0 * M * (w.fc1.K + w.fc1.N));
    ARows al{x, ld};
    EpiRows<true, false> ep{hid, w.fc1.bias, nullptr, ldh, 0};
    if (L.tc && (L.ln_post & 2) && w.fc1_p.tc.img) {
        ep.bias = w.fc1_p.bias;
        ++L.launches, L.note(tc::launch<false, ARows, EpiRows<true, false>, kMlp1Wide, true>(L.st, al, lnpost(L, w.fc1_p), w.fc1_p, M, ep));
    } else if (L.tc) ++L.launches, L.note(tc::launch<true, ARows, EpiRows<true, false>, kMlp1Wide>(L.st, al, lnp(L, w.n2), w.fc1, M, ep));
    else L.note(GemmLauncher<true, ARows, EpiRows<true, false>, 6, 8, 9>::launch(L.st, al, lnp(L, w.n2), w.fc1, M, ep));
}

void op_mlp2(Launcher& L, const BlockW& w, const float* hid, int ldh, long long M, float* x, int ld) {
    L.begin(OP_MLP2, 2.0 * M * w.fc2.N * w.fc2.K, 4.0 * M * (w.fc2.K + 2.0 * w.fc2.N));
    ARows al{hid, ldh};
    EpiRows<false, true> ep{x, w.fc2.bias, x, ld, ld};
    if (L.tc) L.note(tc::launch<false, ARows, EpiRows<false, true>, kMlp2Wide>(L.st, al, noln(L), w.fc2, M, ep));
    else L.note(GemmLauncher<false, ARows, EpiRows<false, true>, 3, 5, 6, 8, 9>::launch(L.st, al, noln(L), w.fc2, M, ep));
}

// LN2 -> fc1 -> GELU -> fc2 -> +x in one launch (mlp_fused.cuh); x is updated in place
void op_mlp_fused(Launcher& L, const BlockW& w, float* x, int ld, long long M, const mf::StatsOut& so) {
    const double C = w.fc1.K, Hd = w.fc1.N;
    L.begin(OP_MLP_FUSED, 4.0 * M * C * Hd, 4.0 * 2.0 * M * C);          // both GEMMs; x in, x out
    cudaError_t e = ld == w.mlpf.plan.ld ? mf::launch(L.st, w.mlpf, x, M, kLnEps, so, L.next_trace()) : cudaErrorInvalidValue;
    L.note(e);
}

void op_merge(Launcher& L, const LayerW& w, const float* x, int ld, int B, int H, int W, float* y, int ldy) {
    AMerge al{x, ld, H, W, w.C};
    EpiRows<false, false> ep{y, nullptr, nullptr, ldy, 0};
    const long long M = (long long)B * (H / 2) * W;
    L.begin(OP_MERGE, 2.0 * M * w.sub.N * w.sub.K, 4.0 * M * (w.sub.K + w.sub.N));
    if (L.tc && (L.ln_post & 8) && w.sub_p.tc.img)
        ++L.launches, L.note(tc::launch<false, AMerge, EpiRows<false, false>, 0, true>(L.st, al, lnpost(L, w.sub_p), w.sub_p, M, ep));
    else if (L.tc) ++L.launches, L.note(tc::launch<true, AMerge, EpiRows<false, false>>(L.st, al, lnp(L, w.sn), w.sub, M, ep));
    else L.note(GemmLauncher<true, AMerge, EpiRows<false, false>, 5, 6, 8, 9>::launch(L.st, al, lnp(L, w.sn), w.sub, M, ep));
}

void op_split(Launcher& L, const LayerW& w, const float* x, int ld, int B, int H, int W, float* y, int ldy, bool stats_ready) {
    ARows al{x, ld};
    EpiSplit ep{y, ldy, H, W, w.out_dim};
    const long long M = (long long)B * H * W;
    L.begin(OP_SPLIT, 2.0 * M * w.sub.N * w.sub.K, 4.0 * M * (w.sub.K + w.sub.N));
    if (L.tc && !stats_ready) ++L.launches;       // the LayerNorm statistics pre-kernel
    if (L.tc && (L.ln_post & 4) && w.sub_p.tc.img)
        L.note(tc::launch<false, ARows, EpiSplit, kSplitWide, true>(L.st, al, lnpost(L, w.sub_p), w.sub_p, M, ep, stats_ready));
    else if (L.tc) L.note(tc::launch<true, ARows, EpiSplit, kSplitWide>(L.st, al, lnp(L, w.sn), w.sub, M, ep, stats_ready));
    else L.note(GemmLauncher<true, ARows, EpiSplit, 6, 8, 9>::launch(L.st, al, lnp(L, w.sn), w.sub, M, ep));
}

// ------------------------------------------------------------------------------------------------ attention
template <int HD, int HDP>
static cudaError_t launch_attn(cudaStream_t st, const float* qkv, int ldq, float* att, int ldo, const float* relbias,
                               int heads, int C, long long nwin, bool masked, const WindowGeom& g) {
    static int target = -1;            // threads per block to aim for (ESCB_ATTN_THREADS, default 128: more, smaller blocks overlap their load and compute phases)
    if (target < 0) { const char* e = getenv("ESCB_ATTN_THREADS"); target = e ? atoi(e) : 128; }
    int wpb = target / (16 * heads);
    if (wpb < 1) wpb = 1;
    while ((wpb * 16 * heads) % 32) ++wpb;
    const int threads = wpb * 16 * heads;
    if (threads > 1024) return cudaErrorInvalidConfiguration;
    const size_t smem = (size_t)wpb * 16 * std::max(ldq, ldo + 4) * sizeof(float);
    const float scale = (float)(1.0 / sqrt((double)HD));
    const long long blocks = (nwin + wpb - 1) / wpb;
    window_attn_kernel<HD, HDP><<<(unsigned)blocks, threads, smem, st>>>(qkv, ldq, att, ldo, relbias, heads, C, nwin, wpb,
                                                                        scale, masked ? 1 : 0, g.nW, g.nWw, g.Hp, g.Wp);
    return cudaGetLastError();
}

#define ESCB_ATTN_HDS(X) X(4) X(6) X(8) X(12) X(15) X(16) X(24) X(32)

void op_attention(Launcher& L, const float* qkv, int ldq, float* att, int ldo, const float* relbias, int heads,
                  int hd, int hdp, int C, long long nwin, bool masked, const WindowGeom& g) {
    L.begin(OP_ATTN, 1024.0 * nwin * C, 4.0 * 16.0 * nwin * 4.0 * C);
    cudaError_t e = cudaErrorInvalidValue;
    if (hdp == head_pad(hd)) switch (hd) {
#define X(n) case n: e = launch_attn<n, (n == 6 ? 6 : (n + 3) & ~3)>(L.st, qkv, ldq, att, ldo, relbias, heads, C, nwin, masked, g); break;
        ESCB_ATTN_HDS(X)
#undef X
        default: break;
    }
    L.note(e);
}

bool attention_supported(int hd) {
    switch (hd) {
#define X(n) case n: return true;
        ESCB_ATTN_HDS(X)
#undef X
        default: return false;
    }
}

cudaError_t swin_init() {
    cudaError_t e = cudaSuccess;
#define X(n) if (e == cudaSuccess) e = cudaFuncSetAttribute(window_attn_kernel<n, (n == 6 ? 6 : (n + 3) & ~3)>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    ESCB_ATTN_HDS(X)
#undef X
    return e;
}

}  // namespace escb
