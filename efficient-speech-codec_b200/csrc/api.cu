// C ABI of libescb200 (include/escb200.h): handle, checkpoint tensors, weight packing and the
// encode / decode / forward drivers that sequence the kernels of swin.cu / pvq.cu / frontend.cu.
//
// Reference control flow restated here (paths under the reference root):
//   ESC.encode / decode / forward ............ esc/models/codecs.py:30-94
//   Encoder.forward .......................... esc/models/base.py:143-158
//   CrossScaleRVQDecoder.encode/decode/forward esc/models/csrvq.py:97-182
//   TransformerLayer.forward ................. esc/modules/transformer/attention.py:48-91
#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

#include "internal.h"

namespace escb {
bool attention_supported(int hd);
bool argmin_supported(int d);

static thread_local std::string g_err;

static int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

struct Weight {
    std::string name;
    std::vector<int64_t> shape;
    int64_t numel = 1;
    std::vector<float> host;
    bool set = false;
};

struct Level { int C, H; };   // H = frequency patches at this scale

}  // namespace escb

using namespace escb;

struct escb_handle {
    escb_config cfg;
    int device = 0;
    int L = 0;                 // levels == streams
    int F = 0, n_fft = 0, win = 0, hop = 0, pf = 0, pt = 0, C0 = 0, nov = 0;
    Level lev[ESCB_MAX_LEVELS];
    std::vector<Weight> weights;
    std::unordered_map<std::string, int> index;
    bool finalized = false;
    float* arena = nullptr;
    LayerW layers[2 * ESCB_MAX_LEVELS];
    QuantW quants[ESCB_MAX_LEVELS];
    RvqW rvq;                  // RVQCodecs only (cfg.num_rvqs > 0)
    FrontW front;
    std::atomic<long long> launches{0};
    bool use_tc = true;        // ESCB_GEMM=simt selects the fp32 SIMT engine for the dense layers (A/B debugging)
    bool pvq_tc = true;        // ESCB_PVQ=simt keeps the product-VQ projections on the SIMT engine
    int ln_post = kLnPostDefault;   // ESCB_LN_POST bit mask (internal.h)
    bool fuse_pvq = true;      // ESCB_FUSE_PVQ=0: three launches per RVQ stream step (down GEMM, argmin, up GEMM)
    bool emit_stats = true;    // ESCB_EMIT_STATS=0: every LayerNorm GEMM runs its own statistics pre-kernel (A/B debugging)
    int fuse_mlp_max_c = 1 << 30;   // ESCB_FUSE_MLP_MAXC: widest level that takes the fused MLP kernel (where a plan exists)
    bool fuse_mlp = true;      // ESCB_FUSE_MLP=0 keeps the unfused mlp1 + mlp2 pair everywhere (A/B debugging, variant tests)
    int fuse_attn_max_c = 1 << 20;   // ESCB_FUSE_ATTN_MAXC: widest layer whose qkv GEMM runs the attention core in its epilogue (0: never)
    Profiler* prof = nullptr;  // escb_profile_begin .. escb_profile_end (debug facility, single caller)
    unsigned long long* trace = nullptr;   // ESCB_TC_TRACE builds only
    int* err_host = nullptr;   // host-mapped latch written by kernels that meet an out-of-range code index (ACodes)
    int* err_dev = nullptr;    // the same word as the device sees it
    // grow-only scratch for the *_host entry points
    std::mutex host_mu;
    void* host_scratch = nullptr;
    size_t host_scratch_bytes = 0;
};

namespace escb {

static inline int round_up(int v, int m) { return (v + m - 1) / m * m; }
static inline size_t align_up(size_t v, size_t m) { return (v + m - 1) / m * m; }

// ------------------------------------------------------------------------------------------------ manifest
static void add_weight(escb_handle* h, const std::string& name, std::vector<int64_t> shape) {
    Weight w;
    w.name = name;
    w.shape = shape;
    for (int64_t s : shape) w.numel *= s;
    h->index[name] = (int)h->weights.size();
    h->weights.push_back(std::move(w));
}

static void add_swin_layer(escb_handle* h, const std::string& p, int C, int heads, int depth, int scale, int out_dim) {
    const int hidden = C * h->cfg.mlp_hidden_mult;
    for (int j = 0; j < depth; ++j) {
        const std::string b = p + ".swint_blocks." + std::to_string(j);
        add_weight(h, b + ".norm1.weight", {C});
        add_weight(h, b + ".norm1.bias", {C});
        add_weight(h, b + ".attn.relative_position_bias_table", {49, heads});
        add_weight(h, b + ".attn.qkv.weight", {3 * C, C});
        add_weight(h, b + ".attn.qkv.bias", {3 * C});
        add_weight(h, b + ".attn.proj.weight", {C, C});
        add_weight(h, b + ".attn.proj.bias", {C});
        add_weight(h, b + ".norm2.weight", {C});
        add_weight(h, b + ".norm2.bias", {C});
        add_weight(h, b + ".mlp.linear_1.weight", {hidden, C});
        add_weight(h, b + ".mlp.linear_1.bias", {hidden});
        add_weight(h, b + ".mlp.linear_2.weight", {C, hidden});
        add_weight(h, b + ".mlp.linear_2.bias", {C});
    }
    if (scale == 1) {
        add_weight(h, p + ".subsample.norm.weight", {2 * C});
        add_weight(h, p + ".subsample.norm.bias", {2 * C});
        add_weight(h, p + ".subsample.down.weight", {out_dim, 2 * C});
    } else if (scale == 2) {
        add_weight(h, p + ".subsample.norm.weight", {C});
        add_weight(h, p + ".subsample.norm.bias", {C});
        add_weight(h, p + ".subsample.up.weight", {2 * out_dim, C});
    }
}

struct LayerDesc { std::string prefix; int C, heads, scale, out_dim, H; };

// layer_index convention of escb_swin_layer: 0 pre_nn, 1..L-1 encoder.blocks, L..2L-2 decoder.blocks, 2L-1 post_nn
static LayerDesc layer_desc(const escb_handle* h, int li) {
    const int L = h->L;
    const escb_config& c = h->cfg;
    if (li == 0) return {"encoder.pre_nn", c.h_dims[0], c.swin_heads[0], 0, c.h_dims[0], h->lev[0].H};
    if (li < L) {
        const int i = li - 1;
        return {"encoder.blocks." + std::to_string(i), c.h_dims[i], c.swin_heads[i], 1, c.h_dims[i + 1], h->lev[i].H};
    }
    if (li < 2 * L - 1) {
        const int i = li - L;                  // decoder block i works at level L-1-i
        const int lv = L - 1 - i;
        return {"decoder.blocks." + std::to_string(i), c.h_dims[lv], c.swin_heads[L - 2 - i], 2, c.h_dims[lv - 1],
                h->lev[lv].H};
    }
    return {"decoder.post_nn", c.h_dims[0], c.swin_heads[0], 0, c.h_dims[0], h->lev[0].H};
}

struct QuantDesc { int in_dim, in_freq, d, level; };
static QuantDesc quant_desc(const escb_handle* h, int q) {
    const int L = h->L;
    const int lv = q == 0 ? L - 1 : L - q;        // base.py:55-68
    return {h->cfg.h_dims[lv], h->lev[lv].H, h->cfg.codebook_dims[q], lv};
}

static void split_dims(int total, int parts, int* out) {   // quantization.py:380-386
    const int base = total / parts;
    for (int i = 0; i < parts; ++i) out[i] = base;
    out[parts - 1] = total - base * (parts - 1);
}

static void build_manifest(escb_handle* h) {
    const escb_config& c = h->cfg;
    if (c.num_rvqs > 0) {
        // RVQCodecs: quantizers = ProductResidualVectorQuantize at the bottleneck (base.py:73-85, quantization.py:139-168, 276-297)
        const QuantDesc d = quant_desc(h, 0);
        int vq[3];
        split_dims(d.in_dim * d.in_freq * c.overlap, 3, vq);
        for (int m = 0; m < 3; ++m) {
            const std::string p = "quantizers.vqs." + std::to_string(m);
            add_weight(h, p + ".proj_down.weight", {d.d, vq[m]});
            add_weight(h, p + ".proj_up.weight", {vq[m], d.d});
            for (int i = 0; i < c.num_rvqs; ++i) add_weight(h, p + ".vqs." + std::to_string(i) + ".embedding.weight", {c.codebook_size, d.d});
        }
    }
    for (int q = 0; q < (c.num_rvqs > 0 ? 0 : h->L); ++q) {
        const QuantDesc d = quant_desc(h, q);
        int vq[3];
        split_dims(d.in_dim * d.in_freq * c.overlap, 3, vq);
        const std::string p = "quantizers." + std::to_string(q);
        for (int g = 0; g < 3; ++g) add_weight(h, p + ".vqs." + std::to_string(g) + ".embedding.weight", {c.codebook_size, d.d});
        for (int g = 0; g < 3; ++g) add_weight(h, p + ".down_projs." + std::to_string(g) + ".weight", {d.d, vq[g]});
        for (int g = 0; g < 3; ++g) add_weight(h, p + ".up_projs." + std::to_string(g) + ".weight", {vq[g], d.d});
    }
    add_weight(h, "encoder.patch_embed.proj.weight", {h->C0, 2, h->pf, h->pt});
    add_weight(h, "encoder.patch_embed.proj.bias", {h->C0});
    add_weight(h, "encoder.patch_embed.norm.weight", {h->C0});
    add_weight(h, "encoder.patch_embed.norm.bias", {h->C0});
    for (int li = 0; li < 2 * h->L; ++li) {
        const LayerDesc d = layer_desc(h, li);
        add_swin_layer(h, d.prefix, d.C, d.heads, c.swin_depth, d.scale, d.out_dim);
    }
    const int n1 = h->C0 * h->pf * h->pt;
    add_weight(h, "decoder.patch_deembed.de_proj1.weight", {n1, h->C0, 5, 5});
    add_weight(h, "decoder.patch_deembed.de_proj1.bias", {n1});
    add_weight(h, "decoder.patch_deembed.de_proj2.weight", {2, h->C0, 3, 3});
    add_weight(h, "decoder.patch_deembed.de_proj2.bias", {2});
}

// ------------------------------------------------------------------------------------------------ packing
struct Arena {
    std::vector<float> data;
    size_t push(const std::vector<float>& v) {
        const size_t off = align_up(data.size(), 32);
        data.resize(off + v.size(), 0.f);
        if (!v.empty()) memcpy(data.data() + off, v.data(), v.size() * sizeof(float));
        return off;
    }
};

struct Fix { const float** slot; size_t off; };

struct Packer {
    escb_handle* h;
    Arena arena;
    std::vector<Fix> fixes;
    const std::vector<float>& w(const std::string& name) const { return h->weights[h->index.at(name)].host; }
    void put(const float** slot, const std::vector<float>& v) { fixes.push_back({slot, arena.push(v)}); }
    void put_ln(LnW& ln, const std::string& p, int C) {
        std::vector<float> g(round_up(C, 4), 0.f), b(round_up(C, 4), 0.f);
        memcpy(g.data(), w(p + ".weight").data(), C * sizeof(float));
        memcpy(b.data(), w(p + ".bias").data(), C * sizeof(float));
        put(&ln.g, g);
        put(&ln.b, b);
    }
    // Wt[k][n] = W[n][k] for a reference nn.Linear weight W [N][K]
    void put_linear(GemmWeight& gw, const std::string& wname, const char* bname, int wide = 0) {
        const Weight& W = h->weights[h->index.at(wname)];
        put_matrix(gw, W.host.data(), (int)W.shape[0], (int)W.shape[1], bname ? &w(bname) : nullptr, wide);
    }
    // W [N][K] row-major (nn.Linear layout), optional bias [N]
    void put_matrix(GemmWeight& gw, const float* W, int N, int K, const std::vector<float>* bias, int wide = 0,
                    const tc::Tiling* forced = nullptr) {
        std::vector<float> t;
        init_gemm(gw, N, K, t);
        for (int n = 0; n < N; ++n)
            for (int k = 0; k < K; ++k) t[(size_t)k * gw.ldw + n] = W[(size_t)n * K + k];
        put(&gw.wt, t);
        put_tc(gw, t, wide, forced);
        gw.bias = nullptr;
        if (bias) put(&gw.bias, *bias);
    }
    // Post-GEMM LayerNorm variant of a [N][K] weight whose input is LayerNorm(gamma, beta): the image holds
    // W'[n][k] = gamma[k] W[n][k]; cs[n] = sum_k W'[n][k] and bw[n] = sum_k beta[k] W[n][k] (accumulated in double) feed
    // the epilogue's rstd * (acc - mean * cs) + bw (tc_gemm.cuh LNP).
    void put_ln_post(GemmWeight& gw, const float* W, int N, int K, const std::vector<float>* bias, const std::string& ln,
                     int wide, const tc::Tiling* forced = nullptr) {
        const std::vector<float>& g = w(ln + ".weight");
        const std::vector<float>& b = w(ln + ".bias");
        std::vector<float> Wg((size_t)N * K), cs(round_up(N, 4), 0.f), bw(round_up(N, 4), 0.f);
        for (int n = 0; n < N; ++n) {
            double s = 0.0, t = 0.0;
            for (int k = 0; k < K; ++k) {
                const float v = W[(size_t)n * K + k] * g[k];
                Wg[(size_t)n * K + k] = v;
                s += (double)v;
                t += (double)b[k] * (double)W[(size_t)n * K + k];
            }
            cs[n] = (float)s;
            bw[n] = (float)t;
        }
        put_matrix(gw, Wg.data(), N, K, bias, wide, forced);
        put(&gw.cs, cs);
        put(&gw.bw, bw);
    }
    // qkv projection with its output columns laid out [3][heads][hdp] (internal.h head_pad): rows of the reference
    // weight (3C, C) are moved to n' = (part*heads + h)*hdp + d, padding rows and their bias are zero
    void put_qkv(GemmWeight& gw, const std::string& wname, const std::string& bname, int C, int heads) {
        const std::vector<float>& W = w(wname);
        const std::vector<float>& B = w(bname);
        const int hd = C / heads, hdp = head_pad(hd), Np = 3 * heads * hdp;
        std::vector<float> Wp((size_t)Np * C, 0.f), Bp((size_t)Np, 0.f);
        for (int part = 0; part < 3; ++part)
            for (int hh = 0; hh < heads; ++hh)
                for (int d = 0; d < hd; ++d) {
                    const int n = part * C + hh * hd + d, np = (part * heads + hh) * hdp + d;
                    memcpy(&Wp[(size_t)np * C], &W[(size_t)n * C], (size_t)C * sizeof(float));
                    Bp[np] = B[n];
                }
        put_matrix(gw, Wp.data(), Np, C, &Bp);
    }
    // the same projection for the fused attention epilogue (tc_gemm.cuh): columns [head slot][q | k | v][hdp] in
    // sub-tiles of 144 = whole heads, zero rows / bias in the padding slots.  Returns false when the head width
    // does not divide the sub-tile (the layer then keeps the unfused qkv + window_attn_kernel pair).
    bool put_qkv_heads(GemmWeight& gw, const std::string& wname, const std::string& bname, int C, int heads,
                       GemmWeight* gp = nullptr, const std::string& lnname = std::string()) {
        const int hd = C / heads, hdp = head_pad(hd);
        gw = GemmWeight{};
        if (gp) *gp = GemmWeight{};
        if (kAttnBN % (3 * hdp) != 0 || !attention_fusable(hd)) return false;
        const int hpb = kAttnBN / (3 * hdp);
        int nsubs = 0;
        const tc::Tiling tl = tc::attn_tiling((heads + hpb - 1) / hpb, C, &nsubs);
        const std::vector<float>& W = w(wname);
        const std::vector<float>& B = w(bname);
        const int Np = nsubs * kAttnBN;
        std::vector<float> Wp((size_t)Np * C, 0.f), Bp((size_t)Np, 0.f);
        for (int hh = 0; hh < heads; ++hh)
            for (int part = 0; part < 3; ++part)
                for (int d = 0; d < hd; ++d) {
                    const int n = part * C + hh * hd + d, np = (hh * 3 + part) * hdp + d;
                    memcpy(&Wp[(size_t)np * C], &W[(size_t)n * C], (size_t)C * sizeof(float));
                    Bp[np] = B[n];
                }
        put_matrix(gw, Wp.data(), Np, C, &Bp, tc::role_code(tc::kAttnE), &tl);
        if (gp) put_ln_post(*gp, Wp.data(), Np, C, &Bp, lnname, tc::role_code(tc::kAttnE), &tl);
        return true;
    }
    static float tf32_rna(float x) {           // cvt.rna.tf32.f32: nearest, ties away, 10 mantissa bits kept
        uint32_t u;
        memcpy(&u, &x, 4);
        if ((u & 0x7f800000u) != 0x7f800000u) u += 0x1000u;
        u &= 0xffffe000u;
        memcpy(&x, &u, 4);
        return x;
    }
    // tcgen05 operand images from the packed Wt [Kpad][ldw] (see TcWeight in gemm.cuh)
    void put_tc(GemmWeight& gw, const std::vector<float>& t, int wide = 0, const tc::Tiling* forced = nullptr, int max_accs = 4) {
        const tc::Tiling tl = forced ? *forced : tc::choose_tiling(gw.N, gw.K, wide, max_accs);
        put_tc_image(gw, gw.tc, t, wide, tl);
        // Second tiling with one sub-tile per output tile (2-3x the tiles): picked at launch when the default one would
        // leave most SMs idle in its last round (small row counts: the C = 384 / 192 levels at 36 clips), tc::pick.
        gw.tc_alt = TcWeight{nullptr, gw.N, gw.K, 0, 0, 0, 0, 0, 0, 1, 0};
        if (forced) return;
        tc::Tiling alt = tl;
        if (tl.nsub > 1) { alt.nsub = 1; alt.ntn = tl.ntn * tl.nsub; }
        else if (tl.ntn == 1 && tl.BN >= 160) { alt.BN = ((gw.N + 1) / 2 + 15) / 16 * 16; alt.ntn = 2; }
        else return;
        const long long stage = (long long)alt.BN * 256;
        alt.resident = (stage * alt.nkb <= tc::b_budget(wide) && alt.nkb <= tc::MAX_NB) ? 1 : 0;
        put_tc_image(gw, gw.tc_alt, t, wide, alt);
    }
    void put_tc_image(const GemmWeight& gw, TcWeight& w, const std::vector<float>& t, int wide, const tc::Tiling& tl) {
        w.N = gw.N;
        w.K = gw.K;
        w.wide = wide;
        w.ntn = tl.ntn;
        w.nsub = tl.nsub;
        w.BN = tl.BN;
        w.nkb = tl.nkb;
        w.resident = tl.resident;
        w.nmain = tl.nmain > 0 ? tl.nmain : 1;
        w.corr = tl.corr;
        const size_t img = (size_t)w.BN * 32;                 // floats per image
        const int nst = w.ntn * w.nsub;                       // sub-tiles; stage order (tile, kb, sub)
        std::vector<float> out((size_t)nst * w.nkb * 2 * img, 0.f);
        for (int st = 0; st < nst; ++st)
            for (int kb = 0; kb < w.nkb; ++kb) {
                const int nt = st / w.nsub, sub = st % w.nsub;
                float* hi = out.data() + (((size_t)nt * w.nkb + kb) * w.nsub + sub) * 2 * img;
                float* lo = hi + img;
                for (int r = 0; r < w.BN; ++r) {
                    const int n = st * w.BN + r;
                    if (n >= gw.N) continue;
                    for (int kk = 0; kk < 32; ++kk) {
                        const int k = kb * 32 + kk;
                        if (k >= gw.K) break;
                        const float v = t[(size_t)k * gw.ldw + n];
                        const float h = tf32_rna(v);
                        const size_t pos = (size_t)r * 32 + (size_t)(((kk >> 2) ^ (r & 7)) << 2) + (kk & 3);
                        hi[pos] = h;
                        lo[pos] = tf32_rna(v - h);
                    }
                }
            }
        put(&w.img, out);
    }
    // Operands of the fused MLP kernel (mlp_fused.cuh): per 64-column hidden chunk j the fc1 stages (K blocks of the
    // input channels, rows = hidden units 64j..64j+63) followed by the two fc2 stages (K blocks = hidden units
    // 64j + 32kb.., rows = output channels), each stage [hi image | lo image] in the 128-byte-swizzled K-major layout.
    void put_mlp_fused(mf::Weights& mw, const std::string& b, int C) {
        const Weight& W1 = h->weights[h->index.at(b + ".mlp.linear_1.weight")];     // (4C, C)
        const Weight& W2 = h->weights[h->index.at(b + ".mlp.linear_2.weight")];     // (C, 4C)
        const int hidden = (int)W1.shape[0];
        mw = mf::Weights{};
        const mf::Plan pl = mf::make_plan(C, hidden);
        mw.plan = pl;
        if (!pl.ok) return;
        std::vector<float> img(pl.img_floats, 0.f);
        auto put_el = [&](float* hi, float* lo, int r, int kk, float v) {
            const float hv = tf32_rna(v);
            const size_t pos = (size_t)r * 32 + (size_t)(((kk >> 2) ^ (r & 7)) << 2) + (kk & 3);
            hi[pos] = hv;
            lo[pos] = tf32_rna(v - hv);
        };
        for (int j = 0; j < pl.nch; ++j) {
            float* chunk = img.data() + (size_t)j * pl.chunk_bytes / 4;
            const bool packed = mf::fc1_tail_packed(C);
            for (int kb = 0; kb < pl.nkb1; ++kb) {
                float* hi = chunk + (size_t)kb * mf::ST1_BYTES / 4;
                float* lo = hi + mf::HC * 32;
                const bool tail = packed && kb + 1 == pl.nkb1;      // one image: [hi(16 k) | lo(16 k)] per row
                for (int r = 0; r < mf::HC; ++r) {
                    const int n = j * mf::HC + r;
                    if (n >= hidden) break;
                    for (int kk = 0; kk < (tail ? 16 : 32); ++kk) {
                        const int k = kb * 32 + kk;
                        if (k >= C) break;
                        const float v = W1.host[(size_t)n * C + k];
                        if (!tail) { put_el(hi, lo, r, kk, v); continue; }
                        const float hv = tf32_rna(v);
                        auto pos = [&](int col) { return (size_t)r * 32 + (size_t)(((col >> 2) ^ (r & 7)) << 2) + (col & 3); };
                        hi[pos(kk)] = hv;
                        hi[pos(16 + kk)] = tf32_rna(v - hv);
                    }
                }
            }
            const size_t fc1_floats = ((size_t)(pl.nkb1 - 1) * mf::ST1_BYTES + (packed ? mf::ST1T_BYTES : mf::ST1_BYTES)) / 4;
            for (int kb = 0; kb < 2; ++kb) {
                float* hi = chunk + fc1_floats + (size_t)kb * pl.st2_bytes / 4;
                float* lo = hi + (size_t)pl.N2 * 32;
                for (int r = 0; r < C; ++r)
                    for (int kk = 0; kk < 32; ++kk) {
                        const int k = j * mf::HC + kb * 32 + kk;
                        if (k >= hidden) break;
                        put_el(hi, lo, r, kk, W2.host[(size_t)r * hidden + k]);
                    }
            }
        }
        std::vector<float> b1((size_t)pl.nch * mf::HC, 0.f), b2((size_t)pl.N2, 0.f), g((size_t)pl.Kp16, 0.f), be((size_t)pl.Kp16, 0.f);
        memcpy(b1.data(), w(b + ".mlp.linear_1.bias").data(), (size_t)hidden * sizeof(float));
        memcpy(b2.data(), w(b + ".mlp.linear_2.bias").data(), (size_t)C * sizeof(float));
        memcpy(g.data(), w(b + ".norm2.weight").data(), (size_t)C * sizeof(float));
        memcpy(be.data(), w(b + ".norm2.bias").data(), (size_t)C * sizeof(float));
        put(&mw.img, img);
        put(&mw.b1, b1);
        put(&mw.b2, b2);
        put(&mw.gamma, g);
        put(&mw.beta, be);
    }
    static void init_gemm(GemmWeight& gw, int N, int K, std::vector<float>& t) {
        gw.N = N;
        gw.K = K;
        gw.Kpad = round_up(K, kBK);
        gw.ldw = round_up(N, 4);
        gw.bias = nullptr;
        gw.tc = TcWeight{nullptr, N, K, 0, 0, 0, 0, 0, 0, 1, 0};
        t.assign((size_t)gw.Kpad * gw.ldw, 0.f);
    }
};

static void pack_layer(Packer& P, int li) {
    escb_handle* h = P.h;
    const LayerDesc d = layer_desc(h, li);
    LayerW& lw = h->layers[li];
    lw.C = d.C;
    lw.heads = d.heads;
    lw.hd = d.C / d.heads;
    lw.hdp = head_pad(lw.hd);
    lw.depth = h->cfg.swin_depth;
    lw.scale = d.scale;
    lw.out_dim = d.out_dim;
    for (int j = 0; j < lw.depth; ++j) {
        const std::string b = d.prefix + ".swint_blocks." + std::to_string(j);
        BlockW& bw = lw.blk[j];
        P.put_ln(bw.n1, b + ".norm1", d.C);
        P.put_ln(bw.n2, b + ".norm2", d.C);
        P.put_qkv(bw.qkv, b + ".attn.qkv.weight", b + ".attn.qkv.bias", d.C, d.heads);
        P.put_qkv_heads(bw.qkvh, b + ".attn.qkv.weight", b + ".attn.qkv.bias", d.C, d.heads, &bw.qkvh_p, b + ".norm1");
        P.put_linear(bw.proj, b + ".attn.proj.weight", (b + ".attn.proj.bias").c_str(), kProjWide);
        P.put_linear(bw.fc1, b + ".mlp.linear_1.weight", (b + ".mlp.linear_1.bias").c_str(), kMlp1Wide);
        {
            const Weight& W1 = h->weights[h->index.at(b + ".mlp.linear_1.weight")];
            P.put_ln_post(bw.fc1_p, W1.host.data(), (int)W1.shape[0], (int)W1.shape[1], &P.w(b + ".mlp.linear_1.bias"),
                          b + ".norm2", kMlp1Wide);
        }
        P.put_linear(bw.fc2, b + ".mlp.linear_2.weight", (b + ".mlp.linear_2.bias").c_str(), kMlp2Wide);
        P.put_mlp_fused(bw.mlpf, b, d.C);
        // relative-position bias gathered to [heads][16][16] (attention.py:190-205, 229-232)
        const std::vector<float>& table = P.w(b + ".attn.relative_position_bias_table");
        std::vector<float> rb((size_t)d.heads * 256);
        for (int hh = 0; hh < d.heads; ++hh)
            for (int i = 0; i < 16; ++i)
                for (int j2 = 0; j2 < 16; ++j2) {
                    const int idx = (i / 4 - j2 / 4 + 3) * 7 + (i % 4 - j2 % 4 + 3);
                    rb[((size_t)hh * 16 + i) * 16 + j2] = table[(size_t)idx * d.heads + hh];
                }
        P.put(&bw.relbias, rb);
    }
    if (d.scale == 1) {
        P.put_ln(lw.sn, d.prefix + ".subsample.norm", 2 * d.C);
        P.put_linear(lw.sub, d.prefix + ".subsample.down.weight", nullptr);
        {
            const Weight& Wd = h->weights[h->index.at(d.prefix + ".subsample.down.weight")];
            P.put_ln_post(lw.sub_p, Wd.host.data(), (int)Wd.shape[0], (int)Wd.shape[1], nullptr, d.prefix + ".subsample.norm", 0);
        }
    } else if (d.scale == 2) {
        P.put_ln(lw.sn, d.prefix + ".subsample.norm", d.C);
        P.put_linear(lw.sub, d.prefix + ".subsample.up.weight", nullptr, kSplitWide);
        {
            const Weight& Wu = h->weights[h->index.at(d.prefix + ".subsample.up.weight")];
            P.put_ln_post(lw.sub_p, Wu.host.data(), (int)Wu.shape[0], (int)Wu.shape[1], nullptr, d.prefix + ".subsample.norm",
                          kSplitWide);
        }
    }
}

// normalised / transposed / raw forms of one codebook (codebook.py:32-40), appended to the three vectors
static void pack_codebook(const std::vector<float>& e, int K, int dd, bool l2norm, std::vector<float>& raw,
                          std::vector<float>& cbt, std::vector<float>& cn) {
    const size_t r0 = raw.size(), t0 = cbt.size(), n0 = cn.size();
    raw.resize(r0 + (size_t)K * dd);
    cbt.resize(t0 + (size_t)K * dd);
    cn.resize(n0 + (size_t)K);
    for (int c = 0; c < K; ++c) {
        float ss = 0.f;
        for (int j = 0; j < dd; ++j) { const float v = e[(size_t)c * dd + j]; ss = fmaf(v, v, ss); }
        const float den = l2norm ? fmaxf(sqrtf(ss), 1e-12f) : 1.0f;     // l2norm=False: the table itself (codebook.py:31-40)
        float s2 = 0.f;
        for (int j = 0; j < dd; ++j) {
            const float v = e[(size_t)c * dd + j];
            const float nv = v / den;
            raw[r0 + (size_t)c * dd + j] = v;
            cbt[t0 + (size_t)j * K + c] = nv;
            s2 = fmaf(nv, nv, s2);
        }
        cn[n0 + c] = s2;
    }
}

// The projections of one product quantizer (ESC's ProductVectorQuantize and RVQCodecs' ProductResidualVectorQuantize share
// the frame geometry); `down_name(g)` / `up_name(g)` give the checkpoint keys of group g's weights.
template <class FD, class FU>
static void pack_projections(Packer& P, QuantW& qw, int C, int Hq, int dd, int K, FD down_name, FU up_name) {
    const int frame = 2 * C * Hq;
    int vq[3], start[3];
    split_dims(frame, 3, vq);
    start[0] = 0; start[1] = vq[0]; start[2] = vq[0] + vq[1];
    qw.in_dim = C; qw.in_freq = Hq; qw.d = dd; qw.frame_dim = frame; qw.ncodes = K; qw.l2norm = P.h->cfg.l2norm ? 1 : 0;
    std::vector<float> down, up;
    Packer::init_gemm(qw.down, 3 * dd, frame, down);
    Packer::init_gemm(qw.up, frame, 3 * dd, up);
    // reference frame index kref = o*(C*Hq) + c*Hq + h (quantization.py:400-409); ours k' = h*(2C) + o*C + c
    for (int hh = 0; hh < Hq; ++hh)
        for (int o = 0; o < 2; ++o)
            for (int c = 0; c < C; ++c) {
                const int kref = o * (C * Hq) + c * Hq + hh;
                const int kp = hh * 2 * C + o * C + c;
                const int g = kref >= start[2] ? 2 : (kref >= start[1] ? 1 : 0);
                const int kl = kref - start[g];
                const std::vector<float>& dw = P.w(down_name(g));   // [d][vq_g]
                const std::vector<float>& uw = P.w(up_name(g));     // [vq_g][d]
                for (int j = 0; j < dd; ++j) {
                    down[(size_t)kp * qw.down.ldw + g * dd + j] = dw[(size_t)j * vq[g] + kl];
                    up[(size_t)(g * dd + j) * qw.up.ldw + kp] = uw[(size_t)kl * dd + j];
                }
            }
    P.put(&qw.down.wt, down);
    P.put(&qw.up.wt, up);
    P.put_tc(qw.up, up, kPvqUpWide);
    // per-group down-projections: group g owns kref in [start_g, start_g + vq_g), i.e. (with equal thirds that are
    // multiples of Hq) the (o, c) range [g*run, (g+1)*run) of every h run; its K index is kg = h*run + (o*C + c - g*run)
    qw.run = 0;
    if ((2 * C) % 3 == 0 && ((2 * C / 3) % 2) == 0 && vq[0] == vq[1] && vq[1] == vq[2] && vq[0] == (2 * C / 3) * Hq) {
        const int run = 2 * C / 3;
        qw.run = run;
        for (int g = 0; g < 3; ++g) {
            std::vector<float> dg;
            Packer::init_gemm(qw.down_g[g], dd, run * Hq, dg);
            const std::vector<float>& dw = P.w(down_name(g));   // [d][vq_g]
            for (int hh = 0; hh < Hq; ++hh)
                for (int oc = g * run; oc < (g + 1) * run; ++oc) {
                    const int o = oc / C, c = oc - o * C;
                    const int kl = o * (C * Hq) + c * Hq + hh - start[g];
                    const int kg = hh * run + (oc - g * run);
                    for (int j = 0; j < dd; ++j) dg[(size_t)kg * qw.down_g[g].ldw + j] = dw[(size_t)j * vq[g] + kl];
                }
            P.put(&qw.down_g[g].wt, dg);
        }
    }
}

static void pack_quant(Packer& P, int q) {
    escb_handle* h = P.h;
    const QuantDesc d = quant_desc(h, q);
    QuantW& qw = h->quants[q];
    const int K = h->cfg.codebook_size;
    const std::string p = "quantizers." + std::to_string(q);
    pack_projections(P, qw, d.in_dim, d.in_freq, d.d, K,
                     [&](int g) { return p + ".down_projs." + std::to_string(g) + ".weight"; },
                     [&](int g) { return p + ".up_projs." + std::to_string(g) + ".weight"; });
    // codebooks: raw, L2-normalised (F.normalize, eps 1e-12) and squared norms of the normalised rows (codebook.py:32-40)
    std::vector<float> raw, cbt, cn;
    for (int g = 0; g < 3; ++g) pack_codebook(P.w(p + ".vqs." + std::to_string(g) + ".embedding.weight"), K, d.d, h->cfg.l2norm != 0, raw, cbt, cn);
    P.put(&qw.raw, raw);
    P.put(&qw.cbt, cbt);
    P.put(&qw.cnorm, cn);
}

// RVQCodecs: one ProductResidualVectorQuantize at the bottleneck, codebooks ordered [group][stream]
static void pack_rvq(Packer& P) {
    escb_handle* h = P.h;
    const QuantDesc d = quant_desc(h, 0);
    RvqW& w = h->rvq;
    const int K = h->cfg.codebook_size, S = h->cfg.num_rvqs;
    w.S = S;
    pack_projections(P, w.q, d.in_dim, d.in_freq, d.d, K,
                     [&](int g) { return "quantizers.vqs." + std::to_string(g) + ".proj_down.weight"; },
                     [&](int g) { return "quantizers.vqs." + std::to_string(g) + ".proj_up.weight"; });
    w.q.raw = w.q.cbt = w.q.cnorm = nullptr;
    std::vector<float> raw, cbt, cn;
    for (int g = 0; g < 3; ++g)
        for (int i = 0; i < S; ++i)
            pack_codebook(P.w("quantizers.vqs." + std::to_string(g) + ".vqs." + std::to_string(i) + ".embedding.weight"), K, d.d, h->cfg.l2norm != 0, raw, cbt, cn);
    P.put(&w.raw, raw);
    P.put(&w.cbt, cbt);
    P.put(&w.cnorm, cn);
}

static void pack_front(Packer& P) {
    escb_handle* h = P.h;
    FrontW& f = h->front;
    const int F = h->F, N = h->n_fft, win = h->win, hop = h->hop, C0 = h->C0;
    const int padl = (N - win) / 2;
    f.F = F; f.win = win; f.hop = hop; f.nov = h->nov; f.C0 = C0; f.pf = h->pf; f.pt = h->pt;
    const double PI = 3.14159265358979323846;
    std::vector<double> wd(win);
    for (int k = 0; k < win; ++k) wd[k] = 0.5 - 0.5 * cos(2.0 * PI * k / win);   // periodic hann (base.py:22-24)
    // forward: Sf[t][f] = sum_k x[t*hop - win/2 + k] w[k] e^{-2 pi i f (k + padl) / N}
    std::vector<float> dft;
    Packer::init_gemm(f.dft, 2 * F, win, dft);
    for (int k = 0; k < win; ++k)
        for (int ff = 0; ff < F; ++ff) {
            const long long ph = ((long long)ff * (k + padl)) % N;
            const double a = 2.0 * PI * (double)ph / N;
            dft[(size_t)k * f.dft.ldw + ff] = (float)(wd[k] * cos(a));
            dft[(size_t)k * f.dft.ldw + F + ff] = (float)(-wd[k] * sin(a));
        }
    P.put(&f.dft.wt, dft);
    // inverse: chunk j of `hop` samples sums frames j-dt, dt < nov; row = dt*2F + cf, col = r; window tap k = dt*hop + r
    std::vector<float> idft;
    Packer::init_gemm(f.idft, hop, h->nov * 2 * F, idft);
    for (int dt = 0; dt < h->nov; ++dt)
        for (int r = 0; r < hop; ++r) {
            const int k = dt * hop + r, n = k + padl;
            for (int ff = 0; ff < F; ++ff) {
                const long long ph = ((long long)ff * n) % N;
                const double a = 2.0 * PI * (double)ph / N;
                const bool edge = (ff == 0) || (2 * ff == N);
                const double cr = (edge ? 1.0 : 2.0) * cos(a) / N;
                const double ci = edge ? 0.0 : -2.0 * sin(a) / N;
                idft[(size_t)(dt * 2 * F + ff) * f.idft.ldw + r] = (float)(wd[k] * cr);
                idft[(size_t)(dt * 2 * F + F + ff) * f.idft.ldw + r] = (float)(wd[k] * ci);
            }
        }
    P.put(&f.idft.wt, idft);
    std::vector<float> wsq(win);
    for (int k = 0; k < win; ++k) { const float wf = (float)wd[k]; wsq[k] = wf * wf; }
    P.put(&f.wsq, wsq);
    // patch embedding (scale.py:38,44): conv weight (C0, 2, pf, pt) is already [C0][k], k = (c*pf + s1)*pt + s2
    P.put(&f.embed_w, P.w("encoder.patch_embed.proj.weight"));
    P.put(&f.embed_b, P.w("encoder.patch_embed.proj.bias"));
    P.put_ln(f.embed_ln, "encoder.patch_embed.norm", C0);
    f.embed_k_ok = 0;
    if (C0 == 45 && h->pf == 3 && h->pt == 2) {
        const std::vector<float>& ew = P.w("encoder.patch_embed.proj.weight");
        const std::vector<float>& eb = P.w("encoder.patch_embed.proj.bias");
        const std::vector<float>& eg = P.w("encoder.patch_embed.norm.weight");
        const std::vector<float>& ebe = P.w("encoder.patch_embed.norm.bias");
        if (ew.size() == 45 * 12 && eb.size() == 45 && eg.size() == 45 && ebe.size() == 45) {
            memcpy(f.embed_k.w, ew.data(), sizeof f.embed_k.w);
            memcpy(f.embed_k.b, eb.data(), sizeof f.embed_k.b);
            memcpy(f.embed_k.g, eg.data(), sizeof f.embed_k.g);
            memcpy(f.embed_k.be, ebe.data(), sizeof f.embed_k.be);
            f.embed_k_ok = 1;
        }
    }
    // de_proj1 (scale.py:66-68): K = tap*ldc(C0) + c
    // Output columns are padded per sub-pixel to the pixel pitch: n' = s*ldc(C0) + c for the reference's n = s*C0 + c
    // (zero weight / bias in the 3 pad channels), so the pixel-shuffle epilogue writes whole aligned float4s and every
    // 32-byte sector of the pixel map is fully written (4-byte scattered stores made L2 fill each sector from DRAM:
    // 1.05 GB of DRAM reads per launch for a 133 MB input).
    const int ld0 = ldc(C0), NS = h->pf * h->pt, N1 = NS * ld0;
    const std::vector<float>& w1 = P.w("decoder.patch_deembed.de_proj1.weight");
    const std::vector<float>& b1 = P.w("decoder.patch_deembed.de_proj1.bias");
    std::vector<float> de1, de1b((size_t)N1, 0.f);
    Packer::init_gemm(f.de1, N1, 25 * ld0, de1);
    for (int sp = 0; sp < NS; ++sp)
        for (int co = 0; co < C0; ++co) {
            const int n = sp * C0 + co, np = sp * ld0 + co;
            de1b[np] = b1[n];
            for (int c = 0; c < C0; ++c)
                for (int tap = 0; tap < 25; ++tap)
                    de1[(size_t)(tap * ld0 + c) * f.de1.ldw + np] = w1[((size_t)n * C0 + c) * 25 + tap];
        }
    P.put(&f.de1.wt, de1);
    P.put_tc(f.de1, de1, 0, nullptr, 1);      // one accumulator: this layer feeds the audio only (tolerance 1e-4), never a code decision
    P.put(&f.de1.bias, de1b);
    // de_proj2 (scale.py:70-71): wp[tap][c][2]
    const std::vector<float>& w2 = P.w("decoder.patch_deembed.de_proj2.weight");
    std::vector<float> de2((size_t)9 * C0 * 2);
    for (int o = 0; o < 2; ++o)
        for (int c = 0; c < C0; ++c)
            for (int tap = 0; tap < 9; ++tap) de2[((size_t)tap * C0 + c) * 2 + o] = w2[((size_t)o * C0 + c) * 9 + tap];
    P.put(&f.de2_w, de2);
    memset(&f.de2_k, 0, sizeof f.de2_k);
    memcpy(f.de2_k.w, de2.data(), de2.size() * sizeof(float));
    P.put(&f.de2_b, P.w("decoder.patch_deembed.de_proj2.bias"));
    f.de2_bias[0] = P.w("decoder.patch_deembed.de_proj2.bias")[0];
    f.de2_bias[1] = P.w("decoder.patch_deembed.de_proj2.bias")[1];
}

// ------------------------------------------------------------------------------------------------ workspace
struct Bump {
    char* base;
    size_t cap, off = 0;
    bool dry;
    Bump(void* p, size_t c) : base((char*)p), cap(c), dry(p == nullptr) {}
    template <class T>
    T* take(size_t n) {
        off = align_up(off, 256);
        T* r = dry ? nullptr : (T*)(base + off);
        off += n * sizeof(T);
        return r;
    }
    bool ok() const { return dry || off <= cap; }
};

// Every buffer one encode / decode / forward call needs for `B` clips of `W` time patches.
struct Work {
    float* Sf = nullptr;                       // [B][T][2F] frame-major spectrum (STFT out / de-embed out)
    float* enc[ESCB_MAX_LEVELS] = {};          // encoder outputs per level, [B*H_l*W][ldc(C_l)]
    float* dec[ESCB_MAX_LEVELS] = {};          // decoder state per level
    float* xw = nullptr;                       // working token map of the running TransformerLayer
    float* post = nullptr;                     // post_nn working map
    float* qkv = nullptr;
    float* att = nullptr;
    float* hid = nullptr;
    float2* stats = nullptr;                   // LayerNorm (mean, rstd) per logical GEMM row
    float* ze = nullptr;                       // projected VQ vectors [B*T][ldc(3d)]
    float* zq = nullptr;                       // RVQCodecs: summed codewords [B*T][ldc(3d)]
    float* se = nullptr;                       // RVQCodecs: eval-loss numerators [B*T][3]
    float* Y1 = nullptr;                       // de-embed pixel map [B][F][2W][ldc(C0)]
    long long* codes = nullptr;                // forward(): internal codes when the caller passes none
    float* dense = nullptr;                    // staging for the unit entry points
    float* stage = nullptr;
};

enum { WK_ENC = 1, WK_DEC = 2, WK_UNIT = 4 };

static size_t plan(const escb_handle* h, int B, int W, int T, int what, Bump& bp, Work& wk) {
    const int L = h->L;
    const int Wp = round_up(W, 4);
    size_t max_tok = 0, max_qkv = 0, max_att = 0, max_hid = 0, max_rows = 0;
    for (int l = 0; l < L; ++l) {
        const int C = h->lev[l].C, H = h->lev[l].H, Hp = round_up(H, 4);
        const size_t M = (size_t)B * H * W, Mw = (size_t)B * Hp * Wp;
        max_tok = std::max(max_tok, M * ldc(C));
        int wq = ldc(3 * C);                  // widest head-padded qkv row among the layers working at this level
        for (int li = 0; li < 2 * L; ++li) {
            const LayerDesc d = layer_desc(h, li);
            if (d.C == C) wq = std::max(wq, 3 * d.heads * head_pad(C / d.heads));
        }
        max_qkv = std::max(max_qkv, Mw * (size_t)wq);
        max_att = std::max(max_att, Mw * ldc(C));
        max_hid = std::max(max_hid, M * (size_t)(C * h->cfg.mlp_hidden_mult));
        max_rows = std::max(max_rows, Mw);
    }
    wk.Sf = bp.take<float>((size_t)B * std::max(T, h->pt * W) * 2 * h->F);
    if (what & WK_ENC)
        for (int l = 0; l < L; ++l) wk.enc[l] = bp.take<float>((size_t)B * h->lev[l].H * W * ldc(h->lev[l].C));
    for (int l = 0; l < L; ++l) wk.dec[l] = bp.take<float>((size_t)B * h->lev[l].H * W * ldc(h->lev[l].C));
    wk.xw = bp.take<float>(max_tok);
    wk.post = bp.take<float>((size_t)B * h->lev[0].H * W * ldc(h->C0));
    wk.qkv = bp.take<float>(max_qkv);
    wk.att = bp.take<float>(max_att);
    wk.hid = bp.take<float>(max_hid);
    wk.stats = bp.take<float2>(max_rows);
    int dmax = 0;
    for (int q = 0; q < L; ++q) dmax = std::max(dmax, h->cfg.codebook_dims[q]);
    wk.ze = bp.take<float>((size_t)B * (W / 2) * ldc(3 * dmax));
    if (h->cfg.num_rvqs > 0) {
        wk.zq = bp.take<float>((size_t)B * (W / 2) * ldc(3 * dmax));
        wk.se = bp.take<float>((size_t)B * (W / 2) * 3);
    }
    if (what & WK_DEC) wk.Y1 = bp.take<float>((size_t)B * h->F * h->pt * W * ldc(h->C0));
    wk.codes = bp.take<long long>((size_t)B * L * 3 * (W / 2));
    if (what & WK_UNIT) {
        wk.dense = bp.take<float>(std::max(max_tok, (size_t)B * std::max(T, h->pt * W) * 2 * h->F));
        wk.stage = bp.take<float>(max_tok);
    }
    return bp.off;
}

// ------------------------------------------------------------------------------------------------ drivers
struct Ctx {
    escb_handle* h;
    Launcher L;
    Work wk;
    int B, W;
};

static WindowGeom geom(int H, int W, int shift) {
    WindowGeom g;
    g.H = H; g.W = W; g.Hp = round_up(H, 4); g.Wp = round_up(W, 4); g.shift = shift;
    g.nWw = g.Wp / 4; g.nW = (g.Hp / 4) * g.nWw;
    g.dW = FastDiv::make((unsigned)g.nW); g.dWw = FastDiv::make((unsigned)g.nWw);
    return g;
}

// TransformerLayer.forward (attention.py:48-91).  Reads x_in (never written), runs the blocks in `xw`, then
// writes the resampled map to `out` (scale != 0) — for scale == 0 the result is left in xw.
static void run_layer(Ctx& c, int li, const float* x_in, float* xw, float* out, int H) {
    const LayerW& lw = c.h->layers[li];
    const int C = lw.C, ld = ldc(C), ldq = 3 * lw.heads * lw.hdp, ldh = C * c.h->cfg.mlp_hidden_mult;
    const int B = c.B, W = c.W;
    const long long M = (long long)B * H * W;
    const float* src = x_in;
    bool have_stats = false;      // wk.stats holds the LayerNorm statistics of `src` for the next consumer (written by the fused MLP)
    for (int j = 0; j < lw.depth; ++j) {
        const WindowGeom g = geom(H, W, (j & 1) ? 2 : 0);
        const long long nwin = (long long)B * g.nW, Mw = nwin * 16;
        const BlockW& bw = lw.blk[j];
        const bool fused_attn = c.L.tc && bw.qkvh.tc.img && C <= c.h->fuse_attn_max_c;
        if (fused_attn)
            op_qkv_attn(c.L, bw, lw.heads, lw.hd, src, ld, g, Mw, c.wk.att, ld, (j & 1) != 0, have_stats);
        else {
            op_qkv(c.L, bw, src, ld, g, Mw, c.wk.qkv, ldq);
            op_attention(c.L, c.wk.qkv, ldq, c.wk.att, ld, bw.relbias, lw.heads, lw.hd, lw.hdp, C, nwin, (j & 1) != 0, g);
        }
        op_proj(c.L, bw, c.wk.att, ld, src, xw, ld, g, Mw);
        have_stats = false;
        if (c.L.fuse_mlp && bw.mlpf.plan.ok && C <= c.h->fuse_mlp_max_c) {
            // the rows this kernel writes are the input of the next LayerNorm: let its epilogue emit their statistics
            mf::StatsOut so{};
            if (c.L.emit_stats) {
                if (j + 1 < lw.depth) {
                    const BlockW& nb = lw.blk[j + 1];
                    if (c.L.tc && nb.qkvh.tc.img && C <= c.h->fuse_attn_max_c && qkv_attn_takes_stats(c.L, nb)) {
                        so.out = c.wk.stats; so.geom = 1; so.H = H; so.W = W;
                        so.ng = geom(H, W, ((j + 1) & 1) ? 2 : 0);
                    }
                } else if (lw.scale == 2) {
                    so.out = c.wk.stats; so.geom = 0; so.H = H; so.W = W;       // PatchSplit normalises token rows
                }
            }
            op_mlp_fused(c.L, bw, xw, ld, M, so);
            have_stats = so.out != nullptr;
        } else {
            op_mlp1(c.L, bw, xw, ld, M, c.wk.hid, ldh);
            op_mlp2(c.L, bw, c.wk.hid, ldh, M, xw, ld);
        }
        src = xw;
    }
    if (lw.scale == 1) op_merge(c.L, lw, xw, ld, B, H, W, out, ldc(lw.out_dim));
    else if (lw.scale == 2) op_split(c.L, lw, xw, ld, B, H, W, out, ldc(lw.out_dim), have_stats);
}

// Encoder.forward (base.py:143-158) from the frame-major spectrum in wk.Sf.
static void run_encoder(Ctx& c, int T) {
    escb_handle* h = c.h;
    const int L = h->L;
    op_patch_embed(c.L, h->front, c.wk.Sf, c.B, T, h->lev[0].H, c.W, c.wk.xw, ldc(h->C0));
    run_layer(c, 0, c.wk.xw, c.wk.enc[0], nullptr, h->lev[0].H);             // pre_nn: result stays in enc[0]
    for (int i = 0; i < L - 1; ++i) run_layer(c, 1 + i, c.wk.enc[i], c.wk.xw, c.wk.enc[i + 1], h->lev[i].H);
}

static void pvq_encode(Ctx& c, int q, const float* enc, const float* dec, long long* codes, int S, int s) {
    const QuantW& qw = c.h->quants[q];
    const int T = c.W / 2, ldz = ldc(3 * qw.d);
    op_pvq_down(c.L, qw, enc, dec, c.B, c.W, c.wk.ze, ldz);
    op_argmin(c.L, qw, 0, 3, c.wk.ze, ldz, (long long)c.B * T, codes + (long long)s * 3 * T, T, (long long)S * 3 * T);
}

// CrossScaleRVQDecoder.encode (csrvq.py:131-158)
static void run_csrvq_encode(Ctx& c, int S, long long* codes) {
    escb_handle* h = c.h;
    const int L = h->L;
    // one stream step: quantize enc - dec, and (unless it is the last transmitted stream) refine dec in place
    auto step = [&](int q, const float* enc, float* dec_in, float* dec_out, bool refine) {
        if (c.L.fuse_pvq && op_pvq_stream(c.L, h->quants[q], enc, dec_in, c.B, c.W, codes, S, q, refine ? dec_out : nullptr, nullptr, 0))
            return;
        pvq_encode(c, q, enc, dec_in, codes, S, q);
        if (refine) op_pvq_up(c.L, h->quants[q], codes, S, q, dec_in, c.B, c.W, dec_out);
    };
    step(0, c.wk.enc[L - 1], nullptr, c.wk.dec[L - 1], S > 1);
    if (S == 1) return;
    for (int i = 0; i < S - 1; ++i) {
        const int lv = L - 1 - i;
        const bool last = i + 2 == S;
        step(i + 1, c.wk.enc[lv], c.wk.dec[lv], c.wk.dec[lv], !last);
        if (last) break;
        run_layer(c, L + i, c.wk.dec[lv], c.wk.xw, c.wk.dec[lv - 1], h->lev[lv].H);
    }
}

// post_nn + PatchDeEmbed + inverse STFT (csrvq.py:181-182, scale.py:73-81, base.py:39-47) from wk.dec[0]
static void run_backend(Ctx& c, float* audio, float* recon_feat) {
    escb_handle* h = c.h;
    const int H0 = h->lev[0].H, T2 = h->pt * c.W;
    run_layer(c, 2 * h->L - 1, c.wk.dec[0], c.wk.post, nullptr, H0);
    op_deembed(c.L, h->front, c.wk.post, ldc(h->C0), c.B, H0, c.W, c.wk.Y1, c.wk.Sf);
    if (recon_feat) op_transpose(c.L, c.wk.Sf, recon_feat, c.B, T2, 2 * h->F);
    if (audio) op_istft(c.L, h->front, c.wk.Sf, c.B, T2, audio);
}

// CrossScaleRVQDecoder.decode (csrvq.py:160-182)
static void run_csrvq_decode(Ctx& c, int S, const long long* codes) {
    escb_handle* h = c.h;
    const int L = h->L;
    auto up = [&](int q, const float* dec, float* out) {
        if (c.L.fuse_pvq && op_pvq_stream(c.L, h->quants[q], nullptr, dec, c.B, c.W, const_cast<long long*>(codes), S, q, out, nullptr, 0))
            return;
        op_pvq_up(c.L, h->quants[q], codes, S, q, dec, c.B, c.W, out);
    };
    up(0, nullptr, c.wk.dec[L - 1]);
    for (int i = 0; i < L - 1; ++i) {
        const int lv = L - 1 - i;
        if (i < S - 1) up(i + 1, c.wk.dec[lv], c.wk.dec[lv]);
        run_layer(c, L + i, c.wk.dec[lv], c.wk.xw, c.wk.dec[lv - 1], h->lev[lv].H);
    }
}

// CrossScaleRVQDecoder.forward in eval mode (csrvq.py:97-129, 23-48): quantize and decode in one sweep.
static void run_csrvq_forward(Ctx& c, int S, long long* codes, float* loss) {
    escb_handle* h = c.h;
    const int L = h->L, T = c.W / 2;
    auto vq = [&](int q, const float* enc, const float* dec, float* out) {
        const QuantW& qw = h->quants[q];
        if (c.L.fuse_pvq && op_pvq_stream(c.L, qw, enc, dec, c.B, c.W, codes, S, q, out, loss ? c.wk.ze : nullptr, ldc(3 * qw.d))) {
            if (loss) op_vq_loss(c.L, qw, c.wk.ze, ldc(3 * qw.d), codes, S, q, c.B, T, loss);
            return;
        }
        pvq_encode(c, q, enc, dec, codes, S, q);
        if (loss) op_vq_loss(c.L, qw, c.wk.ze, ldc(3 * qw.d), codes, S, q, c.B, T, loss);
        op_pvq_up(c.L, qw, codes, S, q, dec, c.B, c.W, out);
    };
    vq(0, c.wk.enc[L - 1], nullptr, c.wk.dec[L - 1]);
    for (int i = 0; i < L - 1; ++i) {
        const int lv = L - 1 - i;
        if (i < S - 1) vq(i + 1, c.wk.enc[lv], c.wk.dec[lv], c.wk.dec[lv]);
        run_layer(c, L + i, c.wk.dec[lv], c.wk.xw, c.wk.dec[lv - 1], h->lev[lv].H);
    }
}

// ---- RVQCodecs (codecs.py:96-181): encoder -> ProductResidualVectorQuantize at the bottleneck -> plain Decoder
static void run_rvq_quantize(Ctx& c, int S, long long* codes, float* zq, float* se) {
    const RvqW& w = c.h->rvq;
    const int T = c.W / 2, ldz = ldc(3 * w.q.d);
    op_pvq_down(c.L, w.q, c.wk.enc[c.h->L - 1], nullptr, c.B, c.W, c.wk.ze, ldz);            // pre_process + proj_down
    op_rvq_chain(c.L, w, c.wk.ze, ldz, (long long)c.B * T, S, codes, T, zq, se);
}

// Decoder.forward (base.py:194-203) from the summed codewords: proj_up + post_process, then the up-sampling layers
static void run_rvq_decoder(Ctx& c, const float* zq) {
    escb_handle* h = c.h;
    const int L = h->L;
    op_rvq_up(c.L, h->rvq, zq, ldc(3 * h->rvq.q.d), c.B, c.W, c.wk.dec[L - 1]);
    for (int i = 0; i < L - 1; ++i) {
        const int lv = L - 1 - i;
        run_layer(c, L + i, c.wk.dec[lv], c.wk.xw, c.wk.dec[lv - 1], h->lev[lv].H);
    }
}

static int max_streams_of(const escb_handle* h) { return h->cfg.num_rvqs > 0 ? h->cfg.num_rvqs : h->L; }

static int check_ready(const escb_handle* h) {
    if (!h) return fail(ESCB_EINVAL, "null handle");
    if (!h->finalized) return fail(ESCB_ESTATE, "escb_finalize() has not been called since the last weight change");
    return ESCB_OK;
}

static int frames_of(const escb_handle* h, int64_t num_samples) { return (int)(1 + num_samples / h->hop); }

static int time_patches(const escb_handle* h, int64_t num_samples, int* W) {
    if (num_samples <= h->n_fft / 2)
        return fail(ESCB_EINVAL, "clip of %lld samples is too short for reflect padding of %d", (long long)num_samples,
                    h->n_fft / 2);
    const int T = frames_of(h, num_samples);
    const int w = (T - h->pt) / h->pt + 1;       // conv stride pt, kernel pt (scale.py:38)
    if (w <= 0 || w % h->cfg.overlap)
        return fail(ESCB_EINVAL, "Time dimension must be multiple of overlap (W=%d, overlap=%d)", w, h->cfg.overlap);
    *W = w;
    return ESCB_OK;
}

static int finish(Ctx& c, const char* what) {
    c.h->launches += c.L.launches;
    if (c.L.err != cudaSuccess) return fail(ESCB_ECUDA, "%s: %s", what, cudaGetErrorString(c.L.err));
    return ESCB_OK;
}

static int begin(escb_handle* h, Ctx& c, int B, int W, int T, int what, void* ws, size_t ws_bytes, void* stream) {
    if (int e = check_ready(h)) return e;
    if (B <= 0) return fail(ESCB_EINVAL, "batch must be positive");
    if (W <= 0 || W % h->cfg.overlap) return fail(ESCB_EINVAL, "Time dimension must be multiple of overlap (W=%d)", W);
    c.h = h;
    c.B = B;
    c.W = W;
    c.L.st = (cudaStream_t)stream;
    c.L.prof = h->prof;
    c.L.ln_post = h->ln_post;
    c.L.tc = h->use_tc;
    c.L.pvq_tc = h->use_tc && h->pvq_tc;
    c.L.fuse_mlp = h->use_tc && h->fuse_mlp;
    c.L.emit_stats = h->emit_stats;
    c.L.fuse_pvq = h->fuse_pvq;
    Bump dry(nullptr, 0);
    Work tmp;
    const size_t need = plan(h, B, W, T, what, dry, tmp);
    if (!ws || ws_bytes < need)
        return fail(ESCB_ENOMEM, "workspace of %zu bytes is too small, %zu needed", ws_bytes, need);
    Bump bp(ws, ws_bytes);
    plan(h, B, W, T, what, bp, c.wk);
    c.L.ln_stats = c.wk.stats;
    c.L.code_err = h->err_dev;
#ifdef ESCB_TC_TRACE
    c.L.trace = h->trace;
    if (h->trace) cudaMemsetAsync(h->trace, 0, 1024 * 16 * 8, c.L.st);
#endif
    return ESCB_OK;
}

}  // namespace escb

// ================================================================================================ C ABI
extern "C" {

int escb_abi_version(void) { return ESCB_ABI_VERSION; }
const char* escb_last_error(void) { return g_err.c_str(); }

int escb_create(const escb_config* cfg, escb_handle** out) {
    if (!cfg || !out) return fail(ESCB_EINVAL, "null argument");
    *out = nullptr;
    const escb_config& c = *cfg;
    if (c.num_levels < 2 || c.num_levels > ESCB_MAX_LEVELS) return fail(ESCB_EINVAL, "num_levels must be in [2, %d]", ESCB_MAX_LEVELS);
    if (c.swin_depth < 1 || c.swin_depth > ESCB_MAX_DEPTH) return fail(ESCB_EINVAL, "swin_depth must be in [1, %d]", ESCB_MAX_DEPTH);
    if (c.window_size != 4) return fail(ESCB_EINVAL, "window_size must be 4");
    if (c.group_size != 3) return fail(ESCB_EINVAL, "group_size must be 3");
    if (c.overlap != 2) return fail(ESCB_EINVAL, "overlap must be 2");
    if (c.patch_freq < 1 || c.patch_time < 1 || 2 * c.patch_freq * c.patch_time > kEmbedMaxK)
        return fail(ESCB_EINVAL, "unsupported patch size (%d, %d)", c.patch_freq, c.patch_time);
    if (c.in_freq < 2 || c.in_freq % c.patch_freq) return fail(ESCB_EINVAL, "in_freq must be a multiple of patch_freq");
    if (c.mlp_hidden_mult < 1) return fail(ESCB_EINVAL, "mlp_hidden_mult must be >= 1");
    if (c.codebook_size < 1) return fail(ESCB_EINVAL, "codebook_size must be positive");
    if (c.num_rvqs < 0 || c.num_rvqs > 64) return fail(ESCB_EINVAL, "num_rvqs must be in [0, 64]");
    const int n_fft = 2 * (c.in_freq - 1);
    if (c.win_length > n_fft || c.win_length < 1 || c.hop_length < 1 || c.win_length % c.hop_length ||
        (n_fft - c.win_length) % 2 || (c.hop_length & 3))
        return fail(ESCB_EINVAL, "unsupported STFT geometry (n_fft=%d win=%d hop=%d)", n_fft, c.win_length, c.hop_length);
    if ((n_fft / 2 - (n_fft - c.win_length) / 2) % c.hop_length)
        return fail(ESCB_EINVAL, "window support must start on a hop boundary");
    if (c.h_dims[0] > kEmbedMaxC) return fail(ESCB_EINVAL, "h_dims[0] must be <= %d", kEmbedMaxC);
    int top = c.in_freq / c.patch_freq;
    for (int l = 0; l < c.num_levels; ++l) {
        if (c.h_dims[l] < 1) return fail(ESCB_EINVAL, "h_dims[%d] must be positive", l);
        if (l < c.num_levels - 1 && (top >> l) % 2) return fail(ESCB_EINVAL, "odd frequency-patch counts are not supported");
        if (l > 0 && (c.h_dims[l] & 3)) return fail(ESCB_EINVAL, "h_dims[%d] must be a multiple of 4", l);
        // LayerNorm rows are held in registers up to tc::LN_MAX_K floats (PatchMerge normalises 2 * h_dims[l])
        if (c.h_dims[l] > tc::LN_MAX_K || (l < c.num_levels - 1 && 2 * c.h_dims[l] > tc::LN_MAX_K))
            return fail(ESCB_EINVAL, "h_dims[%d]=%d is wider than the LayerNorm kernels support (%d, or %d below the last level)",
                        l, c.h_dims[l], tc::LN_MAX_K, tc::LN_MAX_K / 2);
        if (!argmin_supported(c.codebook_dims[l]))
            return fail(ESCB_EINVAL, "codebook_dims[%d]=%d has no argmin kernel (6, 8, 12, 16, 24, 32)", l, c.codebook_dims[l]);
        if (c.codebook_size % 4) return fail(ESCB_EINVAL, "codebook_size must be a multiple of 4");
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) {
        cudaGetLastError();
        return fail(ESCB_ENODEV, "no usable CUDA device (libescb200 has no CPU fallback)");
    }
    escb_handle* h = new escb_handle();
    h->cfg = c;
    if (const char* e = getenv("ESCB_GEMM")) h->use_tc = strcmp(e, "simt") != 0;
    if (const char* e = getenv("ESCB_PVQ")) h->pvq_tc = strcmp(e, "simt") != 0;
    if (const char* e = getenv("ESCB_FUSE_ATTN_MAXC")) h->fuse_attn_max_c = atoi(e);
    if (const char* e = getenv("ESCB_FUSE_MLP")) h->fuse_mlp = atoi(e) != 0;
    if (const char* e = getenv("ESCB_FUSE_MLP_MAXC")) h->fuse_mlp_max_c = atoi(e);
    if (const char* e = getenv("ESCB_EMIT_STATS")) h->emit_stats = atoi(e) != 0;
    if (const char* e = getenv("ESCB_FUSE_PVQ")) h->fuse_pvq = atoi(e) != 0;
    if (const char* e = getenv("ESCB_LN_POST")) h->ln_post = atoi(e);
    cudaGetDevice(&h->device);
    h->L = c.num_levels;
    h->F = c.in_freq; h->n_fft = n_fft; h->win = c.win_length; h->hop = c.hop_length;
    h->pf = c.patch_freq; h->pt = c.patch_time; h->C0 = c.h_dims[0]; h->nov = c.win_length / c.hop_length;
    for (int l = 0; l < h->L; ++l) h->lev[l] = {c.h_dims[l], top >> l};
    for (int li = 0; li < 2 * h->L; ++li) {
        const LayerDesc d = layer_desc(h, li);
        if (d.heads < 1 || d.C % d.heads || 16 * d.heads > 1024 || !attention_supported(d.C / d.heads)) {
            const int code = fail(ESCB_EINVAL, "%s: %d channels / %d heads has no attention kernel", d.prefix.c_str(), d.C, d.heads);
            delete h;
            return code;
        }
    }
    build_manifest(h);
    cudaError_t e = cudaHostAlloc((void**)&h->err_host, sizeof(int), cudaHostAllocMapped);
    if (e == cudaSuccess) { *h->err_host = 0; e = cudaHostGetDevicePointer((void**)&h->err_dev, h->err_host, 0); }
    if (e == cudaSuccess) e = swin_init();
    if (e == cudaSuccess) e = frontend_init();
    if (e == cudaSuccess) e = pvq_init();
    if (e != cudaSuccess) {
        delete h;
        return fail(ESCB_ECUDA, "kernel attribute setup failed: %s", cudaGetErrorString(e));
    }
    *out = h;
    return ESCB_OK;
}

void escb_destroy(escb_handle* h) {
    if (!h) return;
    if (h->arena) cudaFree(h->arena);
    if (h->host_scratch) cudaFree(h->host_scratch);
    if (h->err_host) cudaFreeHost(h->err_host);
    delete h;
}

int escb_num_weights(const escb_handle* h) { return h ? (int)h->weights.size() : 0; }
const char* escb_weight_name(const escb_handle* h, int i) {
    return (h && i >= 0 && i < (int)h->weights.size()) ? h->weights[i].name.c_str() : nullptr;
}
int64_t escb_weight_numel(const escb_handle* h, int i) {
    return (h && i >= 0 && i < (int)h->weights.size()) ? h->weights[i].numel : -1;
}

int escb_set_weight(escb_handle* h, const char* name, const float* data, int64_t numel, int is_device) {
    if (!h || !name || !data) return fail(ESCB_EINVAL, "null argument");
    auto it = h->index.find(name);
    if (it == h->index.end()) return fail(ESCB_EKEY, "unknown weight '%s'", name);
    Weight& w = h->weights[it->second];
    if (numel != w.numel) return fail(ESCB_EINVAL, "size mismatch for %s: got %lld elements, expected %lld", name,
                                      (long long)numel, (long long)w.numel);
    w.host.resize((size_t)numel);
    if (is_device) {
        const cudaError_t e = cudaMemcpy(w.host.data(), data, (size_t)numel * sizeof(float), cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) return fail(ESCB_ECUDA, "copy of %s failed: %s", name, cudaGetErrorString(e));
    } else {
        memcpy(w.host.data(), data, (size_t)numel * sizeof(float));
    }
    w.set = true;
    h->finalized = false;
    return ESCB_OK;
}

int escb_finalize(escb_handle* h) {
    if (!h) return fail(ESCB_EINVAL, "null handle");
    for (const Weight& w : h->weights)
        if (!w.set) return fail(ESCB_ESTATE, "weight '%s' has not been set", w.name.c_str());
    Packer P{h};
    for (int li = 0; li < 2 * h->L; ++li) pack_layer(P, li);
    if (h->cfg.num_rvqs > 0) pack_rvq(P);
    else for (int q = 0; q < h->L; ++q) pack_quant(P, q);
    pack_front(P);
    cudaSetDevice(h->device);
    if (h->arena) { cudaFree(h->arena); h->arena = nullptr; }
    const size_t bytes = P.arena.data.size() * sizeof(float);
    cudaError_t e = cudaMalloc((void**)&h->arena, bytes);
    if (e != cudaSuccess) return fail(ESCB_ENOMEM, "cudaMalloc of %zu weight bytes failed: %s", bytes, cudaGetErrorString(e));
    e = cudaMemcpy(h->arena, P.arena.data.data(), bytes, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) return fail(ESCB_ECUDA, "weight upload failed: %s", cudaGetErrorString(e));
    for (const Fix& f : P.fixes) *f.slot = h->arena + f.off;
    h->finalized = true;
    return ESCB_OK;
}

// ------------------------------------------------------------------------------------------------ geometry
int escb_time_patches(const escb_handle* h, int64_t num_samples, int32_t* W) {
    if (!h || !W) return fail(ESCB_EINVAL, "null argument");
    int w = 0;
    if (int e = time_patches(h, num_samples, &w)) return e;
    *W = w;
    return ESCB_OK;
}

int64_t escb_decoded_samples(const escb_handle* h, int32_t W) { return h ? (int64_t)h->hop * ((int64_t)h->pt * W - 1) : -1; }

int escb_workspace_bytes(const escb_handle* h, int32_t batch, int32_t W, size_t* bytes) {
    if (!h || !bytes) return fail(ESCB_EINVAL, "null argument");
    if (batch <= 0 || W <= 0) return fail(ESCB_EINVAL, "batch and W must be positive");
    Bump dry(nullptr, 0);
    Work tmp;
    // frames of the longest clip that yields W patches: pt*W + pt - 1
    *bytes = plan(h, batch, W, h->pt * W + h->pt - 1, WK_ENC | WK_DEC | WK_UNIT, dry, tmp) + 256;
    return ESCB_OK;
}

// ------------------------------------------------------------------------------------------------ hot path
int escb_encode(escb_handle* h, const float* audio, int32_t B, int64_t Ls, int32_t S, int64_t* codes, void* ws,
                size_t ws_bytes, void* stream) {
    if (int e = check_ready(h)) return e;
    if (!audio || !codes) return fail(ESCB_EINVAL, "null argument");
    if (S < 1 || S > max_streams_of(h)) return fail(ESCB_EINVAL, "num_streams must be in [1, %d]", max_streams_of(h));
    int W = 0;
    if (int e = time_patches(h, Ls, &W)) return e;
    const int T = frames_of(h, Ls);
    Ctx c;
    if (int e = begin(h, c, B, W, T, WK_ENC, ws, ws_bytes, stream)) return e;
    op_stft(c.L, h->front, audio, B, Ls, T, c.wk.Sf);
    run_encoder(c, T);
    if (h->cfg.num_rvqs > 0) run_rvq_quantize(c, S, (long long*)codes, nullptr, nullptr);
    else run_csrvq_encode(c, S, (long long*)codes);
    return finish(c, "escb_encode");
}

int escb_decode(escb_handle* h, const int64_t* codes, int32_t B, int32_t S, int32_t W, float* audio, float* recon_feat,
                void* ws, size_t ws_bytes, void* stream) {
    if (int e = check_ready(h)) return e;
    if (!codes || (!audio && !recon_feat)) return fail(ESCB_EINVAL, "null argument");
    if (S < 1 || S > max_streams_of(h)) return fail(ESCB_EINVAL, "num_streams must be in [1, %d]", max_streams_of(h));
    Ctx c;
    if (int e = begin(h, c, B, W, h->pt * W, WK_DEC, ws, ws_bytes, stream)) return e;
    if (h->cfg.num_rvqs > 0) {
        op_rvq_gather(c.L, h->rvq, (const long long*)codes, S, (long long)B * (W / 2), W / 2, c.wk.zq, ldc(3 * h->rvq.q.d));
        run_rvq_decoder(c, c.wk.zq);
    } else run_csrvq_decode(c, S, (const long long*)codes);
    run_backend(c, audio, recon_feat);
    return finish(c, "escb_decode");
}

int escb_forward(escb_handle* h, const float* audio, int32_t B, int64_t Ls, int32_t S, int64_t* codes, float* audio_out,
                 float* raw_feat, float* recon_feat, float* vq_loss, void* ws, size_t ws_bytes, void* stream) {
    if (int e = check_ready(h)) return e;
    if (!audio) return fail(ESCB_EINVAL, "null argument");
    if (S < 1 || S > max_streams_of(h)) return fail(ESCB_EINVAL, "num_streams must be in [1, %d]", max_streams_of(h));
    int W = 0;
    if (int e = time_patches(h, Ls, &W)) return e;
    const int T = frames_of(h, Ls);
    Ctx c;
    if (int e = begin(h, c, B, W, T, WK_ENC | WK_DEC, ws, ws_bytes, stream)) return e;
    op_stft(c.L, h->front, audio, B, Ls, T, c.wk.Sf);
    if (raw_feat) op_transpose(c.L, c.wk.Sf, raw_feat, B, T, 2 * h->F);
    run_encoder(c, T);
    if (vq_loss) {
        const cudaError_t e = cudaMemsetAsync(vq_loss, 0, (size_t)B * sizeof(float), c.L.st);
        if (e != cudaSuccess && c.L.err == cudaSuccess) c.L.err = e;
    }
    if (h->cfg.num_rvqs > 0) {
        run_rvq_quantize(c, S, codes ? (long long*)codes : c.wk.codes, c.wk.zq, vq_loss ? c.wk.se : nullptr);
        if (vq_loss) op_rvq_loss(c.L, c.wk.se, B, W / 2, h->rvq.q.d, vq_loss);
        run_rvq_decoder(c, c.wk.zq);
    } else run_csrvq_forward(c, S, codes ? (long long*)codes : c.wk.codes, vq_loss);
    run_backend(c, audio_out, recon_feat);
    return finish(c, "escb_forward");
}

// forward(eval) from a precomputed spectrum (x_feat of ESC.forward, codecs.py:33-34): the STFT is skipped.
int escb_forward_feat(escb_handle* h, const float* planes, int32_t B, int32_t T, int32_t S, int64_t* codes, float* audio_out,
                      float* recon_feat, float* vq_loss, void* ws, size_t ws_bytes, void* stream) {
    if (int e = check_ready(h)) return e;
    if (!planes) return fail(ESCB_EINVAL, "null argument");
    if (S < 1 || S > max_streams_of(h)) return fail(ESCB_EINVAL, "num_streams must be in [1, %d]", max_streams_of(h));
    if (T < h->pt) return fail(ESCB_EINVAL, "too few frames");
    const int W = (T - h->pt) / h->pt + 1;
    if (W <= 0 || W % h->cfg.overlap) return fail(ESCB_EINVAL, "Time dimension must be multiple of overlap (W=%d, overlap=%d)", W, h->cfg.overlap);
    Ctx c;
    if (int e = begin(h, c, B, W, T, WK_ENC | WK_DEC, ws, ws_bytes, stream)) return e;
    op_transpose(c.L, planes, c.wk.Sf, B, 2 * h->F, T);                 // [B, 2F, T] planes -> frame-major [B, T, 2F]
    run_encoder(c, T);
    if (vq_loss) {
        const cudaError_t e = cudaMemsetAsync(vq_loss, 0, (size_t)B * sizeof(float), c.L.st);
        if (e != cudaSuccess && c.L.err == cudaSuccess) c.L.err = e;
    }
    if (h->cfg.num_rvqs > 0) {
        run_rvq_quantize(c, S, codes ? (long long*)codes : c.wk.codes, c.wk.zq, vq_loss ? c.wk.se : nullptr);
        if (vq_loss) op_rvq_loss(c.L, c.wk.se, B, W / 2, h->rvq.q.d, vq_loss);
        run_rvq_decoder(c, c.wk.zq);
    } else run_csrvq_forward(c, S, codes ? (long long*)codes : c.wk.codes, vq_loss);
    run_backend(c, audio_out, recon_feat);
    return finish(c, "escb_forward_feat");
}

// Host-buffer variants: H2D, run, D2H, synchronise.
static int host_scratch(escb_handle* h, size_t bytes, void** p) {
    if (h->host_scratch_bytes < bytes) {
        if (h->host_scratch) cudaFree(h->host_scratch);
        h->host_scratch = nullptr;
        h->host_scratch_bytes = 0;
        const cudaError_t e = cudaMalloc(&h->host_scratch, bytes);
        if (e != cudaSuccess) return fail(ESCB_ENOMEM, "cudaMalloc of %zu scratch bytes failed: %s", bytes, cudaGetErrorString(e));
        h->host_scratch_bytes = bytes;
    }
    *p = h->host_scratch;
    return ESCB_OK;
}

int escb_encode_host(escb_handle* h, const float* audio_host, int32_t B, int64_t Ls, int32_t S, int64_t* codes_host,
                     void* stream) {
    if (int e = check_ready(h)) return e;
    if (!audio_host || !codes_host) return fail(ESCB_EINVAL, "null argument");
    if (B <= 0) return fail(ESCB_EINVAL, "batch must be positive");
    if (S < 1 || S > max_streams_of(h)) return fail(ESCB_EINVAL, "num_streams must be in [1, %d]", max_streams_of(h));
    int W = 0;
    if (int e = time_patches(h, Ls, &W)) return e;
    size_t ws = 0;
    if (int e = escb_workspace_bytes(h, B, W, &ws)) return e;
    const size_t a_bytes = align_up((size_t)B * Ls * sizeof(float), 256);
    const size_t c_bytes = align_up((size_t)B * S * 3 * (W / 2) * sizeof(int64_t), 256);
    std::lock_guard<std::mutex> lock(h->host_mu);
    void* base = nullptr;
    if (int e = host_scratch(h, a_bytes + c_bytes + ws, &base)) return e;
    float* a_dev = (float*)base;
    int64_t* c_dev = (int64_t*)((char*)base + a_bytes);
    void* w_dev = (char*)base + a_bytes + c_bytes;
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMemcpyAsync(a_dev, audio_host, (size_t)B * Ls * sizeof(float), cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) return fail(ESCB_ECUDA, "H2D copy failed: %s", cudaGetErrorString(e));
    if (int r = escb_encode(h, a_dev, B, Ls, S, c_dev, w_dev, ws, stream)) return r;
    e = cudaMemcpyAsync(codes_host, c_dev, (size_t)B * S * 3 * (W / 2) * sizeof(int64_t), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return fail(ESCB_ECUDA, "escb_encode_host: %s", cudaGetErrorString(e));
    return ESCB_OK;
}

int escb_decode_host(escb_handle* h, const int64_t* codes_host, int32_t B, int32_t S, int32_t W, float* audio_host,
                     void* stream) {
    if (int e = check_ready(h)) return e;
    if (!codes_host || !audio_host) return fail(ESCB_EINVAL, "null argument");
    if (B <= 0 || W <= 0) return fail(ESCB_EINVAL, "batch and W must be positive");
    if (S < 1 || S > max_streams_of(h)) return fail(ESCB_EINVAL, "num_streams must be in [1, %d]", max_streams_of(h));
    size_t ws = 0;
    if (int e = escb_workspace_bytes(h, B, W, &ws)) return e;
    const int64_t n_out = escb_decoded_samples(h, W);
    const size_t a_bytes = align_up((size_t)B * n_out * sizeof(float), 256);
    const size_t c_bytes = align_up((size_t)B * S * 3 * (W / 2) * sizeof(int64_t), 256);
    std::lock_guard<std::mutex> lock(h->host_mu);
    void* base = nullptr;
    if (int e = host_scratch(h, a_bytes + c_bytes + ws, &base)) return e;
    float* a_dev = (float*)base;
    int64_t* c_dev = (int64_t*)((char*)base + a_bytes);
    void* w_dev = (char*)base + a_bytes + c_bytes;
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMemcpyAsync(c_dev, codes_host, (size_t)B * S * 3 * (W / 2) * sizeof(int64_t), cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) return fail(ESCB_ECUDA, "H2D copy failed: %s", cudaGetErrorString(e));
    if (int r = escb_decode(h, c_dev, B, S, W, a_dev, nullptr, w_dev, ws, stream)) return r;
    e = cudaMemcpyAsync(audio_host, a_dev, (size_t)B * n_out * sizeof(float), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return fail(ESCB_ECUDA, "escb_decode_host: %s", cudaGetErrorString(e));
    return escb_poll_error(h);
}

// ------------------------------------------------------------------------------------------------ unit entry points
int escb_stft(escb_handle* h, const float* audio, int32_t B, int64_t Ls, float* planes, void* ws, size_t ws_bytes,
              void* stream) {
    if (int e = check_ready(h)) return e;
    if (!audio || !planes) return fail(ESCB_EINVAL, "null argument");
    if (Ls <= h->n_fft / 2) return fail(ESCB_EINVAL, "clip too short for reflect padding");
    const int T = frames_of(h, Ls);
    Ctx c;
    if (int e = begin(h, c, B, h->cfg.overlap, T, WK_UNIT, ws, ws_bytes, stream)) return e;
    op_stft(c.L, h->front, audio, B, Ls, T, c.wk.dense);
    op_transpose(c.L, c.wk.dense, planes, B, T, 2 * h->F);
    return finish(c, "escb_stft");
}

int escb_istft(escb_handle* h, const float* planes, int32_t B, int32_t T, float* audio, void* ws, size_t ws_bytes,
               void* stream) {
    if (int e = check_ready(h)) return e;
    if (!audio || !planes) return fail(ESCB_EINVAL, "null argument");
    if (T < 2) return fail(ESCB_EINVAL, "need at least 2 frames");
    Ctx c;
    if (int e = begin(h, c, B, h->cfg.overlap, T, WK_UNIT, ws, ws_bytes, stream)) return e;
    op_transpose(c.L, planes, c.wk.dense, B, 2 * h->F, T);
    op_istft(c.L, h->front, c.wk.dense, B, T, audio);
    return finish(c, "escb_istft");
}

int escb_patch_embed(escb_handle* h, const float* planes, int32_t B, int32_t T, float* tokens, void* ws,
                     size_t ws_bytes, void* stream) {
    if (int e = check_ready(h)) return e;
    if (!tokens || !planes) return fail(ESCB_EINVAL, "null argument");
    const int W = (T - h->pt) / h->pt + 1;
    if (W <= 0) return fail(ESCB_EINVAL, "too few frames");
    Ctx c;
    if (int e = begin(h, c, B, round_up(W, 2), T, WK_UNIT, ws, ws_bytes, stream)) return e;
    const int H = h->lev[0].H;
    op_transpose(c.L, planes, c.wk.dense, B, 2 * h->F, T);
    op_patch_embed(c.L, h->front, c.wk.dense, B, T, H, W, c.wk.xw, ldc(h->C0));
    op_repitch(c.L, c.wk.xw, ldc(h->C0), tokens, h->C0, h->C0, (long long)B * H * W);
    return finish(c, "escb_patch_embed");
}

int escb_patch_deembed(escb_handle* h, const float* tokens, int32_t B, int32_t W, float* planes, void* ws,
                       size_t ws_bytes, void* stream) {
    if (int e = check_ready(h)) return e;
    if (!tokens || !planes) return fail(ESCB_EINVAL, "null argument");
    if (W <= 0) return fail(ESCB_EINVAL, "W must be positive");
    Ctx c;
    if (int e = begin(h, c, B, round_up(W, 2), h->pt * W, WK_UNIT | WK_DEC, ws, ws_bytes, stream)) return e;
    const int H = h->lev[0].H;
    op_repitch(c.L, tokens, h->C0, c.wk.post, ldc(h->C0), h->C0, (long long)B * H * W);
    op_deembed(c.L, h->front, c.wk.post, ldc(h->C0), B, H, W, c.wk.Y1, c.wk.Sf);
    op_transpose(c.L, c.wk.Sf, planes, B, h->pt * W, 2 * h->F);
    return finish(c, "escb_patch_deembed");
}

int escb_swin_layer(escb_handle* h, int32_t li, const float* x, int32_t B, int32_t H, int32_t W, float* y, void* ws,
                    size_t ws_bytes, void* stream) {
    if (int e = check_ready(h)) return e;
    if (!x || !y) return fail(ESCB_EINVAL, "null argument");
    if (li < 0 || li >= 2 * h->L) return fail(ESCB_EINVAL, "layer_index out of range");
    const LayerDesc d = layer_desc(h, li);
    if (H != d.H) return fail(ESCB_EINVAL, "layer %s runs at H=%d, got %d", d.prefix.c_str(), d.H, H);
    if (W <= 0) return fail(ESCB_EINVAL, "W must be positive");
    Ctx c;
    if (int e = begin(h, c, B, round_up(W, 2), h->pt * W, WK_UNIT, ws, ws_bytes, stream)) return e;
    c.W = W;
    const long long M = (long long)B * H * W;
    // stage: dense -> padded rows, run, padded -> dense
    float* xin = c.wk.stage;
    op_repitch(c.L, x, d.C, xin, ldc(d.C), d.C, M);
    float* out = c.wk.dense;
    run_layer(c, li, xin, c.wk.xw, out, H);
    const long long Mo = d.scale == 1 ? M / 2 : (d.scale == 2 ? M * 2 : M);
    op_repitch(c.L, d.scale ? out : c.wk.xw, ldc(d.out_dim), y, d.out_dim, d.out_dim, Mo);
    return finish(c, "escb_swin_layer");
}

int escb_pvq_encode(escb_handle* h, int32_t q, const float* enc, const float* dec, int32_t B, int32_t W, int64_t* codes,
                    void* ws, size_t ws_bytes, void* stream) {
    if (int e = check_ready(h)) return e;
    if (!enc || !codes) return fail(ESCB_EINVAL, "null argument");
    if (q < 0 || q >= h->L || h->cfg.num_rvqs > 0) return fail(ESCB_EINVAL, "stream index out of range");
    Ctx c;
    if (int e = begin(h, c, B, W, h->pt * W, WK_UNIT, ws, ws_bytes, stream)) return e;
    pvq_encode(c, q, enc, dec, (long long*)codes, 1, 0);
    return finish(c, "escb_pvq_encode");
}

int escb_pvq_decode(escb_handle* h, int32_t q, const int64_t* codes, const float* dec, int32_t B, int32_t W, float* out,
                    void* ws, size_t ws_bytes, void* stream) {
    if (int e = check_ready(h)) return e;
    if (!codes || !out) return fail(ESCB_EINVAL, "null argument");
    if (q < 0 || q >= h->L || h->cfg.num_rvqs > 0) return fail(ESCB_EINVAL, "stream index out of range (the per-stream entry points are ESC's)");
    Ctx c;
    if (int e = begin(h, c, B, W, h->pt * W, WK_UNIT, ws, ws_bytes, stream)) return e;
    if (!(c.L.fuse_pvq && op_pvq_stream(c.L, h->quants[q], nullptr, dec, B, W, (long long*)codes, 1, 0, out, nullptr, 0)))
        op_pvq_up(c.L, h->quants[q], (const long long*)codes, 1, 0, dec, B, W, out);
    return finish(c, "escb_pvq_decode");
}

int escb_pvq_stream(escb_handle* h, int32_t q, const float* enc, const float* dec, int32_t B, int32_t W, int64_t* codes,
                    float* out, void* ws, size_t ws_bytes, void* stream) {
    if (int e = check_ready(h)) return e;
    if (!enc || !codes) return fail(ESCB_EINVAL, "null argument");
    if (q < 0 || q >= h->L || h->cfg.num_rvqs > 0) return fail(ESCB_EINVAL, "stream index out of range (the per-stream entry points are ESC's)");
    Ctx c;
    if (int e = begin(h, c, B, W, h->pt * W, WK_UNIT, ws, ws_bytes, stream)) return e;
    if (!(c.L.fuse_pvq && op_pvq_stream(c.L, h->quants[q], enc, dec, B, W, (long long*)codes, 1, 0, out, nullptr, 0))) {
        pvq_encode(c, q, enc, dec, (long long*)codes, 1, 0);
        if (out) op_pvq_up(c.L, h->quants[q], (const long long*)codes, 1, 0, dec, B, W, out);
    }
    return finish(c, "escb_pvq_stream");
}

int escb_codebook_argmin(escb_handle* h, int32_t q, int32_t g, const float* z, int64_t rows, int64_t* idx, void* stream) {
    if (int e = check_ready(h)) return e;
    if (!z || !idx) return fail(ESCB_EINVAL, "null argument");
    if (q < 0 || q >= h->L || g < 0 || g >= 3 || h->cfg.num_rvqs > 0) return fail(ESCB_EINVAL, "stream/group index out of range");
    if (rows <= 0) return ESCB_OK;
    Ctx c;
    c.h = h;
    c.L.st = (cudaStream_t)stream;
    c.L.prof = h->prof;
    const QuantW& qw = h->quants[q];
    op_argmin(c.L, qw, g, 1, z, qw.d, rows, (long long*)idx, (int)std::min<int64_t>(rows, 1 << 30), 0);
    return finish(c, "escb_codebook_argmin");
}

int escb_code_histogram(const int64_t* codes, int32_t B, int32_t S, int32_t G, int32_t T, int32_t ncodes, float* counts,
                        void* stream) {
    if (!codes || !counts) return fail(ESCB_EINVAL, "null argument");
    if (B <= 0 || S <= 0 || G <= 0 || T <= 0 || ncodes <= 0)
        return fail(ESCB_EINVAL, "codes must have shape (B, S, G, T) with positive extents");
    if ((size_t)ncodes * sizeof(unsigned) > 48 * 1024) return fail(ESCB_EINVAL, "codebook_size too large for the histogram kernel");
    Launcher L;
    L.st = (cudaStream_t)stream;
    op_code_histogram(L, (const long long*)codes, B, S, G, T, ncodes, counts);
    if (L.err != cudaSuccess) return fail(ESCB_ECUDA, "escb_code_histogram: %s", cudaGetErrorString(L.err));
    return ESCB_OK;
}

static const char* const kOpNames[OP_COUNT] = {
    "stft_gemm", "patch_embed", "qkv_gemm", "window_attention", "proj_gemm", "mlp1_gemm", "mlp2_gemm", "merge_gemm",
    "split_gemm", "pvq_down_gemm", "codebook_argmin", "pvq_up_gemm", "vq_loss", "deembed_conv5x5_gemm",
    "deembed_conv3x3", "istft_gemm", "layout", "qkv_attention_fused", "mlp_fused", "pvq_stream_fused"};

int escb_profile_begin(escb_handle* h) {
    if (!h) return fail(ESCB_EINVAL, "null handle");
    if (h->prof) return fail(ESCB_ESTATE, "profiling is already active");
    h->prof = new Profiler();
    return ESCB_OK;
}

int escb_profile_end(escb_handle* h, escb_op_stat* stats, int32_t* n) {
    if (!h || !stats || !n) return fail(ESCB_EINVAL, "null argument");
    if (!h->prof) return fail(ESCB_ESTATE, "profiling is not active");
    Profiler* p = h->prof;
    h->prof = nullptr;
    for (int i = 0; i < OP_COUNT; ++i) stats[i] = escb_op_stat{kOpNames[i], 0, 0.0, 0.0, 0.0};
    const cudaError_t e = cudaDeviceSynchronize();
    const bool dump = getenv("ESCB_PROFILE_DUMP") != nullptr;      // debug: one stderr line per launch, in launch order
    for (ProfRec& r : p->recs) {
        float ms = 0.f;
        if (e == cudaSuccess && cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) {
            if (dump) fprintf(stderr, "escb_launch %s %.4f ms %.0f flop %.0f B\n", kOpNames[r.op], ms, r.flops, r.bytes);
            stats[r.op].launches += 1;
            stats[r.op].ms += ms;
            stats[r.op].flops += r.flops;
            stats[r.op].bytes += r.bytes;
        }
        cudaEventDestroy(r.a);
        cudaEventDestroy(r.b);
    }
    delete p;
    *n = OP_COUNT;
    if (e != cudaSuccess) return fail(ESCB_ECUDA, "escb_profile_end: %s", cudaGetErrorString(e));
    return ESCB_OK;
}

#ifdef ESCB_TC_TRACE
extern "C" __attribute__((visibility("default"))) int escb_debug_trace(escb_handle* h, unsigned long long* out_host) {
    if (!h) return -1;
    if (!h->trace) { cudaMalloc((void**)&h->trace, 1024 * 16 * 8); cudaMemset(h->trace, 0, 1024 * 16 * 8); return 0; }
    cudaDeviceSynchronize();
    if (out_host) cudaMemcpy(out_host, h->trace, 1024 * 16 * 8, cudaMemcpyDeviceToHost);
    return 0;
}
#endif

int escb_poll_error(escb_handle* h) {
    if (!h) return fail(ESCB_EINVAL, "null handle");
    if (h->err_host && *(volatile int*)h->err_host) {
        *(volatile int*)h->err_host = 0;
        return fail(ESCB_EINVAL, "code index out of range [0, %d): the row was decoded as index 0", h->cfg.codebook_size);
    }
    return ESCB_OK;
}

int64_t escb_launch_count(const escb_handle* h) { return h ? (int64_t)h->launches.load() : 0; }

int escb_tiling_info(int32_t N, int32_t K, int32_t role, int32_t* out) {
    if (!out || N <= 0 || K <= 0 || role < 0 || role > 2) return fail(ESCB_EINVAL, "escb_tiling_info: bad argument");
    const tc::Tiling t = tc::choose_tiling(N, K, role);
    if (t.BN <= 0) return fail(ESCB_EINVAL, "escb_tiling_info: no tiling for N = %d, K = %d", N, K);
    out[0] = t.BN; out[1] = t.nsub; out[2] = t.ntn; out[3] = t.nkb; out[4] = t.resident; out[5] = t.nmain; out[6] = t.corr;
    out[7] = t.BN * t.nsub * (t.nmain + t.corr);
    return ESCB_OK;
}

}  // extern "C"
