// Product-VQ ops of one stream.  Reference: quantization.py:74-136 (ProductVectorQuantize.encode/decode),
// :388-432 (pre/post_process), codebook.py:20-55 (argmin / de-quantisation), csrvq.py:15-21,50-60 (fuse).
#include "internal.h"
#include "kernels.cuh"

namespace escb {

static inline LnParams noln(Launcher& L) { return LnParams{nullptr, nullptr, 0.f, nullptr, L.next_trace(), nullptr, nullptr}; }

void op_pvq_down(Launcher& L, const QuantW& q, const float* enc, const float* dec, int B, int W, float* ze, int ldz) {
    AFrame al{enc, dec, q.in_freq, W, q.in_dim};
    EpiRows<false, false> ep{ze, nullptr, nullptr, ldz, 0};
    const long long M = (long long)B * (W / 2);
    L.begin(OP_PVQ_DOWN, 2.0 * M * q.frame_dim * q.d, 4.0 * M * ((dec ? 2.0 : 1.0) * q.frame_dim + 3.0 * q.d));
    // stays on the fp32 SIMT engine: its output feeds the argmin directly (and is no faster on the tensor cores)
    if (q.run > 0 && q.d <= 48) {
        // block-diagonal form: three GEMMs [M, frame/3] x [frame/3, d] in one grid (gemm.cuh Gemm3)
        Gemm3<AFrameG, EpiRows<false, false>> g;
        for (int i = 0; i < 3; ++i) {
            const int goff = i * q.run;
            g.al[i] = AFrameG{enc, dec, q.in_freq, W, q.in_dim, q.run, goff, (q.run % 4 == 0 && q.in_dim % 4 == 0) ? 1 : 0,
                              FastDiv::make((unsigned)q.run)};
            g.ep[i] = EpiRows<false, false>{ze + i * q.d, nullptr, nullptr, ldz, 0};
            g.wt[i] = q.down_g[i].wt;
        }
        L.note(launch_gemm3<3>(L.st, g, q.down_g[0], M));
        return;
    }
    L.note(GemmLauncher<false, AFrame, EpiRows<false, false>, 3, 6>::launch(L.st, al, noln(L), q.down, M, ep));
}

void op_pvq_up(Launcher& L, const QuantW& q, const long long* codes, int S, int s, const float* dec, int B, int W,
               float* out) {
    ACodes al{codes, q.raw, S, s, W / 2, q.d, q.ncodes, L.code_err};
    EpiFrame ep{out, dec, q.in_freq, W, q.in_dim};
    const long long M = (long long)B * (W / 2);
    L.begin(OP_PVQ_UP, 2.0 * M * q.frame_dim * q.d, 4.0 * M * ((dec ? 2.0 : 1.0) * q.frame_dim) + 24.0 * M);
    if (L.pvq_tc) L.note(tc::launch<false, ACodes, EpiFrame, kPvqUpWide>(L.st, al, noln(L), q.up, M, ep));
    else L.note(GemmLauncher<false, ACodes, EpiFrame, 8, 9>::launch(L.st, al, noln(L), q.up, M, ep));
}

template <int D>
static cudaError_t launch_argmin(cudaStream_t st, const QuantW& q, int g_first, int groups, const float* ze, int ldz,
                                 long long rows, long long* out, int T, long long bstride) {
    dim3 grid((unsigned)((rows + kArgminRows - 1) / kArgminRows), (unsigned)groups);
    codebook_argmin_kernel<D><<<grid, 256, 0, st>>>(ze, ldz, q.d, q.cbt + (long long)g_first * q.ncodes * q.d,
                                                    q.cnorm + (long long)g_first * q.ncodes, q.ncodes, rows, out, T,
                                                    bstride, q.l2norm);
    return cudaGetLastError();
}

bool argmin_supported(int d) { return d == 6 || d == 8 || d == 12 || d == 16 || d == 24 || d == 32; }

void op_argmin(Launcher& L, const QuantW& q, int g_first, int groups, const float* ze, int ldz, long long rows,
               long long* out, int T, long long bstride) {
    L.begin(OP_ARGMIN, 2.0 * rows * groups * q.ncodes * q.d, rows * groups * (4.0 * q.d + 8.0));
    cudaError_t e = cudaErrorInvalidValue;
    switch (q.d) {
        case 6: e = launch_argmin<6>(L.st, q, g_first, groups, ze, ldz, rows, out, T, bstride); break;
        case 8: e = launch_argmin<8>(L.st, q, g_first, groups, ze, ldz, rows, out, T, bstride); break;
        case 12: e = launch_argmin<12>(L.st, q, g_first, groups, ze, ldz, rows, out, T, bstride); break;
        case 16: e = launch_argmin<16>(L.st, q, g_first, groups, ze, ldz, rows, out, T, bstride); break;
        case 24: e = launch_argmin<24>(L.st, q, g_first, groups, ze, ldz, rows, out, T, bstride); break;
        case 32: e = launch_argmin<32>(L.st, q, g_first, groups, ze, ldz, rows, out, T, bstride); break;
        default: break;
    }
    L.note(e);
}

void op_vq_loss(Launcher& L, const QuantW& q, const float* ze, int ldz, const long long* codes, int S, int s, int B,
                int T, float* loss) {
    L.begin(OP_VQLOSS, 3.0 * B * T * 3 * q.d, 4.0 * B * T * 3 * q.d * 2);
    vq_loss_kernel<<<B, 256, 0, L.st>>>(ze, ldz, q.raw, codes, S, s, T, q.d, 3, q.ncodes, loss);
    L.note(cudaGetLastError());
}

template <int D, int FR>
static cudaError_t launch_stream(cudaStream_t st, const PvqStreamArgs& a) {
    constexpr int LDW = (D + 3) & ~3;
    const size_t smem = a.decode_only ? 0 : ((size_t)FR * (a.Kg + 4) + kPvqStages * kPvqKC * LDW) * sizeof(float);
    if (smem > 200 * 1024) return cudaErrorInvalidValue;
    dim3 grid((unsigned)((a.rows + FR - 1) / FR), 3);
    pvq_stream_kernel<D, FR><<<grid, 256, smem, st>>>(a);
    return cudaGetLastError();
}
template <int D>
static cudaError_t launch_stream_d(cudaStream_t st, const PvqStreamArgs& a) {
    // Frames per CTA: 8 while that is what fills the GPU (N = 1024 frames -> 384 CTAs); with more rows 16 or 32 - every CTA
    // streams its group's projections and codebook from L2, so more frames per CTA means fewer bytes per frame - bounded by
    // two co-resident CTAs (~100 KB of residual tile each).
    const long long ctas8 = (a.rows + 7) / 8 * 3;
    if (ctas8 < 4 * 148) return launch_stream<D, 8>(st, a);
    if (ctas8 >= 16 * 148 && (a.decode_only || (size_t)32 * (a.Kg + 4) * sizeof(float) <= 100 * 1024)) return launch_stream<D, 32>(st, a);
    if (a.decode_only || (size_t)16 * (a.Kg + 4) * sizeof(float) <= 100 * 1024) return launch_stream<D, 16>(st, a);
    return launch_stream<D, 8>(st, a);
}
#define ESCB_STREAM_DS(X) X(6) X(8) X(12) X(16) X(24) X(32)

cudaError_t pvq_init() {
    cudaError_t e = cudaSuccess;
#define X(n)                                                                                                                        \
    if (e == cudaSuccess) e = cudaFuncSetAttribute(pvq_stream_kernel<n, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);   \
    if (e == cudaSuccess) e = cudaFuncSetAttribute(pvq_stream_kernel<n, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);   \
    if (e == cudaSuccess) e = cudaFuncSetAttribute(pvq_stream_kernel<n, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    ESCB_STREAM_DS(X)
#undef X
    return e;
}

// enc == nullptr: decode-only form (codes are read, out = vq.decode(codes) + dec) with the fused kernel's up-projection
// arithmetic, so that decode(encode(x)) reproduces forward(x) bit for bit like the reference does.
bool op_pvq_stream(Launcher& L, const QuantW& q, const float* enc, const float* dec, int B, int W, long long* codes, int S,
                   int s, float* out, float* ze, int ldz) {
    if (q.run <= 0 || (q.run & 3) || (q.ncodes & 3) || !argmin_supported(q.d) || (q.run * q.in_freq) % kPvqKC) return false;
    if (!enc && !out) return false;
    const int T = W / 2;
    PvqStreamArgs a;
    a.E = enc; a.Dm = dec; a.out = out;
    a.codes = codes + (long long)s * 3 * T;
    a.cstride = (long long)S * 3 * T;
    a.ze = ze; a.ldz = ldz;
    for (int g = 0; g < 3; ++g) a.wd[g] = q.down_g[g].wt;
    a.ldwd = q.down_g[0].ldw;
    a.wu = q.up.wt; a.ldwu = q.up.ldw;
    a.cbt = q.cbt; a.cnorm = q.cnorm; a.raw = q.raw;
    a.ncodes = q.ncodes; a.Hq = q.in_freq; a.W = W; a.C = q.in_dim; a.run = q.run; a.Kg = q.run * q.in_freq; a.T = T;
    a.rows = (long long)B * T;
    a.decode_only = enc ? 0 : 1;
    a.drun4 = FastDiv::make((unsigned)(q.run / 4));
    a.lgH = -1;
    for (int l = 0; l < 16; ++l) if ((1 << l) == q.in_freq) a.lgH = l;
    a.bad = L.code_err;
    a.l2norm = q.l2norm;
    const double M = (double)a.rows, fd = q.frame_dim;
    if (enc) L.begin(OP_PVQ_STREAM, 2.0 * M * (fd * q.d * (out ? 2.0 : 1.0) + 3.0 * q.ncodes * q.d),
                     4.0 * M * fd * ((dec ? 2.0 : 1.0) + (out ? 1.0 : 0.0)) + 24.0 * M);
    else L.begin(OP_PVQ_UP, 2.0 * M * fd * q.d, 4.0 * M * fd * (dec ? 2.0 : 1.0) + 24.0 * M);
    cudaError_t e = cudaErrorInvalidValue;
    switch (q.d) {
#define X(n) case n: e = launch_stream_d<n>(L.st, a); break;
        ESCB_STREAM_DS(X)
#undef X
        default: break;
    }
    L.note(e);
    return true;
}

// ------------------------------------------------------------------------------------------------ RVQCodecs
void op_rvq_chain(Launcher& L, const RvqW& w, const float* ze, int ldz, long long rows, int S, long long* codes, int T,
                  float* zq, float* se) {
    const QuantW& q = w.q;
    L.begin(OP_ARGMIN, 2.0 * rows * 3 * S * q.ncodes * q.d, rows * 3.0 * (4.0 * q.d + 8.0 * S));
    dim3 grid((unsigned)((rows + kArgminRows - 1) / kArgminRows), 3);
    cudaError_t e = cudaErrorInvalidValue;
    switch (q.d) {
#define X(n) case n: rvq_chain_kernel<n><<<grid, 256, 0, L.st>>>(ze, ldz, w.cbt, w.cnorm, w.raw, q.ncodes, w.S, S, rows, codes, T, zq, se, q.l2norm); e = cudaGetLastError(); break;
        ESCB_STREAM_DS(X)
#undef X
        default: break;
    }
    L.note(e);
}

void op_rvq_gather(Launcher& L, const RvqW& w, const long long* codes, int S, long long rows, int T, float* zq, int ldz) {
    const long long total = rows * 3 * w.q.d;
    L.begin(OP_LAYOUT, 0.0, 8.0 * rows * 3 * S + 4.0 * total);
    rvq_gather_kernel<<<(unsigned)((total + 255) / 256), 256, 0, L.st>>>(codes, w.raw, w.q.ncodes, w.S, S, w.q.d, rows, T, zq, ldz, L.code_err);
    L.note(cudaGetLastError());
}

// proj_up of every group + post_process (quantization.py:366-377): the block-structured up-projection GEMM of the
// product VQ with dense rows as its A operand
void op_rvq_up(Launcher& L, const RvqW& w, const float* zq, int ldz, int B, int W, float* out) {
    const QuantW& q = w.q;
    ARows al{zq, ldz};
    EpiFrame ep{out, nullptr, q.in_freq, W, q.in_dim};
    const long long M = (long long)B * (W / 2);
    L.begin(OP_PVQ_UP, 2.0 * M * q.frame_dim * q.d, 4.0 * M * (q.frame_dim + 3.0 * q.d));
    if (L.pvq_tc) L.note(tc::launch<false, ARows, EpiFrame, kPvqUpWide>(L.st, al, noln(L), q.up, M, ep));
    else L.note(GemmLauncher<false, ARows, EpiFrame, 8, 9>::launch(L.st, al, noln(L), q.up, M, ep));
}

void op_rvq_loss(Launcher& L, const float* se, int B, int T, int d, float* loss) {
    L.begin(OP_VQLOSS, 3.0 * B * T, 4.0 * B * T * 3);
    rvq_loss_kernel<<<B, 256, 0, L.st>>>(se, T, d, loss);
    L.note(cudaGetLastError());
}

void op_code_histogram(Launcher& L, const long long* codes, int B, int S, int G, int T, int ncodes, float* counts) {
    L.begin(OP_LAYOUT, 0.0, 8.0 * B * S * G * T + 8.0 * S * G * ncodes);
    code_histogram_kernel<<<S * G, 256, (size_t)ncodes * sizeof(unsigned), L.st>>>(codes, B, S, G, T, ncodes, counts, L.code_err);
    L.note(cudaGetLastError());
}

}  // namespace escb
