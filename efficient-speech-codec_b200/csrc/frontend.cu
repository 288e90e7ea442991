// STFT front end, patch embedding, patch de-embedding and inverse STFT.  Reference: base.py:22-47,
// scale.py:26-81.  The DFT and inverse DFT are GEMMs against windowed bases packed at load time.
#include "internal.h"
#include "kernels.cuh"

namespace escb {

static inline LnParams noln(Launcher& L) { return LnParams{nullptr, nullptr, 0.f, nullptr, L.next_trace(), nullptr, nullptr}; }

#ifndef ESCB_C3_TPT
#define ESCB_C3_TPT 1
#endif
constexpr int kC3Tpt = ESCB_C3_TPT;      // frames per thread of the specialised output conv (1, 2 or 4; measured in DESIGN.md)
static size_t conv3_smem_bytes(int ld) { return (size_t)(kC3F + 2) * ((kC3T + 2) * ld + 4) * sizeof(float); }

cudaError_t frontend_init() {
    cudaError_t e = cudaFuncSetAttribute(conv3x3_out_kernel<0, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)conv3_smem_bytes(ldc(kEmbedMaxC)));
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(conv3x3_out_kernel<45, kC3Tpt>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)conv3_smem_bytes(ldc(45)));
    return e;
}

void op_stft(Launcher& L, const FrontW& f, const float* audio, int B, long long Ls, int T, float* Sf) {
    L.begin(OP_STFT, 2.0 * B * T * f.win * 2 * f.F, 4.0 * (1.0 * B * Ls + 2.0 * B * T * f.F));
    AStftFrames al{audio, Ls, T, f.hop, f.win / 2};
    EpiRows<false, false> ep{Sf, nullptr, nullptr, 2 * f.F, 0};
    L.note(GemmLauncher<false, AStftFrames, EpiRows<false, false>, 8>::launch(L.st, al, noln(L), f.dft, (long long)B * T, ep));
}

void op_patch_embed(Launcher& L, const FrontW& f, const float* Sf, int B, int T, int H, int W, float* tok, int ld) {
    const long long total = (long long)B * H * W;
    L.begin(OP_EMBED, 2.0 * total * f.C0 * 2 * f.pf * f.pt, 4.0 * total * (2.0 * f.pf * f.pt + f.C0));
    dim3 grid((H + kPeH - 1) / kPeH, (W + kPeW - 1) / kPeW, B);
    if (f.embed_k_ok && ld >= 48)
        patch_embed45_kernel<<<grid, kPeH * kPeW, 0, L.st>>>(Sf, T, f.F, tok, ld, f.embed_k, H, W, kLnEps);
    else
        patch_embed_kernel<<<grid, kPeH * kPeW, 0, L.st>>>(Sf, T, f.F, tok, ld, f.embed_w, f.embed_b, f.embed_ln.g, f.embed_ln.b,
                                                          f.C0, f.pf, f.pt, H, W, kLnEps);
    L.note(cudaGetLastError());
}

void op_deembed(Launcher& L, const FrontW& f, const float* tok, int ld, int B, int H, int W, float* Y1, float* Xf) {
    const double n1 = (double)f.C0 * f.pf * f.pt;          // true output channels (the packed weight pads them to the pixel pitch)
    L.begin(OP_DEEMBED1, 2.0 * B * H * W * n1 * 25 * f.C0, 4.0 * B * H * W * (f.C0 + n1));
    AIm2col al{tok, ld, H, W, f.C0};
    EpiDeembed ep{Y1, f.de1.bias, ld, H, W, f.C0, f.pf, f.pt};
    if (L.tc) L.note(tc::launch<false, AIm2col, EpiDeembed>(L.st, al, noln(L), f.de1, (long long)B * H * W, ep));
    else L.note(GemmLauncher<false, AIm2col, EpiDeembed, 9>::launch(L.st, al, noln(L), f.de1, (long long)B * H * W, ep));
    const int Fq = H * f.pf, T2 = W * f.pt;
    const long long total = (long long)B * Fq * T2;
    L.begin(OP_DEEMBED2, 2.0 * total * 2 * 9 * f.C0, 4.0 * total * (f.C0 + 2.0));
    dim3 grid((Fq + kC3F - 1) / kC3F, (T2 + kC3T - 1) / kC3T, B);
    if (f.C0 == 45 && ld == 48 && !getenv("ESCB_C3_WHOLE"))
        conv3x3_out45_kernel<<<grid, 256, 0, L.st>>>(Y1, Fq, T2, f.de2_k, f.de2_bias[0], f.de2_bias[1], Xf);
    else if (f.C0 == 45)
        conv3x3_out_kernel<45, kC3Tpt><<<grid, 32 * kC3T / kC3Tpt, conv3_smem_bytes(ld), L.st>>>(Y1, ld, f.C0, Fq, T2, f.de2_k,
                                                                                            f.de2_bias[0], f.de2_bias[1], Xf);
    else
        conv3x3_out_kernel<0, 1><<<grid, 32 * kC3T, conv3_smem_bytes(ld), L.st>>>(Y1, ld, f.C0, Fq, T2, f.de2_k, f.de2_bias[0],
                                                                                f.de2_bias[1], Xf);
    L.note(cudaGetLastError());
}

void op_istft(Launcher& L, const FrontW& f, const float* Xf, int B, int T, float* audio) {
    // output sample s lives in chunk j = s/hop + j0 of the window-support axis, j0 = (n_fft/2 - pad_left)/hop
    const int n_fft = 2 * (f.F - 1);
    const int j0 = (n_fft / 2 - (n_fft - f.win) / 2) / f.hop;
    const int nchunks = T - 1;
    L.begin(OP_ISTFT, 2.0 * B * nchunks * f.hop * f.nov * 2 * f.F, 4.0 * (2.0 * B * T * f.F + 1.0 * B * nchunks * f.hop));
    AIstft al{Xf, T, 2 * f.F, j0, nchunks};
    EpiIstft ep{audio, f.wsq, T, f.hop, f.nov, j0, nchunks, (long long)f.hop * (T - 1)};
    L.note(GemmLauncher<false, AIstft, EpiIstft, 5>::launch(L.st, al, noln(L), f.idft, (long long)B * nchunks, ep));
}

void op_transpose(Launcher& L, const float* in, float* out, int B, int R, int C) {
    L.begin(OP_LAYOUT, 0.0, 8.0 * B * R * C);
    dim3 grid((C + 31) / 32, (R + 31) / 32, B), block(32, 8);
    transpose_kernel<<<grid, block, 0, L.st>>>(in, out, R, C);
    L.note(cudaGetLastError());
}

void op_repitch(Launcher& L, const float* src, int lds, float* dst, int ldd, int C, long long rows) {
    const long long total = rows * C;
    L.begin(OP_LAYOUT, 0.0, 8.0 * total);
    repitch_kernel<<<(unsigned)((total + 255) / 256), 256, 0, L.st>>>(src, lds, dst, ldd, C, total);
    L.note(cudaGetLastError());
}

}  // namespace escb
