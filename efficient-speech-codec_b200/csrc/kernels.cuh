// Non-GEMM kernels of the ESC hot path: window-attention core, codebook argmin, patch embedding,
// the 3x3 output convolution and small layout helpers.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "internal.h"

namespace escb {

// ------------------------------------------------------------------------------------------------ attention core
// One thread per (window, head, query token): 16 scores, softmax, 16-term weighted sum of V, all in registers.
// Mirrors WindowAttention.forward (attention.py:222-241): q is scaled before the product, the relative-position
// bias and then the 0/-100 shifted-window mask are added, softmax over keys.
// qkv rows are window-major (row = window*16 + token) with columns [3][heads][HDP] (head_pad: the qkv GEMM writes
// heads padded to a multiple of 4 floats), so the block's windows are staged with one coalesced float4 copy and
// every q / k / v row is read from shared memory with aligned vector loads (the 16 lanes of a head read K / V rows
// as broadcasts).  The outputs go back through shared memory and leave as one coalesced float4 copy.
template <int HD, int HDP>
__global__ void window_attn_kernel(const float* __restrict__ qkv, const int ldq, float* __restrict__ out,
                                   const int ldo, const float* __restrict__ relbias, const int nH, const int C,
                                   const long long nwin_total, const int wpb, const float scale, const int masked,
                                   const int nW, const int nWw, const int Hp, const int Wp) {
    constexpr int VW = (HDP % 4 == 0) ? 4 : 2;              // vector width of the shared-memory row reads
    constexpr int NV = HDP / VW;
    extern __shared__ __align__(16) float sm[];
    const int tid = threadIdx.x;
    const long long win0 = (long long)blockIdx.x * wpb;
    const int nwin = (int)((nwin_total - win0 < wpb) ? (nwin_total - win0) : wpb);
    {
        const int n4 = nwin * 16 * ldq / 4;
        const float4* src = reinterpret_cast<const float4*>(qkv + win0 * 16 * (long long)ldq);
        float4* dst = reinterpret_cast<float4*>(sm);
        for (int i = tid; i < n4; i += blockDim.x) dst[i] = __ldg(src + i);
    }
    __syncthreads();
    const int per_win = 16 * nH;
    const int lw = tid / per_win;
    const int r = tid - lw * per_win;
    const int h = r >> 4, i = r & 15;
    const bool active = lw < nwin;
    float o[HDP];
    if (active) {
        const float* base = sm + (long long)lw * 16 * ldq;
        auto ldrow = [&](const float* p, float (&v)[HDP]) {
#pragma unroll
            for (int u = 0; u < NV; ++u) {
                if (VW == 4) {
                    const float4 t = *reinterpret_cast<const float4*>(p + 4 * u);
                    v[4 * u] = t.x; v[4 * u + 1] = t.y; v[4 * u + 2] = t.z; v[4 * u + 3] = t.w;
                } else {
                    const float2 t = *reinterpret_cast<const float2*>(p + 2 * u);
                    v[2 * u] = t.x; v[2 * u + 1] = t.y;
                }
            }
        };
        float q[HDP];
        ldrow(base + i * ldq + h * HDP, q);
#pragma unroll
        for (int d = 0; d < HD; ++d) q[d] *= scale;

        float s[16];
        const float* kb = base + nH * HDP + h * HDP;
        const float4* bias = reinterpret_cast<const float4*>(relbias + (h * 16 + i) * 16);
#pragma unroll
        for (int j4 = 0; j4 < 4; ++j4) {
            const float4 bv = __ldg(bias + j4);
            const float bb[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
                const int j = j4 * 4 + jj;
                float kr[HDP];
                ldrow(kb + j * ldq, kr);
                float acc = 0.f;
#pragma unroll
                for (int d = 0; d < HD; ++d) acc = fmaf(q[d], kr[d], acc);
                s[j] = acc + bb[jj];
            }
        }
        if (masked) {
            // region ids of the shifted map (attention.py:56-75): 0 | 1 | 2 along each axis, id = 3*rh + rw
            const int win = (int)((win0 + lw) % nW);
            const int wh = win / nWw, ww = win - wh * nWw;
            int rh[4], rw[4];
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                const int hs = wh * 4 + a, ws = ww * 4 + a;
                rh[a] = hs < Hp - 4 ? 0 : (hs < Hp - 2 ? 1 : 2);
                rw[a] = ws < Wp - 4 ? 0 : (ws < Wp - 2 ? 1 : 2);
            }
            const int mine = 3 * rh[i >> 2] + rw[i & 3];
#pragma unroll
            for (int j = 0; j < 16; ++j)
                if (3 * rh[j >> 2] + rw[j & 3] != mine) s[j] += -100.0f;
        }
        float mx = s[0];
#pragma unroll
        for (int j = 1; j < 16; ++j) mx = fmaxf(mx, s[j]);
        float sum = 0.f;
#pragma unroll
        for (int j = 0; j < 16; ++j) { s[j] = expf(s[j] - mx); sum += s[j]; }
        const float inv = 1.0f / sum;
#pragma unroll
        for (int d = 0; d < HD; ++d) o[d] = 0.f;
        const float* vb = base + 2 * nH * HDP + h * HDP;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const float p = s[j] * inv;
            float vr[HDP];
            ldrow(vb + j * ldq, vr);
#pragma unroll
            for (int d = 0; d < HD; ++d) o[d] = fmaf(p, vr[d], o[d]);
        }
    }
    __syncthreads();                                       // every q / k / v read is done: reuse the tile for the outputs
    const int pitch = ldo + 4;
    if (active) {
        float* op = sm + (lw * 16 + i) * pitch + h * HD;
#pragma unroll
        for (int d = 0; d < HD; ++d) op[d] = o[d];
    }
    __syncthreads();
    {
        const int l4 = ldo / 4, n4 = nwin * 16 * l4;
        float4* dst = reinterpret_cast<float4*>(out + win0 * 16 * (long long)ldo);
        for (int e = tid; e < n4; e += blockDim.x) {
            const int row = e / l4, c4 = e - row * l4;
            dst[e] = *reinterpret_cast<const float4*>(sm + row * pitch + c4 * 4);
        }
    }
}

// ------------------------------------------------------------------------------------------------ RVQ argmin
// Codebook.quantize_to_code (codebook.py:20-43): z and the table are L2-normalised (x / max(|x|, 1e-12)), the
// distance is (|z|^2 - (2z).c) + |c|^2 and the FIRST minimum wins.  cbt / cnorm are the normalised table
// (transposed to [group][d][code]) and its squared norms, precomputed once per weight load.
//
// One block = 32 rows x all codes of one group.  The codebook streams through shared memory in chunks of 256
// codes; each thread keeps an 8-row x 4-code register tile, so one k step is 3 LDS.128 for 32 FFMA (the naive
// one-warp-per-row scan was LDS/LDG bound at 0.8 TFLOP/s).  Every dot product is still the same sequential
// k = 0..d-1 FMA chain, each thread scans its codes in increasing order with a strict <, and the final
// lexicographic (dist, index) reduction keeps the lowest index on ties.  Codes land in the [B, S, G, T] tensor.
constexpr int kArgminRows = 32;
constexpr int kArgminChunk = 256;
template <int D>
__global__ void __launch_bounds__(256)
codebook_argmin_kernel(const float* __restrict__ z, const int ldz, const int d_stride, const float* __restrict__ cbt,
                       const float* __restrict__ cnorm, const int ncodes, const long long rows,
                       long long* __restrict__ out, const int T, const long long bstride, const int l2norm) {
    __shared__ __align__(16) float zs[D][kArgminRows];        // 2 * z_hat, transposed
    __shared__ float zzs[kArgminRows];
    __shared__ __align__(16) float cbs[D][kArgminChunk];
    __shared__ float cns[kArgminChunk];
    __shared__ float redv[kArgminRows][2];
    __shared__ int redi[kArgminRows][2];
    const int tid = threadIdx.x, g = blockIdx.y;
    const long long row0 = (long long)blockIdx.x * kArgminRows;
    if (tid < kArgminRows) {
        const long long m = row0 + tid;
        float zn[D];
        float ss = 0.f;
        if (m < rows) {
            const float* zp = z + m * (long long)ldz + g * d_stride;
#pragma unroll
            for (int k = 0; k < D; ++k) { zn[k] = __ldg(zp + k); ss = fmaf(zn[k], zn[k], ss); }
        } else {
#pragma unroll
            for (int k = 0; k < D; ++k) zn[k] = 0.f;
        }
        const float denom = l2norm ? fmaxf(sqrtf(ss), 1e-12f) : 1.0f;      // l2norm=False: plain squared distance (codebook.py:31-40)
        float zz = 0.f;
#pragma unroll
        for (int k = 0; k < D; ++k) {
            const float v = zn[k] / denom;
            zz = fmaf(v, v, zz);
            zs[k][tid] = 2.0f * v;
        }
        zzs[tid] = zz;
    }
    const int cg = tid & 63, rg = tid >> 6;                   // codes cg*4..+3 of the chunk, rows rg*8..+7
    float best[8];
    int besti[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) { best[r] = 3.0e38f; besti[r] = 0x7fffffff; }
    const float* cb = cbt + (long long)g * D * ncodes;
    const float* cn = cnorm + (long long)g * ncodes;
    for (int chunk = 0; chunk < ncodes; chunk += kArgminChunk) {
        __syncthreads();
        for (int i = tid; i < D * (kArgminChunk / 4); i += 256) {
            const int k = i / (kArgminChunk / 4), c4 = (i % (kArgminChunk / 4)) * 4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (chunk + c4 + 3 < ncodes) v = __ldg(reinterpret_cast<const float4*>(cb + (long long)k * ncodes + chunk + c4));
            else {
                if (chunk + c4 + 0 < ncodes) v.x = __ldg(cb + (long long)k * ncodes + chunk + c4 + 0);
                if (chunk + c4 + 1 < ncodes) v.y = __ldg(cb + (long long)k * ncodes + chunk + c4 + 1);
                if (chunk + c4 + 2 < ncodes) v.z = __ldg(cb + (long long)k * ncodes + chunk + c4 + 2);
            }
            *reinterpret_cast<float4*>(&cbs[k][c4]) = v;
        }
        cns[tid] = (chunk + tid < ncodes) ? __ldg(cn + chunk + tid) : 3.0e38f;
        __syncthreads();
        float acc[8][4];
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[r][j] = 0.f;
#pragma unroll
        for (int k = 0; k < D; ++k) {
            const float4 c = *reinterpret_cast<const float4*>(&cbs[k][cg * 4]);
            const float4 z0 = *reinterpret_cast<const float4*>(&zs[k][rg * 8]);
            const float4 z1 = *reinterpret_cast<const float4*>(&zs[k][rg * 8 + 4]);
            const float zr[8] = {z0.x, z0.y, z0.z, z0.w, z1.x, z1.y, z1.z, z1.w};
            const float cj[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
            for (int r = 0; r < 8; ++r)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[r][j] = fmaf(zr[r], cj[j], acc[r][j]);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int code = chunk + cg * 4 + j;
            const float cnj = cns[cg * 4 + j];
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                const float dist = (zzs[rg * 8 + r] - acc[r][j]) + cnj;
                if (dist < best[r]) { best[r] = dist; besti[r] = code; }
            }
        }
    }
    const int lane = tid & 31, wsel = (tid >> 5) & 1;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        float bv = best[r];
        int bi = besti[r];
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov < bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
        }
        if (lane == 0) { redv[rg * 8 + r][wsel] = bv; redi[rg * 8 + r][wsel] = bi; }
    }
    __syncthreads();
    if (tid < kArgminRows) {
        const long long m = row0 + tid;
        if (m < rows) {
            float bv = redv[tid][0];
            int bi = redi[tid][0];
            const float ov = redv[tid][1];
            const int oi = redi[tid][1];
            if (ov < bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
            if (bi == 0x7fffffff) bi = 0;                     // all-NaN row
            const long long b = m / T;
            const int t = (int)(m - b * T);
            out[b * bstride + (long long)g * T + t] = bi;
        }
    }
}

// ------------------------------------------------------------------------------------------------ patch embedding
// PatchEmbed.forward (scale.py:42-50): non-overlapping (pf x pt) patches of the 2-plane spectrum -> C0 channels,
// then LayerNorm(C0).  The spectrum is read frame-major [B, T, 2F]; one thread per token, adjacent threads take
// adjacent frequency patches so the pf-float reads coalesce.  Weight k index = (c*pf + s1)*pt + s2.
constexpr int kPeH = 32, kPeW = 4;       // tokens per block: 32 frequency patches x 4 time patches
static __global__ void __launch_bounds__(kPeH * kPeW)
patch_embed_kernel(const float* __restrict__ Sf, const int T, const int F, float* __restrict__ tok,
                   const int ld, const float* __restrict__ w, const float* __restrict__ bias,
                   const float* __restrict__ gamma, const float* __restrict__ beta, const int C0,
                   const int pf, const int pt, const int H, const int W, const float eps) {
    // One thread per token: lanes run along the frequency axis, so the pf-float reads of a warp cover one contiguous
    // stretch of the frame; the finished rows go through shared memory ([w][h] rows of pitch 49: conflict-free scalar
    // writes) and leave as float4 segments - the kPeW tokens of one frequency row are kPeW * ld contiguous floats of the
    // token map (l = h * W + w).  Writing each row from its own thread (45 scalar stores, W * ld floats between
    // neighbouring lanes) cost 307 us per step for 145 MB.
    __shared__ float sw[kEmbedMaxC * kEmbedMaxK];
    __shared__ float sb[kEmbedMaxC], sg[kEmbedMaxC], sbe[kEmbedMaxC];
    constexpr int PITCH = kEmbedMaxC + 1 > 49 ? kEmbedMaxC + 1 : 49;
    __shared__ float rows[kPeH * kPeW * PITCH];
    const int KP = 2 * pf * pt, tid = threadIdx.x;
    for (int i = tid; i < C0 * KP; i += kPeH * kPeW) sw[i] = w[i];
    for (int i = tid; i < C0; i += kPeH * kPeW) { sb[i] = bias[i]; sg[i] = gamma[i]; sbe[i] = beta[i]; }
    __syncthreads();
    const int hl = tid & (kPeH - 1), wl = tid / kPeH;
    const int h = blockIdx.x * kPeH + hl, wq = blockIdx.y * kPeW + wl;
    const long long b = blockIdx.z;
    if (h < H && wq < W) {
        float in[kEmbedMaxK];
#pragma unroll
        for (int k = 0; k < kEmbedMaxK; ++k) {
            if (k < KP) {
                const int c = k / (pf * pt), rem = k - c * pf * pt;
                const int s1 = rem / pt, s2 = rem - s1 * pt;
                in[k] = __ldg(Sf + (b * T + (long long)wq * pt + s2) * (2 * F) + c * F + h * pf + s1);
            } else in[k] = 0.f;
        }
        float y[kEmbedMaxC];
        float s = 0.f;
#pragma unroll
        for (int n = 0; n < kEmbedMaxC; ++n) {
            if (n < C0) {
                float acc = 0.f;
#pragma unroll
                for (int k = 0; k < kEmbedMaxK; ++k)
                    if (k < KP) acc = fmaf(in[k], sw[n * KP + k], acc);
                y[n] = acc + sb[n];
                s += y[n];
            }
        }
        const float mean = s / (float)C0;
        float q = 0.f;
#pragma unroll
        for (int n = 0; n < kEmbedMaxC; ++n)
            if (n < C0) { const float dlt = y[n] - mean; q = fmaf(dlt, dlt, q); }
        const float rstd = 1.0f / sqrtf(q / (float)C0 + eps);
        float* r = rows + (wl * kPeH + hl) * PITCH;
#pragma unroll
        for (int n = 0; n < kEmbedMaxC; ++n)
            if (n < C0) r[n] = (y[n] - mean) * rstd * sg[n] + sbe[n];
    }
    __syncthreads();
    // float4 j of the block's output: frequency row j / (kPeW * ld4), then the kPeW tokens of that row back to back
    const int ld4 = ld >> 2, per_h = kPeW * ld4;
    for (int j = tid; j < kPeH * per_h; j += kPeH * kPeW) {
        const int hh = j / per_h, rem = j - hh * per_h, ww = rem / ld4, c4 = rem - ww * ld4;
        const int hg = blockIdx.x * kPeH + hh, wg = blockIdx.y * kPeW + ww;
        if (hg >= H || wg >= W) continue;
        const float* r = rows + (ww * kPeH + hh) * PITCH + 4 * c4;
        float4 v;
        v.x = 4 * c4 + 0 < C0 ? r[0] : 0.f;
        v.y = 4 * c4 + 1 < C0 ? r[1] : 0.f;
        v.z = 4 * c4 + 2 < C0 ? r[2] : 0.f;
        v.w = 4 * c4 + 3 < C0 ? r[3] : 0.f;
        *reinterpret_cast<float4*>(tok + ((b * H + hg) * (long long)W + wg) * ld + 4 * c4) = v;
    }
}

// The shipped geometry (C0 = 45, pf x pt = 3 x 2, 2 planes) with every index a compile-time constant: the 540 weights,
// the bias and the LayerNorm vectors come straight from the constant bank (kernel parameter), no shared-memory operand
// per multiply-add.  Same arithmetic and order as patch_embed_kernel.
static __global__ void __launch_bounds__(kPeH * kPeW, 6)
patch_embed45_kernel(const float* __restrict__ Sf, const int T, const int F, float* __restrict__ tok, const int ld,
                     const __grid_constant__ EmbedWeights ew, const int H, const int W, const float eps) {
    constexpr int C0 = 45, PF = 3, PT = 2, KP = 2 * PF * PT, PITCH = 49;
    __shared__ float rows[kPeH * kPeW * PITCH];            // per token: 45 conv outputs, then its mean and rstd
    __shared__ float sg[48], sbe[48];
    const int tid = threadIdx.x, hl = tid & (kPeH - 1), wl = tid / kPeH;
    const int h = blockIdx.x * kPeH + hl, wq = blockIdx.y * kPeW + wl;
    const long long b = blockIdx.z;
    if (tid < 48) { sg[tid] = tid < C0 ? ew.g[tid] : 0.f; sbe[tid] = tid < C0 ? ew.be[tid] : 0.f; }
    if (h < H && wq < W) {
        float in[KP];
#pragma unroll
        for (int k = 0; k < KP; ++k) {
            const int c = k / (PF * PT), rem = k - c * PF * PT, s1 = rem / PT, s2 = rem - s1 * PT;
            in[k] = __ldg(Sf + (b * T + (long long)wq * PT + s2) * (2 * F) + c * F + h * PF + s1);
        }
        // the outputs go to the token's shared-memory row as they are produced (keeping all 45 in registers next to the
        // 540 constant-bank weights made ptxas preload weights: 255 registers or 900 bytes of spills)
        float* r = rows + (wl * kPeH + hl) * PITCH;
        float s = 0.f;
#pragma unroll 5                                            // full unrolling interleaves all 45 chains: 257 spilled registers
        for (int n = 0; n < C0; ++n) {
            float acc = 0.f;
#pragma unroll
            for (int k = 0; k < KP; ++k) acc = fmaf(in[k], ew.w[n * KP + k], acc);
            const float y = acc + ew.b[n];
            r[n] = y;
            s += y;
        }
        const float mean = s / (float)C0;
        float q = 0.f;
#pragma unroll 5
        for (int n = 0; n < C0; ++n) { const float dlt = r[n] - mean; q = fmaf(dlt, dlt, q); }
        r[C0] = mean;
        r[C0 + 1] = 1.0f / sqrtf(q / (float)C0 + eps);
    }
    __syncthreads();
    const int ld4 = ld >> 2, per_h = kPeW * ld4;
    for (int j = tid; j < kPeH * per_h; j += kPeH * kPeW) {
        const int hh = j / per_h, rem = j - hh * per_h, ww = rem / ld4, c4 = rem - ww * ld4;
        const int hg = blockIdx.x * kPeH + hh, wg = blockIdx.y * kPeW + ww;
        if (hg >= H || wg >= W) continue;
        const float* r = rows + (ww * kPeH + hh) * PITCH;
        const float mean = r[C0], rstd = r[C0 + 1];
        const int c = 4 * c4;
        float4 v;
        v.x = c + 0 < C0 ? (r[c + 0] - mean) * rstd * sg[c + 0] + sbe[c + 0] : 0.f;
        v.y = c + 1 < C0 ? (r[c + 1] - mean) * rstd * sg[c + 1] + sbe[c + 1] : 0.f;
        v.z = c + 2 < C0 ? (r[c + 2] - mean) * rstd * sg[c + 2] + sbe[c + 2] : 0.f;
        v.w = c + 3 < C0 ? (r[c + 3] - mean) * rstd * sg[c + 3] + sbe[c + 3] : 0.f;
        *reinterpret_cast<float4*>(tok + ((b * H + hg) * (long long)W + wg) * ld + c) = v;
    }
}

// ------------------------------------------------------------------------------------------------ output conv
// PatchDeEmbed.de_proj2 (scale.py:70-71,79): 3x3 / pad 1 convolution C0 -> 2 over the channels-last pixel map
// Y1 [B, F, T2, ld]; writes the spectrum frame-major Xf[b, t, c2*F + f] (what the inverse STFT reads).
// One block = 32 frequency bins x 8 frames: the (34 x 10) halo is staged in shared memory with coalesced float4
// loads (row pitch padded by 4 floats so the 32 lanes of a warp, one bin each, read conflict-free), the taps
// [tap][c][2] arrive as a by-value kernel parameter so every FMA takes its weight from the constant bank, and
// lanes run along the bin axis so both output planes are written as 128-byte segments.
// C0T > 0: the channel count as a compile-time constant (45 in every shipped config).  Every tap index is then an
// immediate, so each FMA takes its weight straight from the constant bank; with a run-time C0 the compiler fetched the
// weights with 870 uniform constant loads per thread in front of the 1 548 FMAs (0.63 ms per step -> see DESIGN.md).
constexpr int kC3F = 32, kC3T = 8;
// TPT consecutive frames per thread: a staged pixel feeds every output whose 3-wide window covers it, so the shared
// memory reads per output drop from 108 to 36 (TPT + 2) / TPT float4 and the 2 TPT accumulators are independent FMA
// chains.  Every output still accumulates in the order (kh, kw, c): the results do not depend on TPT.
template <int C0T, int TPT>
static __global__ void __launch_bounds__(32 * kC3T / TPT)
conv3x3_out_kernel(const float* __restrict__ Y1, const int LDr, const int C0r, const int F, const int T2,
                   const __grid_constant__ Conv3Weights wk, const float b0, const float b1, float* __restrict__ Xf) {
    constexpr int NT = 32 * kC3T / TPT;
    const int C0 = C0T > 0 ? C0T : C0r;
    const int LD = C0T > 0 ? ((C0T + 3) & ~3) : LDr;       // the row pitch is ldc(C0): a constant too (shared-memory offsets become immediates)
    const int PITCH = (kC3T + 2) * LD + 4;
    extern __shared__ __align__(16) float halo[];          // [kC3F + 2][PITCH]
    const int f0 = blockIdx.x * kC3F, t0 = blockIdx.y * kC3T;
    const long long b = blockIdx.z;
    const int ROW4 = (kC3T + 2) * LD / 4;
    for (int i = threadIdx.x; i < (kC3F + 2) * ROW4; i += NT) {
        const int fr = i / ROW4, j = i - fr * ROW4;
        const int tt = j / (LD / 4), c4 = j - tt * (LD / 4);
        const int f = f0 + fr - 1, t = t0 + tt - 1;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (f >= 0 && f < F && t >= 0 && t < T2)
            v = __ldg(reinterpret_cast<const float4*>(Y1 + ((b * F + f) * (long long)T2 + t) * LD) + c4);
        *reinterpret_cast<float4*>(halo + fr * PITCH + tt * LD + c4 * 4) = v;
    }
    __syncthreads();
    const int fl = threadIdx.x & 31, tl = (threadIdx.x >> 5) * TPT;
    const int f = f0 + fl;
    float a0[TPT], a1[TPT];
#pragma unroll
    for (int j = 0; j < TPT; ++j) { a0[j] = 0.f; a1[j] = 0.f; }
#pragma unroll
    for (int kh = 0; kh < 3; ++kh)
#pragma unroll
        for (int col = 0; col < TPT + 2; ++col) {
            const float* p = halo + (fl + kh) * PITCH + (tl + col) * LD;
#pragma unroll (C0T > 0 ? 16 : 4)
            for (int c = 0; c + 3 < C0; c += 4) {
                const float4 v = *reinterpret_cast<const float4*>(p + c);
#pragma unroll
                for (int j = 0; j < TPT; ++j) {
                    const int kw = col - j;                // output j sees this pixel through tap kw
                    if (kw < 0 || kw > 2) continue;
                    const float* wv = wk.w + (kh * 3 + kw) * C0 * 2;
                    a0[j] = fmaf(v.x, wv[2 * c + 0], a0[j]); a1[j] = fmaf(v.x, wv[2 * c + 1], a1[j]);
                    a0[j] = fmaf(v.y, wv[2 * c + 2], a0[j]); a1[j] = fmaf(v.y, wv[2 * c + 3], a1[j]);
                    a0[j] = fmaf(v.z, wv[2 * c + 4], a0[j]); a1[j] = fmaf(v.z, wv[2 * c + 5], a1[j]);
                    a0[j] = fmaf(v.w, wv[2 * c + 6], a0[j]); a1[j] = fmaf(v.w, wv[2 * c + 7], a1[j]);
                }
            }
#pragma unroll
            for (int c = C0 & ~3; c < C0; ++c) {
                const float v = p[c];
#pragma unroll
                for (int j = 0; j < TPT; ++j) {
                    const int kw = col - j;
                    if (kw < 0 || kw > 2) continue;
                    const float* wv = wk.w + (kh * 3 + kw) * C0 * 2;
                    a0[j] = fmaf(v, wv[2 * c], a0[j]); a1[j] = fmaf(v, wv[2 * c + 1], a1[j]);
                }
            }
        }
    if (f < F) {
#pragma unroll
        for (int j = 0; j < TPT; ++j) {
            const int t = t0 + tl + j;
            if (t >= T2) continue;
            float* o = Xf + (b * T2 + t) * (long long)(2 * F);
            o[f] = a0[j] + b0;
            o[F + f] = a1[j] + b1;
        }
    }
}

// The shipped geometry (C0 = 45, row pitch 48) staged 16 channels at a time through a two-buffer cp.async pipeline:
// 23 KB of shared memory per buffer instead of 66 KB for the whole halo, so five blocks (40 warps) share an SM instead of
// three, and the loads of the next channel group overlap the multiply-adds of the current one (with the whole halo
// staged in one phase the three resident blocks loaded and computed in lockstep: 484 us for a pass that is 130 us of HBM
// traffic).  Weights are immediates from the constant bank.  Accumulation order per output: (channel group, kh, kw, c).
#ifndef ESCB_C3_G
#define ESCB_C3_G 16
#endif
constexpr int kC3G = ESCB_C3_G;                              // channels per stage (48 / kC3G stages)
constexpr int kC3RowP = (kC3T + 2) * kC3G + 4;               // floats per halo row of one stage (+4: conflict-free float4 reads)
constexpr int kC3BufF = (kC3F + 2) * kC3RowP;                // floats per buffer
static __global__ void __launch_bounds__(256)
conv3x3_out45_kernel(const float* __restrict__ Y1, const int F, const int T2, const __grid_constant__ Conv3Weights wk,
                     const float b0, const float b1, float* __restrict__ Xf) {
    constexpr int C0 = 45, LD = 48;
    __shared__ __align__(16) float halo[2 * kC3BufF];
    const int f0 = blockIdx.x * kC3F, t0 = blockIdx.y * kC3T;
    const long long b = blockIdx.z;
    const int tid = threadIdx.x;
    auto issue = [&](int g) {                                // channel group g -> buffer g & 1
        float* dst = halo + (g & 1) * kC3BufF;
        constexpr int PER_PX = kC3G / 4, ITEMS = (kC3F + 2) * (kC3T + 2) * PER_PX;
        for (int i = tid; i < ITEMS; i += 256) {
            const int px = i / PER_PX, c4 = i - px * PER_PX;
            const int fr = px / (kC3T + 2), tt = px - fr * (kC3T + 2);
            const int f = f0 + fr - 1, t = t0 + tt - 1;
            const bool ok = f >= 0 && f < F && t >= 0 && t < T2;
            const float* src = ok ? Y1 + ((b * F + f) * (long long)T2 + t) * LD + g * kC3G + 4 * c4 : Y1;
            const unsigned d = (unsigned)__cvta_generic_to_shared(dst + fr * kC3RowP + tt * kC3G + 4 * c4);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(ok ? 16 : 0) : "memory");   // out of range: zero fill
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    const int fl = tid & 31, tl = tid >> 5;
    float a0 = 0.f, a1 = 0.f;
    auto compute = [&](int g) {
        const float* base = halo + (g & 1) * kC3BufF + fl * kC3RowP + tl * kC3G;
#pragma unroll
        for (int kh = 0; kh < 3; ++kh)
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) {
                const float* p = base + kh * kC3RowP + kw * kC3G;
                const float* wv = wk.w + (kh * 3 + kw) * C0 * 2;
#pragma unroll
                for (int cc = 0; cc < kC3G; cc += 4) {
                    const int c = g * kC3G + cc;
                    if (c >= C0) continue;
                    const float4 v = *reinterpret_cast<const float4*>(p + cc);
                    a0 = fmaf(v.x, wv[2 * c + 0], a0); a1 = fmaf(v.x, wv[2 * c + 1], a1);
                    if (c + 1 < C0) { a0 = fmaf(v.y, wv[2 * c + 2], a0); a1 = fmaf(v.y, wv[2 * c + 3], a1); }
                    if (c + 2 < C0) { a0 = fmaf(v.z, wv[2 * c + 4], a0); a1 = fmaf(v.z, wv[2 * c + 5], a1); }
                    if (c + 3 < C0) { a0 = fmaf(v.w, wv[2 * c + 6], a0); a1 = fmaf(v.w, wv[2 * c + 7], a1); }
                }
            }
    };
    constexpr int NG = LD / kC3G;
    issue(0);
    issue(1);
#pragma unroll
    for (int g = 0; g < NG; ++g) {
        if (g + 1 < NG) asm volatile("cp.async.wait_group 1;" ::: "memory");      // group g has landed (g + 1 may be in flight)
        else asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();
        compute(g);
        if (g + 2 < NG) {
            __syncthreads();                                 // every thread is done with buffer g & 1
            issue(g + 2);
        }
    }
    const int f = f0 + fl, t = t0 + tl;
    if (f < F && t < T2) {
        float* o = Xf + (b * T2 + t) * (long long)(2 * F);
        o[f] = a0 + b0;
        o[F + f] = a1 + b1;
    }
}

// ------------------------------------------------------------------------------------------------ layout helpers
// Batched 2-D transpose: in [B][R][Cc] -> out [B][Cc][R] (frame-major spectrum <-> [B,2,F,T] planes).
static __global__ void transpose_kernel(const float* __restrict__ in, float* __restrict__ out, const int R, const int Cc) {
    __shared__ float tile[32][33];
    const long long b = blockIdx.z;
    const float* ip = in + b * (long long)R * Cc;
    float* op = out + b * (long long)R * Cc;
    const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int r = r0 + i, c = c0 + threadIdx.x;
        if (r < R && c < Cc) tile[i][threadIdx.x] = ip[(long long)r * Cc + c];
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int c = c0 + i, r = r0 + threadIdx.x;
        if (r < R && c < Cc) op[(long long)c * R + r] = tile[threadIdx.x][i];
    }
}

// Row re-pitch: dst[r*ldd + c] = src[r*lds + c] for c < C (dense <-> padded token rows).
static __global__ void repitch_kernel(const float* __restrict__ src, const int lds, float* __restrict__ dst, const int ldd,
                               const int C, const long long total) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const long long r = idx / C;
    const int c = (int)(idx - r * C);
    dst[r * ldd + c] = src[r * lds + c];
}

// Eval-mode VQ "losses" (codebook.py:71-73; quantization.py:69-72): per batch row, mean over (T, d) of
// (table[code] - z_e)^2 summed over groups / groups, accumulated into loss[b].
static __global__ void vq_loss_kernel(const float* __restrict__ ze, const int ldz, const float* __restrict__ tables,
                               const long long* __restrict__ codes, const int S, const int s, const int T,
                               const int d, const int groups, const int ncodes, float* __restrict__ loss) {
    const int b = blockIdx.x;
    float acc = 0.f;
    const int n = groups * T * d;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int g = i / (T * d);
        const int rem = i - g * T * d;
        const int t = rem / d, dd = rem - t * d;
        const long long c = codes[(((long long)b * S + s) * groups + g) * T + t];
        const float diff = tables[((long long)g * ncodes + c) * d + dd] - ze[((long long)b * T + t) * ldz + g * d + dd];
        acc = fmaf(diff, diff, acc);
    }
    __shared__ float red[32];
#pragma unroll
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        float v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
#pragma unroll
        for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (threadIdx.x == 0) loss[b] += v / ((float)T * d) / (float)groups;
    }
}

// ------------------------------------------------------------------------------------------------ code histogram
// EntropyCounter.update (scripts/metrics.py:37-51 of the reference): counts[(s, g)][code] += occurrences over (b, t).
// The reference builds a one-hot [B*T, 1024] tensor per (stream, group) and sums it; here one block per (stream,
// group) histograms its B*T codes in shared memory and adds the 1024 bins to the running fp32 counts.  Indices
// outside [0, ncodes) are skipped and latched (escb_poll_error), like ACodes.
static __global__ void __launch_bounds__(256)
code_histogram_kernel(const long long* __restrict__ codes, const int B, const int S, const int G, const int T,
                      const int ncodes, float* __restrict__ counts, int* __restrict__ bad) {
    extern __shared__ unsigned hist[];
    const int sg = blockIdx.x, s = sg / G, g = sg - s * G;
    for (int i = threadIdx.x; i < ncodes; i += blockDim.x) hist[i] = 0u;
    __syncthreads();
    const long long n = (long long)B * T;
    for (long long i = threadIdx.x; i < n; i += blockDim.x) {
        const long long b = i / T;
        const int t = (int)(i - b * T);
        const long long c = codes[((b * S + s) * G + g) * (long long)T + t];
        if (c < 0 || c >= ncodes) { if (bad) *(volatile int*)bad = 1; continue; }
        atomicAdd(&hist[(int)c], 1u);
    }
    __syncthreads();
    float* out = counts + (long long)sg * ncodes;
    for (int i = threadIdx.x; i < ncodes; i += blockDim.x) out[i] += (float)hist[i];
}

// ------------------------------------------------------------------------------------------------ fused RVQ stream step
// One cross-scale RVQ stream step in ONE launch (csrvq.py:23-60 + quantization.py:74-136 + codebook.py:20-55):
//   residual = enc - dec (pre_process gather) -> per-group down-projection -> L2-normalise -> distance to the 1024
//   normalised codewords, first minimum -> code -> raw codeword -> up-projection -> post_process scatter (+ dec).
// The three product-VQ groups are independent (own projection, own codebook, disjoint (o, c) thirds of the frame), so a
// CTA is (tile of FR VQ frames) x (one group): N = 1024 frames give 384 CTAs.  The residual third lives in shared memory
// (FR x K_g floats, row pitch K_g + 4), z / z_hat / the chosen codewords in a few hundred bytes; enc and dec are read
// once (dec a second time, from L2, for the final add) and the refined map is written once: the step is HBM-bound
// (8.6 flop/B over the six streams of ESC-Base).
// Arithmetic is that of the unfused kernels, operation for operation: every down-projection output is ONE sequential
// k = 0..K_g-1 FMA chain (gemm_tile's order), the normalisation and the distance (|z|^2 - (2z).c) + |c|^2 follow
// codebook_argmin_kernel, and ties go to the lowest index.  The distance loop runs two codes per FFMA2.
// Requires the groups to be equal thirds of every frequency run (QuantW::run > 0) and ld == C.
struct PvqStreamArgs {
    const float* E;             // encoder map of the scale   [B, Hq*W, C]
    const float* Dm;            // decoder state (null: stream 0 quantizes enc itself)
    float* out;                 // dec_refine = vq.decode(code) + dec (null: codes only; may alias Dm)
    long long* codes;           // [B, S, 3, T] (this stream's slice: + s*3*T)
    long long cstride;          // batch stride of codes (S*3*T)
    float* ze;                  // optional [rows][ldz] projected vectors (the eval-mode VQ loss reads them)
    int ldz;
    const float* wd[3];         // per group Wt [Kg_pad][ldwd]
    int ldwd;
    const float* wu;            // up-projection Wt [3d][ldwu], rows g*d + j, columns in frame order (h, o, c)
    int ldwu;
    const float* cbt;           // [3][d][ncodes] normalised, transposed
    const float* cnorm;         // [3][ncodes]
    const float* raw;           // [3][ncodes][d]
    int* bad;                   // host-mapped latch for out-of-range codes given to the decode-only form
    int l2norm;                 // 1: both sides L2-normalised (the shipped configs); 0: plain squared distance
    int ncodes, Hq, W, C, run, Kg, T;
    long long rows;             // B * T
    FastDiv drun4;              // by run / 4 (float4 groups per frequency run)
    int lgH;                    // log2(Hq) when Hq is a power of two, else -1
    int decode_only;            // codes are given (ProductVectorQuantize.decode + post_fuse): steps 5 and 6 only
};
constexpr int kPvqKC = 64;      // k rows of the down-projection staged in shared memory per chunk
constexpr int kPvqStages = 4;   // chunks in flight (cp.async ring): one chunk is 0.25 us of FMA chain, an L2 / DRAM round trip 0.5 - 1.5 us

template <int D, int FR>
__global__ void __launch_bounds__(256, 3)
pvq_stream_kernel(const PvqStreamArgs a) {
    constexpr int LDW = (D + 3) & ~3;                           // row pitch of the packed down-projection Wt [k][LDW]
    constexpr int WL4 = (kPvqKC * LDW / 4 + 255) / 256;         // float4 loads per thread and weight chunk
    constexpr int SUB = FR < 8 ? FR : 8;                        // frames per pass of the distance loop (register tile)
    extern __shared__ __align__(16) float rs[];                 // [FR][Kg + 4], then the weight ring [kPvqStages][kPvqKC][LDW]
    __shared__ __align__(16) float zs[FR][D];                   // z
    __shared__ __align__(16) unsigned long long z2s[FR][D];     // (2 z_hat, 2 z_hat) pairs: the FFMA2 operand of the distance loop
    __shared__ float zzs[FR];
    __shared__ __align__(16) float es[FR][D];                   // chosen raw codewords
    __shared__ float redv[FR][8];
    __shared__ int redi[FR][8];
    __shared__ int best_code[FR];
    __shared__ long long fbase[FR];                             // element offset of (frame, h = 0) + this group's third
    const int tid = threadIdx.x, g = blockIdx.y;
    const long long row0 = (long long)blockIdx.x * FR;
    const int Kg = a.Kg, pitch = Kg + 4, run = a.run, goff = g * run;
    const int nval = a.rows - row0 >= FR ? FR : (int)(a.rows - row0);
    const long long hstride = (long long)a.W * a.C;
    if (tid < FR) {
        const long long m = row0 + (tid < nval ? tid : 0), b = m / a.T;
        const int t = (int)(m - b * a.T);
        fbase[tid] = ((long long)b * a.Hq * a.W + 2 * t) * a.C + goff;
        if (a.decode_only) {
            int code = 0;
            if (tid < nval) {
                const long long cv = a.codes[b * a.cstride + (long long)g * a.T + t];
                code = (cv < 0 || cv >= a.ncodes) ? 0 : (int)cv;
                if (cv != code && a.bad) *(volatile int*)a.bad = 1;      // caller data: clamp and latch (escb_poll_error)
            }
            best_code[tid] = code;
        }
    }
    __syncthreads();
    if (!a.decode_only) {
        // ---- 1. residual third of FR frames -> shared memory (runs of `run` contiguous floats per frequency row).
        // Item = one float4 of one (frame, frequency row) run; four items per thread are loaded before any is consumed.
        {
            const int run4 = run >> 2, nitem = FR * a.Hq * run4;
            auto locate = [&](int idx, int& f, int& k, long long& goffs) {   // item -> frame, k, offset inside the frame
                const int row = (int)a.drun4.div((unsigned)idx), x4 = idx - row * run4;
                int h;
                if (a.lgH >= 0) { f = row >> a.lgH; h = row & (a.Hq - 1); }
                else { f = row / a.Hq; h = row - f * a.Hq; }
                k = h * run + 4 * x4;
                goffs = h * hstride + 4 * x4;
            };
            for (int base = tid; base < nitem; base += 4 * 256) {
                float4 e[4], d[4];
                int f[4], k[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int idx = base + 256 * u;
                    e[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                    d[u] = e[u];
                    f[u] = -1;
                    if (idx < nitem) {
                        long long goffs;
                        locate(idx, f[u], k[u], goffs);
                        if (f[u] < nval) {
                            const long long off = fbase[f[u]] + goffs;
                            e[u] = ldg4(a.E + off);
                            if (a.Dm) d[u] = ldg4(a.Dm + off);
                        }
                    }
                }
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (f[u] >= 0)
                        *reinterpret_cast<float4*>(rs + f[u] * pitch + k[u]) =
                            make_float4(e[u].x - d[u].x, e[u].y - d[u].y, e[u].z - d[u].z, e[u].w - d[u].w);
            }
        }
        // ---- 2. down-projection: one sequential FMA chain per (frame, component), the order of gemm_tile.  The weight
        // streams through a cp.async ring of kPvqKC-row chunks in shared memory (all 256 threads copy).
        {
            float* wsm = rs + FR * pitch;
            const float* wg = a.wd[g];
            // thread = (group of PF frames, component j): PF independent chains share every weight read, and the residual
            // is read four k at a time (1.5 instructions per multiply-add instead of 3; the chains fill the FMA latency)
            constexpr int PF = 4;
            static_assert(FR % PF == 0, "frames per CTA must be a multiple of 4");
            const bool active = tid < (FR / PF) * D;
            const int fg = active ? tid / D : 0, oj = active ? tid - (tid / D) * D : 0;
            float acc[PF];
#pragma unroll
            for (int i = 0; i < PF; ++i) acc[i] = 0.f;
            const int nchunks = Kg / kPvqKC;
            auto issue = [&](int ch) {                         // chunk ch -> ring slot ch % kPvqStages (16-byte cp.async, all threads)
                if (ch < nchunks) {
                    const float* src = wg + (long long)ch * kPvqKC * LDW;
                    float* dst = wsm + (ch % kPvqStages) * kPvqKC * LDW;
#pragma unroll
                    for (int i = 0; i < WL4; ++i) {
                        const int idx = tid + 256 * i;
                        if (idx < kPvqKC * LDW / 4)
                            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(dst + 4 * idx)), "l"(src + 4 * idx) : "memory");
                    }
                }
                asm volatile("cp.async.commit_group;" ::: "memory");
            };
#pragma unroll
            for (int st = 0; st < kPvqStages - 1; ++st) issue(st);
            for (int ch = 0; ch < nchunks; ++ch) {
                asm volatile("cp.async.wait_group %0;" ::"n"(kPvqStages - 2) : "memory");
                __syncthreads();                               // chunk ch has landed for everyone; slot (ch - 1) % stages is free
                issue(ch + kPvqStages - 1);
                const float* wc = wsm + (ch % kPvqStages) * kPvqKC * LDW + oj;
                if (active) {
                    const float* r = rs + (fg * PF) * pitch + ch * kPvqKC;
#pragma unroll 4
                    for (int kk = 0; kk < kPvqKC; kk += 4) {
                        float4 rv[PF];
#pragma unroll
                        for (int i = 0; i < PF; ++i) rv[i] = *reinterpret_cast<const float4*>(r + i * pitch + kk);
                        const float w0 = wc[kk * LDW], w1 = wc[(kk + 1) * LDW], w2 = wc[(kk + 2) * LDW], w3 = wc[(kk + 3) * LDW];
#pragma unroll
                        for (int i = 0; i < PF; ++i) {
                            acc[i] = fmaf(rv[i].x, w0, acc[i]);
                            acc[i] = fmaf(rv[i].y, w1, acc[i]);
                            acc[i] = fmaf(rv[i].z, w2, acc[i]);
                            acc[i] = fmaf(rv[i].w, w3, acc[i]);
                        }
                    }
                }
            }
            if (active) {
#pragma unroll
                for (int i = 0; i < PF; ++i) {
                    const int f = fg * PF + i;
                    zs[f][oj] = acc[i];
                    if (a.ze && f < nval) a.ze[(row0 + f) * (long long)a.ldz + g * D + oj] = acc[i];
                }
            }
        }
        __syncthreads();
        // ---- 3. L2 normalisation (codebook.py:32-33), one thread per frame
        if (tid < FR) {
            float zn[D];
            float ss = 0.f;
#pragma unroll
            for (int k = 0; k < D; ++k) { zn[k] = zs[tid][k]; ss = fmaf(zn[k], zn[k], ss); }
            const float denom = a.l2norm ? fmaxf(sqrtf(ss), 1e-12f) : 1.0f;
            float zz = 0.f;
#pragma unroll
            for (int k = 0; k < D; ++k) {
                const float v = zn[k] / denom;
                zz = fmaf(v, v, zz);
                z2s[tid][k] = pack2(2.0f * v, 2.0f * v);
            }
            zzs[tid] = zz;
        }
        __syncthreads();
        // ---- 4. distances to all codewords: 4 codes per thread, SUB frames per pass, two codes per FFMA2
        const float* cb = a.cbt + (long long)g * D * a.ncodes;
        const float* cn = a.cnorm + (long long)g * a.ncodes;
        const int lane = tid & 31, wid = tid >> 5;
        for (int f0 = 0; f0 < FR; f0 += SUB) {
            float best[SUB];
            int besti[SUB];
#pragma unroll
            for (int f = 0; f < SUB; ++f) { best[f] = 3.0e38f; besti[f] = 0x7fffffff; }
            for (int c0 = tid * 4; c0 < a.ncodes; c0 += 1024) {
                unsigned long long acc[SUB][2];
#pragma unroll
                for (int f = 0; f < SUB; ++f) { acc[f][0] = 0ull; acc[f][1] = 0ull; }
#pragma unroll
                for (int k = 0; k < D; ++k) {
                    const float4 c4 = __ldg(reinterpret_cast<const float4*>(cb + (long long)k * a.ncodes + c0));
                    const unsigned long long c01 = pack2(c4.x, c4.y), c23 = pack2(c4.z, c4.w);
#pragma unroll
                    for (int f = 0; f < SUB; ++f) {
                        const unsigned long long z2 = z2s[f0 + f][k];
                        acc[f][0] = ffma2(z2, c01, acc[f][0]);
                        acc[f][1] = ffma2(z2, c23, acc[f][1]);
                    }
                }
                const float4 cn4 = __ldg(reinterpret_cast<const float4*>(cn + c0));
                const float cnj[4] = {cn4.x, cn4.y, cn4.z, cn4.w};
#pragma unroll
                for (int f = 0; f < SUB; ++f) {
                    float d4[4];
                    unpack2(acc[f][0], d4[0], d4[1]);
                    unpack2(acc[f][1], d4[2], d4[3]);
                    const float zz = zzs[f0 + f];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float dist = (zz - d4[j]) + cnj[j];
                        if (dist < best[f]) { best[f] = dist; besti[f] = c0 + j; }
                    }
                }
            }
#pragma unroll
            for (int f = 0; f < SUB; ++f) {
                float bv = best[f];
                int bi = besti[f];
#pragma unroll
                for (int o = 16; o; o >>= 1) {
                    const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
                    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                    if (ov < bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
                }
                if (lane == 0) { redv[f0 + f][wid] = bv; redi[f0 + f][wid] = bi; }
            }
        }
        __syncthreads();
        if (tid < FR) {
            float bv = redv[tid][0];
            int bi = redi[tid][0];
#pragma unroll
            for (int w = 1; w < 8; ++w) {
                const float ov = redv[tid][w];
                const int oi = redi[tid][w];
                if (ov < bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
            }
            if (bi == 0x7fffffff) bi = 0;                       // all-NaN row
            best_code[tid] = bi;
            if (tid < nval) {
                const long long m = row0 + tid, b = m / a.T;
                const int t = (int)(m - b * a.T);
                a.codes[b * a.cstride + (long long)g * a.T + t] = bi;
            }
        }
        if (!a.out) return;
        __syncthreads();
    }
    // ---- 5. de-quantise from the RAW table (codebook.py:45-55)
    for (int o = tid; o < FR * D; o += 256) {
        const int f = o / D, j = o - f * D;
        es[f][j] = __ldg(a.raw + ((long long)g * a.ncodes + best_code[f]) * D + j);
    }
    __syncthreads();
    // ---- 6. up-projection + post_process scatter + post_fuse: thread = column of this group's third
    for (int kg = tid; kg < Kg; kg += 256) {
        const int h = kg / run, x = kg - h * run;
        const float* u = a.wu + (long long)(g * D) * a.ldwu + (long long)h * 2 * a.C + goff + x;
        float uj[D];
#pragma unroll
        for (int j = 0; j < D; ++j) uj[j] = __ldg(u + (long long)j * a.ldwu);
        const long long col = h * hstride + x;
#pragma unroll 2
        for (int f = 0; f < nval; ++f) {
            float acc = 0.f;
            if constexpr (D % 4 == 0) {
#pragma unroll
                for (int j = 0; j < D; j += 4) {
                    const float4 e4 = *reinterpret_cast<const float4*>(&es[f][j]);
                    acc = fmaf(e4.x, uj[j], acc); acc = fmaf(e4.y, uj[j + 1], acc);
                    acc = fmaf(e4.z, uj[j + 2], acc); acc = fmaf(e4.w, uj[j + 3], acc);
                }
            } else {
#pragma unroll
                for (int j = 0; j < D; j += 2) {
                    const float2 e2 = *reinterpret_cast<const float2*>(&es[f][j]);
                    acc = fmaf(e2.x, uj[j], acc); acc = fmaf(e2.y, uj[j + 1], acc);
                }
            }
            const long long off = fbase[f] + col;
            a.out[off] = a.Dm ? acc + a.Dm[off] : acc;
        }
    }
}

// ------------------------------------------------------------------------------------------------ RVQCodecs residual chain
// ResidualVectorQuantize.quantize_to_code / residual_vector_quantization in eval mode (quantization.py:170-195, 223-237)
// on the down-projected vectors of one group: for stream i = 0..S-1  code_i = argmin_i(residual)  (Codebook.quantize_to_code:
// both sides L2-normalised, first minimum), residual -= raw_i[code_i].  One block = 32 rows x one group, the whole chain
// stays in shared memory; the distance tile is codebook_argmin_kernel's.  Optionally emits z_q = sum_i raw_i[code_i]
// (accumulated in stream order from 0, like `z_q = z_q + z_q_i`) and the eval-mode loss numerator
// sum_i sum_j (raw_i[code_i][j] - residual_i[j])^2 per row.
template <int D>
__global__ void __launch_bounds__(256)
rvq_chain_kernel(const float* __restrict__ ze, const int ldz, const float* __restrict__ cbt, const float* __restrict__ cnorm,
                 const float* __restrict__ raw, const int ncodes, const int Stot, const int S, const long long rows,
                 long long* __restrict__ codes, const int T, float* __restrict__ zq, float* __restrict__ se, const int l2norm) {
    __shared__ float res[kArgminRows][D + 1];                  // running residual
    __shared__ float zqa[kArgminRows][D + 1];
    __shared__ float sea[kArgminRows];
    __shared__ __align__(16) float zs[D][kArgminRows];         // 2 * normalised residual, transposed
    __shared__ float zzs[kArgminRows];
    __shared__ __align__(16) float cbs[D][kArgminChunk];
    __shared__ float cns[kArgminChunk];
    __shared__ float redv[kArgminRows][2];
    __shared__ int redi[kArgminRows][2];
    const int tid = threadIdx.x, g = blockIdx.y;
    const long long row0 = (long long)blockIdx.x * kArgminRows;
    for (int i = tid; i < kArgminRows * D; i += 256) {
        const int r = i / D, k = i - r * D;
        const long long m = row0 + r;
        res[r][k] = m < rows ? __ldg(ze + m * (long long)ldz + g * D + k) : 0.f;
        zqa[r][k] = 0.f;
    }
    if (tid < kArgminRows) sea[tid] = 0.f;
    __syncthreads();
    const int cg = tid & 63, rg = tid >> 6;
    const int lane = tid & 31, wsel = (tid >> 5) & 1;
    for (int si = 0; si < S; ++si) {
        if (tid < kArgminRows) {
            float zn[D];
            float ss = 0.f;
#pragma unroll
            for (int k = 0; k < D; ++k) { zn[k] = res[tid][k]; ss = fmaf(zn[k], zn[k], ss); }
            const float denom = l2norm ? fmaxf(sqrtf(ss), 1e-12f) : 1.0f;
            float zz = 0.f;
#pragma unroll
            for (int k = 0; k < D; ++k) {
                const float v = zn[k] / denom;
                zz = fmaf(v, v, zz);
                zs[k][tid] = 2.0f * v;
            }
            zzs[tid] = zz;
        }
        float best[8];
        int besti[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) { best[r] = 3.0e38f; besti[r] = 0x7fffffff; }
        const float* cb = cbt + ((long long)g * Stot + si) * D * ncodes;
        const float* cn = cnorm + ((long long)g * Stot + si) * ncodes;
        for (int chunk = 0; chunk < ncodes; chunk += kArgminChunk) {
            __syncthreads();
            for (int i = tid; i < D * (kArgminChunk / 4); i += 256) {
                const int k = i / (kArgminChunk / 4), c4 = (i % (kArgminChunk / 4)) * 4;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (chunk + c4 + 3 < ncodes) v = __ldg(reinterpret_cast<const float4*>(cb + (long long)k * ncodes + chunk + c4));
                *reinterpret_cast<float4*>(&cbs[k][c4]) = v;
            }
            cns[tid] = (chunk + tid < ncodes) ? __ldg(cn + chunk + tid) : 3.0e38f;
            __syncthreads();
            float acc[8][4];
#pragma unroll
            for (int r = 0; r < 8; ++r)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[r][j] = 0.f;
#pragma unroll
            for (int k = 0; k < D; ++k) {
                const float4 c = *reinterpret_cast<const float4*>(&cbs[k][cg * 4]);
                const float4 z0 = *reinterpret_cast<const float4*>(&zs[k][rg * 8]);
                const float4 z1 = *reinterpret_cast<const float4*>(&zs[k][rg * 8 + 4]);
                const float zr[8] = {z0.x, z0.y, z0.z, z0.w, z1.x, z1.y, z1.z, z1.w};
                const float cj[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
                for (int r = 0; r < 8; ++r)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[r][j] = fmaf(zr[r], cj[j], acc[r][j]);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int code = chunk + cg * 4 + j;
                const float cnj = cns[cg * 4 + j];
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    const float dist = (zzs[rg * 8 + r] - acc[r][j]) + cnj;
                    if (dist < best[r]) { best[r] = dist; besti[r] = code; }
                }
            }
        }
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            float bv = best[r];
            int bi = besti[r];
#pragma unroll
            for (int o = 16; o; o >>= 1) {
                const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                if (ov < bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
            }
            if (lane == 0) { redv[rg * 8 + r][wsel] = bv; redi[rg * 8 + r][wsel] = bi; }
        }
        __syncthreads();
        if (tid < kArgminRows) {
            float bv = redv[tid][0];
            int bi = redi[tid][0];
            const float ov = redv[tid][1];
            const int oi = redi[tid][1];
            if (ov < bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
            if (bi == 0x7fffffff) bi = 0;
            const long long m = row0 + tid;
            if (m < rows) {
                const long long b = m / T;
                const int t = (int)(m - b * T);
                codes[((b * S + si) * 3 + g) * (long long)T + t] = bi;
            }
            const float* e = raw + (((long long)g * Stot + si) * ncodes + bi) * D;
            float sq = sea[tid];
#pragma unroll
            for (int k = 0; k < D; ++k) {
                const float ev = __ldg(e + k), rv = res[tid][k];
                const float d = ev - rv;
                sq = fmaf(d, d, sq);
                zqa[tid][k] = zqa[tid][k] + ev;
                res[tid][k] = rv - ev;
            }
            sea[tid] = sq;
        }
        __syncthreads();
    }
    if (zq)
        for (int i = tid; i < kArgminRows * D; i += 256) {
            const int r = i / D, k = i - r * D;
            if (row0 + r < rows) zq[(row0 + r) * (long long)ldz + g * D + k] = zqa[r][k];
        }
    if (se && tid < kArgminRows && row0 + tid < rows) se[(row0 + tid) * 3 + g] = sea[tid];
}

// RVQCodecs decode: z_q[row][g*d + j] = sum_i raw_i[codes[b, i, g, t]][j], in stream order from 0 (dequantize_code,
// quantization.py:239-245).  Out-of-range indices are clamped and latched like ACodes.
static __global__ void rvq_gather_kernel(const long long* __restrict__ codes, const float* __restrict__ raw, const int ncodes,
                                         const int Stot, const int S, const int d, const long long rows, const int T,
                                         float* __restrict__ zq, const int ldz, int* __restrict__ bad) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= rows * 3 * d) return;
    const int j = (int)(idx % d);
    const int g = (int)((idx / d) % 3);
    const long long m = idx / (3 * d), b = m / T;
    const int t = (int)(m - b * T);
    float acc = 0.f;
    for (int si = 0; si < S; ++si) {
        long long c = codes[((b * S + si) * 3 + g) * (long long)T + t];
        if (c < 0 || c >= ncodes) { if (bad) *(volatile int*)bad = 1; c = 0; }
        acc = acc + __ldg(raw + (((long long)g * Stot + si) * ncodes + c) * d + j);
    }
    zq[m * ldz + g * d + j] = acc;
}

// eval-mode VQ loss of RVQCodecs: loss[b] = sum_{t, g} se[b, t, g] / (T * d) / 3 (Codebook.forward mse over (T, d) per
// stream, summed over streams, averaged over the 3 groups: quantization.py:184-186, 343-346), one block per clip
static __global__ void rvq_loss_kernel(const float* __restrict__ se, const int T, const int d, float* __restrict__ loss) {
    const int b = blockIdx.x;
    float acc = 0.f;
    for (int i = threadIdx.x; i < T * 3; i += blockDim.x) acc += se[(long long)b * T * 3 + i];
    __shared__ float red[32];
#pragma unroll
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        float v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
#pragma unroll
        for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (threadIdx.x == 0) loss[b] = v / ((float)T * d) / 3.0f;
    }
}

}  // namespace escb
