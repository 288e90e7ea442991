// A-loaders (how a logical GEMM row is gathered from HBM) and epilogues (where outputs land) for gemm.cuh.
// Each one replaces a permute/reshape/contiguous chain of the reference; the cited lines are the behaviour
// restated as index arithmetic.
#pragma once
#include "gemm.cuh"

namespace escb {

// Division by a launch-invariant divisor as multiply-high + shift: exact for every n < 2^31 with
// magic = ceil(2^(31 + s) / d), s = ceil(log2 d) (d = 1 is flagged by magic = 0).  The window index arithmetic below
// runs once per row per tile in the A producers and epilogues, where a hardware-emulated 32-bit division (~25
// instructions) was a third of the producers' time.
struct FastDiv {
    unsigned d, magic, shift;
    __host__ __device__ static FastDiv make(unsigned d) {
        FastDiv f{d, 0u, 0u};
        if (d > 1) {
            unsigned s = 0;
            while ((1ull << s) < d) ++s;
            f.magic = (unsigned)(((1ull << (31 + s)) + d - 1) / d);
            f.shift = s - 1;
        }
        return f;
    }
    __device__ __forceinline__ unsigned div(unsigned n) const { return magic ? __umulhi(n, magic) >> shift : n; }
};

// =============================================================================================== A loaders
// Contract: init(m, M, row) fills the per-row context; valid(row) == false means the whole row is zero
// (out-of-range row or a zero-padded token); load1(row,k) returns element k (< K); load4(row,k,K) returns
// elements k..k+3 with those >= K zeroed (k % 4 == 0, k < K).  load4_raw(row,k,K) is the same load WITHOUT any
// instruction that consumes the loaded registers (kRawMask == true: the caller zeroes the elements >= K itself, later):
// the tcgen05 producers issue it several K blocks ahead, and a select on the loaded value would park the warp on the
// load right at the issue point (it did: mask4 inside load4 cost the producers half their time).

// Dense rows X[m*ld + k].
struct ARows {
    const float* X;
    int ld;
    struct Row { const float* p; };
    __device__ __forceinline__ void init(long long m, long long M, Row& r) const {
        r.p = (m < M) ? X + m * (long long)ld : nullptr;
    }
    __device__ __forceinline__ bool valid(const Row& r) const { return r.p != nullptr; }
    __device__ __forceinline__ float load1(const Row& r, int k) const { return __ldg(r.p + k); }
    __device__ __forceinline__ float4 load4(const Row& r, int k, int K) const { return mask4(ldg4(r.p + k), k, K); }
    static constexpr bool kRawMask = true;
    __device__ __forceinline__ float4 load4_raw(const Row& r, int k, int) const { return ldg4(r.p + k); }
    __device__ __forceinline__ void prefetch(const Row& r, int K, int part) const { prefetch_row(r.p, K, part); }
};

// Window partition of a [B,H,W,C] token map zero-padded to (Hp,Wp) and cyclically shifted by `shift`
// (attention.py:137-153, 246-250).  Row m = ((b*nWh + wh)*nWw + ww)*16 + (i*4 + j) reads the token at
// h = (4*wh + i + shift) % Hp, w = (4*ww + j + shift) % Wp; tokens in the padding are all-zero rows
// (the reference pads AFTER norm1, so they bypass LayerNorm).
struct WindowGeom {
    int H, W, Hp, Wp, shift, nWw, nW;   // nW = (Hp/4)*(Wp/4)
    FastDiv dW, dWw;                    // by nW, by nWw
    __device__ __forceinline__ long long token(long long m) const {   // m < 2^31 (checked by the launchers)
        const unsigned mm = (unsigned)m;
        const int t = (int)(mm & 15u);
        const unsigned wi = mm >> 4;
        const unsigned b = dW.div(wi);
        const unsigned win = wi - b * (unsigned)nW;
        const int wh = (int)dWw.div(win), ww = (int)(win - (unsigned)wh * (unsigned)nWw);
        int h = wh * 4 + (t >> 2) + shift, w = ww * 4 + (t & 3) + shift;
        if (h >= Hp) h -= Hp;
        if (w >= Wp) w -= Wp;
        if (h >= H || w >= W) return -1;
        return ((long long)b * H + h) * (long long)W + w;
    }
};

struct AWindow {
    const float* X;
    int ld;
    WindowGeom g;
    struct Row { const float* p; };
    __device__ __forceinline__ void init(long long m, long long M, Row& r) const {
        r.p = nullptr;
        if (m < M) {
            const long long t = g.token(m);
            if (t >= 0) r.p = X + t * (long long)ld;
        }
    }
    __device__ __forceinline__ bool valid(const Row& r) const { return r.p != nullptr; }
    __device__ __forceinline__ float load1(const Row& r, int k) const { return __ldg(r.p + k); }
    __device__ __forceinline__ float4 load4(const Row& r, int k, int K) const { return mask4(ldg4(r.p + k), k, K); }
    static constexpr bool kRawMask = true;
    __device__ __forceinline__ float4 load4_raw(const Row& r, int k, int) const { return ldg4(r.p + k); }
    __device__ __forceinline__ void prefetch(const Row& r, int K, int part) const { prefetch_row(r.p, K, part); }
};

// PatchMerge gather (scale.py:7-14,104-112): row m = (b, h2, w) is [ x[b,2*h2,w,:] ; x[b,2*h2+1,w,:] ], K = 2C.
struct AMerge {
    const float* X;
    int ld, H, W, C;
    struct Row { const float* p0; const float* p1; };
    __device__ __forceinline__ void init(long long m, long long M, Row& r) const {
        r.p0 = r.p1 = nullptr;
        if (m < M) {
            const unsigned H2 = (unsigned)H >> 1, mm = (unsigned)m;
            const unsigned bh = mm / (unsigned)W, w = mm - bh * (unsigned)W;
            const unsigned b = bh / H2, h2 = bh - b * H2;
            r.p0 = X + (((long long)b * H + 2 * h2) * (long long)W + w) * ld;
            r.p1 = r.p0 + (long long)W * ld;
        }
    }
    __device__ __forceinline__ bool valid(const Row& r) const { return r.p0 != nullptr; }
    __device__ __forceinline__ float load1(const Row& r, int k) const { return k < C ? __ldg(r.p0 + k) : __ldg(r.p1 + (k - C)); }
    __device__ __forceinline__ float4 load4(const Row& r, int k, int K) const {
        if ((C & 3) == 0) return k < C ? ldg4(r.p0 + k) : ldg4(r.p1 + (k - C));
        float4 v;
        v.x = load1(r, k);
        v.y = (k + 1 < K) ? load1(r, k + 1) : 0.f;
        v.z = (k + 2 < K) ? load1(r, k + 2) : 0.f;
        v.w = (k + 3 < K) ? load1(r, k + 3) : 0.f;
        return v;
    }
    static constexpr bool kRawMask = false;
    __device__ __forceinline__ float4 load4_raw(const Row& r, int k, int K) const { return load4(r, k, K); }
    __device__ __forceinline__ void prefetch(const Row& r, int K, int part) const {
        for (int k = (part >> 1) * 32; k < C; k += 64) prefetch_l2(((part & 1) ? r.p1 : r.p0) + k);
    }
};

// Product-VQ frame of the residual enc - dec (csrvq.py:15-17; quantization.py:400-409).  The reference's frame
// vector is ordered (o, c, h); the packed down-projection weight is permuted to k' = (h, o, c) so that a frame
// is H contiguous runs of 2C floats: x[b, h*W + 2t + o, c].  Requires ld == C.
struct AFrame {
    const float* E;
    const float* D;     // may be null (stream 0 quantizes enc itself)
    int Hq, W, C;
    struct Row { long long base; };
    __device__ __forceinline__ void init(long long m, long long M, Row& r) const {
        r.base = -1;
        if (m < M) {
            const unsigned T = (unsigned)W >> 1, mm = (unsigned)m;
            const unsigned b = mm / T, t = mm - b * T;
            r.base = ((long long)b * Hq * (long long)W + 2 * t) * C;
        }
    }
    __device__ __forceinline__ bool valid(const Row& r) const { return r.base >= 0; }
    __device__ __forceinline__ long long off(const Row& r, int k) const {
        const int h = k / (2 * C);
        return r.base + (long long)h * W * C + (k - h * 2 * C);
    }
    __device__ __forceinline__ float load1(const Row& r, int k) const {
        const long long o = off(r, k);
        return D ? __ldg(E + o) - __ldg(D + o) : __ldg(E + o);
    }
    __device__ __forceinline__ float4 load4(const Row& r, int k, int K) const {
        const long long o = off(r, k);
        float4 e = ldg4(E + o);
        if (D) {
            const float4 d = ldg4(D + o);
            e.x -= d.x; e.y -= d.y; e.z -= d.z; e.w -= d.w;
        }
        return e;
    }
    static constexpr bool kRawMask = false;
    __device__ __forceinline__ float4 load4_raw(const Row& r, int k, int K) const { return load4(r, k, K); }
    __device__ __forceinline__ void prefetch(const Row&, int, int) const {}
};

// One group's third of that frame: in the (h, o, c) order a group is the (o, c) sub-range [goff, goff + run) of each
// of the Hq runs (run = 2C/3), so row m, element k reads x[b, h*W + 2t, goff + k - h*run] with h = k / run.
struct AFrameG {
    const float* E;
    const float* D;     // may be null
    int Hq, W, C, run, goff, vec4;      // vec4: run and goff are multiples of 4 (one aligned 16-byte load per call)
    FastDiv drun;
    struct Row { long long base; };
    __device__ __forceinline__ void init(long long m, long long M, Row& r) const {
        r.base = -1;
        if (m < M) {
            const unsigned T = (unsigned)W >> 1, mm = (unsigned)m;
            const unsigned b = mm / T, t = mm - b * T;
            r.base = ((long long)b * Hq * (long long)W + 2 * t) * C + goff;
        }
    }
    __device__ __forceinline__ bool valid(const Row& r) const { return r.base >= 0; }
    __device__ __forceinline__ long long off(const Row& r, int k) const {
        const int h = (int)drun.div((unsigned)k);
        return r.base + (long long)h * W * C + (k - h * run);
    }
    __device__ __forceinline__ float load1(const Row& r, int k) const {
        const long long o = off(r, k);
        return D ? __ldg(E + o) - __ldg(D + o) : __ldg(E + o);
    }
    __device__ __forceinline__ float2 load2(const Row& r, int k) const {     // k even: never straddles a run
        const long long o = off(r, k);
        float2 e = __ldg(reinterpret_cast<const float2*>(E + o));
        if (D) {
            const float2 d = __ldg(reinterpret_cast<const float2*>(D + o));
            e.x -= d.x; e.y -= d.y;
        }
        return e;
    }
    __device__ __forceinline__ float4 load4(const Row& r, int k, int K) const {
        if (vec4) {
            const long long o = off(r, k);
            float4 e = ldg4(E + o);
            if (D) {
                const float4 d = ldg4(D + o);
                e.x -= d.x; e.y -= d.y; e.z -= d.z; e.w -= d.w;
            }
            return e;
        }
        const float2 a = load2(r, k);
        const float2 b = (k + 2 < K) ? load2(r, k + 2) : make_float2(0.f, 0.f);
        return make_float4(a.x, a.y, b.x, b.y);
    }
    static constexpr bool kRawMask = false;
    __device__ __forceinline__ float4 load4_raw(const Row& r, int k, int K) const { return load4(r, k, K); }
    __device__ __forceinline__ void prefetch(const Row&, int, int) const {}
};

// Rows of de-quantised codebook vectors [z_q0 ; z_q1 ; z_q2] gathered from the RAW tables by code
// (codebook.py:45-55; quantization.py:124-136).  codes layout [B, S, 3, T] int64, this stream at index s.
struct ACodes {
    const long long* codes;
    const float* tables;   // [3][K][d] raw
    int S, s, T, d, ncodes;
    int* bad;              // host-mapped error latch (escb_poll_error): set when a code index is outside [0, ncodes)
    struct Row { int c0, c1, c2; };
    // The codes are caller data (saved encoded_*.pth files): an index outside the table would be an out-of-bounds read
    // where the reference's F.embedding raises, so it is clamped to entry 0 and latched for the host to report.
    __device__ __forceinline__ int checked(long long c) const {
        if (c < 0 || c >= (long long)ncodes) {
            if (bad) *(volatile int*)bad = 1;
            return 0;
        }
        return (int)c;
    }
    __device__ __forceinline__ void init(long long m, long long M, Row& r) const {
        r.c0 = -1; r.c1 = r.c2 = 0;
        if (m < M) {
            const int t = (int)(m % T);
            const long long b = m / T;
            const long long* p = codes + ((b * S + s) * 3) * (long long)T + t;
            r.c0 = checked(p[0]); r.c1 = checked(p[T]); r.c2 = checked(p[2 * (long long)T]);
        }
    }
    __device__ __forceinline__ bool valid(const Row& r) const { return r.c0 >= 0; }
    __device__ __forceinline__ float load1(const Row& r, int k) const {
        const int g = k / d, dd = k - g * d;
        const int c = g == 0 ? r.c0 : (g == 1 ? r.c1 : r.c2);
        return __ldg(tables + ((long long)g * ncodes + c) * d + dd);
    }
    __device__ __forceinline__ float4 load4(const Row& r, int k, int K) const {
        float4 v;
        v.x = load1(r, k);
        v.y = (k + 1 < K) ? load1(r, k + 1) : 0.f;
        v.z = (k + 2 < K) ? load1(r, k + 2) : 0.f;
        v.w = (k + 3 < K) ? load1(r, k + 3) : 0.f;
        return v;
    }
    static constexpr bool kRawMask = false;
    __device__ __forceinline__ float4 load4_raw(const Row& r, int k, int K) const { return load4(r, k, K); }
    __device__ __forceinline__ void prefetch(const Row&, int, int) const {}
};

// im2col of the 5x5 / pad 2 convolution over channels-last tokens [B,H,W,ld] (scale.py:66-68,77).
// k = tap*ldc + c with tap = kh*5 + kw and ldc the padded channel count (48): the packed weight uses the same order.
struct AIm2col {
    const float* X;
    int ld, H, W, C;
    struct Row { const float* p; int h, w; };
    __device__ __forceinline__ void init(long long m, long long M, Row& r) const {
        r.p = nullptr; r.h = r.w = 0;
        if (m < M) {
            const unsigned mm = (unsigned)m, q = mm / (unsigned)W;
            r.w = (int)(mm - q * (unsigned)W);
            r.h = (int)(q % (unsigned)H);
            r.p = X + m * (long long)ld;
        }
    }
    __device__ __forceinline__ bool valid(const Row& r) const { return r.p != nullptr; }
    __device__ __forceinline__ float load1(const Row& r, int k) const {
        const float4 v = load4(r, k & ~3, 1 << 30);
        const int j = k & 3;
        return j == 0 ? v.x : (j == 1 ? v.y : (j == 2 ? v.z : v.w));
    }
    __device__ __forceinline__ float4 load4(const Row& r, int k, int K) const {
        const int tap = k / ld, c = k - tap * ld;
        const int dh = tap / 5 - 2, dw = tap % 5 - 2;
        const int hh = r.h + dh, ww = r.w + dw;
        if (hh < 0 || hh >= H || ww < 0 || ww >= W) return zero4();
        // every token row is read by 25 taps of ~25 neighbouring rows: let L1 keep it (ldg4 is the no-allocate stream load)
        return mask4(__ldg(reinterpret_cast<const float4*>(r.p + ((long long)dh * W + dw) * ld + c)), c, C);
    }
    static constexpr bool kRawMask = false;
    __device__ __forceinline__ float4 load4_raw(const Row& r, int k, int K) const { return load4(r, k, K); }
    __device__ __forceinline__ void prefetch(const Row&, int, int) const {}
};

// STFT framing with centre/reflect padding (base.py:22-24,36 -> torch.stft): row m = (b, t), element k is
// x[b, reflect(t*hop - win/2 + k)], k in [0, win).  (The window is applied by the packed DFT basis.)
struct AStftFrames {
    const float* X;
    long long L;
    int T, hop, half_win;
    struct Row { const float* p; long long i0; };
    __device__ __forceinline__ void init(long long m, long long M, Row& r) const {
        r.p = nullptr; r.i0 = 0;
        if (m < M) {
            const int t = (int)(m % T);
            r.p = X + (m / T) * L;
            r.i0 = (long long)t * hop - half_win;
        }
    }
    __device__ __forceinline__ bool valid(const Row& r) const { return r.p != nullptr; }
    __device__ __forceinline__ float load1(const Row& r, int k) const {
        long long i = r.i0 + k;
        if (i < 0) i = -i;
        if (i >= L) i = 2 * (L - 1) - i;
        return __ldg(r.p + i);
    }
    __device__ __forceinline__ float4 load4(const Row& r, int k, int K) const {
        float4 v;
        v.x = load1(r, k);
        v.y = (k + 1 < K) ? load1(r, k + 1) : 0.f;
        v.z = (k + 2 < K) ? load1(r, k + 2) : 0.f;
        v.w = (k + 3 < K) ? load1(r, k + 3) : 0.f;
        return v;
    }
    static constexpr bool kRawMask = false;
    __device__ __forceinline__ float4 load4_raw(const Row& r, int k, int K) const { return load4(r, k, K); }
    __device__ __forceinline__ void prefetch(const Row&, int, int) const {}
};

// Inverse STFT as one GEMM (base.py:25-27,46-47 -> torch.istft): output chunk j (hop samples) sums the
// `nov` = win/hop frames j-dt that overlap it: row m = (b, j), k = dt*F2 + cf reads Xf[b, j-dt, cf]
// (frame-major spectrum, F2 = 2*in_freq); frames outside [0,T) contribute zero.
struct AIstft {
    const float* Xf;
    int T, F2, j0, nchunks;
    struct Row { const float* p; int j; };
    __device__ __forceinline__ void init(long long m, long long M, Row& r) const {
        r.p = nullptr; r.j = 0;
        if (m < M) {
            r.j = (int)(m % nchunks) + j0;
            r.p = Xf + (m / nchunks) * (long long)T * F2;
        }
    }
    __device__ __forceinline__ bool valid(const Row& r) const { return r.p != nullptr; }
    __device__ __forceinline__ float load1(const Row& r, int k) const {
        const int dt = k / F2, cf = k - dt * F2;
        const int t = r.j - dt;
        return (t >= 0 && t < T) ? __ldg(r.p + (long long)t * F2 + cf) : 0.f;
    }
    __device__ __forceinline__ float4 load4(const Row& r, int k, int K) const {
        const int dt = k / F2, cf = k - dt * F2;
        const int t = r.j - dt;
        return (t >= 0 && t < T) ? ldg4(r.p + (long long)t * F2 + cf) : zero4();
    }
    static constexpr bool kRawMask = false;
    __device__ __forceinline__ float4 load4_raw(const Row& r, int k, int K) const { return load4(r, k, K); }
    __device__ __forceinline__ void prefetch(const Row&, int, int) const {}
};

// =============================================================================================== epilogues
// Contract: row(m, ctx) -> false skips the row; store(ctx, n, v) writes element n (< N).

template <bool GELU, bool RES>
struct EpiRows {   // Y[m*ldy + n] = act(v + bias[n]) (+ R[m*ldr + n])
    float* Y;
    const float* bias;
    const float* R;
    int ldy, ldr;
    struct Row { long long y, r; };   // element offsets (not pointers: keeps the accesses in the global window)
    __device__ __forceinline__ bool row(long long m, Row& c) const {
        c.y = m * (long long)ldy;
        c.r = RES ? m * (long long)ldr : 0;
        return true;
    }
    __device__ __forceinline__ void store(const Row& c, int n, float v) const {
        if (bias) v += __ldg(bias + n);
        if (GELU) v = gelu_erf(v);
        if (RES) v = R[c.r + n] + v;
        Y[c.y + n] = v;
    }
    __device__ __forceinline__ void store4(const Row& c, int n, float4 v) const {   // n % 4 == 0, ldy % 4 == 0
        const float4 b = bias4(n);
        v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
        store4_nb(c, n, v);
    }
    // tcgen05 epilogue: the bias of a column chunk is fetched once (bias4) and added by the caller
    __device__ __forceinline__ float4 bias4(int n) const { return bias ? ldg4(bias + n) : zero4(); }
    __device__ __forceinline__ void store4_nb(const Row& c, int n, float4 v) const { fin4(c, n, v, resid4(c, n)); }
    // split form: the residual is fetched early (resid4, before the accumulator is read) and applied by fin4
    __device__ __forceinline__ float4 resid4(const Row& c, int n) const {
        return RES ? *reinterpret_cast<const float4*>(R + c.r + n) : zero4();
    }
    __device__ __forceinline__ void fin4(const Row& c, int n, float4 v, const float4 r) const {
        if (GELU) { gelu_erf2(v.x, v.y); gelu_erf2(v.z, v.w); }   // (scalar tails in store() use erff: same function to ~1e-7)
        if (RES) { v.x = r.x + v.x; v.y = r.y + v.y; v.z = r.z + v.z; v.w = r.w + v.w; }
        // the MLP hidden map (4x the token map, read once by mlp2) is stored streaming so that it does not push the
        // token map - mlp2's residual - out of L2
        if (GELU) __stcs(reinterpret_cast<float4*>(Y + c.y + n), v);
        else *reinterpret_cast<float4*>(Y + c.y + n) = v;
    }
    __device__ __forceinline__ void prefetch(const Row& c, int N) const {
        if (RES) for (int k = 0; k < N; k += 32) prefetch_l2(R + c.r + k);
    }
};

// window reverse + reverse shift + crop + residual (attention.py:158-175): Y[token] = R[token] + (v + bias).
struct EpiWindow {
    float* Y;
    const float* R;
    const float* bias;
    int ld;
    WindowGeom g;
    struct Row { long long off; };
    __device__ __forceinline__ bool row(long long m, Row& c) const {
        const long long t = g.token(m);
        c.off = t * (long long)ld;
        return t >= 0;
    }
    __device__ __forceinline__ void store(const Row& c, int n, float v) const {
        Y[c.off + n] = R[c.off + n] + (v + __ldg(bias + n));
    }
    __device__ __forceinline__ void store4(const Row& c, int n, float4 v) const {
        const float4 b = ldg4(bias + n);
        const float4 r = *reinterpret_cast<const float4*>(R + c.off + n);
        v.x = r.x + (v.x + b.x); v.y = r.y + (v.y + b.y); v.z = r.z + (v.z + b.z); v.w = r.w + (v.w + b.w);
        *reinterpret_cast<float4*>(Y + c.off + n) = v;
    }
    __device__ __forceinline__ float4 bias4(int n) const { return ldg4(bias + n); }
    __device__ __forceinline__ void store4_nb(const Row& c, int n, float4 v) const { fin4(c, n, v, resid4(c, n)); }
    __device__ __forceinline__ float4 resid4(const Row& c, int n) const { return *reinterpret_cast<const float4*>(R + c.off + n); }
    __device__ __forceinline__ void fin4(const Row& c, int n, float4 v, const float4 r) const {
        v.x = r.x + v.x; v.y = r.y + v.y; v.z = r.z + v.z; v.w = r.w + v.w;
        *reinterpret_cast<float4*>(Y + c.off + n) = v;
    }
    __device__ __forceinline__ void prefetch(const Row& c, int N) const {
        for (int k = 0; k < N; k += 32) prefetch_l2(R + c.off + k);
    }
};

// Fused window-attention epilogue of the tcgen05 engine (tc_gemm.cuh): the qkv accumulator never leaves the SM.
// Column layout of the GEMM is [head slot][q | k | v][HDP] in sub-tiles of kAttnBN columns = HPB whole heads; the
// engine's epilogue warps compute softmax(q k^T * scale + relbias + mask) v per (window, head) and write
// out[row][h*HD + d] (rows in window order, like window_attn_kernel).  WindowAttention.forward attention.py:222-241.
constexpr int kAttnBN = 144;
template <int HD_, int HDP_>
struct EpiAttn {
    static constexpr bool kAttn = true;
    static constexpr int HD = HD_, HDP = HDP_, HPB = kAttnBN / (3 * HDP_);
    static_assert(HPB * 3 * HDP_ == kAttnBN, "head width must divide the sub-tile");
    float* out;
    int ldo;
    const float* bias;      // [slots][3][HDP], zero in the padding
    const float* relbias;   // [heads][16][16]
    int heads;
    float scale;
    int masked, nW, nWw, Hp, Wp;
    FastDiv dW, dWw;
    struct Row { int unused; };
};

// PatchSplit pixel shuffle (scale.py:16-23,142-144): row m = (b,h,w); n < Co goes to freq row 2h, the rest to 2h+1.
struct EpiSplit {
    float* Y;
    int ldy, H, W, Co;
    struct Row { long long y0; };
    __device__ __forceinline__ bool row(long long m, Row& c) const {
        const unsigned mm = (unsigned)m, bh = mm / (unsigned)W, w = mm - bh * (unsigned)W;     // bh = b*H + h
        c.y0 = ((2LL * bh) * (long long)W + w) * ldy;
        return true;
    }
    __device__ __forceinline__ void store(const Row& c, int n, float v) const {
        if (n < Co) Y[c.y0 + n] = v;
        else Y[c.y0 + (long long)W * ldy + (n - Co)] = v;
    }
    __device__ __forceinline__ void store4(const Row& c, int n, float4 v) const {
        if ((Co & 3) == 0) {
            float* p = n < Co ? Y + c.y0 + n : Y + c.y0 + (long long)W * ldy + (n - Co);
            *reinterpret_cast<float4*>(p) = v;
        } else {
            store(c, n, v.x); store(c, n + 1, v.y); store(c, n + 2, v.z); store(c, n + 3, v.w);
        }
    }
    __device__ __forceinline__ float4 bias4(int) const { return zero4(); }
    __device__ __forceinline__ void store4_nb(const Row& c, int n, float4 v) const { store4(c, n, v); }
    __device__ __forceinline__ float4 resid4(const Row&, int) const { return zero4(); }
    __device__ __forceinline__ void fin4(const Row& c, int n, float4 v, const float4) const { store4(c, n, v); }
    __device__ __forceinline__ void prefetch(const Row&, int) const {}
};

// product-VQ post_process + post_fuse (quantization.py:411-432; csrvq.py:19-21): column n' = (h, o, c) of frame
// (b, t) lands on token (h*W + 2t + o), channel c; out = v + dec.
struct EpiFrame {
    float* Y;
    const float* D;   // may be null
    int Hq, W, C;
    struct Row { long long base; };
    __device__ __forceinline__ bool row(long long m, Row& c) const {
        const unsigned T = (unsigned)W >> 1, mm = (unsigned)m;
        const unsigned b = mm / T, t = mm - b * T;
        c.base = ((long long)b * Hq * (long long)W + 2 * t) * C;
        return true;
    }
    __device__ __forceinline__ void store(const Row& c, int n, float v) const {
        const int h = n / (2 * C);
        const long long o = c.base + (long long)h * W * C + (n - h * 2 * C);
        Y[o] = D ? v + D[o] : v;
    }
    __device__ __forceinline__ void store4(const Row& c, int n, float4 v) const {   // 2C % 4 == 0
        const int h = n / (2 * C);
        const long long o = c.base + (long long)h * W * C + (n - h * 2 * C);
        if (D) { const float4 d = *reinterpret_cast<const float4*>(D + o); v.x += d.x; v.y += d.y; v.z += d.z; v.w += d.w; }
        *reinterpret_cast<float4*>(Y + o) = v;
    }
    __device__ __forceinline__ float4 bias4(int) const { return zero4(); }
    __device__ __forceinline__ void store4_nb(const Row& c, int n, float4 v) const { store4(c, n, v); }
    __device__ __forceinline__ float4 resid4(const Row&, int) const { return zero4(); }
    __device__ __forceinline__ void fin4(const Row& c, int n, float4 v, const float4) const { store4(c, n, v); }
    __device__ __forceinline__ void prefetch(const Row&, int) const {}
};

// conv5x5 bias + pixel shuffle (3,2) to channels-last [B, pf*H, pt*W, ldy] (scale.py:77-78).  The GEMM's columns are
// padded to the pixel pitch: n = (s1*pt + s2)*ldy + c -> pixel (pf*h + s1, pt*w + s2), channel c (c >= C: zero pad).
struct EpiDeembed {
    float* Y;
    const float* bias;      // [pf*pt*ldy], zero in the pad channels
    int ldy, H, W, C, pf, pt;
    struct Row { long long y; };
    __device__ __forceinline__ bool row(long long m, Row& c) const {
        const unsigned mm = (unsigned)m, bh = mm / (unsigned)W, w = mm - bh * (unsigned)W;
        c.y = (((long long)bh * pf) * (long long)(W * pt) + (long long)w * pt) * ldy;
        return true;
    }
    __device__ __forceinline__ long long at(const Row& c, int n) const {
        const int s = n / ldy, ch = n - s * ldy;
        const int s1 = s / pt, s2 = s - s1 * pt;
        return c.y + ((long long)s1 * (W * pt) + s2) * ldy + ch;
    }
    __device__ __forceinline__ void store(const Row& c, int n, float v) const { Y[at(c, n)] = v + __ldg(bias + n); }
    __device__ __forceinline__ void store4(const Row& c, int n, float4 v) const {   // n % 4 == 0: one pixel, aligned
        const float4 b = ldg4(bias + n);
        v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
        *reinterpret_cast<float4*>(Y + at(c, n)) = v;
    }
    __device__ __forceinline__ float4 bias4(int n) const { return ldg4(bias + n); }
    // streaming store: the 800 MB pixel map must not push the 133 MB token map (re-read 25 times) out of L2
    __device__ __forceinline__ void store4_nb(const Row& c, int n, float4 v) const { __stcs(reinterpret_cast<float4*>(Y + at(c, n)), v); }
    __device__ __forceinline__ float4 resid4(const Row&, int) const { return zero4(); }
    __device__ __forceinline__ void fin4(const Row& c, int n, float4 v, const float4) const { store4_nb(c, n, v); }
    __device__ __forceinline__ void prefetch(const Row&, int) const {}
};

// overlap-add normalisation + trim of torch.istft: sample s = hop*(j - j0) + n of clip b gets v / envelope, where
// the envelope is the sum of squared window taps of the frames that overlap it.
struct EpiIstft {
    float* Y;
    const float* wsq;   // [win] squared window
    int T, hop, nov, j0, nchunks;
    long long out_len;
    struct Row { float* y; int j; };
    __device__ __forceinline__ bool row(long long m, Row& c) const {
        const int jj = (int)(m % nchunks);
        c.j = jj + j0;
        c.y = Y + (m / nchunks) * out_len + (long long)jj * hop;
        return true;
    }
    __device__ __forceinline__ void store(const Row& c, int n, float v) const {
        float env = 0.f;
        for (int dt = 0; dt < nov; ++dt) {
            const int t = c.j - dt;
            if (t >= 0 && t < T) env += __ldg(wsq + dt * hop + n);
        }
        c.y[n] = v / env;
    }
    __device__ __forceinline__ void store4(const Row& c, int n, float4 v) const {
        store(c, n, v.x); store(c, n + 1, v.y); store(c, n + 2, v.z); store(c, n + 3, v.w);
    }
    __device__ __forceinline__ float4 bias4(int) const { return zero4(); }
    __device__ __forceinline__ void store4_nb(const Row& c, int n, float4 v) const { store4(c, n, v); }
    __device__ __forceinline__ float4 resid4(const Row&, int) const { return zero4(); }
    __device__ __forceinline__ void fin4(const Row& c, int n, float4 v, const float4) const { store4(c, n, v); }
    __device__ __forceinline__ void prefetch(const Row&, int) const {}
};

}  // namespace escb
