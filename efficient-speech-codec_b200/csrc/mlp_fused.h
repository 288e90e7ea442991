// Fused Swin MLP:  x <- x + fc2( GELU( fc1( LayerNorm2(x) ) ) )   (attention.py:176-178, 258-272 of the reference)
// ONE launch, the hidden map never leaves the SM.  The unfused pair (mlp1_gemm + mlp2_gemm, swin.cu) writes and re-reads
// the 4C-wide hidden map through HBM - 40 % of the step's DRAM traffic at 36 clips - and runs a LayerNorm-statistics
// pre-kernel in front; this kernel reads the token map once and writes it once.
//
// Per persistent CTA (one per SM, 20 warps), looping over 128-row tiles of the token map:
//   warp 0      x loader : one elected lane, TMA tensor loads (cp.async.bulk.tensor.2d -> SASS UTMALDG) of the 128 x ld fp32
//                          tile into a 128-byte-swizzled shared-memory slot (2-3 slots), completion on an mbarrier;
//   warp 1      w loader : pre-swizzled [hi | lo] weight images by cp.async.bulk (UBLKCP) - resident for the life of the
//                          CTA when both layers fit (C = 45), else streamed from L2 through a ring in the MMA's order;
//   warp 2      MMA      : one elected lane issues tcgen05.mma.kind::tf32 with the A operand in TENSOR MEMORY
//                          (D[tmem] += A[tmem] * B[smem]^T), 3 MMAs per k-step (A_lo B_hi + A_hi B_lo + A_hi B_hi);
//   warps 4-7   LN       : thread = row.  Reads its row of the x tile from shared memory, two-pass LayerNorm in
//                          registers (no statistics pre-kernel, no shuffles), cvt.rna.tf32 split, tcgen05.st of the
//                          hi and lo images into TMEM (lane = row, column = k: the TS-mode A layout);
//   warps 8-15  GELU     : thread = row x 32 of the 64 hidden columns of a chunk.  tcgen05.ld the fc1 accumulator, + bias,
//                          exact-erf GELU (FFMA2 polynomial), split, tcgen05.st hi IN PLACE over the accumulator and lo
//                          next to it: the chunk is now the A operand of fc2's k-steps;
//   warps 16-19 OUT      : thread = row.  tcgen05.ld the fc2 accumulator, + bias + residual (the x tile still in shared
//                          memory), writes the result over the x tile and stores it with a TMA tensor store (UTMASTG);
//                          optionally emits the NEXT LayerNorm's (mean, rstd) of the rows it just produced.
// The hidden dimension is processed in chunks of 64 columns: G1(c) = fc1 chunk -> R[c % 2], GELU(c), G2(c) = fc2
// partial product of that chunk accumulated into ACC2.  The MMA warp issues ... G2(c-2), G1(c), G2(c-1), G1(c+1) ... so
// the tensor pipe works on G1 of the next chunks while the GELU warps convert the current one.
//
// TMEM columns (512): A1 images 2 * Kp16 per buffer | R0, R1 (64 each) | L0[, L1] (64 each) | ACC2 (N2 per buffer) [| ACC2 corrections].
// Precision: the arithmetic of the unfused tcgen05 path (same 3xTF32 split, same k order, same bias / GELU / residual
// formulas); the LayerNorm sums are accumulated per thread instead of by 8 lanes, and fc2's partial products go to ONE
// main accumulator (+ one for the corrections at C = 45 / 72) where the unfused mlp2 may alternate K blocks between two
// mains: the two paths agree to fp32 rounding, not bit for bit (tests/test_gpu_parity.py::test_engine_variants_agree).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

#include "loaders.cuh"
#include "tc_gemm.cuh"

namespace escb {
namespace mf {

constexpr int BM = 128;
constexpr int HC = 64;                  // hidden columns per chunk = UMMA N of fc1 = two K blocks of fc2
// Warp ids: the schedulers favour high warp ids (measured in round 1 on the GEMM engine), and the single MMA-issuing
// warp must not starve behind the ALU-heavy GELU warps that share its scheduler - it sits on top.
#ifndef ESCB_MF_GELU_WARPS
#define ESCB_MF_GELU_WARPS 8
#endif
constexpr int LN_BASE = 0, OUT_BASE = 4, GELU_BASE = 8, GELU_WARPS = ESCB_MF_GELU_WARPS;   // 8 or 16: 32 or 16 hidden columns per thread and chunk
constexpr int GELU_COLS = HC / (GELU_WARPS / 4);
// Two MMA-issuing warps: W_MMA1 issues the fc1 chunks (G1), W_MMA2 the fc2 partial products (G2).  tcgen05.mma issue is
// nearly synchronous with the tensor pipe (the issuing thread stalls while its MMAs execute), so with ONE issuer the
// per-op overhead (barrier probes, descriptor arithmetic, commits: ~1 kclk per chunk at C = 45) ran with the pipe idle;
// with two, one warp's overhead hides behind the other's MMAs.  Each has its own weight loader warp and ring.
constexpr int W_XLOAD = GELU_BASE + GELU_WARPS, W_WLOAD1 = W_XLOAD + 1, W_WLOAD2 = W_XLOAD + 2, W_MMA1 = W_XLOAD + 3, W_MMA2 = W_XLOAD + 4;
constexpr int W_ALLOC = W_XLOAD;        // the x loader warp also allocates / frees tensor memory
constexpr int WARPS = W_MMA2 + 1, THREADS = WARPS * 32;
constexpr int MAX_WST = 8;              // weight stage barriers per layer (ring slots, or all stages of a tile when resident)
constexpr int MAX_NX = 3;
constexpr int BOX_BYTES = BM * 128;     // one 32-column box of the x tile: 128 rows x 128 bytes
constexpr int ST1_BYTES = 2 * HC * 128; // fc1 stage: [hi | lo] images of 64 hidden rows x one 32-wide K block
// When the last K block of fc1 holds at most two k-steps (16 input channels: C = 45 -> k = 32..47, C = 72 -> k = 64..71)
// its hi and lo values share ONE image: row r = [hi(k0..k0+15) | lo(k0..k0+15)], so the lo operand of k-step ks is the hi
// descriptor advanced by 64 bytes.  Saves 8 KB per chunk: the third x slot at C = 45, 17 % of the fc1 stream at C = 72.
constexpr int ST1T_BYTES = HC * 128;
inline int fc1_tail_ksteps(int C) { return ((C + 7) / 8) % 4; }          // k-steps in the last K block (0: it is full)
inline bool fc1_tail_packed(int C) { const int t = fc1_tail_ksteps(C); return t == 1 || t == 2; }
constexpr int MAX_C = 96;

// barrier indices
constexpr int B_XFULL = 0, B_XFREE = B_XFULL + MAX_NX, B_A1FULL = B_XFREE + MAX_NX, B_A1FREE = B_A1FULL + 2,
              B_RFULL = B_A1FREE + 2, B_HFULL = B_RFULL + 2, B_LFREE = B_HFULL + 2, B_ACCFULL = B_LFREE + 1,
              B_ACCFREE = B_ACCFULL + 2, B_RFREE = B_ACCFREE + 2, B_W1FULL = B_RFREE + 2, B_W1FREE = B_W1FULL + MAX_WST,
              B_W2FULL = B_W1FREE + MAX_WST, B_W2FREE = B_W2FULL + MAX_WST, NBARS = B_W2FREE + MAX_WST;

// ---------------------------------------------------------------------------------------------------- host side
struct Plan {                       // geometry of the fused kernel for one channel width (filled at pack time)
    int ok = 0;
    int C, ld, Kp16, ksteps1, nkb1, nch, N2, nx, na1, nl, nacc, resident, ns1, ns2, nboxf, rem;
    unsigned st2_bytes, slot_bytes, chunk_bytes, xslot_bytes;
    int col_a1, col_r, col_l, col_acc;
    int corr2;                      // fc2's lo*hi + hi*lo corrections accumulate in a region of their own behind ACC2 (tc_gemm.cuh acc_policy)
    size_t smem_bytes;
    size_t img_floats;              // weight blob size
};

inline Plan make_plan(int C, int hidden) {
    Plan pl;
    if ((C != 45 && C != 72 && C != 96) || hidden != 4 * C) return pl;      // widths the kernel is instantiated for (mlp_fused.cu)
    pl.C = C;
    pl.ld = (C + 3) & ~3;
    pl.Kp16 = (C + 15) & ~15;
    pl.ksteps1 = (C + 7) / 8;
    pl.nkb1 = (pl.ksteps1 + 3) / 4;
    pl.nch = (hidden + HC - 1) / HC;
    pl.N2 = (C + 15) & ~15;
    pl.nboxf = pl.ld / 32;
    pl.rem = pl.ld % 32;
    pl.st2_bytes = (unsigned)pl.N2 * 256u;
    pl.slot_bytes = pl.st2_bytes > (unsigned)ST1_BYTES ? pl.st2_bytes : (unsigned)ST1_BYTES;
    pl.chunk_bytes = (unsigned)(pl.nkb1 - 1) * ST1_BYTES + (fc1_tail_packed(C) ? ST1T_BYTES : ST1_BYTES) + 2u * pl.st2_bytes;
    pl.xslot_bytes = (unsigned)((pl.nboxf * BOX_BYTES + BM * pl.rem * 4 + 1023) & ~1023);
    pl.img_floats = (size_t)pl.nch * pl.chunk_bytes / 4;
    // TMEM: A1 buffers | R0 R1 | L buffers | ACC2 buffers
    const int fixed = 2 * HC;
    pl.na1 = 1; pl.nl = 1; pl.nacc = 1;
    // fc2 reduces over 4C = 180..384 hidden units (68..144 MMAs in one accumulator): its corrections get their own
    // accumulator where TMEM has the columns (C = 45, 72; ESCB_MF_CORR=0 keeps one accumulator, A-B builds)
    pl.corr2 = 1;
    if (const char* e = getenv("ESCB_MF_CORR")) pl.corr2 = atoi(e) ? 1 : 0;
    if (2 * pl.Kp16 + fixed + HC + 2 * pl.N2 > 512) pl.corr2 = 0;
    auto cols = [&](int na1, int nl, int nacc) { return na1 * 2 * pl.Kp16 + fixed + nl * HC + nacc * pl.N2 * (1 + pl.corr2); };
    if (cols(1, 1, 1) > 512) return pl;
    if (cols(1, 2, 1) <= 512) pl.nl = 2;
    if (cols(2, pl.nl, 1) <= 512) pl.na1 = 2;
    else if (cols(1, pl.nl, 2) <= 512) pl.nacc = 2;
    if (const char* e = getenv("ESCB_MF_BUF")) {           // experiments: "na1,nl,nacc" buffer counts (taken when they fit)
        int a = 1, l = 1, c = 1;
        if (sscanf(e, "%d,%d,%d", &a, &l, &c) == 3 && a >= 1 && a <= 2 && l >= 1 && l <= 2 && c >= 1 && c <= 2 && cols(a, l, c) <= 512) {
            pl.na1 = a; pl.nl = l; pl.nacc = c;
        }
    }
    pl.col_a1 = 0;
    pl.col_r = pl.na1 * 2 * pl.Kp16;
    pl.col_l = pl.col_r + fixed;
    pl.col_acc = pl.col_l + pl.nl * HC;
    // shared memory: x slots | weights | barriers
    const size_t tail = NBARS * 8 + 64, budget = tc::SMEM_MAX - 1024;
    const size_t w_all = (size_t)pl.nch * pl.chunk_bytes;
    const size_t fc1_chunk = pl.chunk_bytes - 2 * (size_t)pl.st2_bytes;
    pl.nx = 2;
    pl.resident = (2 * (size_t)pl.xslot_bytes + w_all + tail <= budget && pl.nch * pl.nkb1 <= MAX_WST && pl.nch * 2 <= MAX_WST) ? 1 : 0;
    if (pl.resident) {
        pl.ns1 = pl.nch * pl.nkb1;
        pl.ns2 = pl.nch * 2;
        if (3 * (size_t)pl.xslot_bytes + w_all + tail <= budget) pl.nx = 3;
        pl.smem_bytes = 1024 + (size_t)pl.nx * pl.xslot_bytes + w_all + tail;
    } else {
        // two rings (fc1 slots of ST1_BYTES, fc2 slots of st2_bytes): at least one chunk of each, then grow them in turn
        const size_t left = budget - tail - 2 * (size_t)pl.xslot_bytes;
        int n1 = pl.nkb1, n2 = 2;
        if ((size_t)n1 * ST1_BYTES + (size_t)n2 * pl.st2_bytes > left) return pl;
        for (;;) {
            const size_t used = (size_t)n1 * ST1_BYTES + (size_t)n2 * pl.st2_bytes;
            const bool can1 = n1 < MAX_WST && n1 < 2 * pl.nkb1 + 1 && used + ST1_BYTES <= left;
            const bool can2 = n2 < MAX_WST && n2 < 4 && used + pl.st2_bytes <= left;
            if (can1 && (n1 - pl.nkb1 <= (n2 - 2) || !can2)) ++n1;
            else if (can2) ++n2;
            else break;
        }
        pl.ns1 = n1;
        pl.ns2 = n2;
        (void)fc1_chunk;
        pl.smem_bytes = 1024 + (size_t)pl.nx * pl.xslot_bytes + (size_t)n1 * ST1_BYTES + (size_t)n2 * pl.st2_bytes + tail;
    }
    pl.ok = 1;
    return pl;
}

struct Weights {                    // device pointers of one block's fused-MLP operands (api.cu put_mlp_fused)
    Plan plan;
    const float* img = nullptr;
    const float* b1 = nullptr;
    const float* b2 = nullptr;
    const float* gamma = nullptr;
    const float* beta = nullptr;
};

struct StatsOut { float2* out = nullptr; int geom = 0; int H = 0, W = 0; WindowGeom ng; };

// mlp_fused.cu
cudaError_t launch(cudaStream_t st, const Weights& w, float* x, long long M, float eps, const StatsOut& so,
                   unsigned long long* trace = nullptr);

}  // namespace mf
}  // namespace escb
