// Causal attention BACKWARD on tcgen05 tensor cores, "wide" scheme: 128 x 128 tiles, one tile set, interleaved MMA order.
// bf16 / fp16, head size 64 or 128, dense [B,H,S,D] or strided (packed qkv projection) operands through 4-D TMA maps.
// The reference has no attention backward at all (only AddGradFunction exists, src/core/binary_ops.cpp:16-43); the
// oracle is the float64 restatement in oracle/oracle.py (causal_attention_bwd).
//
// Same deterministic two-kernel decomposition as attention_bwd_tc.cu (no atomics, bit-reproducible), one template, two modes:
//   MODE_DKV : one CTA per 128-row KV block j.  X0 = K_j, X1 = V_j stay in shared memory; 128-row query tiles Y0 = Q_t, Y1 = dO_t stream.
//              T0 = S^T = K_j Q_t^T,  T1 = dP^T = V_j dO_t^T          (128 x 128 fp32 in TMEM, lanes = kv rows)
//              P^T = exp2(S^T c - lse_q),  dS^T = P^T o (dP^T - delta_q)  -> 16-bit, written back over T0 / T1 in TMEM
//              dV_j += P^T dO_t  (A0),   dK_j += dS^T Q_t  (A1)        (A from TENSOR MEMORY, B = the streamed tile, MN-major)
//   MODE_DQ  : one CTA per 128-row query block i.  X0 = Q_i, X1 = dO_i stay; Y0 = K_t, Y1 = V_t stream.
//              T0 = S = Q_i K_t^T,  T1 = dP = dO_i V_t^T,  dS as above with per-row lse / delta,   dQ_i += dS K_t  (A1).
// What changed against the 64-wide two-set kernel (measured there: tensor pipe 57-59 % active):
//   * N = 128 MMAs.  An M128 N64 K16 MMA reads its whole 4 KB A slice from shared memory for 32 tensor cycles of math (48 cycles of
//     shared-memory traffic): S^T and dP^T cost 1.5x their math.  At N = 128 the operand reads (8 KB, 64 cycles) match the math.
//   * Sixteen element-wise warps on ONE tile (thread = stationary row x 32-column quarter) instead of four on each of two tiles,
//     and every second pair of exponentials evaluated on the FMA pipe (the MUFU unit's 16 ex2 / clk / SM alone is 1024 clk per
//     128 x 128 tile): measured with eight warps and MUFU only, phase A took 1430 clk and phase B 700 clk of a 2760 clk tile period
//     (profiles/r2_attn_bwd_wide_trace.log).  The MMA order below hides the phases without a second tile set in TMEM:
//         DKV:  T0(n) T1(n) | A0(n) T0(n+1) A1(n) T1(n+1) | A0(n+1) T0(n+2) A1(n+1) T1(n+2) | ...
//         DQ :  T0(n) T1(n) |       T0(n+1) A1(n) T1(n+1) |         T0(n+2) A1(n+1) T1(n+2) | ...
//     phase A of tile n+1 (exp2, needs T0) runs under A1(n) + T1(n+1); phase B (dS, needs T1) under A0(n+1) + T0(n+2).  P stays
//     in registers (fp32) between the phases, so T0 can be overwritten as soon as A0 has read P^T (DQ: as soon as it is loaded).
//   * Separate rings for the two streamed operands (Y0: 3 slots, it is read by the first and the last MMA of a tile; Y1: 2); the
//     per-column -lse2 / -delta vectors of a streamed query tile arrive by bulk copy from the producer warp (no named barrier).
// TMEM (512 columns): T0 | T1 | ACC0 = dV | ACC1 = dK or dQ.  16-bit P^T / dS^T of streamed columns [0,64) land on TMEM columns
// [32 c, 32 c + 16) for the column quarter c: every element-wise warp only overwrites columns it has already read.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "ew_common.cuh"
#include "tc_common.cuh"

namespace kf {
using namespace tc;

namespace {

constexpr int W_EW_WARPS = 16;
constexpr int W_THREADS = (W_EW_WARPS + 2) * 32;  // warps 0-15 element-wise, 16 MMA issuer, 17 TMA producer
constexpr int W_NS0 = 3, W_NS1 = 2;
enum { W_DKV = 0, W_DQ = 1 };

struct WParams {
    int64_t Sq, Skv;
    int H;               // heads per batch entry: (b, h) = (bh / H, bh % H)
    const float *nlse2;   // [BH, Sq_pad] MINUS the row log-sum-exp in the exp2 domain (rows padded to a multiple of 128 with zeros)
    const float *ndelta;  // [BH, Sq_pad] MINUS rowsum(dO o O)
    int64_t Sq_pad;
    void *out0, *out1;   // DKV: dV, dK      DQ: unused, dQ
    AttnLayout l0, l1;
    float scale_log2;    // softmax scale * log2(e)
    float scale;         // softmax scale (applied to dK / dQ in the epilogue)
    int nblk;            // 128-row blocks of the stationary operand per (b, h)
    int is_bf16;
    int hg, BH;          // heads per scheduling group (weight-major inside a group, see attention_bwd_tc.cu) / batch-heads
    long long *trace;    // KF_ATTN_TRACE=1: clock64() stamps of CTA 0's pipeline events, [tile][16]; null otherwise
};
#define W_TRACE(n, k)                                                                                   \
    do {                                                                                                \
        if (p.trace != nullptr && blockIdx.x == 0 && (n) < 64 && (threadIdx.x & 31) == 0) p.trace[(n) * 24 + (k)] = clock64(); \
    } while (0)

template <bool BF16>
__device__ __forceinline__ uint32_t w_pack(float2 v) {
    if (BF16) {
        __nv_bfloat162 h = __float22bfloat162_rn(v);
        return *reinterpret_cast<uint32_t *>(&h);
    }
    __half2 h = __float22half2_rn(v);
    return *reinterpret_cast<uint32_t *>(&h);
}
__device__ __forceinline__ float w_ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float4 w_lds128(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}

// Shared-memory descriptor (128B swizzle) as two 32-bit words, so that the per-k-step variants are one 32-bit add on the low word
// (start address >> 4 in bits [0,14), leading-dim offset >> 4 in [16,30)); high word: stride-dim offset >> 4, version 1, swizzle 2.
__device__ __forceinline__ uint32_t w_desc_lo(uint32_t smem_addr, uint32_t lbo_bytes) { return ((smem_addr & 0x3FFFFu) >> 4) | (((lbo_bytes >> 4) & 0x3FFFu) << 16); }
constexpr uint32_t W_DESC_HI = (1024u >> 4) | (1u << 14) | (2u << 29);  // SBO = 1024 B
__device__ __forceinline__ void w_umma_ss(uint32_t tmem_d, uint32_t alo, uint32_t blo, uint32_t idesc, uint32_t accumulate, bool leader) {
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t.reg .b64 da, db;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "setp.ne.b32 q, %5, 0;\n\t"
        "mov.b64 da, {%1, %6};\n\t"
        "mov.b64 db, {%2, %6};\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(alo), "r"(blo), "r"(idesc), "r"(accumulate), "r"((uint32_t)leader), "r"(W_DESC_HI)
        : "memory");
}
__device__ __forceinline__ void w_umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint32_t blo, uint32_t idesc, uint32_t accumulate, bool leader) {
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t.reg .b64 db;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "setp.ne.b32 q, %5, 0;\n\t"
        "mov.b64 db, {%2, %6};\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "r"(blo), "r"(idesc), "r"(accumulate), "r"((uint32_t)leader), "r"(W_DESC_HI)
        : "memory");
}

// 2^x for a pair on the FMA pipe: round-to-nearest split x = j + f via the 1.5 * 2^23 magic add, degree-3 minimax of 2^f on
// [-1/2, 1/2] (max rel. error 7.5e-5, far below the 16-bit rounding of P), j added into the exponent field by one integer
// multiply-add.  x is clamped at -126 so the exponent arithmetic cannot wrap (masked entries are selected to 0 afterwards).
__device__ __forceinline__ float2 w_ex2_poly2(float2 x) {
    x.x = fmaxf(x.x, -126.f);
    x.y = fmaxf(x.y, -126.f);
    const float2 magic = make_float2(12582912.f, 12582912.f);
    const float2 t = __fadd2_rn(x, magic);
    const float2 r = __fadd2_rn(t, make_float2(-12582912.f, -12582912.f));
    const float2 f = __ffma2_rn(r, make_float2(-1.f, -1.f), x);
    float2 q = __ffma2_rn(f, make_float2(0.0551716648042202f, 0.0551716648042202f), make_float2(0.2426111251115799f, 0.2426111251115799f));
    q = __ffma2_rn(q, f, make_float2(0.6932609677314758f, 0.6932609677314758f));
    q = __ffma2_rn(q, f, make_float2(0.9999280571937561f, 0.9999280571937561f));
    return make_float2(__uint_as_float(__float_as_uint(t.x) * 8388608u + __float_as_uint(q.x)),
                       __uint_as_float(__float_as_uint(t.y) * 8388608u + __float_as_uint(q.y)));
}

// Phase A for one thread (one stationary row, 32 streamed columns): P = exp2(T0 c - lse2), rounded to 16 bits, into `pk` (kept for
// phase B: dS is formed from the SAME rounded P the dV MMA consumes); DKV also stores it over the first 16 of the 32 TMEM columns read.
// vec_smem (DKV): -lse2 of this thread's 32 streamed columns; nl_row (DQ): -lse2 of this thread's row.
template <int MODE, bool MASKED, bool BF16>
__device__ __forceinline__ void w_phase_a(const uint32_t t0_addr, const uint32_t vec_smem, const float nl_row, const float sc, const int lo, const int hi,
                                          uint32_t (&pk)[16], uint64_t *t0_free, long long *tr) {
    const float2 sc2 = make_float2(sc, sc);
    uint32_t s[32];
    tmem_ld32(t0_addr, s);
    tmem_ld_wait();
    if (tr) tr[6] = clock64();
    if (MODE == W_DQ) {  // T0 is in registers: the next S may overwrite it
        tc_fence_before();
        __syncwarp();
        if ((threadIdx.x & 31) == 0) mbar_arrive(t0_free);
    }
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
        float4 nl;
        if (MODE == W_DKV) nl = w_lds128(vec_smem + 4u * i);
        else nl = make_float4(nl_row, nl_row, nl_row, nl_row);
        const float2 xa = __ffma2_rn(make_float2(__uint_as_float(s[i]), __uint_as_float(s[i + 1])), sc2, make_float2(nl.x, nl.y));
        const float2 xb = __ffma2_rn(make_float2(__uint_as_float(s[i + 2]), __uint_as_float(s[i + 3])), sc2, make_float2(nl.z, nl.w));
        float p0 = w_ex2(xa.x), p1 = w_ex2(xa.y);     // MUFU
        const float2 pb = w_ex2_poly2(xb);            // FMA pipe
        float p2 = pb.x, p3 = pb.y;
        if (MASKED) {
            if (i < lo || i >= hi) p0 = 0.f;
            if (i + 1 < lo || i + 1 >= hi) p1 = 0.f;
            if (i + 2 < lo || i + 2 >= hi) p2 = 0.f;
            if (i + 3 < lo || i + 3 >= hi) p3 = 0.f;
        }
        pk[i >> 1] = w_pack<BF16>(make_float2(p0, p1));
        pk[(i >> 1) + 1] = w_pack<BF16>(make_float2(p2, p3));
    }
    if (tr) tr[7] = clock64();
    if (MODE == W_DKV) tmem_st16(t0_addr, pk);
}

// Phase B: dS = P o (T1 - delta) -> 16 bits over the first 16 of the 32 TMEM columns of T1 this thread has just read.
template <bool BF16>
__device__ __forceinline__ float2 w_unpack(uint32_t v) {
    if (BF16) return make_float2(__uint_as_float(v << 16), __uint_as_float(v & 0xffff0000u));
    return __half22float2(*reinterpret_cast<const __half2 *>(&v));
}
template <int MODE, bool BF16>
__device__ __forceinline__ void w_phase_b(const uint32_t t1_addr, const uint32_t ds_addr, const uint32_t vec_smem, const float nd_row, const uint32_t (&pk)[16],
                                          uint64_t *t1_free, uint64_t *ds_free, const uint32_t ds_free_parity, long long *tr) {
    uint32_t d[32];
    tmem_ld32(t1_addr, d);
    tmem_ld_wait();
    if (tr) tr[12] = clock64();
    if (MODE == W_DQ) {  // T1 is in registers: the next dP may overwrite it
        tc_fence_before();
        __syncwarp();
        if ((threadIdx.x & 31) == 0) mbar_arrive(t1_free);
    }
    uint32_t dk[16];
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
        float4 nd;
        if (MODE == W_DKV) nd = w_lds128(vec_smem + 4u * i);
        else nd = make_float4(nd_row, nd_row, nd_row, nd_row);
        const float2 da = __fmul2_rn(w_unpack<BF16>(pk[i >> 1]),
                                     __fadd2_rn(make_float2(__uint_as_float(d[i]), __uint_as_float(d[i + 1])), make_float2(nd.x, nd.y)));
        const float2 db = __fmul2_rn(w_unpack<BF16>(pk[(i >> 1) + 1]),
                                     __fadd2_rn(make_float2(__uint_as_float(d[i + 2]), __uint_as_float(d[i + 3])), make_float2(nd.z, nd.w)));
        dk[i >> 1] = w_pack<BF16>(da);
        dk[(i >> 1) + 1] = w_pack<BF16>(db);
    }
    if (tr) tr[13] = clock64();
    if (MODE == W_DQ) {  // the dS buffer of two tiles ago must have been consumed by its dQ MMAs
        mbar_wait(ds_free, ds_free_parity);
        tc_fence_after();
    }
    tmem_st16(ds_addr, dk);
}

// MINUS rowsum(dO o O) and MINUS the exp2-domain LSE for [B, H, S, D] operands in arbitrary layouts, written as [BH, Sq_pad] rows
// (padding = 0, so that a masked P = 0 times (dP + 0) stays 0).  16-byte loads, D / 8 lanes per query row.
template <typename T>
__global__ void __launch_bounds__(256) w_prep_kernel(const T *__restrict__ o, const T *__restrict__ dout, const float *__restrict__ lse,
                                                     float *__restrict__ ndelta, float *__restrict__ nlse2, const int64_t rows_pad, const int64_t Sq,
                                                     const int64_t Sq_pad, const int H, const int D, const AttnLayout lo, const AttnLayout ldo) {
    const int lpr = D >> 3;  // lanes per row: 8 or 16
    const int64_t row = (int64_t)blockIdx.x * (256 / lpr) + threadIdx.x / lpr;  // index into the padded layout
    const int sub = threadIdx.x % lpr;
    const int64_t bh = row / Sq_pad, s = row % Sq_pad;
    const bool live = row < rows_pad && s < Sq;
    float acc = 0.f;
    if (live) {
        const int64_t b = bh / H, h = bh % H;
        const uint4 vo = __ldg(reinterpret_cast<const uint4 *>(o + b * lo.sb + h * lo.sh + s * lo.ss) + sub);
        const uint4 vd = __ldg(reinterpret_cast<const uint4 *>(dout + b * ldo.sb + h * ldo.sh + s * ldo.ss) + sub);
        const T *po = reinterpret_cast<const T *>(&vo), *pd = reinterpret_cast<const T *>(&vd);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc = fmaf(cvt_in<float>(po[i]), cvt_in<float>(pd[i]), acc);
    }
    for (int off = lpr >> 1; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    if (row < rows_pad && sub == 0) {
        ndelta[row] = live ? -acc : 0.f;
        nlse2[row] = live ? -lse[bh * Sq + s] * 1.4426950408889634f : 0.f;
    }
}

__device__ __forceinline__ void w_bulk_load(void *smem_dst, const void *gsrc, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(gsrc)), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

template <int D, int MODE>
__global__ void __launch_bounds__(W_THREADS, 1)
attn_bwd_wide_kernel(const __grid_constant__ CUtensorMap tmap_x0, const __grid_constant__ CUtensorMap tmap_x1,
                     const __grid_constant__ CUtensorMap tmap_y0, const __grid_constant__ CUtensorMap tmap_y1, const WParams p) {
    constexpr int ATOMS = D / 64;
    constexpr int TILE = 128 * D * 2, ATOM = 128 * 128;  // every tile: 128 rows, ATOMS column atoms of 64 elements (128 B rows)
    constexpr uint32_t TMEM_COLS = 512;
    // DKV: T0 | T1 | dV | dK (P^T over T0, dS^T over T1).   DQ: T0 | T1 | dS[2] (64 columns each, tiles alternate) | dQ
    constexpr uint32_t T0_COL = 0, T1_COL = 128, ACC0_COL = 256, ACC1_COL = MODE == W_DKV ? 256 + D : 384, DS_COL = 256;
    constexpr int W_MMA = W_EW_WARPS, W_TMA = W_EW_WARPS + 1;
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char *sX = smem;                        // X0 | X1
    unsigned char *sY0 = smem + 2 * TILE;            // W_NS0 slots
    unsigned char *sY1 = sY0 + W_NS0 * TILE;         // W_NS1 slots
    float *svec = reinterpret_cast<float *>(sY1 + W_NS1 * TILE);  // [2 slots][-lse2[128] | -delta[128]]
    uint64_t *bars = reinterpret_cast<uint64_t *>(svec + 2 * 256);
    uint64_t *x_full = bars;
    uint64_t *y0_full = bars + 1, *y0_empty = y0_full + W_NS0;
    uint64_t *y1_full = y0_empty + W_NS0, *y1_empty = y1_full + W_NS1;
    uint64_t *t0_full = y1_empty + W_NS1, *t1_full = t0_full + 1;
    uint64_t *p_full = t1_full + 1, *ds_full = p_full + 1, *acc_full = ds_full + 1;
    uint64_t *vec_full = acc_full + 1, *vec_empty = vec_full + 2;  // [2] each
    uint64_t *t1_free = vec_empty + 2, *ds_free = t1_free + 1;      // DQ only: T1 read into registers; dS buffer [2] consumed
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(ds_free + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int bh, blk;
    {
        const int per_group = p.hg * p.nblk;
        const int g = blockIdx.x / per_group, idx = blockIdx.x - g * per_group;
        const int heads = min(p.hg, p.BH - g * p.hg);  // the last group may be partial
        const int lvl = idx / heads;                    // 0 = heaviest
        bh = g * p.hg + (idx - lvl * heads);
        blk = MODE == W_DKV ? lvl : (p.nblk - 1 - lvl);
    }
    const int b_idx = bh / p.H, h_idx = bh % p.H;
    const int x0_row = blk * 128;
    int t_lo, t_hi;  // streamed 128-row tiles [t_lo, t_hi)
    if (MODE == W_DKV) {  // query tiles at or below the diagonal of this KV block
        t_lo = blk;
        t_hi = (int)((p.Sq + 127) / 128);
    } else {  // KV tiles up to the diagonal of this query block
        t_lo = 0;
        t_hi = (int)((min((int64_t)p.Skv, (int64_t)x0_row + 128) + 127) / 128);
    }
    const int ntile = max(t_hi - t_lo, 0);

    if (threadIdx.x == 0 && (smem_u32(smem) & 1023u) != 0) {
        printf("kfunca_b200: attn_bwd_wide_kernel needs 1024-byte aligned dynamic shared memory\n");
        __trap();
    }
    if (warp == W_TMA && lane == 0) {
        prefetch_tmap(&tmap_x0);
        prefetch_tmap(&tmap_x1);
        prefetch_tmap(&tmap_y0);
        prefetch_tmap(&tmap_y1);
        mbar_init(x_full, 1);
        for (int s = 0; s < W_NS0; ++s) {
            mbar_init(&y0_full[s], 1);
            mbar_init(&y0_empty[s], 1);
        }
        for (int s = 0; s < W_NS1; ++s) {
            mbar_init(&y1_full[s], 1);
            mbar_init(&y1_empty[s], 1);
        }
        mbar_init(t0_full, 1);
        mbar_init(t1_full, 1);
        mbar_init(p_full, W_EW_WARPS);
        mbar_init(ds_full, W_EW_WARPS);
        mbar_init(acc_full, 1);
        for (int s = 0; s < 2; ++s) {
            mbar_init(&vec_full[s], 1);
            mbar_init(&vec_empty[s], W_EW_WARPS);
            mbar_init(&ds_free[s], 1);
        }
        mbar_init(t1_free, W_EW_WARPS);
        fence_barrier_init();
    }
    if (warp == W_MMA) tmem_alloc(tmem_slot, TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == W_TMA) {
        // ===================================================== TMA producer
        if (lane == 0 && ntile > 0) {
            mbar_arrive_expect_tx(x_full, 2 * TILE);
#pragma unroll
            for (int a = 0; a < ATOMS; ++a) {
                tma_load_4d(sX + a * ATOM, &tmap_x0, x_full, a * 64, x0_row, h_idx, b_idx);
                tma_load_4d(sX + TILE + a * ATOM, &tmap_x1, x_full, a * 64, x0_row, h_idx, b_idx);
            }
            const float *nl = p.nlse2 + (int64_t)bh * p.Sq_pad, *nd = p.ndelta + (int64_t)bh * p.Sq_pad;
            for (int n = 0; n < ntile; ++n) {
                const int row = (t_lo + n) * 128;
                const int s0 = n % W_NS0, s1 = n % W_NS1;
                if (MODE == W_DKV) {  // the 2 x 128 per-column scalars of this query tile
                    const int sv = n & 1;
                    mbar_wait(&vec_empty[sv], (uint32_t)(((n >> 1) & 1) ^ 1));
                    mbar_arrive_expect_tx(&vec_full[sv], 1024);
                    w_bulk_load(svec + sv * 256, nl + row, 512, &vec_full[sv]);
                    w_bulk_load(svec + sv * 256 + 128, nd + row, 512, &vec_full[sv]);
                }
                mbar_wait(&y0_empty[s0], (uint32_t)(((n / W_NS0) & 1) ^ 1));
                mbar_arrive_expect_tx(&y0_full[s0], TILE);
#pragma unroll
                for (int a = 0; a < ATOMS; ++a) tma_load_4d(sY0 + s0 * TILE + a * ATOM, &tmap_y0, &y0_full[s0], a * 64, row, h_idx, b_idx);
                if (p.trace != nullptr && blockIdx.x == 0 && n < 64) p.trace[n * 24 + 22] = clock64();
                mbar_wait(&y1_empty[s1], (uint32_t)(((n / W_NS1) & 1) ^ 1));
                mbar_arrive_expect_tx(&y1_full[s1], TILE);
#pragma unroll
                for (int a = 0; a < ATOMS; ++a) tma_load_4d(sY1 + s1 * TILE + a * ATOM, &tmap_y1, &y1_full[s1], a * 64, row, h_idx, b_idx);
                if (p.trace != nullptr && blockIdx.x == 0 && n < 64) p.trace[n * 24 + 23] = clock64();
            }
        }
        __syncwarp();
    } else if (warp == W_MMA) {
        // ===================================================== MMA issuer (converged warp, the elected lane issues)
        const bool leader = elect_one();
        if (ntile > 0) {
            const int fmt = p.is_bf16 ? 1 : 0;
            const uint32_t idesc_t = make_idesc_f16(fmt, 0, 0, 128, 128);  // T = X Y^T : both operands K-major
            const uint32_t idesc_a = make_idesc_f16(fmt, 0, 1, 128, D);    // ACC += (TMEM) Y : B MN-major
            const uint32_t x_addr = smem_u32(sX), y0_addr = smem_u32(sY0), y1_addr = smem_u32(sY1);
            // Everything that does not depend on the element-wise warps (the TMA-full waits, the descriptor words) is done BEFORE the
            // wait for P / dS, so that only the tcgen05.mma instructions themselves sit between that wait and the tensor pipe (measured
            // before: ~300 clk to issue 8 MMAs plus ~250 clk in an already-complete mbarrier wait, twice per tile, all on the critical path).
            auto issue_t = [&](int which, uint32_t xlo, uint32_t ylo) {  // T_which = X_which Y_which^T; 64-element atoms ATOM bytes apart
#pragma unroll
                for (int kk = 0; kk < D / 16; ++kk) {
                    const uint32_t off = (uint32_t)(((kk >> 2) * ATOM + (kk & 3) * 32) >> 4);
                    w_umma_ss(tmem_base + (uint32_t)(which * 128), xlo + off, ylo + off, idesc_t, kk ? 1u : 0u, leader);
                }
            };
            // ACC += (16-bit tile in TMEM) Y.  K = 128 streamed rows = 8 x 16; 16-bit A in TMEM: 8 columns per step, at a_col + step_col(kk)
            auto issue_acc = [&](uint32_t acc_col, uint32_t a_col, uint32_t ylo, bool accumulate) {
#pragma unroll
                for (int kk = 0; kk < 8; ++kk) {
                    const uint32_t step_col = MODE == W_DKV ? (uint32_t)((kk >> 1) * 32 + (kk & 1) * 8) : (uint32_t)(kk * 8);
                    w_umma_ts(tmem_base + acc_col, tmem_base + a_col + step_col, ylo + (uint32_t)(kk * 2048 >> 4), idesc_a, (accumulate || kk) ? 1u : 0u, leader);
                }
            };
            const uint32_t x0lo = w_desc_lo(x_addr, 0), x1lo = w_desc_lo(x_addr + TILE, 0);
            auto y0_wait = [&](int n) { mbar_wait(&y0_full[n % W_NS0], (uint32_t)((n / W_NS0) & 1)); };
            auto y1_wait = [&](int n) { mbar_wait(&y1_full[n % W_NS1], (uint32_t)((n / W_NS1) & 1)); };
            auto y0_at = [&](int n) { return y0_addr + (uint32_t)((n % W_NS0) * TILE); };
            auto y1_at = [&](int n) { return y1_addr + (uint32_t)((n % W_NS1) * TILE); };
            mbar_wait(x_full, 0);
            y0_wait(0);
            tc_fence_after();
            issue_t(0, x0lo, w_desc_lo(y0_at(0), 0));
            umma_commit_p(t0_full, leader);
            y1_wait(0);
            tc_fence_after();
            issue_t(1, x1lo, w_desc_lo(y1_at(0), 0));
            umma_commit_p(t1_full, leader);
            if (MODE == W_DQ) umma_commit_p(&y1_empty[0], leader);
            for (int n = 0; n < ntile; ++n) {
                const uint32_t par = (uint32_t)(n & 1);
                const bool more = n + 1 < ntile;
                // ---- first half: (DKV: dV += P^T dO_t,) next T0
                const uint32_t a0_ylo = w_desc_lo(y1_at(n), ATOM);                 // B of A0: dO_t, MN-major
                const uint32_t t0_ylo = w_desc_lo(y0_at(more ? n + 1 : n), 0);    // B of the next T0
                if (more) y0_wait(n + 1);
                W_TRACE(n, 8);
                mbar_wait(p_full, par);  // DKV: P^T(n) is in TMEM; DQ: T0(n) has been read into registers
                tc_fence_after();
                W_TRACE(n, 9);
                if (MODE == W_DKV) {
                    issue_acc(ACC0_COL, T0_COL, a0_ylo, n > 0);
                    umma_commit_p(&y1_empty[n % W_NS1], leader);
                }
                if (more) {
                    issue_t(0, x0lo, t0_ylo);
                    umma_commit_p(t0_full, leader);
                }
                W_TRACE(n, 16);
                // ---- second half.  DKV: dK += dS^T Q_t, next T1 (dS^T lives in T1).  DQ: next T1 as soon as T1(n) is in registers,
                // then dQ += dS K_t from the separate dS buffer.
                const uint32_t a1_ylo = w_desc_lo(y0_at(n), ATOM);                 // B of A1: Q_t / K_t, MN-major
                const uint32_t t1_ylo = w_desc_lo(y1_at(more ? n + 1 : n), 0);
                if (more) y1_wait(n + 1);
                if (MODE == W_DQ && more) {
                    mbar_wait(t1_free, par);
                    tc_fence_after();
                    issue_t(1, x1lo, t1_ylo);
                    umma_commit_p(t1_full, leader);
                    umma_commit_p(&y1_empty[(n + 1) % W_NS1], leader);
                }
                W_TRACE(n, 10);
                mbar_wait(ds_full, par);
                tc_fence_after();
                W_TRACE(n, 11);
                issue_acc(ACC1_COL, MODE == W_DKV ? T1_COL : DS_COL + (uint32_t)((n & 1) * 64), a1_ylo, n > 0);
                umma_commit_p(&y0_empty[n % W_NS0], leader);
                if (MODE == W_DQ) umma_commit_p(&ds_free[n & 1], leader);
                if (MODE == W_DKV && more) {
                    issue_t(1, x1lo, t1_ylo);
                    umma_commit_p(t1_full, leader);
                }
                W_TRACE(n, 18);
            }
            umma_commit_p(acc_full, leader);
        }
        __syncwarp();
    } else {
        // ===================================================== element-wise math + epilogue: thread = stationary row x column quarter
        const int q4 = warp & 3, cq = warp >> 2;
        const int r = q4 * 32 + lane;
        const int64_t row_g = (int64_t)x0_row + r;  // DKV: kv index, DQ: query index
        const uint32_t lane_addr = tmem_base + ((uint32_t)(q4 * 32) << 16);
        const uint32_t t0_addr = lane_addr + T0_COL + (uint32_t)(cq * 32), t1_addr = lane_addr + T1_COL + (uint32_t)(cq * 32);
        const uint32_t ds_base = lane_addr + DS_COL + (uint32_t)(cq * 16);  // DQ: this thread's 16 packed columns of a dS buffer
        const float sc = p.scale_log2;
        float nl_row = 0.f, nd_row = 0.f;  // DQ: per-row scalars
        if (MODE == W_DQ && row_g < p.Sq) {
            nl_row = p.nlse2[(int64_t)bh * p.Sq_pad + row_g];
            nd_row = p.ndelta[(int64_t)bh * p.Sq_pad + row_g];
        }
        uint32_t pk[16];
        for (int n = 0; n < ntile; ++n) {
            const int64_t y0_row = (int64_t)(t_lo + n) * 128;
            const int sv = n & 1;
            const uint32_t vl = smem_u32(svec + sv * 256) + (uint32_t)(cq * 32 * 4), vd = vl + 128 * 4;
            if (warp == 0) W_TRACE(n, 0);
            if (MODE == W_DKV) mbar_wait(&vec_full[sv], (uint32_t)((n >> 1) & 1));
            if (warp == 0) W_TRACE(n, 1);
            // streamed columns c of the whole tile kept iff lo <= c < hi; this thread owns columns [cq * 32, cq * 32 + 32)
            int lo = 0, hi = 128;
            if (MODE == W_DKV) {  // keep q >= kv, q < Sq
                lo = (int)max((int64_t)0, min((int64_t)128, row_g - y0_row));
                hi = (int)max((int64_t)0, min((int64_t)128, p.Sq - y0_row));
            } else {  // keep kv <= q, kv < Skv
                hi = (int)max((int64_t)0, min((int64_t)128, min(row_g + 1, (int64_t)p.Skv) - y0_row));
            }
            lo = max(0, min(32, lo - cq * 32));
            hi = max(0, min(32, hi - cq * 32));
            const bool need_mask = __any_sync(0xffffffffu, lo > 0 || hi < 32);
            const uint32_t par = (uint32_t)(n & 1);
            long long *tr = (p.trace != nullptr && blockIdx.x == 0 && n < 64 && threadIdx.x == 0) ? p.trace + n * 24 : nullptr;
            // ---- phase A
            mbar_wait(t0_full, par);
            tc_fence_after();
            if (warp == 0) W_TRACE(n, 2);
            if (p.is_bf16) {
                if (need_mask) w_phase_a<MODE, true, true>(t0_addr, vl, nl_row, sc, lo, hi, pk, p_full, tr);
                else w_phase_a<MODE, false, true>(t0_addr, vl, nl_row, sc, lo, hi, pk, p_full, tr);
            } else {
                if (need_mask) w_phase_a<MODE, true, false>(t0_addr, vl, nl_row, sc, lo, hi, pk, p_full, tr);
                else w_phase_a<MODE, false, false>(t0_addr, vl, nl_row, sc, lo, hi, pk, p_full, tr);
            }
            if (MODE == W_DKV) {
                tmem_st_wait();
                if (tr) tr[14] = clock64();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(p_full);
            }
            if (warp == 0) W_TRACE(n, 3);
            // ---- phase B
            mbar_wait(t1_full, par);
            tc_fence_after();
            if (warp == 0) W_TRACE(n, 4);
            const uint32_t ds_addr = MODE == W_DKV ? t1_addr : ds_base + (uint32_t)((n & 1) * 64);
            const uint32_t dsf_par = (uint32_t)(((n >> 1) & 1) ^ 1);
            if (p.is_bf16) w_phase_b<MODE, true>(t1_addr, ds_addr, vd, nd_row, pk, t1_free, &ds_free[n & 1], dsf_par, tr);
            else w_phase_b<MODE, false>(t1_addr, ds_addr, vd, nd_row, pk, t1_free, &ds_free[n & 1], dsf_par, tr);
            tmem_st_wait();
            if (tr) tr[15] = clock64();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(ds_full);
                if (MODE == W_DKV) mbar_arrive(&vec_empty[sv]);
            }
            if (warp == 0) W_TRACE(n, 5);
        }
        // ---- epilogue: this thread stores columns [cq * D/4, (cq + 1) * D/4) of its row
        constexpr int EC = D / 4;  // 32 or 16 columns
        const bool is_bf16 = p.is_bf16;
        auto store_acc = [&](uint32_t col0, void *outp, const AttnLayout &l, int64_t nrows_valid, float mul) {
            const bool row_ok = row_g < nrows_valid;
            uint16_t *orow = reinterpret_cast<uint16_t *>(outp) + (int64_t)b_idx * l.sb + (int64_t)h_idx * l.sh + (row_ok ? row_g : 0) * l.ss + cq * EC;
            uint32_t a[EC];
            if (ntile > 0) {
                if constexpr (EC == 32) tmem_ld32(lane_addr + col0 + (uint32_t)(cq * EC), a);
                else tmem_ld16(lane_addr + col0 + (uint32_t)(cq * EC), a);
                tmem_ld_wait();
            } else {
#pragma unroll
                for (int i = 0; i < EC; ++i) a[i] = 0u;  // nothing attends to this block: exact zeros
            }
            if (row_ok) {
                uint4 *dst = reinterpret_cast<uint4 *>(orow);
#pragma unroll
                for (int i = 0; i < EC / 8; ++i) {
                    uint32_t wv[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const float2 f = make_float2(__uint_as_float(a[8 * i + 2 * k]) * mul, __uint_as_float(a[8 * i + 2 * k + 1]) * mul);
                        wv[k] = is_bf16 ? w_pack<true>(f) : w_pack<false>(f);
                    }
                    dst[i] = make_uint4(wv[0], wv[1], wv[2], wv[3]);
                }
            }
        };
        if (ntile > 0) {
            mbar_wait(acc_full, 0);
            tc_fence_after();
        }
        if (MODE == W_DKV) {
            store_acc(ACC0_COL, p.out0, p.l0, p.Skv, 1.f);      // dV
            store_acc(ACC1_COL, p.out1, p.l1, p.Skv, p.scale);  // dK
        } else {
            store_acc(ACC1_COL, p.out1, p.l1, p.Sq, p.scale);   // dQ
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == W_MMA) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

struct WLayouts {
    AttnLayout q, k, v, o, dout, dq, dk, dv;
    int64_t H, B;
};

template <int D, int MODE>
void launch_wide_mode(const AttnBwdPlan &a, const WLayouts &L, const float *nlse2, const float *ndelta, int64_t Sq_pad) {
    Runtime &rt = Runtime::get();
    const bool bf16 = a.dtype == KF_BFLOAT16;
    auto map = [&](const void *ptr, int64_t S, const AttnLayout &l) {
        return make_tmap_4d_16bit(ptr, bf16, D, (uint64_t)S, (uint64_t)L.H, (uint64_t)L.B, (uint64_t)l.ss, (uint64_t)l.sh, (uint64_t)l.sb, 64, 128);
    };
    const CUtensorMap x0 = MODE == W_DKV ? map(a.k, a.Skv, L.k) : map(a.q, a.Sq, L.q);
    const CUtensorMap x1 = MODE == W_DKV ? map(a.v, a.Skv, L.v) : map(a.dout, a.Sq, L.dout);
    const CUtensorMap y0 = MODE == W_DKV ? map(a.q, a.Sq, L.q) : map(a.k, a.Skv, L.k);
    const CUtensorMap y1 = MODE == W_DKV ? map(a.dout, a.Sq, L.dout) : map(a.v, a.Skv, L.v);
    WParams p{};
    p.Sq = a.Sq; p.Skv = a.Skv; p.H = (int)L.H;
    p.nlse2 = nlse2; p.ndelta = ndelta; p.Sq_pad = Sq_pad;
    p.out0 = a.dv; p.l0 = L.dv;
    p.out1 = MODE == W_DKV ? a.dk : a.dq;
    p.l1 = MODE == W_DKV ? L.dk : L.dq;
    const double scale = 1.0 / std::sqrt((double)D);
    p.scale = (float)scale;
    p.scale_log2 = (float)(scale * 1.4426950408889634);
    p.nblk = (int)(((MODE == W_DKV ? a.Skv : a.Sq) + 127) / 128);
    p.is_bf16 = bf16;
    {  // heads per scheduling group.  Measured at C3 and at S = 1024 (tools/gpu_attn_bwd_ab.py): unlike the forward, the backward
       // is no faster in weight-major order (groups that fit the L2, or one global list) than head-major, so 1 stays the default
        int64_t hg = 1;
        if (const char *e = std::getenv("KF_ATTN_HG")) hg = std::max(1, std::atoi(e));
        p.hg = (int)std::min<int64_t>(hg, a.BH);
        p.BH = (int)a.BH;
    }
    static const bool want_trace = std::getenv("KF_ATTN_TRACE") != nullptr;
    Scratch trace_buf(want_trace ? 64 * 24 * 8 : 16);
    p.trace = nullptr;
    if (want_trace) {
        rt.memset_async(trace_buf.p, 0, 64 * 24 * 8);
        p.trace = trace_buf.as<long long>();
    }
    constexpr int SMEM = (2 + W_NS0 + W_NS1) * 128 * D * 2 + 2 * 256 * 4 + 256;
    auto kern = attn_bwd_wide_kernel<D, MODE>;
    static bool attr_done = false;
    if (!attr_done) {
        KF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
        attr_done = true;
    }
    const int64_t grid = a.BH * p.nblk;
    KF_CHECK(grid < (int64_t)0x7FFFFFFF);
    kern<<<(unsigned)grid, W_THREADS, SMEM, rt.stream()>>>(x0, x1, y0, y1, p);
    rt.post_launch(MODE == W_DKV ? "attn_bwd_wide_dkv_kernel" : "attn_bwd_wide_dq_kernel");
    if (want_trace) {  // bring-up aid: per-tile gaps between the pipeline events of CTA 0 (clocks)
        static int dumps = 0;
        std::vector<long long> h(64 * 24);
        KF_CUDA(cudaStreamSynchronize(rt.stream()));
        KF_CUDA(cudaMemcpy(h.data(), trace_buf.p, h.size() * 8, cudaMemcpyDeviceToHost));
        if (dumps++ < 2) {
            std::printf("[attn trace %s] tile: loop->vec vec->t0 | A: ld  math  st+wait  arrive | ->t1 | B: ld  math  st+wait  arrive | mma: wait_p  p->ds_wait  wait_ds | period\n", MODE == W_DKV ? "dkv" : "dq");
            for (int n = 0; n < 64 && h[n * 24 + 5] != 0; ++n) {
                const long long *e = &h[n * 24];
                const long long a_st = MODE == W_DKV ? e[14] : e[7];
                std::printf("  %2d: %5lld %5lld | %5lld %5lld %5lld %5lld | %5lld | %5lld %5lld %5lld %5lld | %5lld %5lld %5lld | %5lld\n", n, e[1] - e[0], e[2] - e[1],
                            e[6] - e[2], e[7] - e[6], a_st - e[7], e[3] - a_st, e[4] - e[3], e[12] - e[4], e[13] - e[12], e[15] - e[13], e[5] - e[15],
                            e[9] - e[8], e[10] - e[9], e[11] - e[10], n ? e[5] - h[(n - 1) * 24 + 5] : 0ll);
            }
            std::printf("[attn trace %s] absolute stamps relative to the MMA warp passing wait_p(n): A1 issue(n) | t0_full(n+1) seen, P(n+1) arrive | t1_full(n+1) seen, dS(n+1) arrive | wait_p(n+1) passed\n", MODE == W_DKV ? "dkv" : "dq");
            for (int n = 8; n < 14 && h[(n + 1) * 24 + 5] != 0; ++n) {
                const long long *e = &h[n * 24], *f = &h[(n + 1) * 24];
                const long long b = e[9];
                std::printf("  %2d: %5lld | %5lld %5lld | %5lld %5lld | %5lld   mma: first half issued %5lld, at ds wait %5lld, second half issued %5lld\n", n, e[11] - b, f[2] - b, f[3] - b, f[4] - b, f[5] - b, f[9] - b,
                            e[16] - b, e[10] - b, e[18] - b);
                std::printf("      producer: Y0(n+1) issued %5lld, Y1(n+1) issued %5lld\n", f[22] - b, f[23] - b);
            }
        }
    }
}

}  // namespace

bool launch_attention_bwd_wide(const AttnBwdPlan &a) {
    if (a.dtype != KF_HALF && a.dtype != KF_BFLOAT16) return false;
    if (a.D != 64 && a.D != 128) return false;
    if (a.Sq < 1 || a.Skv < 1 || a.BH < 1) return false;
    WLayouts L;
    const bool dense = a.H <= 0;
    L.H = dense ? a.BH : a.H;
    if (a.BH % L.H != 0) return false;
    L.B = a.BH / L.H;
    if (L.B >= 65536 || L.H >= 65536) return false;
    const AttnLayout dq_d{L.H * a.Sq * a.D, a.Sq * a.D, a.D}, dkv_d{L.H * a.Skv * a.D, a.Skv * a.D, a.D};
    L.q = dense ? dq_d : a.lq; L.k = dense ? dkv_d : a.lk; L.v = dense ? dkv_d : a.lv; L.o = dense ? dq_d : a.lo;
    L.dout = dense ? dq_d : a.ldo; L.dq = dense ? dq_d : a.ldq; L.dk = dense ? dkv_d : a.ldk; L.dv = dense ? dkv_d : a.ldv;
    auto ok = [&](const void *p, const AttnLayout &l) {
        return reinterpret_cast<uintptr_t>(p) % 16 == 0 && l.sb % 8 == 0 && l.sh % 8 == 0 && l.ss % 8 == 0 && l.ss >= a.D;
    };
    if (!ok(a.q, L.q) || !ok(a.k, L.k) || !ok(a.v, L.v) || !ok(a.out, L.o) || !ok(a.dout, L.dout) || !ok(a.dq, L.dq) || !ok(a.dk, L.dk) ||
        !ok(a.dv, L.dv))
        return false;
    Runtime &rt = Runtime::get();
    const bool bf16 = a.dtype == KF_BFLOAT16;
    const int64_t Sq_pad = (a.Sq + 127) / 128 * 128;
    const int64_t rows_pad = a.BH * Sq_pad;
    Scratch ndelta((size_t)rows_pad * 4), nlse2((size_t)rows_pad * 4);
    const int64_t rows_per_cta = 256 / (a.D / 8);
    KF_CHECK((rows_pad + rows_per_cta - 1) / rows_per_cta < (int64_t)0x7FFFFFFF);
    const unsigned pgrid = (unsigned)((rows_pad + rows_per_cta - 1) / rows_per_cta);
    if (bf16)
        w_prep_kernel<__nv_bfloat16><<<pgrid, 256, 0, rt.stream()>>>((const __nv_bfloat16 *)a.out, (const __nv_bfloat16 *)a.dout, (const float *)a.lse,
                                                                     ndelta.as<float>(), nlse2.as<float>(), rows_pad, a.Sq, Sq_pad, (int)L.H, (int)a.D, L.o,
                                                                     L.dout);
    else
        w_prep_kernel<__half><<<pgrid, 256, 0, rt.stream()>>>((const __half *)a.out, (const __half *)a.dout, (const float *)a.lse, ndelta.as<float>(),
                                                              nlse2.as<float>(), rows_pad, a.Sq, Sq_pad, (int)L.H, (int)a.D, L.o, L.dout);
    rt.post_launch("attn_bwd_prep_kernel");
    if (a.D == 64) {
        launch_wide_mode<64, W_DKV>(a, L, nlse2.as<float>(), ndelta.as<float>(), Sq_pad);
        launch_wide_mode<64, W_DQ>(a, L, nlse2.as<float>(), ndelta.as<float>(), Sq_pad);
    } else {
        launch_wide_mode<128, W_DKV>(a, L, nlse2.as<float>(), ndelta.as<float>(), Sq_pad);
        launch_wide_mode<128, W_DQ>(a, L, nlse2.as<float>(), ndelta.as<float>(), Sq_pad);
    }
    return true;
}

}  // namespace kf
