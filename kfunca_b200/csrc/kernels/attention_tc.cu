// Causal attention forward on tcgen05 tensor cores (bf16 / fp16, head size 64 or 128), hand-written PTX.
// The reference's forward is fp32 scalar-FMA only and walks every KV block (src/device/utils/causal_attention.h:73-208;
// SURVEY F2); this kernel is its B200-native replacement for the 16-bit dtypes the north star asks for.
//
// One CTA = a PAIR of 128-row query tiles of one (batch, head); KV consumed in blocks of 128 keys up to each tile's diagonal.
//   warp 9    : TMA producer (Q tiles once; K_j, V_j through a 4-slot ring; 128B-swizzled tiles)
//   warp 8    : MMA issuer   S_t = Q_t K_j^T  (128x128xD, both operands K-major in smem, fp32 accumulator in TMEM)
//                            O_t += P_t V_j   (128xDx128, P read from TENSOR MEMORY, V MN-major in smem)
//               issue order  S_0, S_1, then per block { P_0 V + next S_0 ; P_1 V + next S_1 } so the tensor pipe works on
//               one tile while the softmax warps of the other tile run (ping-pong).  P_t is handed over in two 64-key halves, so
//               the first four k-steps of P V run under the second half's exponentials.
//               (Tried and rejected, 1082 -> 811 TFLOP/s: issuing the upper 64 columns of the next S early as an N = 64 MMA.
//               An M128 N64 K16 MMA still reads the whole 4 KB A slice from shared memory, so two of them cost 1.5x one N = 128.)
//   warps 0-3 / 4-7 : softmax of tile 0 / 1.  thread = query row: one tcgen05.ld of the whole S row (128 fp32 registers),
//               max, exp2 in packed fp32x2 FMAs, P (16-bit) written back over S in TMEM (tcgen05.st).  The running max is
//               LAZY: it only moves (and O, l are only rescaled) when the row max grew by more than 2^8, which is exact
//               because the final 1/l normalisation uses the same reference.  Final 1/l scaling + 16-byte stores + row LSE.
// TMEM: 512 columns = S_0 | S_1 | O_0 | O_1.  Shared memory: 2 Q tiles + 4 K/V slots (192 KB at D = 128), one CTA per SM.
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <algorithm>
#include <cstdio>

#include "ew_common.cuh"
#include "tc_common.cuh"

namespace kf {
using namespace tc;

constexpr int FA_BQ = 128, FA_BKV = 128;
// warps 0-3 softmax(tile 0), 4-7 softmax(tile 1), 8 MMA issuer, 9 TMA producer, 10-11 idle.  12 warps = 3 whole warpgroups so that
// setmaxnreg can move registers from the control warpgroup (40 / thread) to the two softmax warpgroups (232 / thread): a
// softmax thread holds a whole 128-column S row plus the packed P chunk, which does not fit the 168 the launch gives everyone.
constexpr int FA_NSTAGE = 4;     // K/V ring slots (K_0, V_0, K_1, V_1, ... in consumption order)

struct AttnTcParams {
    int64_t BH, Sq, Skv;
    void *out;
    float *lse;
    float scale_log2;  // softmax scale * log2(e)
    int npairs;        // 256-row query pairs per (b, h)
    int is_bf16;
    int H;             // heads per batch entry: (b, h) = (bh / H, bh % H) for the 4-D tensor maps and the output layout
    AttnLayout lo;     // layout of `out`
    int hg;            // heads per scheduling group (see fwd_decode_item)
    unsigned int *work_counter;  // persistent kernel: items claimed beyond the first one of every CTA (zeroed before the launch)
    int stale;         // 1: blocks after the first take their exponentials relative to the running reference (fwd_softmax_block_stale)
    long long *trace;  // KF_ATTN_TRACE=1 (persistent kernel): clock64() stamps of CTA 0, [256 blocks][16] + [64 items][4] at 4096; else null
};

// Work item w -> (batch-head, query pair).  Heads are taken in groups of p.hg (a group's K and V take ~ an eighth of the L2); inside a group
// the items run length-major: every head's longest pair first, then the next length ... — longest-processing-time order for the
// load balance, at most hg heads' K / V live at a time for the L2.  hg = BH: one global length-major list; hg = 1: head-major.
__device__ __forceinline__ void fwd_decode_item(const AttnTcParams &p, const int w, int &bh, int &pr) {
    const int per_group = p.hg * p.npairs;
    const int g = w / per_group, idx = w - g * per_group;
    const int heads = min(p.hg, (int)p.BH - g * p.hg);  // the last group may be partial
    const int lvl = idx / heads;
    bh = g * p.hg + (idx - lvl * heads);
    pr = p.npairs - 1 - lvl;
}

__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
          "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]),
          "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]),
          "r"(r[30]), "r"(r[31])
        : "memory");
}

__device__ __forceinline__ uint32_t pack16(float a, float b, int is_bf16) {
    if (is_bf16) {
        __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
        return *reinterpret_cast<uint32_t *>(&h);
    }
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t *>(&h);
}

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
template <bool BF16>
__device__ __forceinline__ uint32_t pack16t(float2 v) {
    if (BF16) {
        __nv_bfloat162 h = __float22bfloat162_rn(v);
        return *reinterpret_cast<uint32_t *>(&h);
    }
    __half2 h = __float22half2_rn(v);
    return *reinterpret_cast<uint32_t *>(&h);
}
// 2^x for a pair on the FMA pipe (the MUFU unit does 16 ex2 / clk / SM, exactly as many cycles per 128x128 tile as the two
// tile MMAs take, so a share of the exponentials is moved off it): round-to-nearest split x = j + f via the 1.5*2^23 magic
// add, degree-3 minimax of 2^f on [-1/2, 1/2] (max rel. error 7.5e-5, far below the 16-bit P rounding), then j is added
// into the exponent field.  x is clamped at -126 so the exponent arithmetic cannot wrap.
__device__ __forceinline__ float2 ex2_poly2(float2 x) {
    x.x = fmaxf(x.x, -126.f);
    x.y = fmaxf(x.y, -126.f);
    const float2 magic = make_float2(12582912.f, 12582912.f);
    const float2 t = __fadd2_rn(x, magic);
    const float2 r = __fadd2_rn(t, make_float2(-12582912.f, -12582912.f));
    const float2 f = __ffma2_rn(r, make_float2(-1.f, -1.f), x);
    float2 q = __ffma2_rn(f, make_float2(0.0551716648042202f, 0.0551716648042202f), make_float2(0.2426111251115799f, 0.2426111251115799f));
    q = __ffma2_rn(q, f, make_float2(0.6932609677314758f, 0.6932609677314758f));
    q = __ffma2_rn(q, f, make_float2(0.9999280571937561f, 0.9999280571937561f));
    return make_float2(__uint_as_float(__float_as_uint(q.x) + (__float_as_uint(t.x) << 23)),
                       __uint_as_float(__float_as_uint(q.y) + (__float_as_uint(t.y) << 23)));
}


// Softmax of one 128-key block for one thread.  NH = 1 (default): thread = one query row (128 columns).  NH = 2 (experiment): TWO
// threads per row (warps w and w + 4 of the tile own columns [0, 64) and [64, 128)), so every dependent step (TMEM load, max
// tree, exp, TMEM store) is half as long.  The halves agree on the row maximum through shared memory and one 64-thread named
// barrier per block; that barrier also orders "both halves have loaded S" before either overwrites S with P.
// S row from tensor memory, [mask], row max, lazy rescale of O / l, P = exp2(S c - m) written back over S as 16-bit (each 64-key
// half is signalled on its own mbarrier `p_half[..]` as soon as it is stored), running
// row sum (per half).  `lim`: columns > lim (relative to this half) are masked.  POLY: of every 8 column pairs, this many take
// the FMA-pipe exp2.
template <int D, bool BF16, bool MASKED, int POLY, int NH, bool TR = false>
__device__ __forceinline__ void fwd_softmax_block(const uint32_t s_addr, const uint32_t p_addr, const uint32_t o_addr, const float sc, const int lim,
                                                  const bool first, float &m_ref, float &l_run, float *xch_mine, const float *xch_peer,
                                                  const int bar_id, uint64_t *p_half, long long *tr = nullptr) {
    constexpr int NC = 4 / NH;  // 32-column chunks per thread
    uint32_t s[NC][32];
#pragma unroll
    for (int c = 0; c < NC; ++c) tmem_ld32(s_addr + (uint32_t)(c * 32), s[c]);
    tmem_ld_wait();
    if (TR && tr) tr[0] = clock64();
    // Diagonal / ragged blocks: 32-column chunks that are masked for EVERY row of this warp (warp q of the diagonal block keeps
    // chunks 0 .. q only) take no mask, no maximum and no exponentials: their P is zero.  `nact` is warp-uniform.
    const int nact = (MASKED && NH == 1) ? ((__reduce_max_sync(0xffffffffu, lim) + 32) >> 5) : NC;
    if (MASKED) {
#pragma unroll
        for (int c = 0; c < NC; ++c)
            if (c < nact) {
#pragma unroll
                for (int i = 0; i < 32; ++i)
                    if (c * 32 + i > lim) s[c][i] = 0xff800000u;  // -inf
            }
    }
    float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
    for (int c = 0; c < NC; ++c)
        if (c < nact) {
#pragma unroll
            for (int i = 0; i < 32; i += 8) {
                mx0 = fmaxf(mx0, fmaxf(__uint_as_float(s[c][i]), __uint_as_float(s[c][i + 1])));
                mx1 = fmaxf(mx1, fmaxf(__uint_as_float(s[c][i + 2]), __uint_as_float(s[c][i + 3])));
                mx2 = fmaxf(mx2, fmaxf(__uint_as_float(s[c][i + 4]), __uint_as_float(s[c][i + 5])));
                mx3 = fmaxf(mx3, fmaxf(__uint_as_float(s[c][i + 6]), __uint_as_float(s[c][i + 7])));
            }
        }
    float mx = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
    if (NH == 2) {
        *xch_mine = mx;
        asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory");
        mx = fmaxf(mx, *xch_peer);
    }
    const float m_new = fmaxf(m_ref, mx * sc);
    // lazy reference max: only move it (and rescale O, l) when the row max grew by more than 2^8
    const bool grow = first ? true : (m_new - m_ref > 8.f);
    if (!first && __any_sync(0xffffffffu, grow)) {
        const float f = grow ? ((m_ref == -INFINITY) ? 0.f : ex2_approx(m_ref - m_new)) : 1.f;
        l_run *= f;
#pragma unroll 1
        for (int c = 0; c < D / 32 / NH; ++c) {
            uint32_t orr[32];
            tmem_ld32(o_addr + (uint32_t)(c * 32), orr);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) orr[i] = __float_as_uint(__uint_as_float(orr[i]) * f);
            tmem_st32(o_addr + (uint32_t)(c * 32), orr);
        }
    }
    if (grow) m_ref = m_new;
    const float m_use = (m_ref == -INFINITY) ? 0.f : m_ref;
    if (TR && tr) tr[1] = clock64() + (long long)(m_use == 12345.f);  // (depends on the maximum: the stamp cannot be hoisted)
    const float2 sc2 = make_float2(sc, sc), nm2 = make_float2(-m_use, -m_use);
    float2 rs2 = make_float2(0.f, 0.f), rs3 = make_float2(0.f, 0.f);
#pragma unroll
    for (int c = 0; c < NC; c += 2) {
        uint32_t pk[32];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            if (c + h < nact) {
#pragma unroll
                for (int i = 0; i < 32; i += 2) {
                    float2 x = __ffma2_rn(make_float2(__uint_as_float(s[c + h][i]), __uint_as_float(s[c + h][i + 1])), sc2, nm2);
                    if (((i >> 1) & 7) < POLY) {
                        x = ex2_poly2(x);
                    } else {
                        x.x = ex2_approx(x.x);
                        x.y = ex2_approx(x.y);
                    }
                    if (i & 2) rs3 = __fadd2_rn(rs3, x);  // two accumulators: half the dependent-add chain
                    else rs2 = __fadd2_rn(rs2, x);
                    pk[h * 16 + (i >> 1)] = pack16t<BF16>(x);
                }
            } else {
#pragma unroll
                for (int i = 0; i < 16; ++i) pk[h * 16 + i] = 0u;
            }
        }
        if (TR && tr) tr[2 + c] = clock64() + (long long)(pk[31] == 0x12345u);
        tmem_st32(p_addr + (uint32_t)(c * 16), pk);
        // hand this 64-key half of P to the MMA warp right away: the first four k-steps of P V run under the second half's exps
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if ((threadIdx.x & 31) == 0) mbar_arrive(p_half + (NH == 1 ? c / 2 : 0));
        if (TR && tr) tr[3 + c] = clock64();
    }
    l_run += (rs2.x + rs2.y) + (rs3.x + rs3.y);
}

// Blocks after the first, one thread per row: the exponentials do not wait for this block's row maximum.  The running reference m_ref
// is already known when S arrives, so each 64-key half is exponentiated relative to it straight away; the half's own maximum is formed
// on the side and only CHECKED (max <= m_ref + 2^8, the same lazy-rescale bound as the classic path).  The TMEM load of the second
// half runs under the first half's arithmetic.  When a half fails the check (rare: a key much larger than everything before it) the
// warp takes the slow path for that half: wait until every P V issued into O so far has completed (half 1: the extra `pv0_done`
// commit), rescale O and l to the new reference, redo the half's exponentials from the S registers.  Nothing stored to TMEM or handed
// to the MMA is ever computed with an overflowing argument: the check precedes the store.
template <int D, bool BF16, bool MASKED, int POLY>
__device__ __forceinline__ void fwd_softmax_block_stale(const uint32_t s_addr, const uint32_t p_addr, const uint32_t o_addr, const float sc, const int lim,
                                                        float &m_ref, float &l_run, uint64_t *p_half, uint64_t *pv0_done, const uint32_t pv0_parity) {
    uint32_t s[4][32];
    tmem_ld32(s_addr, s[0]);
    tmem_ld32(s_addr + 32, s[1]);
    tmem_ld_wait();
    tmem_ld32(s_addr + 64, s[2]);  // in flight under the first half
    tmem_ld32(s_addr + 96, s[3]);
    const float2 sc2 = make_float2(sc, sc);
    float2 rs2 = make_float2(0.f, 0.f), rs3 = make_float2(0.f, 0.f);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        if (h == 1) tmem_ld_wait();
        if (MASKED) {
#pragma unroll
            for (int cc = 0; cc < 2; ++cc)
#pragma unroll
                for (int i = 0; i < 32; ++i)
                    if ((2 * h + cc) * 32 + i > lim) s[2 * h + cc][i] = 0xff800000u;  // -inf
        }
        float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
        for (int cc = 0; cc < 2; ++cc)
#pragma unroll
            for (int i = 0; i < 32; i += 8) {
                mx0 = fmaxf(mx0, fmaxf(__uint_as_float(s[2 * h + cc][i]), __uint_as_float(s[2 * h + cc][i + 1])));
                mx1 = fmaxf(mx1, fmaxf(__uint_as_float(s[2 * h + cc][i + 2]), __uint_as_float(s[2 * h + cc][i + 3])));
                mx2 = fmaxf(mx2, fmaxf(__uint_as_float(s[2 * h + cc][i + 4]), __uint_as_float(s[2 * h + cc][i + 5])));
                mx3 = fmaxf(mx3, fmaxf(__uint_as_float(s[2 * h + cc][i + 6]), __uint_as_float(s[2 * h + cc][i + 7])));
            }
        const float mxh = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)) * sc;
        const bool grow = mxh > m_ref + 8.f;  // false for NaN-free data whenever the half stays within 2^8 of the reference
        if (__any_sync(0xffffffffu, grow)) {
            // ---- slow path: move the reference (growing rows only), rescale O and l once every earlier P V has completed
            const float m_new = grow ? mxh : m_ref;
            if (h == 1) {
                mbar_wait(pv0_done, pv0_parity);
                tc_fence_after();
            }
            const float f = grow ? ((m_ref == -INFINITY) ? 0.f : ex2_approx(m_ref - m_new)) : 1.f;
            l_run *= f;
            rs2.x *= f; rs2.y *= f; rs3.x *= f; rs3.y *= f;  // this block's first-half sums were taken relative to the old reference
#pragma unroll 1
            for (int c = 0; c < D / 32; ++c) {
                uint32_t orr[32];
                tmem_ld32(o_addr + (uint32_t)(c * 32), orr);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i) orr[i] = __float_as_uint(__uint_as_float(orr[i]) * f);
                tmem_st32(o_addr + (uint32_t)(c * 32), orr);
            }
            tmem_st_wait();
            m_ref = m_new;
        }
        const float m_use = (m_ref == -INFINITY) ? 0.f : m_ref;
        const float2 nm2 = make_float2(-m_use, -m_use);
        uint32_t pk[32];
#pragma unroll
        for (int cc = 0; cc < 2; ++cc)
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
                float2 x = __ffma2_rn(make_float2(__uint_as_float(s[2 * h + cc][i]), __uint_as_float(s[2 * h + cc][i + 1])), sc2, nm2);
                if (((i >> 1) & 7) < POLY) {
                    x = ex2_poly2(x);
                } else {
                    x.x = ex2_approx(x.x);
                    x.y = ex2_approx(x.y);
                }
                if (i & 2) rs3 = __fadd2_rn(rs3, x);
                else rs2 = __fadd2_rn(rs2, x);
                pk[cc * 16 + (i >> 1)] = pack16t<BF16>(x);
            }
        tmem_st32(p_addr + (uint32_t)(h * 32), pk);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if ((threadIdx.x & 31) == 0) mbar_arrive(p_half + h);
    }
    l_run += (rs2.x + rs2.y) + (rs3.x + rs3.y);
}

// Two threads per query row WITHOUT a barrier in front of the exponentials (KF_ATTN_SPLIT=2).  Each thread owns 64 of the 128 keys of
// the block.  It takes its exponentials relative to a reference it knows on its own: the running reference m_ref, or — when its own
// 64-key maximum exceeds m_ref by more than 2^8 (always in the first block) — that maximum.  The two halves of a row exchange their
// maxima through shared memory AFTER P has been stored (one 64-thread named barrier, normally already satisfied), agree on the new
// reference, and only in the rare case that a thread's own reference differs from it rescale their stored 16-bit P (a factor <= 1) and,
// when the reference moved, O and l.  Every intermediate stays finite: P <= 2^8 relative to m_ref, <= 1 relative to an own maximum.
// P of keys [0, 64) lands on TMEM columns [0, 32) of the tile's S region and P of keys [64, 128) on [64, 96): a thread only overwrites
// S columns it has read itself.  POLY of every 8 column pairs take the FMA-pipe exp2.
template <int D, bool BF16, bool MASKED, int POLY>
__device__ __forceinline__ void fwd_softmax_block_split(const uint32_t s_addr, const uint32_t o_addr, const float sc, const int lim, const bool first,
                                                        float &m_ref, float &l_run, float *xch_mine, const float *xch_peer, const int bar_id,
                                                        uint64_t *p_bar) {
    uint32_t s[2][32];
    tmem_ld32(s_addr, s[0]);
    tmem_ld32(s_addr + 32, s[1]);
    tmem_ld_wait();
    if (MASKED) {
#pragma unroll
        for (int c = 0; c < 2; ++c)
#pragma unroll
            for (int i = 0; i < 32; ++i)
                if (c * 32 + i > lim) s[c][i] = 0xff800000u;  // -inf
    }
    float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
    for (int c = 0; c < 2; ++c)
#pragma unroll
        for (int i = 0; i < 32; i += 8) {
            mx0 = fmaxf(mx0, fmaxf(__uint_as_float(s[c][i]), __uint_as_float(s[c][i + 1])));
            mx1 = fmaxf(mx1, fmaxf(__uint_as_float(s[c][i + 2]), __uint_as_float(s[c][i + 3])));
            mx2 = fmaxf(mx2, fmaxf(__uint_as_float(s[c][i + 4]), __uint_as_float(s[c][i + 5])));
            mx3 = fmaxf(mx3, fmaxf(__uint_as_float(s[c][i + 6]), __uint_as_float(s[c][i + 7])));
        }
    const float mx = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)) * sc;  // this half's maximum in the exp2 domain
    *xch_mine = mx;
    const float r_own = (mx > m_ref + 8.f) ? mx : m_ref;  // m_ref = -inf in the first block: then the own maximum (or -inf when fully masked)
    const float m_use = (r_own == -INFINITY) ? 0.f : r_own;
    const float2 sc2 = make_float2(sc, sc), nm2 = make_float2(-m_use, -m_use);
    float2 rs2 = make_float2(0.f, 0.f), rs3 = make_float2(0.f, 0.f);
    uint32_t pk[32];
#pragma unroll
    for (int c = 0; c < 2; ++c)
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
            float2 x = __ffma2_rn(make_float2(__uint_as_float(s[c][i]), __uint_as_float(s[c][i + 1])), sc2, nm2);
            if (((i >> 1) & 7) < POLY) {
                x = ex2_poly2(x);
            } else {
                x.x = ex2_approx(x.x);
                x.y = ex2_approx(x.y);
            }
            if (i & 2) rs3 = __fadd2_rn(rs3, x);
            else rs2 = __fadd2_rn(rs2, x);
            pk[c * 16 + (i >> 1)] = pack16t<BF16>(x);
        }
    tmem_st32(s_addr, pk);  // over this thread's own S columns [0, 32) of its 64
    float rs = (rs2.x + rs2.y) + (rs3.x + rs3.y);
    asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory");
    const float m_blk = fmaxf(mx, *xch_peer);
    const bool grow = m_blk > m_ref + 8.f;
    const float m_new = grow ? m_blk : m_ref;  // m_blk > m_ref whenever grow
    const bool fix = r_own != m_new;           // this thread's exponentials used another reference than the row's new one
    if (__any_sync(0xffffffffu, fix)) {
        const float f = fix ? ((r_own == -INFINITY) ? 0.f : ex2_approx(r_own - m_new)) : 1.f;
        rs *= f;
        tmem_st_wait();
        uint32_t q[32];
        tmem_ld32(s_addr, q);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            float2 v;
            if (BF16) v = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162 *>(&q[i]));
            else v = __half22float2(*reinterpret_cast<__half2 *>(&q[i]));
            q[i] = pack16t<BF16>(make_float2(v.x * f, v.y * f));
        }
        tmem_st32(s_addr, q);
    }
    if (!first && __any_sync(0xffffffffu, grow)) {
        const float f = grow ? ((m_ref == -INFINITY) ? 0.f : ex2_approx(m_ref - m_new)) : 1.f;
        l_run *= f;
#pragma unroll 1
        for (int c = 0; c < D / 64; ++c) {
            uint32_t orr[32];
            tmem_ld32(o_addr + (uint32_t)(c * 32), orr);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) orr[i] = __float_as_uint(__uint_as_float(orr[i]) * f);
            tmem_st32(o_addr + (uint32_t)(c * 32), orr);
        }
    }
    m_ref = m_new;
    l_run += rs;
    tmem_st_wait();
    tc_fence_before();
    __syncwarp();
    if ((threadIdx.x & 31) == 0) mbar_arrive(p_bar);
}

// Epilogue of one query tile, thread = row: O / l -> 16-bit -> a 128B-swizzled shared-memory tile -> TMA store.  A thread-per-row
// st.global of 16 B per lane touches 32 different 128 B lines per instruction (measured 3300 - 4000 clk per 128 x 128 tile, LSU
// bound); staged through shared memory the rows leave as full lines and the warps are free as soon as the tile is staged.
// `stage` holds SA 64-column atoms (128 rows x 128 B each): the whole tile at once (SA = D / 64, the per-CTA kernel stages in the
// tile's own Q buffer, dead after the tile's last S MMA) or one atom per pass (the persistent kernel's 16 KB per tile).  WAIT_PREV:
// an earlier store may still be reading `stage`.  Rows >= Sq are clipped by the tensor map.  `o_free`: arrived on (one lane per
// warp) once the last O columns are in registers.  The issuer thread is the tile's thread 0; bulk-group completion is per thread.
template <int D, int SA, bool WAIT_PREV>
__device__ __forceinline__ void fwd_store_tile(const uint32_t o_addr, const float inv_l, const bool is_bf16, unsigned char *stage,
                                               const CUtensorMap *tmap_o, const int row0, const int h_idx, const int b_idx, const int r,
                                               const int bar_id, uint64_t *o_free, long long *tr = nullptr) {
    constexpr int ATOMS = D / 64, NC = D / 32;
    const bool issuer = r == 0;
    uint32_t orr[NC][32];  // the whole O row: every load is in flight before the first wait, and O is released at once
#pragma unroll
    for (int c = 0; c < NC; ++c) tmem_ld32(o_addr + (uint32_t)(c * 32), orr[c]);
    tmem_ld_wait();
    if (o_free != nullptr) {  // O is in registers: the next item's first P V may overwrite it
        tc_fence_before();
        __syncwarp();
        if ((threadIdx.x & 31) == 0) mbar_arrive(o_free);
    }
    const float2 il2 = make_float2(inv_l, inv_l);
#pragma unroll
    for (int a0 = 0; a0 < ATOMS; a0 += SA) {
        if (WAIT_PREV || a0 > 0) {
            if (issuer) tma_store_wait_read<0>();
            asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
        }
        if (tr) tr[(a0 / SA) * 4 + 0] = clock64();
#pragma unroll
        for (int c = a0 * 2; c < (a0 + SA) * 2; ++c) {
            unsigned char *dst = stage + ((c >> 1) - a0) * (128 * 128) + r * 128;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                uint32_t wd[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float2 v = __fmul2_rn(make_float2(__uint_as_float(orr[c][8 * i + 2 * k]), __uint_as_float(orr[c][8 * i + 2 * k + 1])), il2);
                    wd[k] = is_bf16 ? pack16t<true>(v) : pack16t<false>(v);
                }
                const int chunk = (c & 1) * 4 + i;  // 16-byte chunk of the 128 B row; the 128B swizzle XORs it with the row's low 3 bits
                *reinterpret_cast<uint4 *>(dst + ((chunk ^ (r & 7)) << 4)) = make_uint4(wd[0], wd[1], wd[2], wd[3]);
            }
        }
        if (tr) tr[(a0 / SA) * 4 + 1] = clock64();
        fence_proxy_async();  // generic-proxy stores -> visible to the bulk copy's async-proxy reads
        asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
        if (tr) tr[(a0 / SA) * 4 + 2] = clock64();
        if (issuer) {
#pragma unroll
            for (int a = 0; a < SA; ++a) tma_store_4d(tmap_o, stage + a * (128 * 128), (a0 + a) * 64, row0, h_idx, b_idx);
            tma_store_commit();
        }
        if (tr) tr[(a0 / SA) * 4 + 3] = clock64();
    }
}

// One CTA = two 128-row query tiles (a 256-row "pair") of one (batch, head); the two tiles ping-pong on the
// tensor pipe: while softmax warps work on S of tile t, the MMA warp runs P V + the next Q K^T of tile 1-t.
// Warp roles: [0, 8 NH) softmax (tile = w / (4 NH), column half = (w / 4) % NH, TMEM lane quarter = w % 4), then the MMA issuer,
// the TMA producer and two idle warps that complete the control warpgroup (setmaxnreg works on whole warpgroups).
template <int NH>
struct FaCfg {
    static constexpr int NSW = 8 * NH;             // softmax warps
    static constexpr int THREADS = (NSW + 4) * 32;  // 384 (NH = 1) or 640 (NH = 2): whole warpgroups, setmaxnreg moves registers
    // NH = 1: launch at 168, control warpgroup 88, softmax 208 (a thread holds a whole 128-column S row).
    // NH = 2: launch at 96 (640 threads), control warpgroup 64 (frees 128 x 32), the four softmax warpgroups 104 (takes 512 x 8).
    static constexpr int REG_CTRL = NH == 1 ? 88 : 64;
    static constexpr int REG_SOFTMAX = NH == 1 ? 208 : 104;
};

template <int D, int POLY, int NH>
__device__ __forceinline__ void attn_fwd_tc_body(const CUtensorMap &tmap_q, const CUtensorMap &tmap_k, const CUtensorMap &tmap_v,
                                                 const CUtensorMap &tmap_o, const AttnTcParams &p) {
    constexpr int ATOMS = D / 64;              // 64-element (128 B) swizzle atoms along the head dimension
    constexpr int TILE_BYTES = 128 * D * 2;    // one 128 x D 16-bit tile
    constexpr int ATOM_BYTES = 128 * 128;      // 128 rows x 128 B
    constexpr uint32_t TMEM_COLS = 512;        // S0 | S1 | O0 | O1 (P_t aliases the first 64 columns of S_t)
    constexpr uint32_t O_COL = 256;
    constexpr int NS = FA_NSTAGE;
    constexpr int NSW = FaCfg<NH>::NSW, W_MMA = NSW, W_TMA = NSW + 1;
    extern __shared__ unsigned char smem_raw[];
    unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    unsigned char *sQ = smem;                    // 2 tiles
    unsigned char *sKV = smem + 2 * TILE_BYTES;  // NS tiles
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + (2 + NS) * TILE_BYTES);
    uint64_t *q_full = bars + 0;
    uint64_t *kv_full = bars + 1, *kv_empty = bars + 1 + NS;
    uint64_t *s_full = bars + 1 + 2 * NS;  // [2]
    uint64_t *p_full = s_full + 2;         // [2 tiles][2 key halves]
    uint64_t *pv0_done = p_full + 4;       // [2 tiles]: the first-half P V MMAs of the block have completed (slow path of the stale softmax)
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(pv0_done + 2);
    float *xch = reinterpret_cast<float *>(smem + (2 + NS) * TILE_BYTES + 256);  // [tile][half][parity][row] row-max / row-sum exchange (NH = 2)

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int bh, pr;
    fwd_decode_item(p, (int)blockIdx.x, bh, pr);  // heaviest (longest KV range) pairs of a head group first
    const int b_idx = bh / p.H, h_idx = bh % p.H;
    const int q0 = pr * 2 * FA_BQ;
    auto blocks_of = [&](int t) {
        const int64_t q0t = (int64_t)q0 + t * FA_BQ;
        const int64_t kv_end = min((int64_t)p.Skv, q0t + FA_BQ);
        return q0t < p.Sq ? (int)((kv_end + FA_BKV - 1) / FA_BKV) : 0;
    };
    const int nblk0 = blocks_of(0), nblk1 = blocks_of(1);  // scalars: a dynamically indexed local array would live in local memory
    const int nmax = max(nblk0, nblk1);

    if (warp == W_TMA && lane == 0) {
        prefetch_tmap(&tmap_q);
        prefetch_tmap(&tmap_k);
        prefetch_tmap(&tmap_v);
        mbar_init(q_full, 1);
        for (int s = 0; s < NS; ++s) {
            mbar_init(&kv_full[s], 1);
            mbar_init(&kv_empty[s], 1);
        }
        for (int t = 0; t < 2; ++t) {
            mbar_init(&s_full[t], 1);
            mbar_init(&p_full[2 * t], 4);
            mbar_init(&p_full[2 * t + 1], 4);
            mbar_init(&pv0_done[t], 1);
        }
        fence_barrier_init();
    }
    if (warp == W_MMA) tmem_alloc(tmem_slot, TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp >= NSW) {
      // control warpgroup (last two warps idle): the register hand-over must sit INSIDE the role branch, ptxas budgets the
      // code that follows a setmaxnreg by the value it names
      asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(FaCfg<NH>::REG_CTRL));
      if (warp == W_TMA) {
        // ===================================================== TMA producer
        if (lane == 0) {
            const int ntile_q = nblk1 > 0 ? 2 : 1;
            mbar_arrive_expect_tx(q_full, ntile_q * TILE_BYTES);
            for (int t = 0; t < ntile_q; ++t)
#pragma unroll
                for (int a = 0; a < ATOMS; ++a) tma_load_4d(sQ + t * TILE_BYTES + a * ATOM_BYTES, &tmap_q, q_full, a * 64, q0 + t * FA_BQ, h_idx, b_idx);
            int s = 0;
            uint32_t ph = 0;
            for (int i = 0; i < 2 * nmax; ++i) {  // K_0, V_0, K_1, V_1, ...
                const int kv0 = (i >> 1) * FA_BKV;
                const CUtensorMap *tm = (i & 1) ? &tmap_v : &tmap_k;
                mbar_wait(&kv_empty[s], ph ^ 1);
                mbar_arrive_expect_tx(&kv_full[s], TILE_BYTES);
#pragma unroll
                for (int a = 0; a < ATOMS; ++a) tma_load_4d(sKV + s * TILE_BYTES + a * ATOM_BYTES, tm, &kv_full[s], a * 64, kv0, h_idx, b_idx);
                if (++s == NS) {
                    s = 0;
                    ph ^= 1;
                }
            }
        }
        __syncwarp();
      } else if (warp == W_MMA) {
        // ===================================================== MMA issuer (converged warp, elected lane issues: see umma_f16_p)
        const bool leader = elect_one();
        {
            const int fmt = p.is_bf16 ? 1 : 0;
            const uint32_t idesc_s = make_idesc_f16(fmt, 0, 0, FA_BQ, FA_BKV);
            const uint32_t idesc_pv = make_idesc_f16(fmt, 0, 1, FA_BQ, D);
            const uint32_t q_addr = smem_u32(sQ), kv_addr = smem_u32(sKV);
            auto issue_s = [&](int t, uint32_t k_addr) {
#pragma unroll
                for (int kk = 0; kk < D / 16; ++kk) {
                    const uint32_t off = (uint32_t)((kk >> 2) * ATOM_BYTES + (kk & 3) * 32);
                    umma_f16_p(tmem_base + (uint32_t)(t * 128), make_sw128_desc(q_addr + t * TILE_BYTES + off, 0, 1024),
                               make_sw128_desc(k_addr + off, 0, 1024), idesc_s, kk ? 1u : 0u, leader);
                }
            };
            auto issue_pv = [&](int t, uint32_t v_addr, bool accumulate, int half) {
#pragma unroll
                for (int kk = half * (FA_BKV / 32); kk < (half + 1) * (FA_BKV / 32); ++kk) {
                    // A = P from tensor memory: 16 k-values of 16 bits = 8 columns per step;
                    // B = V, MN-major: 16 kv rows = 2 x 1024 B per step, 64-wide d atoms ATOM_BYTES apart
                    const uint32_t pcol = NH == 1 ? (uint32_t)(kk * 8) : (uint32_t)((kk >> 2) * 64 + (kk & 3) * 8);  // NH = 2: P halves at columns 0 / 64
                    umma_f16_ts_p(tmem_base + O_COL + (uint32_t)(t * D), tmem_base + (uint32_t)(t * 128) + pcol,
                                  make_sw128_desc(v_addr + kk * 2048, ATOM_BYTES, 1024), idesc_pv, (accumulate || kk) ? 1u : 0u, leader);
                }
            };
            int s = 0;
            uint32_t ph = 0;
            auto next_slot = [&]() {
                mbar_wait(&kv_full[s], ph);
                const int cur = s;
                if (++s == NS) {
                    s = 0;
                    ph ^= 1;
                }
                return cur;
            };
            mbar_wait(q_full, 0);
            {
                const int sk = next_slot();  // K_0
                tc_fence_after();
#pragma unroll
                for (int t = 0; t < 2; ++t)
                    if ((t ? nblk1 : nblk0) > 0) {
                        issue_s(t, kv_addr + sk * TILE_BYTES);
                        umma_commit_p(&s_full[t], leader);
                    }
                umma_commit_p(&kv_empty[sk], leader);
            }
            for (int j = 1; j <= nmax; ++j) {
                const int sv = next_slot();  // V_{j-1}
                const bool has_k = j < nmax;
                const int sk = has_k ? next_slot() : 0;  // K_j
#pragma unroll
                for (int t = 0; t < 2; ++t) {
                    const int nb_t = t ? nblk1 : nblk0;
                    if (j - 1 < nb_t) {
                        mbar_wait(&p_full[2 * t], (uint32_t)((j - 1) & 1));
                        tc_fence_after();
                        issue_pv(t, kv_addr + sv * TILE_BYTES, j > 1, 0);
                        if (p.stale) umma_commit_p(&pv0_done[t], leader);  // only the stale-reference softmax ever waits on it
                        mbar_wait(&p_full[2 * t + 1], (uint32_t)((j - 1) & 1));
                        tc_fence_after();
                        issue_pv(t, kv_addr + sv * TILE_BYTES, true, 1);
                        if (j < nb_t) issue_s(t, kv_addr + sk * TILE_BYTES);
                        umma_commit_p(&s_full[t], leader);
                    }
                }
                umma_commit_p(&kv_empty[sv], leader);
                if (has_k) umma_commit_p(&kv_empty[sk], leader);
            }
        }
        __syncwarp();
      }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(FaCfg<NH>::REG_SOFTMAX));
        // ===================================================== softmax + epilogue: thread = one query row (or half of it) of tile t
        const int t = warp / (4 * NH), h = (warp >> 2) % NH, q = warp & 3;
        const int n_t = t ? nblk1 : nblk0;
        if (n_t > 0) {
            constexpr int HC = 128 / NH;   // S columns per thread
            constexpr int HD = D / NH;     // O columns per thread
            const int r = q * 32 + lane;
            const int64_t q0t = (int64_t)q0 + t * FA_BQ;
            const int64_t m_row = q0t + r;
            const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
            const uint32_t s_addr = lane_addr + (uint32_t)(t * 128 + h * HC);
            const uint32_t p_addr = lane_addr + (uint32_t)(t * 128 + h * (HC / 2));
            const uint32_t o_addr = lane_addr + O_COL + (uint32_t)(t * D + h * HD);
            float *xm = xch + ((t * 2 + h) * 2) * 128 + r;        // [parity] stride 128
            const float *xp = xch + ((t * 2 + (h ^ 1)) * 2) * 128 + r;
            const int bar_id = 1 + t * 4 + q;                      // named barrier of this row quarter's two half-warps
            const float sc = p.scale_log2;
            float m_ref = -INFINITY, l_run = 0.f;
            for (int j = 0; j < n_t; ++j) {
                const int kv0 = j * FA_BKV;
                mbar_wait(&s_full[t], (uint32_t)(j & 1));
                tc_fence_after();
                const bool masked = (kv0 + FA_BKV - 1 > q0t) || (kv0 + FA_BKV > p.Skv);  // diagonal / ragged block (CTA-uniform per tile)
                const int64_t lim64 = min(m_row, p.Skv - 1) - kv0 - h * HC;             // columns i > lim (of this half) are masked
                const int lim = (int)max((int64_t)-1, min(lim64, (int64_t)127));
                uint64_t *p_bar = &p_full[2 * t + h];  // NH = 1: halves 0 and 1 in turn; NH = 2: this warp's half
                float *xm_j = xm + (j & 1) * 128;
                const float *xp_j = xp + (j & 1) * 128;
                if constexpr (NH == 2) {
                    if (p.is_bf16) {
                        if (masked) fwd_softmax_block_split<D, true, true, POLY>(s_addr, o_addr, sc, lim, j == 0, m_ref, l_run, xm_j, xp_j, bar_id, p_bar);
                        else fwd_softmax_block_split<D, true, false, POLY>(s_addr, o_addr, sc, lim, j == 0, m_ref, l_run, xm_j, xp_j, bar_id, p_bar);
                    } else {
                        if (masked) fwd_softmax_block_split<D, false, true, POLY>(s_addr, o_addr, sc, lim, j == 0, m_ref, l_run, xm_j, xp_j, bar_id, p_bar);
                        else fwd_softmax_block_split<D, false, false, POLY>(s_addr, o_addr, sc, lim, j == 0, m_ref, l_run, xm_j, xp_j, bar_id, p_bar);
                    }
                } else if (NH == 1 && p.stale && j > 0) {
                    const uint32_t pvp = (uint32_t)(j & 1);  // the (j + 1)-th completion of pv0_done = the first-half P V of THIS block
                    if (p.is_bf16) {
                        if (masked) fwd_softmax_block_stale<D, true, true, POLY>(s_addr, p_addr, o_addr, sc, lim, m_ref, l_run, p_bar, &pv0_done[t], pvp);
                        else fwd_softmax_block_stale<D, true, false, POLY>(s_addr, p_addr, o_addr, sc, lim, m_ref, l_run, p_bar, &pv0_done[t], pvp);
                    } else {
                        if (masked) fwd_softmax_block_stale<D, false, true, POLY>(s_addr, p_addr, o_addr, sc, lim, m_ref, l_run, p_bar, &pv0_done[t], pvp);
                        else fwd_softmax_block_stale<D, false, false, POLY>(s_addr, p_addr, o_addr, sc, lim, m_ref, l_run, p_bar, &pv0_done[t], pvp);
                    }
                } else if (p.is_bf16) {
                    if (masked) fwd_softmax_block<D, true, true, POLY, NH>(s_addr, p_addr, o_addr, sc, lim, j == 0, m_ref, l_run, xm_j, xp_j, bar_id, p_bar);
                    else fwd_softmax_block<D, true, false, POLY, NH>(s_addr, p_addr, o_addr, sc, lim, j == 0, m_ref, l_run, xm_j, xp_j, bar_id, p_bar);
                } else {
                    if (masked) fwd_softmax_block<D, false, true, POLY, NH>(s_addr, p_addr, o_addr, sc, lim, j == 0, m_ref, l_run, xm_j, xp_j, bar_id, p_bar);
                    else fwd_softmax_block<D, false, false, POLY, NH>(s_addr, p_addr, o_addr, sc, lim, j == 0, m_ref, l_run, xm_j, xp_j, bar_id, p_bar);
                }
            }
            // ---- epilogue: O / l -> 16-bit -> global, row LSE
            if (NH == 2) {  // total row sum = sum of the two halves (slot parity n_t: the last block used parity (n_t - 1) & 1)
                xm[(n_t & 1) * 128] = l_run;
                asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory");
                l_run += xp[(n_t & 1) * 128];
            }
            mbar_wait(&s_full[t], (uint32_t)(n_t & 1));
            tc_fence_after();
            const float inv_l = 1.f / l_run;
            const bool row_ok = m_row < p.Sq;
            if constexpr (NH == 1) {  // staged in the tile's own Q buffer (every S MMA of the tile has completed), stored by TMA
                fwd_store_tile<D, ATOMS, false>(o_addr, inv_l, p.is_bf16, sQ + t * TILE_BYTES, &tmap_o, (int)q0t, h_idx, b_idx, r, 1 + t, nullptr);
                if (row_ok && p.lse) p.lse[(int64_t)bh * p.Sq + m_row] = (m_ref + log2f(l_run)) * 0.6931471805599453f;
                if (r == 0) tma_store_wait_all<0>();  // the staging buffer must outlive the bulk copy's reads
            } else {
            uint16_t *orow = reinterpret_cast<uint16_t *>(p.out) + (int64_t)b_idx * p.lo.sb + (int64_t)h_idx * p.lo.sh + (row_ok ? m_row : 0) * p.lo.ss + h * HD;
#pragma unroll 1
            for (int c = 0; c < HD / 32; ++c) {
                uint32_t orr[32];
                tmem_ld32(o_addr + (uint32_t)(c * 32), orr);  // .sync.aligned: every lane takes part, only the stores are predicated
                tmem_ld_wait();
                if (row_ok) {
                    uint4 *dst = reinterpret_cast<uint4 *>(orow + c * 32);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        uint32_t w[4];
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            w[k] = pack16(__uint_as_float(orr[8 * i + 2 * k]) * inv_l, __uint_as_float(orr[8 * i + 2 * k + 1]) * inv_l, p.is_bf16);
                        dst[i] = make_uint4(w[0], w[1], w[2], w[3]);
                    }
                }
            }
            if (h == 0 && row_ok && p.lse) p.lse[(int64_t)bh * p.Sq + m_row] = (m_ref + log2f(l_run)) * 0.6931471805599453f;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == W_MMA) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

// =====================================================================================================================
// Forward with 64-KEY blocks and the next S issued BEFORE P V (KF_ATTN_FWD=k64).  In attn_fwd_tc_body P overwrites S in tensor
// memory, so per tile  S(j) -> softmax(j) -> [P V(j), S(j+1)] -> softmax(j+1)  is one serial chain (2900 clk per 128 keys for the two
// tiles against 2048 clk of tensor work).  With 64-key blocks tensor memory has room for S and P side by side —
//   S0 | S1 (64 columns each) | P0 x2 | P1 x2 (32 columns each, double-buffered) | O0 | O1 = 512 columns —
// so the issuer queues S_t(j+1) as soon as the softmax warps hold S_t(j) in registers (s_free) and P V_t(j) when P_t(j) is stored
// (p_full): the tensor pipe always has the other tile's and the next block's work queued and never waits for a softmax, and a
// softmax never waits for its own tile's P V.  Price: the score MMAs run at N = 64 (48 clk instead of 32 for half the columns: the
// Q operand is re-read from shared memory for half as much math), 1280 clk of tensor work per 64 keys for the two tiles
// (2560 per 128 keys).  P is double-buffered because P V_t(j - 1) may still be queued when softmax(j) stores; P V_t(j - 2) precedes
// S_t(j) on the in-order pipe, so the buffer of block j - 2 is free once S_t(j) has been seen.  A rescale of O (row maximum grown
// by more than 2^8) first waits for the tile's last P V (pv_done).
template <int D, bool BF16, bool MASKED, int POLY>
__device__ __forceinline__ void fwd_softmax_block64(const uint32_t s_addr, const uint32_t p_addr, const uint32_t o_addr, const float sc, const int lim,
                                                    const bool first, float &m_ref, float &l_run, uint64_t *s_free, uint64_t *p_full,
                                                    uint64_t *pv_done, const uint32_t pv_parity) {
    uint32_t s[2][32];
    tmem_ld32(s_addr, s[0]);
    tmem_ld32(s_addr + 32, s[1]);
    tmem_ld_wait();
    tc_fence_before();
    __syncwarp();
    if ((threadIdx.x & 31) == 0) mbar_arrive(s_free);  // S is in registers: the issuer may queue the tile's next S
    if (MASKED) {
#pragma unroll
        for (int c = 0; c < 2; ++c)
#pragma unroll
            for (int i = 0; i < 32; ++i)
                if (c * 32 + i > lim) s[c][i] = 0xff800000u;  // -inf
    }
    float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
    for (int c = 0; c < 2; ++c)
#pragma unroll
        for (int i = 0; i < 32; i += 8) {
            mx0 = fmaxf(mx0, fmaxf(__uint_as_float(s[c][i]), __uint_as_float(s[c][i + 1])));
            mx1 = fmaxf(mx1, fmaxf(__uint_as_float(s[c][i + 2]), __uint_as_float(s[c][i + 3])));
            mx2 = fmaxf(mx2, fmaxf(__uint_as_float(s[c][i + 4]), __uint_as_float(s[c][i + 5])));
            mx3 = fmaxf(mx3, fmaxf(__uint_as_float(s[c][i + 6]), __uint_as_float(s[c][i + 7])));
        }
    const float m_new = fmaxf(m_ref, fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)) * sc);
    const bool grow = first ? true : (m_new - m_ref > 8.f);  // lazy reference maximum, as in fwd_softmax_block
    if (!first && __any_sync(0xffffffffu, grow)) {
        mbar_wait(pv_done, pv_parity);  // the tile's previous P V may still be accumulating into O
        tc_fence_after();
        const float f = grow ? ((m_ref == -INFINITY) ? 0.f : ex2_approx(m_ref - m_new)) : 1.f;
        l_run *= f;
#pragma unroll 1
        for (int c = 0; c < D / 32; ++c) {
            uint32_t orr[32];
            tmem_ld32(o_addr + (uint32_t)(c * 32), orr);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) orr[i] = __float_as_uint(__uint_as_float(orr[i]) * f);
            tmem_st32(o_addr + (uint32_t)(c * 32), orr);
        }
    }
    if (grow) m_ref = m_new;
    const float m_use = (m_ref == -INFINITY) ? 0.f : m_ref;
    const float2 sc2 = make_float2(sc, sc), nm2 = make_float2(-m_use, -m_use);
    float2 rs2 = make_float2(0.f, 0.f), rs3 = make_float2(0.f, 0.f);
    uint32_t pk[32];
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
            float2 x = __ffma2_rn(make_float2(__uint_as_float(s[h][i]), __uint_as_float(s[h][i + 1])), sc2, nm2);
            if (((i >> 1) & 7) < POLY) {
                x = ex2_poly2(x);
            } else {
                x.x = ex2_approx(x.x);
                x.y = ex2_approx(x.y);
            }
            if (i & 2) rs3 = __fadd2_rn(rs3, x);
            else rs2 = __fadd2_rn(rs2, x);
            pk[h * 16 + (i >> 1)] = pack16t<BF16>(x);
        }
    tmem_st32(p_addr, pk);
    tmem_st_wait();
    tc_fence_before();
    __syncwarp();
    if ((threadIdx.x & 31) == 0) mbar_arrive(p_full);
    l_run += (rs2.x + rs2.y) + (rs3.x + rs3.y);
}

template <int D, int POLY>
__global__ void __launch_bounds__(FaCfg<1>::THREADS, 1)
attn_fwd_k64_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                    const __grid_constant__ CUtensorMap tmap_v, const __grid_constant__ CUtensorMap tmap_o, const AttnTcParams p) {
    constexpr int ATOMS = D / 64;
    constexpr int TILE_Q = 128 * D * 2, ATOM_Q = 128 * 128;  // a 128-row query tile / one of its 64-column atoms
    constexpr int TILE_KV = 64 * D * 2, ATOM_KV = 64 * 128;  // a 64-row K or V tile
    constexpr int BKV = 64, NS = 8;                           // ring: K_0, K_1, V_0, K_2, V_1, ... in consumption order
    constexpr uint32_t TMEM_COLS = 512, P_COL = 128, O_COL = 256;
    constexpr int W_MMA = 8, W_TMA = 9;
    extern __shared__ unsigned char smem_raw[];
    unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    unsigned char *sQ = smem;
    unsigned char *sKV = smem + 2 * TILE_Q;
    uint64_t *bars = reinterpret_cast<uint64_t *>(sKV + NS * TILE_KV);
    uint64_t *q_full = bars + 0;
    uint64_t *kv_full = bars + 1, *kv_empty = bars + 1 + NS;
    uint64_t *s_full = bars + 1 + 2 * NS;  // [2]
    uint64_t *s_free = s_full + 2;         // [2]
    uint64_t *p_full = s_free + 2;         // [2 tiles][2 P buffers]: a tile's softmax may store P(j + 1) before the issuer has looked at
                                           // P(j) (it only needs S(j + 1), queued on s_free(j)); one barrier per buffer keeps a waiter
                                           // at most one phase behind
    uint64_t *pv_done = p_full + 4;        // [2]
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(pv_done + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int bh, pr;
    fwd_decode_item(p, (int)blockIdx.x, bh, pr);
    const int b_idx = bh / p.H, h_idx = bh % p.H;
    const int q0 = pr * 2 * FA_BQ;
    auto blocks_of = [&](int t) {
        const int64_t q0t = (int64_t)q0 + t * FA_BQ;
        const int64_t kv_end = min((int64_t)p.Skv, q0t + FA_BQ);
        return q0t < p.Sq ? (int)((kv_end + BKV - 1) / BKV) : 0;
    };
    const int nblk0 = blocks_of(0), nblk1 = blocks_of(1);
    const int nmax = max(nblk0, nblk1);

    if (warp == W_TMA && lane == 0) {
        prefetch_tmap(&tmap_q);
        prefetch_tmap(&tmap_k);
        prefetch_tmap(&tmap_v);
        prefetch_tmap(&tmap_o);
        mbar_init(q_full, 1);
        for (int s = 0; s < NS; ++s) {
            mbar_init(&kv_full[s], 1);
            mbar_init(&kv_empty[s], 1);
        }
        for (int t = 0; t < 2; ++t) {
            mbar_init(&s_full[t], 1);
            mbar_init(&s_free[t], 4);
            mbar_init(&p_full[2 * t], 4);
            mbar_init(&p_full[2 * t + 1], 4);
            mbar_init(&pv_done[t], 1);
        }
        fence_barrier_init();
    }
    if (warp == W_MMA) tmem_alloc(tmem_slot, TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp >= 8) {
      asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(FaCfg<1>::REG_CTRL));
      if (warp == W_TMA) {
        // ===================================================== TMA producer
        if (lane == 0) {
            const int ntile_q = nblk1 > 0 ? 2 : 1;
            mbar_arrive_expect_tx(q_full, ntile_q * TILE_Q);
            for (int t = 0; t < ntile_q; ++t)
#pragma unroll
                for (int a = 0; a < ATOMS; ++a) tma_load_4d(sQ + t * TILE_Q + a * ATOM_Q, &tmap_q, q_full, a * 64, q0 + t * FA_BQ, h_idx, b_idx);
            int s = 0;
            uint32_t ph = 0;
            const int nloads = 2 * nmax;  // K_0, K_1, V_0, K_2, V_1, ..., K_{n-1}, V_{n-2}, V_{n-1}
            for (int i = 0; i < nloads; ++i) {
                const bool is_k = i == 0 || (i < nloads - 1 && (i & 1));
                const int blk = i == 0 ? 0 : (is_k ? (i + 1) / 2 : (i == nloads - 1 ? nmax - 1 : i / 2 - 1));
                const CUtensorMap *tm = is_k ? &tmap_k : &tmap_v;
                mbar_wait(&kv_empty[s], ph ^ 1);
                mbar_arrive_expect_tx(&kv_full[s], TILE_KV);
#pragma unroll
                for (int a = 0; a < ATOMS; ++a) tma_load_4d(sKV + s * TILE_KV + a * ATOM_KV, tm, &kv_full[s], a * 64, blk * BKV, h_idx, b_idx);
                if (++s == NS) {
                    s = 0;
                    ph ^= 1;
                }
            }
        }
        __syncwarp();
      } else if (warp == W_MMA) {
        // ===================================================== MMA issuer (converged warp, elected lane issues)
        const bool leader = elect_one();
        const int fmt = p.is_bf16 ? 1 : 0;
        const uint32_t idesc_s = make_idesc_f16(fmt, 0, 0, FA_BQ, BKV);
        const uint32_t idesc_pv = make_idesc_f16(fmt, 0, 1, FA_BQ, D);
        const uint32_t q_addr = smem_u32(sQ), kv_addr = smem_u32(sKV);
        auto issue_s = [&](int t, uint32_t k_addr) {
#pragma unroll
            for (int kk = 0; kk < D / 16; ++kk)
                umma_f16_p(tmem_base + (uint32_t)(t * 64), make_sw128_desc(q_addr + t * TILE_Q + (uint32_t)((kk >> 2) * ATOM_Q + (kk & 3) * 32), 0, 1024),
                           make_sw128_desc(k_addr + (uint32_t)((kk >> 2) * ATOM_KV + (kk & 3) * 32), 0, 1024), idesc_s, kk ? 1u : 0u, leader);
        };
        auto issue_pv = [&](int t, uint32_t v_addr, int j) {
#pragma unroll
            for (int kk = 0; kk < BKV / 16; ++kk)
                umma_f16_ts_p(tmem_base + O_COL + (uint32_t)(t * D), tmem_base + P_COL + (uint32_t)(t * 64 + (j & 1) * 32 + kk * 8),
                              make_sw128_desc(v_addr + kk * 2048, ATOM_KV, 1024), idesc_pv, (j > 0 || kk) ? 1u : 0u, leader);
        };
        int s = 0;
        uint32_t ph = 0;
        auto next_slot = [&]() {
            mbar_wait(&kv_full[s], ph);
            const int cur = s;
            if (++s == NS) {
                s = 0;
                ph ^= 1;
            }
            return cur;
        };
        mbar_wait(q_full, 0);
        {
            const int sk = next_slot();  // K_0
            tc_fence_after();
#pragma unroll
            for (int t = 0; t < 2; ++t)
                if ((t ? nblk1 : nblk0) > 0) {
                    issue_s(t, kv_addr + sk * TILE_KV);
                    umma_commit_p(&s_full[t], leader);
                }
            umma_commit_p(&kv_empty[sk], leader);
        }
        for (int j = 0; j < nmax; ++j) {
            if (j + 1 < nmax) {
                const int sk = next_slot();  // K_{j+1}
#pragma unroll
                for (int t = 0; t < 2; ++t)
                    if (j + 1 < (t ? nblk1 : nblk0)) {
                        mbar_wait(&s_free[t], (uint32_t)(j & 1));  // S_t(j) is in the softmax warps' registers
                        tc_fence_after();
                        issue_s(t, kv_addr + sk * TILE_KV);
                        umma_commit_p(&s_full[t], leader);
                    }
                umma_commit_p(&kv_empty[sk], leader);
            }
            const int sv = next_slot();  // V_j
#pragma unroll
            for (int t = 0; t < 2; ++t)
                if (j < (t ? nblk1 : nblk0)) {
                    mbar_wait(&p_full[2 * t + (j & 1)], (uint32_t)((j >> 1) & 1));
                    tc_fence_after();
                    issue_pv(t, kv_addr + sv * TILE_KV, j);
                    umma_commit_p(&pv_done[t], leader);
                }
            umma_commit_p(&kv_empty[sv], leader);
        }
        __syncwarp();
      }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(FaCfg<1>::REG_SOFTMAX));
        // ===================================================== softmax + epilogue: thread = one query row of tile t
        const int t = warp >> 2, q = warp & 3;
        const int n_t = t ? nblk1 : nblk0;
        if (n_t > 0) {
            const int r = q * 32 + lane;
            const int64_t q0t = (int64_t)q0 + t * FA_BQ;
            const int64_t m_row = q0t + r;
            const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
            const uint32_t s_addr = lane_addr + (uint32_t)(t * 64);
            const uint32_t o_addr = lane_addr + O_COL + (uint32_t)(t * D);
            const float sc = p.scale_log2;
            float m_ref = -INFINITY, l_run = 0.f;
            for (int j = 0; j < n_t; ++j) {
                const int kv0 = j * BKV;
                mbar_wait(&s_full[t], (uint32_t)(j & 1));
                tc_fence_after();
                const bool masked = (kv0 + BKV - 1 > q0t) || (kv0 + BKV > p.Skv);  // diagonal / ragged block (CTA-uniform per tile)
                const int64_t lim64 = min(m_row, p.Skv - 1) - kv0;
                const int lim = (int)max((int64_t)-1, min(lim64, (int64_t)63));
                const uint32_t p_addr = lane_addr + P_COL + (uint32_t)(t * 64 + (j & 1) * 32);
                const uint32_t pvp = (uint32_t)((j - 1) & 1);  // completion of the tile's previous P V (only looked at when j > 0)
                if (p.is_bf16) {
                    if (masked) fwd_softmax_block64<D, true, true, POLY>(s_addr, p_addr, o_addr, sc, lim, j == 0, m_ref, l_run, &s_free[t], &p_full[2 * t + (j & 1)], &pv_done[t], pvp);
                    else fwd_softmax_block64<D, true, false, POLY>(s_addr, p_addr, o_addr, sc, lim, j == 0, m_ref, l_run, &s_free[t], &p_full[2 * t + (j & 1)], &pv_done[t], pvp);
                } else {
                    if (masked) fwd_softmax_block64<D, false, true, POLY>(s_addr, p_addr, o_addr, sc, lim, j == 0, m_ref, l_run, &s_free[t], &p_full[2 * t + (j & 1)], &pv_done[t], pvp);
                    else fwd_softmax_block64<D, false, false, POLY>(s_addr, p_addr, o_addr, sc, lim, j == 0, m_ref, l_run, &s_free[t], &p_full[2 * t + (j & 1)], &pv_done[t], pvp);
                }
            }
            // ---- epilogue: O / l -> 16-bit -> the tile's own (dead) Q buffer -> TMA store; row LSE
            mbar_wait(&pv_done[t], (uint32_t)((n_t - 1) & 1));
            tc_fence_after();
            const float inv_l = 1.f / l_run;
            const bool row_ok = m_row < p.Sq;
            fwd_store_tile<D, ATOMS, false>(o_addr, inv_l, p.is_bf16, sQ + t * TILE_Q, &tmap_o, (int)q0t, h_idx, b_idx, r, 1 + t, nullptr);
            if (row_ok && p.lse) p.lse[(int64_t)bh * p.Sq + m_row] = (m_ref + log2f(l_run)) * 0.6931471805599453f;
            if (r == 0) tma_store_wait_all<0>();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == W_MMA) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

// =====================================================================================================================
// Forward, variant "P through shared memory" (KF_ATTN_FWD=ps): the next S of a tile no longer waits for the tile's P V.
// In the kernel above P overwrites S in tensor memory, so the chain  S(j) -> softmax(j) -> [P V(j), S(j+1)] -> softmax(j+1)
// is serial per tile: with two tiles ping-ponging, the tensor pipe has 2048 clk of work per KV block against a per-tile cycle of
// softmax + 1024 + two ~250 clk hand-overs (measured ~3300 clk per block).  Here the softmax warps signal "S is in registers"
// right after their tcgen05.ld, the MMA warp issues S(j+1) at once into the same columns, and P travels through a 16 KB
// shared-memory buffer per tile in two 64-key halves (K-major, 128B swizzle, written with st.shared.v4 + fence.proxy.async,
// read by the P V MMAs as their A operand).  The softmax warps of a tile then run block after block without waiting for the
// tensor pipe; the pipe's work (S of the next block, the four P V halves of this one) fills in underneath.
// Shared memory: Q 2 x 32 KB, K/V ring 4 x 32 KB, P 2 x 16 KB = 224 KB at D = 128.  TMEM: S0 | S1 | O0 | O1 as before.
template <int D, bool BF16, bool MASKED, int POLY>
__device__ __forceinline__ void fwd_softmax_block_ps(const uint32_t s_addr, const uint32_t o_addr, const float sc, const int lim, const bool first,
                                                     float &m_ref, float &l_run, uint64_t *s_free, uint64_t *p_full, uint64_t *p_free, const int j,
                                                     const uint32_t p_row, const uint32_t rsw) {
    uint32_t s[4][32];
#pragma unroll
    for (int c = 0; c < 4; ++c) tmem_ld32(s_addr + (uint32_t)(c * 32), s[c]);
    tmem_ld_wait();
    tc_fence_before();
    __syncwarp();
    if ((threadIdx.x & 31) == 0) mbar_arrive(s_free);  // S(j) is in registers: the MMA warp may overwrite it with S(j+1)
    if (MASKED) {
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
            for (int i = 0; i < 32; ++i)
                if (c * 32 + i > lim) s[c][i] = 0xff800000u;  // -inf
    }
    float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int i = 0; i < 32; i += 8) {
            mx0 = fmaxf(mx0, fmaxf(__uint_as_float(s[c][i]), __uint_as_float(s[c][i + 1])));
            mx1 = fmaxf(mx1, fmaxf(__uint_as_float(s[c][i + 2]), __uint_as_float(s[c][i + 3])));
            mx2 = fmaxf(mx2, fmaxf(__uint_as_float(s[c][i + 4]), __uint_as_float(s[c][i + 5])));
            mx3 = fmaxf(mx3, fmaxf(__uint_as_float(s[c][i + 6]), __uint_as_float(s[c][i + 7])));
        }
    const float mx = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
    const float m_new = fmaxf(m_ref, mx * sc);
    const bool grow = first ? true : (m_new - m_ref > 8.f);  // lazy reference: only moves when the row max grew by more than 2^8
    if (!first && __any_sync(0xffffffffu, grow)) {
        // O may only be touched once every P V of the previous blocks has completed: 2 j completions of p_free
        mbar_wait(p_free, (uint32_t)((2 * j - 1) & 1));
        tc_fence_after();
        const float f = grow ? ((m_ref == -INFINITY) ? 0.f : ex2_approx(m_ref - m_new)) : 1.f;
        l_run *= f;
#pragma unroll 1
        for (int c = 0; c < D / 32; ++c) {
            uint32_t orr[32];
            tmem_ld32(o_addr + (uint32_t)(c * 32), orr);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) orr[i] = __float_as_uint(__uint_as_float(orr[i]) * f);
            tmem_st32(o_addr + (uint32_t)(c * 32), orr);
        }
        tmem_st_wait();
        tc_fence_before();
    }
    if (grow) m_ref = m_new;
    const float m_use = (m_ref == -INFINITY) ? 0.f : m_ref;
    const float2 sc2 = make_float2(sc, sc), nm2 = make_float2(-m_use, -m_use);
    float2 rs2 = make_float2(0.f, 0.f), rs3 = make_float2(0.f, 0.f);
#pragma unroll
    for (int h = 0; h < 2; ++h) {  // 64-key halves
        uint32_t pk[32];
#pragma unroll
        for (int cc = 0; cc < 2; ++cc)
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
                float2 x = __ffma2_rn(make_float2(__uint_as_float(s[2 * h + cc][i]), __uint_as_float(s[2 * h + cc][i + 1])), sc2, nm2);
                if (((i >> 1) & 7) < POLY) {
                    x = ex2_poly2(x);
                } else {
                    x.x = ex2_approx(x.x);
                    x.y = ex2_approx(x.y);
                }
                if (i & 2) rs3 = __fadd2_rn(rs3, x);
                else rs2 = __fadd2_rn(rs2, x);
                pk[cc * 16 + (i >> 1)] = pack16t<BF16>(x);
            }
        // the 16 KB buffer is free once the previous half's P V MMAs have read it: 2 j + h completions of p_free
        const int need = 2 * j + h;
        if (need > 0) mbar_wait(p_free, (uint32_t)((need - 1) & 1));
#pragma unroll
        for (int c = 0; c < 8; ++c)  // this row's 128 bytes: eight 16-byte chunks, XOR-swizzled by the row index (128B swizzle atom)
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(p_row + ((((uint32_t)c) ^ rsw) << 4)), "r"(pk[4 * c]), "r"(pk[4 * c + 1]),
                         "r"(pk[4 * c + 2]), "r"(pk[4 * c + 3])
                         : "memory");
        fence_proxy_async();  // generic-proxy stores -> visible to the MMA's async-proxy reads
        __syncwarp();
        if ((threadIdx.x & 31) == 0) mbar_arrive(p_full + h);
    }
    l_run += (rs2.x + rs2.y) + (rs3.x + rs3.y);
}

template <int D, int POLY>
__global__ void __launch_bounds__(FaCfg<1>::THREADS, 1)
attn_fwd_ps_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                   const __grid_constant__ CUtensorMap tmap_v, const __grid_constant__ CUtensorMap /*tmap_o: direct stores here*/,
                   const AttnTcParams p) {
    constexpr int ATOMS = D / 64;
    constexpr int TILE_BYTES = 128 * D * 2;
    constexpr int ATOM_BYTES = 128 * 128;
    constexpr int P_BYTES = 128 * 128;  // one 64-key half: 128 rows x 128 B
    constexpr uint32_t TMEM_COLS = 512, O_COL = 256;
    constexpr int NS = FA_NSTAGE;
    constexpr int W_MMA = 8, W_TMA = 9;
    extern __shared__ unsigned char smem_raw[];
    unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    unsigned char *sQ = smem;
    unsigned char *sKV = smem + 2 * TILE_BYTES;
    unsigned char *sP = sKV + NS * TILE_BYTES;  // [2 tiles] 16 KB
    uint64_t *bars = reinterpret_cast<uint64_t *>(sP + 2 * P_BYTES);
    uint64_t *q_full = bars + 0;
    uint64_t *kv_full = bars + 1, *kv_empty = bars + 1 + NS;
    uint64_t *s_full = bars + 1 + 2 * NS;  // [2]
    uint64_t *s_free = s_full + 2;         // [2]
    uint64_t *p_full = s_free + 2;         // [2 tiles][2 halves]
    uint64_t *p_free = p_full + 4;         // [2]: one completion per P V half
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(p_free + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int bh = blockIdx.x / p.npairs;
    const int b_idx = bh / p.H, h_idx = bh % p.H;
    const int pr = p.npairs - 1 - (blockIdx.x % p.npairs);  // heaviest pairs first
    const int q0 = pr * 2 * FA_BQ;
    auto blocks_of = [&](int t) {
        const int64_t q0t = (int64_t)q0 + t * FA_BQ;
        const int64_t kv_end = min((int64_t)p.Skv, q0t + FA_BQ);
        return q0t < p.Sq ? (int)((kv_end + FA_BKV - 1) / FA_BKV) : 0;
    };
    const int nblk0 = blocks_of(0), nblk1 = blocks_of(1);
    const int nmax = max(nblk0, nblk1);

    if (warp == W_TMA && lane == 0) {
        prefetch_tmap(&tmap_q);
        prefetch_tmap(&tmap_k);
        prefetch_tmap(&tmap_v);
        mbar_init(q_full, 1);
        for (int s = 0; s < NS; ++s) {
            mbar_init(&kv_full[s], 1);
            mbar_init(&kv_empty[s], 1);
        }
        for (int t = 0; t < 2; ++t) {
            mbar_init(&s_full[t], 1);
            mbar_init(&s_free[t], 4);
            mbar_init(&p_full[2 * t], 4);
            mbar_init(&p_full[2 * t + 1], 4);
            mbar_init(&p_free[t], 1);
        }
        fence_barrier_init();
    }
    if (warp == W_MMA) tmem_alloc(tmem_slot, TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp >= 8) {
      asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(FaCfg<1>::REG_CTRL));
      if (warp == W_TMA) {
        // ===================================================== TMA producer: Q tiles once; K_0, V_0, K_1, V_1, ... through the ring
        if (lane == 0) {
            const int ntile_q = nblk1 > 0 ? 2 : 1;
            mbar_arrive_expect_tx(q_full, ntile_q * TILE_BYTES);
            for (int t = 0; t < ntile_q; ++t)
#pragma unroll
                for (int a = 0; a < ATOMS; ++a) tma_load_4d(sQ + t * TILE_BYTES + a * ATOM_BYTES, &tmap_q, q_full, a * 64, q0 + t * FA_BQ, h_idx, b_idx);
            for (int i = 0; i < 2 * nmax; ++i) {
                const int s = i % NS;
                const int kv0 = (i >> 1) * FA_BKV;
                const CUtensorMap *tm = (i & 1) ? &tmap_v : &tmap_k;
                mbar_wait(&kv_empty[s], (uint32_t)(((i / NS) & 1) ^ 1));
                mbar_arrive_expect_tx(&kv_full[s], TILE_BYTES);
#pragma unroll
                for (int a = 0; a < ATOMS; ++a) tma_load_4d(sKV + s * TILE_BYTES + a * ATOM_BYTES, tm, &kv_full[s], a * 64, kv0, h_idx, b_idx);
            }
        }
        __syncwarp();
      } else if (warp == W_MMA) {
        // ===================================================== MMA issuer
        const bool leader = elect_one();
        {
            const int fmt = p.is_bf16 ? 1 : 0;
            const uint32_t idesc_s = make_idesc_f16(fmt, 0, 0, FA_BQ, FA_BKV);
            const uint32_t idesc_pv = make_idesc_f16(fmt, 0, 1, FA_BQ, D);  // A = P K-major (smem), B = V MN-major
            const uint32_t q_addr = smem_u32(sQ), kv_addr = smem_u32(sKV), p_addr = smem_u32(sP);
            auto slot_of = [&](int i) {  // ring load i (K_j = 2 j, V_j = 2 j + 1): wait until it has landed, return its address
                mbar_wait(&kv_full[i % NS], (uint32_t)((i / NS) & 1));
                return kv_addr + (uint32_t)((i % NS) * TILE_BYTES);
            };
            auto issue_s = [&](int t, uint32_t k_addr) {
#pragma unroll
                for (int kk = 0; kk < D / 16; ++kk) {
                    const uint32_t off = (uint32_t)((kk >> 2) * ATOM_BYTES + (kk & 3) * 32);
                    umma_f16_p(tmem_base + (uint32_t)(t * 128), make_sw128_desc(q_addr + t * TILE_BYTES + off, 0, 1024),
                               make_sw128_desc(k_addr + off, 0, 1024), idesc_s, kk ? 1u : 0u, leader);
                }
            };
            auto issue_pv_half = [&](int t, uint32_t v_addr, bool accumulate, int half) {
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {  // 64 keys = 4 k-steps: A = 32 B further along the P row, B = 16 more rows of V
                    umma_f16_p(tmem_base + O_COL + (uint32_t)(t * D), make_sw128_desc(p_addr + t * P_BYTES + kk * 32, 0, 1024),
                               make_sw128_desc(v_addr + (half * 4 + kk) * 2048, ATOM_BYTES, 1024), idesc_pv, (accumulate || kk) ? 1u : 0u, leader);
                }
            };
            mbar_wait(q_full, 0);
            {
                const uint32_t k0 = slot_of(0);
                tc_fence_after();
                if (nblk0 > 0) {
                    issue_s(0, k0);
                    umma_commit_p(&s_full[0], leader);
                }
                if (nblk1 > 0) {
                    issue_s(1, k0);
                    umma_commit_p(&s_full[1], leader);
                }
                umma_commit_p(&kv_empty[0], leader);
            }
            for (int j = 0; j < nmax; ++j) {
                const uint32_t par = (uint32_t)(j & 1);
                // ---- S of the next block for both tiles, as soon as their softmax warps hold S(j) in registers
                if (j + 1 < nmax) {
                    const uint32_t kn = slot_of(2 * j + 2);
#pragma unroll
                    for (int t = 0; t < 2; ++t)
                        if (j + 1 < (t ? nblk1 : nblk0)) {
                            mbar_wait(&s_free[t], par);
                            tc_fence_after();
                            issue_s(t, kn);
                            umma_commit_p(&s_full[t], leader);
                        }
                    umma_commit_p(&kv_empty[(2 * j + 2) % NS], leader);
                }
                // ---- P V of this block: half 0 of both tiles, then half 1
                const uint32_t vj = slot_of(2 * j + 1);
#pragma unroll
                for (int half = 0; half < 2; ++half)
#pragma unroll
                    for (int t = 0; t < 2; ++t)
                        if (j < (t ? nblk1 : nblk0)) {
                            mbar_wait(&p_full[2 * t + half], par);
                            tc_fence_after();
                            issue_pv_half(t, vj, j > 0 || half > 0, half);
                            umma_commit_p(&p_free[t], leader);
                        }
                umma_commit_p(&kv_empty[(2 * j + 1) % NS], leader);
            }
        }
        __syncwarp();
      }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(FaCfg<1>::REG_SOFTMAX));
        // ===================================================== softmax + epilogue: thread = one query row of tile t
        const int t = warp >> 2, q = warp & 3;
        const int n_t = t ? nblk1 : nblk0;
        if (n_t > 0) {
            const int r = q * 32 + lane;
            const int64_t q0t = (int64_t)q0 + t * FA_BQ;
            const int64_t m_row = q0t + r;
            const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
            const uint32_t s_addr = lane_addr + (uint32_t)(t * 128);
            const uint32_t o_addr = lane_addr + O_COL + (uint32_t)(t * D);
            const uint32_t p_row = smem_u32(sP) + (uint32_t)(t * P_BYTES + r * 128);
            const uint32_t rsw = (uint32_t)(r & 7);
            const float sc = p.scale_log2;
            float m_ref = -INFINITY, l_run = 0.f;
            for (int j = 0; j < n_t; ++j) {
                const int kv0 = j * FA_BKV;
                mbar_wait(&s_full[t], (uint32_t)(j & 1));
                tc_fence_after();
                const bool masked = (kv0 + FA_BKV - 1 > q0t) || (kv0 + FA_BKV > p.Skv);
                const int64_t lim64 = min(m_row, p.Skv - 1) - kv0;
                const int lim = (int)max((int64_t)-1, min(lim64, (int64_t)127));
                if (p.is_bf16) {
                    if (masked) fwd_softmax_block_ps<D, true, true, POLY>(s_addr, o_addr, sc, lim, j == 0, m_ref, l_run, &s_free[t], &p_full[2 * t], &p_free[t], j, p_row, rsw);
                    else fwd_softmax_block_ps<D, true, false, POLY>(s_addr, o_addr, sc, lim, j == 0, m_ref, l_run, &s_free[t], &p_full[2 * t], &p_free[t], j, p_row, rsw);
                } else {
                    if (masked) fwd_softmax_block_ps<D, false, true, POLY>(s_addr, o_addr, sc, lim, j == 0, m_ref, l_run, &s_free[t], &p_full[2 * t], &p_free[t], j, p_row, rsw);
                    else fwd_softmax_block_ps<D, false, false, POLY>(s_addr, o_addr, sc, lim, j == 0, m_ref, l_run, &s_free[t], &p_full[2 * t], &p_free[t], j, p_row, rsw);
                }
            }
            // ---- epilogue: every P V half has completed (2 n_t completions of p_free) -> O / l -> 16-bit -> global, row LSE
            mbar_wait(&p_free[t], (uint32_t)((2 * n_t - 1) & 1));
            tc_fence_after();
            const float inv_l = 1.f / l_run;
            const bool row_ok = m_row < p.Sq;
            uint16_t *orow = reinterpret_cast<uint16_t *>(p.out) + (int64_t)b_idx * p.lo.sb + (int64_t)h_idx * p.lo.sh + (row_ok ? m_row : 0) * p.lo.ss;
#pragma unroll 1
            for (int c = 0; c < D / 32; ++c) {
                uint32_t orr[32];
                tmem_ld32(o_addr + (uint32_t)(c * 32), orr);
                tmem_ld_wait();
                if (row_ok) {
                    uint4 *dst = reinterpret_cast<uint4 *>(orow + c * 32);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        uint32_t w[4];
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            w[k] = pack16(__uint_as_float(orr[8 * i + 2 * k]) * inv_l, __uint_as_float(orr[8 * i + 2 * k + 1]) * inv_l, p.is_bf16);
                        dst[i] = make_uint4(w[0], w[1], w[2], w[3]);
                    }
                }
            }
            if (row_ok && p.lse) p.lse[(int64_t)bh * p.Sq + m_row] = (m_ref + log2f(l_run)) * 0.6931471805599453f;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == W_MMA) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

// =====================================================================================================================
// Forward, PERSISTENT form (KF_ATTN_FWD=pers; the default for KV lengths <= 2048): one CTA per SM takes (batch-head, query pair)
// work items from a dynamic queue (see get_item below) in the order of fwd_decode_item: head groups that fit the L2, longest pairs
// first inside a group.  The roles are those of attn_fwd_tc_body (NH = 1), but the K/V ring, the TMEM allocation, the tensor maps
// and the barrier set-up outlive a work item:
//   * the TMA warp loads the next item's Q tiles as soon as the last S MMA of the current item has completed (q_empty) and keeps
//     the K/V ring running across the item boundary, so a new item never waits for cold loads;
//   * the MMA warp issues S_0, S_1 of the next item straight after the last P V of the current one: they run under the epilogue;
//   * the softmax warps write O out, release the tile's O columns (o_free) and start on the next item's S, which is already there.
// Barrier phases are running counters per role (a tile can be idle in an item when Sq is ragged).  "P V of the last block done"
// has its own barrier (o_full) so that s_full never completes two phases between two looks of a waiter.
template <int D, int POLY, bool TRACE>
__global__ void __launch_bounds__(FaCfg<1>::THREADS, 1)
attn_fwd_pers_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                     const __grid_constant__ CUtensorMap tmap_v, const __grid_constant__ CUtensorMap tmap_o, const AttnTcParams p) {
    long long *const trace = TRACE ? p.trace : nullptr;  // the stamps exist only in the TRACE instance (registers of the control warps)
    constexpr int ATOMS = D / 64;
    constexpr int TILE_BYTES = 128 * D * 2;
    constexpr int ATOM_BYTES = 128 * 128;
    constexpr uint32_t TMEM_COLS = 512, O_COL = 256;  // S0 | S1 | O0 | O1 (P_t aliases the first 64 columns of S_t)
    constexpr int NS = FA_NSTAGE;
    constexpr int W_MMA = 8, W_TMA = 9;
    extern __shared__ unsigned char smem_raw[];
    unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    unsigned char *sQ = smem;
    unsigned char *sKV = smem + 2 * TILE_BYTES;
    unsigned char *sO = smem + (2 + NS) * TILE_BYTES;  // epilogue staging: one 64-column atom (128 rows x 128 B) per tile
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + (2 + NS) * TILE_BYTES + 2 * ATOM_BYTES);
    uint64_t *q_full = bars + 0, *q_empty = bars + 1;
    uint64_t *kv_full = bars + 2, *kv_empty = bars + 2 + NS;
    uint64_t *s_full = bars + 2 + 2 * NS;  // [2]
    uint64_t *p_full = s_full + 2;         // [2 tiles][2 key halves]
    uint64_t *o_full = p_full + 4;         // [2]: every P V of the item has completed
    uint64_t *o_free = o_full + 2;         // [2]: the epilogue has read O out of tensor memory
    uint64_t *work_full = o_free + 2;      // [4]: the producer has published the work item of this queue slot
    uint64_t *work_empty = work_full + 4;  // [4]: the MMA warp and the eight softmax warps have read it
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(work_empty + 4);
    volatile int *work = reinterpret_cast<volatile int *>(tmem_slot + 1);  // [4] item index, -1 = no more work

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nwork = (int)p.BH * p.npairs;
    // Work items are handed out DYNAMICALLY in head-major order (item w = pair npairs - 1 - w % npairs of batch-head w / npairs: a
    // head's pairs longest first, then the next head), the order the hardware scheduler gives the per-CTA kernel.  CTA c starts with
    // item c; its producer thread claims every further item from a global counter (one atomicAdd per item, issued before the item's
    // Q loads so the round trip is hidden) and publishes it through a four-slot queue in shared memory.  A static length-major list
    // (every head's longest pair first) balances as well but has all 148 CTAs streaming the K / V of 148 DIFFERENT heads at once:
    // 296 MB of live K / V at S = 4096 against 126 MB of L2, i.e. 64 KB per SM and block from HBM — the kernel ran at the HBM limit
    // (measured 1.12 ms at C3 against 1.03 ms for the per-CTA kernel, which keeps ~5 heads live).
    auto get_item = [&](int it) {  // consumers (whole warp): the it-th item of this CTA
        const int sl = it & 3;
        mbar_wait(&work_full[sl], (uint32_t)((it >> 2) & 1));
        const int w = work[sl];
        __syncwarp();
        if (lane == 0) mbar_arrive(&work_empty[sl]);
        return w;
    };
    auto blocks_of = [&](int q0, int t) {
        const int64_t q0t = (int64_t)q0 + t * FA_BQ;
        const int64_t kv_end = min((int64_t)p.Skv, q0t + FA_BQ);
        return q0t < p.Sq ? (int)((kv_end + FA_BKV - 1) / FA_BKV) : 0;
    };

    if (warp == W_TMA && lane == 0) {
        prefetch_tmap(&tmap_q);
        prefetch_tmap(&tmap_k);
        prefetch_tmap(&tmap_v);
        prefetch_tmap(&tmap_o);
        mbar_init(q_full, 1);
        mbar_init(q_empty, 1);
        for (int sl = 0; sl < 4; ++sl) {
            mbar_init(&work_full[sl], 1);
            mbar_init(&work_empty[sl], 9);
        }
        for (int s = 0; s < NS; ++s) {
            mbar_init(&kv_full[s], 1);
            mbar_init(&kv_empty[s], 1);
        }
        for (int t = 0; t < 2; ++t) {
            mbar_init(&s_full[t], 1);
            mbar_init(&p_full[2 * t], 4);
            mbar_init(&p_full[2 * t + 1], 4);
            mbar_init(&o_full[t], 1);
            mbar_init(&o_free[t], 4);
        }
        fence_barrier_init();
    }
    if (warp == W_MMA) tmem_alloc(tmem_slot, TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp >= 8) {
      asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(104)  /* the MMA warp keeps two items' loop state: 88 spills it; 128 x 104 + 256 x 200 = 384 x 168 */);
      if (warp == W_TMA) {
        // ===================================================== TMA producer
        if (lane == 0) {
            int s = 0, it = 0;
            uint32_t ph = 0;
            auto claim = [&]() {
                const int c = (int)atomicAdd(p.work_counter, 1u) + (int)gridDim.x;
                return c < nwork ? c : -1;
            };
            auto publish = [&](int idx, int item) {  // queue slot free once the nine consumer warps have read its previous content
                const int sl = idx & 3;
                mbar_wait(&work_empty[sl], (uint32_t)(((idx >> 2) & 1) ^ 1));
                work[sl] = item;
                mbar_arrive(&work_full[sl]);
            };
            int w = (int)blockIdx.x;  // grid <= nwork
            publish(0, w);
            int wn = claim();  // the only claim whose round trip is exposed (once per CTA)
            publish(1, wn);
            for (; w >= 0; ++it) {
                const int wnn = claim();  // item it + 2: claimed two items ahead, used (published) at the end of this item's loads
                int bh, pr;
                fwd_decode_item(p, w, bh, pr);
                const int b_idx = bh / p.H, h_idx = bh % p.H, q0 = pr * 2 * FA_BQ;
                const int nblk0 = blocks_of(q0, 0), nblk1 = blocks_of(q0, 1), nmax = max(nblk0, nblk1);
                const int ntile_q = nblk1 > 0 ? 2 : 1;
                mbar_wait(q_empty, (uint32_t)((it & 1) ^ 1));  // the previous item's last S MMA has read Q (first item: passes at once)
                mbar_arrive_expect_tx(q_full, ntile_q * TILE_BYTES);
                for (int t = 0; t < ntile_q; ++t)
#pragma unroll
                    for (int a = 0; a < ATOMS; ++a) tma_load_4d(sQ + t * TILE_BYTES + a * ATOM_BYTES, &tmap_q, q_full, a * 64, q0 + t * FA_BQ, h_idx, b_idx);
                for (int i = 0; i < 2 * nmax; ++i) {  // K_0, V_0, K_1, V_1, ...
                    const int kv0 = (i >> 1) * FA_BKV;
                    const CUtensorMap *tm = (i & 1) ? &tmap_v : &tmap_k;
                    mbar_wait(&kv_empty[s], ph ^ 1);
                    mbar_arrive_expect_tx(&kv_full[s], TILE_BYTES);
#pragma unroll
                    for (int a = 0; a < ATOMS; ++a) tma_load_4d(sKV + s * TILE_BYTES + a * ATOM_BYTES, tm, &kv_full[s], a * 64, kv0, h_idx, b_idx);
                    if (++s == NS) {
                        s = 0;
                        ph ^= 1;
                    }
                }
                publish(it + 2, wnn);
                w = wn;
                wn = wnn;
            }
        }
        __syncwarp();
      } else if (warp == W_MMA) {
        // ===================================================== MMA issuer (converged warp, elected lane issues)
        const bool leader = elect_one();
        const int fmt = p.is_bf16 ? 1 : 0;
        const uint32_t idesc_s = make_idesc_f16(fmt, 0, 0, FA_BQ, FA_BKV);
        const uint32_t idesc_pv = make_idesc_f16(fmt, 0, 1, FA_BQ, D);
        const uint32_t q_addr = smem_u32(sQ), kv_addr = smem_u32(sKV);
        auto issue_s = [&](int t, uint32_t k_addr) {
#pragma unroll
            for (int kk = 0; kk < D / 16; ++kk) {
                const uint32_t off = (uint32_t)((kk >> 2) * ATOM_BYTES + (kk & 3) * 32);
                umma_f16_p(tmem_base + (uint32_t)(t * 128), make_sw128_desc(q_addr + t * TILE_BYTES + off, 0, 1024),
                           make_sw128_desc(k_addr + off, 0, 1024), idesc_s, kk ? 1u : 0u, leader);
            }
        };
        auto issue_pv = [&](int t, uint32_t v_addr, bool accumulate, int half) {
#pragma unroll
            for (int kk = half * (FA_BKV / 32); kk < (half + 1) * (FA_BKV / 32); ++kk)
                umma_f16_ts_p(tmem_base + O_COL + (uint32_t)(t * D), tmem_base + (uint32_t)(t * 128) + (uint32_t)(kk * 8),
                              make_sw128_desc(v_addr + kk * 2048, ATOM_BYTES, 1024), idesc_pv, (accumulate || kk) ? 1u : 0u, leader);
        };
        int s = 0, it = 0;
        uint32_t ph = 0;
        auto next_slot = [&]() {
            mbar_wait(&kv_full[s], ph);
            const int cur = s;
            if (++s == NS) {
                s = 0;
                ph ^= 1;
            }
            return cur;
        };
        uint32_t cp0 = 0, cp1 = 0;    // blocks of tile 0 / 1 handed over so far (phase of p_full)
        uint32_t act0 = 0, act1 = 0;  // items in which tile 0 / 1 was active so far (phase of o_free)
        for (int w = get_item(0); w >= 0; w = get_item(++it)) {
            int bh_unused, pr;
            fwd_decode_item(p, w, bh_unused, pr);
            const int q0 = pr * 2 * FA_BQ;
            const int nblk0 = blocks_of(q0, 0), nblk1 = blocks_of(q0, 1), nmax = max(nblk0, nblk1);
            {
                mbar_wait(q_full, (uint32_t)(it & 1));
                const int sk = next_slot();  // K_0
                tc_fence_after();
#pragma unroll
                for (int t = 0; t < 2; ++t)
                    if ((t ? nblk1 : nblk0) > 0) {
                        issue_s(t, kv_addr + sk * TILE_BYTES);
                        umma_commit_p(&s_full[t], leader);
                    }
                umma_commit_p(&kv_empty[sk], leader);
                if (nmax == 1) umma_commit_p(q_empty, leader);
            }
            for (int j = 1; j <= nmax; ++j) {
                const bool tri = trace != nullptr && blockIdx.x == 0 && cp1 < 256 && leader;
                if (tri) trace[cp1 * 16 + 12] = clock64();
                const int sv = next_slot();  // V_{j-1}
                if (tri) trace[cp1 * 16 + 13] = clock64();
                const bool has_k = j < nmax;
                const int sk = has_k ? next_slot() : 0;  // K_j
                if (tri) trace[cp1 * 16 + 14] = clock64();
#pragma unroll
                for (int t = 0; t < 2; ++t) {
                    const int nb_t = t ? nblk1 : nblk0;
                    if (j - 1 < nb_t) {
                        uint32_t &cp = t ? cp1 : cp0;
                        const uint32_t act = t ? act1 : act0;
                        if (j == 1 && act > 0) mbar_wait(&o_free[t], (act - 1) & 1);  // the first P V overwrites O: the previous epilogue must have read it
                        const bool tr = trace != nullptr && blockIdx.x == 0 && cp < 256 && leader;
                        if (tr) trace[cp * 16 + 4 + t * 4] = clock64();
                        mbar_wait(&p_full[2 * t], cp & 1);
                        tc_fence_after();
                        if (tr) trace[cp * 16 + 5 + t * 4] = clock64();
                        issue_pv(t, kv_addr + sv * TILE_BYTES, j > 1, 0);
                        mbar_wait(&p_full[2 * t + 1], cp & 1);
                        tc_fence_after();
                        if (tr) trace[cp * 16 + 6 + t * 4] = clock64();
                        issue_pv(t, kv_addr + sv * TILE_BYTES, true, 1);
                        if (j < nb_t) {
                            issue_s(t, kv_addr + sk * TILE_BYTES);
                            umma_commit_p(&s_full[t], leader);
                        } else {
                            umma_commit_p(&o_full[t], leader);
                        }
                        if (tr) trace[cp * 16 + 7 + t * 4] = clock64();
                        ++cp;
                    }
                }
                umma_commit_p(&kv_empty[sv], leader);
                if (has_k) umma_commit_p(&kv_empty[sk], leader);
                if (j == nmax - 1) umma_commit_p(q_empty, leader);  // the item's last S has been issued
            }
            if (nblk0 > 0) ++act0;
            if (nblk1 > 0) ++act1;
        }
        __syncwarp();
      }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(200));
        // ===================================================== softmax + epilogue: thread = one query row of tile t
        const int t = warp >> 2, q = warp & 3;
        const int r = q * 32 + lane;
        const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
        const uint32_t s_addr = lane_addr + (uint32_t)(t * 128);
        const uint32_t o_addr = lane_addr + O_COL + (uint32_t)(t * D);
        const float sc = p.scale_log2;
        uint32_t cs = 0, co = 0;  // completions of s_full[t] / o_full[t] consumed so far
        int it = 0;
        for (int w = get_item(0); w >= 0; w = get_item(++it)) {
            int bh, pr;
            fwd_decode_item(p, w, bh, pr);
            const int b_idx = bh / p.H, h_idx = bh % p.H, q0 = pr * 2 * FA_BQ;
            const int n_t = blocks_of(q0, t);
            if (n_t == 0) continue;
            const int64_t q0t = (int64_t)q0 + t * FA_BQ;
            const int64_t m_row = q0t + r;
            float m_ref = -INFINITY, l_run = 0.f;
            for (int j = 0; j < n_t; ++j) {
                const int kv0 = j * FA_BKV;
                mbar_wait(&s_full[t], cs & 1);
                ++cs;
                tc_fence_after();
                if (trace != nullptr && blockIdx.x == 0 && cs <= 256 && q == 0 && lane == 0) trace[(cs - 1) * 16 + t * 2] = clock64();
                const bool masked = (kv0 + FA_BKV - 1 > q0t) || (kv0 + FA_BKV > p.Skv);  // diagonal / ragged block (CTA-uniform per tile)
                const int64_t lim64 = min(m_row, p.Skv - 1) - kv0;                       // columns i > lim are masked
                const int lim = (int)max((int64_t)-1, min(lim64, (int64_t)127));
                uint64_t *p_bar = &p_full[2 * t];
                if (trace != nullptr) {  // traced run (bf16, unmasked arithmetic only on full blocks): stamps inside the block for CTA 0, warp 0 of the tile
                    long long *tr = (blockIdx.x == 0 && cs <= 256 && q == 0 && lane == 0) ? trace + 8192 + (cs - 1) * 16 + t * 8 : nullptr;
                    if (masked) fwd_softmax_block<D, true, true, POLY, 1, true>(s_addr, s_addr, o_addr, sc, lim, j == 0, m_ref, l_run, nullptr, nullptr, 0, p_bar, tr);
                    else fwd_softmax_block<D, true, false, POLY, 1, true>(s_addr, s_addr, o_addr, sc, lim, j == 0, m_ref, l_run, nullptr, nullptr, 0, p_bar, tr);
                } else if (p.is_bf16) {
                    if (masked) fwd_softmax_block<D, true, true, POLY, 1>(s_addr, s_addr, o_addr, sc, lim, j == 0, m_ref, l_run, nullptr, nullptr, 0, p_bar);
                    else fwd_softmax_block<D, true, false, POLY, 1>(s_addr, s_addr, o_addr, sc, lim, j == 0, m_ref, l_run, nullptr, nullptr, 0, p_bar);
                } else {
                    if (masked) fwd_softmax_block<D, false, true, POLY, 1>(s_addr, s_addr, o_addr, sc, lim, j == 0, m_ref, l_run, nullptr, nullptr, 0, p_bar);
                    else fwd_softmax_block<D, false, false, POLY, 1>(s_addr, s_addr, o_addr, sc, lim, j == 0, m_ref, l_run, nullptr, nullptr, 0, p_bar);
                }
                if (trace != nullptr && blockIdx.x == 0 && cs <= 256 && q == 0 && lane == 0) trace[(cs - 1) * 16 + t * 2 + 1] = clock64();
            }
            // ---- epilogue: O / l -> 16-bit -> global, row LSE; then hand the O columns back to the MMA warp
            mbar_wait(&o_full[t], co & 1);
            ++co;
            tc_fence_after();
            if (trace != nullptr && blockIdx.x == 0 && co <= 64 && q == 0 && lane == 0) trace[4096 + (co - 1) * 4 + t * 2] = clock64();
            const float inv_l = 1.f / l_run;
            const bool row_ok = m_row < p.Sq;
            fwd_store_tile<D, 1, true>(o_addr, inv_l, p.is_bf16, sO + t * ATOM_BYTES, &tmap_o, (int)q0t, h_idx, b_idx, r, 1 + t, &o_free[t],
                                       (trace != nullptr && blockIdx.x == 0 && co <= 64 && r == 0) ? trace + 6144 + (co - 1) * 16 + t * 8 : nullptr);
            if (row_ok && p.lse) p.lse[(int64_t)bh * p.Sq + m_row] = (m_ref + log2f(l_run)) * 0.6931471805599453f;
            if (trace != nullptr && blockIdx.x == 0 && co <= 64 && q == 0 && lane == 0) trace[4096 + (co - 1) * 4 + t * 2 + 1] = clock64();
        }
        if (r == 0) tma_store_wait_all<0>();  // the staging buffer must outlive the last bulk copy
    }
    tc_fence_before();
    __syncthreads();
    if (warp == W_MMA) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

template <int D, int POLY>
__global__ void __launch_bounds__(FaCfg<1>::THREADS, 1)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                   const __grid_constant__ CUtensorMap tmap_v, const __grid_constant__ CUtensorMap tmap_o, const AttnTcParams p) {
    attn_fwd_tc_body<D, POLY, 1>(tmap_q, tmap_k, tmap_v, tmap_o, p);
}
template <int D, int POLY>
__global__ void __launch_bounds__(FaCfg<2>::THREADS, 1)
attn_fwd_tc2_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                    const __grid_constant__ CUtensorMap tmap_v, const __grid_constant__ CUtensorMap tmap_o, const AttnTcParams p) {
    attn_fwd_tc_body<D, POLY, 2>(tmap_q, tmap_k, tmap_v, tmap_o, p);
}

template <int D, int POLY, int NH>
static void launch_fwd_tc(const AttnPlan &a) {
    Runtime &rt = Runtime::get();
    const bool bf16 = a.dtype == KF_BFLOAT16;
    // dense [BH, S, D] operands are the H = BH, B = 1 case of the strided 4-D maps
    const bool dense = a.H <= 0;
    const int64_t H = dense ? a.BH : a.H, B = a.BH / H;
    const AttnLayout lq = dense ? AttnLayout{a.BH * a.Sq * D, a.Sq * D, D} : a.lq, lkv_d = AttnLayout{a.BH * a.Skv * D, a.Skv * D, D};
    const AttnLayout lk = dense ? lkv_d : a.lk, lv = dense ? lkv_d : a.lv, lo = dense ? lq : a.lo;
    auto map = [&](const void *ptr, int64_t S, const AttnLayout &l) {
        return make_tmap_4d_16bit(ptr, bf16, D, (uint64_t)S, (uint64_t)H, (uint64_t)B, (uint64_t)l.ss, (uint64_t)l.sh, (uint64_t)l.sb, 64, 128);
    };
    const CUtensorMap tq = map(a.q, a.Sq, lq), tk = map(a.k, a.Skv, lk), tv = map(a.v, a.Skv, lv), to = map(a.out, a.Sq, lo);
    AttnTcParams p{};
    p.H = (int)H;
    p.lo = lo;
    p.BH = a.BH; p.Sq = a.Sq; p.Skv = a.Skv;
    p.out = a.out;
    p.lse = reinterpret_cast<float *>(a.lse);
    p.scale_log2 = (float)(1.4426950408889634 / std::sqrt((double)D));
    p.npairs = (int)((a.Sq + 2 * FA_BQ - 1) / (2 * FA_BQ));
    p.is_bf16 = bf16;
    {  // KF_ATTN_STALE=1: exponentials relative to the running reference, per-half maximum only checked (read per call for A/B runs).
       // Measured at C3 (run r2v): 1.090 ms against 1.039 ms for the classic order, so it stays off by default.
        const char *st = std::getenv("KF_ATTN_STALE");
        p.stale = (st && st[0] == '1') ? 1 : 0;
    }
    p.trace = nullptr;
    {  // heads per scheduling group: K + V of a group ~ an eighth of the L2 (16 MB), the best of a sweep at S = 1024 / 4096 / 8192
       // (tools/gpu_attn_hg_sweep.py: S = 4096 fastest call 1.058 ms head-major, 1.020 at 8 - 16 heads, 1.033 at 32, 1.174 for one
       // global list; S = 8192 best at 4 heads, S = 1024 at >= 16).  KF_ATTN_HG overrides (read per call).
        const int64_t kv_bytes = 2 * a.Skv * D * 2;
        // the persistent kernel (short sequences: everything fits the L2 anyway) prefers the longer longest-first runs of big groups:
        // S = 1024, 126 heads per group 0.388 ms, 31 heads 0.432 ms (gpurun_out/r6c, r21)
        const int64_t budget = (int64_t)rt.props().l2_bytes / (NH == 4 ? 2 : 8);
        int64_t hg = std::max<int64_t>(1, budget / std::max<int64_t>(1, kv_bytes));
        if (const char *e = std::getenv("KF_ATTN_HG")) hg = std::max(1, std::atoi(e));
        p.hg = (int)std::min<int64_t>(hg, a.BH);
    }
    if constexpr (NH == 5) {  // 64-key blocks, next S issued before P V (attn_fwd_k64_kernel)
        auto map64 = [&](const void *ptr, int64_t S, const AttnLayout &l) {
            return make_tmap_4d_16bit(ptr, bf16, D, (uint64_t)S, (uint64_t)H, (uint64_t)B, (uint64_t)l.ss, (uint64_t)l.sh, (uint64_t)l.sb, 64, 64);
        };
        const CUtensorMap tk64 = map64(a.k, a.Skv, lk), tv64 = map64(a.v, a.Skv, lv);
        constexpr int SMEM_K = 2 * 128 * D * 2 + 8 * 64 * D * 2 + 256 + 1024;
        static bool attr_k = false;
        if (!attr_k) {
            KF_CUDA(cudaFuncSetAttribute(attn_fwd_k64_kernel<D, POLY>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_K));
            attr_k = true;
        }
        const int64_t grid = a.BH * p.npairs;
        KF_CHECK(grid < (int64_t)0x7FFFFFFF);
        attn_fwd_k64_kernel<D, POLY><<<(unsigned)grid, FaCfg<1>::THREADS, SMEM_K, rt.stream()>>>(tq, tk64, tv64, to, p);
        rt.post_launch("attn_fwd_k64_kernel");
        return;
    }
    if constexpr (NH == 4) {  // persistent kernel: one CTA per SM over the longest-first work list
        constexpr int SMEM_P = (2 + FA_NSTAGE) * 128 * D * 2 + 2 * 128 * 128 + 512 + 1024;  // tiles + epilogue staging + barriers, work queue + alignment slack
        static bool attr_p = false;
        if (!attr_p) {
            KF_CUDA(cudaFuncSetAttribute(attn_fwd_pers_kernel<D, POLY, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_P));
            if constexpr (D == 128 && POLY == 2)
                KF_CUDA(cudaFuncSetAttribute(attn_fwd_pers_kernel<D, POLY, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_P));
            attr_p = true;
        }
        const int64_t nwork = a.BH * p.npairs;
        KF_CHECK(nwork < (int64_t)0x7FFFFFFF);
        static const bool want_trace = std::getenv("KF_ATTN_TRACE") != nullptr && D == 128 && POLY == 2;  // the one traced instance
        constexpr size_t TRACE_WORDS = 8192 + 4096;
        Scratch trace_buf(want_trace ? TRACE_WORDS * 8 : 16);
        if (want_trace) {
            rt.memset_async(trace_buf.p, 0, TRACE_WORDS * 8);
            p.trace = trace_buf.as<long long>();
        }
        const int64_t grid = std::min<int64_t>(nwork, rt.props().sm_count);
        Scratch counter(16);
        rt.memset_async(counter.p, 0, 16);
        p.work_counter = counter.as<unsigned int>();
        if constexpr (D == 128 && POLY == 2) {
            if (want_trace) attn_fwd_pers_kernel<D, POLY, true><<<(unsigned)grid, FaCfg<1>::THREADS, SMEM_P, rt.stream()>>>(tq, tk, tv, to, p);
            else attn_fwd_pers_kernel<D, POLY, false><<<(unsigned)grid, FaCfg<1>::THREADS, SMEM_P, rt.stream()>>>(tq, tk, tv, to, p);
        } else {
            attn_fwd_pers_kernel<D, POLY, false><<<(unsigned)grid, FaCfg<1>::THREADS, SMEM_P, rt.stream()>>>(tq, tk, tv, to, p);
        }
        rt.post_launch("attn_fwd_pers_kernel");
        if (want_trace) {  // bring-up aid: CTA 0's pipeline, per 128-key block (clocks)
            rt.sync();
            std::vector<long long> h(TRACE_WORDS);
            KF_CUDA(cudaMemcpy(h.data(), trace_buf.p, h.size() * 8, cudaMemcpyDeviceToHost));
            std::printf("[attn fwd trace] blk | tile0: S seen (delta to previous)  softmax | tile1: S seen (delta)  softmax | mma tile0: wait_p0 wait_p1 issue | mma tile1: wait_p0 wait_p1 issue\n");
            for (int n = 0; n < 256 && (h[n * 16] || h[n * 16 + 2]); ++n) {
                const long long *e = &h[n * 16], *pe = &h[(n ? n - 1 : 0) * 16];
                std::printf("[attn fwd trace] %3d | %6lld %6lld | %6lld %6lld | %6lld %6lld %6lld | %6lld %6lld %6lld\n", n, e[0] - pe[0], e[1] - e[0], e[2] - pe[2],
                            e[3] - e[2], e[5] - e[4], e[6] - e[5], e[7] - e[6], e[9] - e[8], e[10] - e[9], e[11] - e[10]);
            }
            std::printf("[attn fwd trace] blk | mma: ring wait V, ring wait K, end of previous issue -> loop top | softmax tile0: S seen->ld done, max, exps h0, st+arrive h0, exps h1, st+arrive h1 | tile1 same\n");
            for (int n = 0; n < 256 && (h[n * 16] || h[n * 16 + 2]); ++n) {
                const long long *e = &h[n * 16], *pe = &h[(n ? n - 1 : 0) * 16], *x = &h[8192 + n * 16];
                std::printf("[attn fwd trace2] %3d | %6lld %6lld %6lld | %5lld %5lld %5lld %5lld %5lld %5lld | %5lld %5lld %5lld %5lld %5lld %5lld\n", n, e[13] - e[12], e[14] - e[13],
                            e[12] - pe[11], x[0] - e[0], x[1] - x[0], x[2] - x[1], x[3] - x[2], x[4] - x[3], x[5] - x[4], x[8] - e[2], x[9] - x[8], x[10] - x[9],
                            x[11] - x[10], x[12] - x[11], x[13] - x[12]);
            }
            for (int n = 0; n < 64 && (h[4096 + n * 4] || h[4096 + n * 4 + 2]); ++n) {
                std::printf("[attn fwd trace] item %2d epilogue tile0 %6lld clk, tile1 %6lld clk, tile1 start - tile0 start %6lld |", n, h[4096 + n * 4 + 1] - h[4096 + n * 4],
                            h[4096 + n * 4 + 3] - h[4096 + n * 4 + 2], h[4096 + n * 4 + 2] - h[4096 + n * 4]);
                for (int t = 0; t < 2; ++t) {  // per pass: wait for the staging buffer, O -> shared memory, fence + barrier, store issue
                    const long long *e = &h[6144 + n * 16 + t * 8], st = h[4096 + n * 4 + t * 2];
                    std::printf(" t%d: %5lld %5lld %5lld %5lld | %5lld %5lld %5lld %5lld ;", t, e[0] - st, e[1] - e[0], e[2] - e[1], e[3] - e[2], e[4] - e[3], e[5] - e[4],
                                e[6] - e[5], e[7] - e[6]);
                }
                std::printf("\n");
            }
        }
        return;
    }
    // NH = 3 stands for the "P through shared memory" kernel (one thread per row): tiles + 2 x 16 KB of P + barriers + alignment slack
    constexpr int SMEM = NH == 3 ? (2 + FA_NSTAGE) * 128 * D * 2 + 2 * 128 * 128 + 256 + 1024 : (2 + FA_NSTAGE) * 128 * D * 2 + 256 + 4096 + 1024;
    void (*kern)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap, const AttnTcParams);
    if constexpr (NH == 1) kern = attn_fwd_tc_kernel<D, POLY>;
    else if constexpr (NH == 3) kern = attn_fwd_ps_kernel<D, POLY>;
    else if constexpr (NH == 2) kern = attn_fwd_tc2_kernel<D, POLY>;
    else kern = nullptr;
    static bool attr_done = false;
    if (!attr_done) {
        KF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
        attr_done = true;
    }
    const int64_t grid = a.BH * p.npairs;
    KF_CHECK(grid < (int64_t)0x7FFFFFFF);
    kern<<<(unsigned)grid, FaCfg<NH == 2 ? 2 : 1>::THREADS, SMEM, rt.stream()>>>(tq, tk, tv, to, p);
    rt.post_launch(NH == 1 ? "attn_fwd_tc_kernel" : NH == 3 ? "attn_fwd_ps_kernel" : "attn_fwd_tc2_kernel");
}

bool launch_attention_fwd_tc(const AttnPlan &a) {
    static const bool force_simt = std::getenv("KF_ATTN_FORCE_SIMT") != nullptr;
    if (force_simt) return false;
    if (a.dtype != KF_HALF && a.dtype != KF_BFLOAT16) return false;
    if (a.D != 64 && a.D != 128) return false;
    if (a.Sq < 1 || a.Skv < 1 || a.BH < 1 || a.BH >= 65536) return false;
    auto al = [](const void *p) { return reinterpret_cast<uintptr_t>(p) % 16 == 0; };
    if (!al(a.q) || !al(a.k) || !al(a.v) || !al(a.out)) return false;
    if (a.H > 0) {  // strided operands: every stride a multiple of 8 elements (16 bytes), rows at least D apart
        auto lok = [&](const AttnLayout &l) { return l.sb % 8 == 0 && l.sh % 8 == 0 && l.ss % 8 == 0 && l.ss >= a.D; };
        if (a.BH % a.H != 0 || !lok(a.lq) || !lok(a.lk) || !lok(a.lv) || !lok(a.lo)) return false;
    }
    // share of the exponentials computed on the FMA pipe instead of the MUFU unit, in eighths (KF_ATTN_POLY; measured at C3:
    // 0 -> 1066, 2 -> 1082, 3 -> 1064 TFLOP/s with the split P hand-off; default 2)
    static const int poly = std::getenv("KF_ATTN_POLY") ? std::atoi(std::getenv("KF_ATTN_POLY")) : 2;
    // threads per query row in the softmax (see fwd_softmax_block).  Measured at C3: one thread per row 1.08 ms (1017 TFLOP/s), two
    // threads per row 1.34 ms (818 TFLOP/s) — the extra named barrier, doubled polling and 96-register budget cost more than the
    // shorter dependent chains gain, so KF_ATTN_SPLIT=2 stays an opt-in experiment
    static const int nh = std::getenv("KF_ATTN_SPLIT") ? std::atoi(std::getenv("KF_ATTN_SPLIT")) : 1;
    // KF_ATTN_FWD (read per call): "pers" = persistent kernel (one CTA per SM, dynamic work queue), "cta" = one CTA per query pair,
    // "ps" = P through shared memory with the next S issued early; unset = by KV length.  Both kernels walk the items in the order
    // of fwd_decode_item.  Measured with the variants alternating call by call (gpurun_out/r6c_fwd_ab.log, B H S constant, fastest
    // call in ms, persistent / per-CTA): S = 1024 0.388 / 0.421 (head-major order: 0.478 / 0.479), 2048 0.600 / 0.624, 4096
    // 1.016 / 1.026, 8192 1.880 / 1.821, 16384 equal within the power-cap noise (the first ~20 ms of a series run at 1270 TFLOP/s,
    // later calls at ~1050 for either kernel).  The persistent kernel's kept-alive ring, barriers and tensor-memory allocation
    // pay for short items; for long ones the per-CTA kernel is as fast and simpler.
    const char *fwd_mode = std::getenv("KF_ATTN_FWD");
    const bool ps = fwd_mode && std::strcmp(fwd_mode, "ps") == 0;
    const bool k64 = fwd_mode && std::strcmp(fwd_mode, "k64") == 0;
    const bool per_cta = fwd_mode ? std::strcmp(fwd_mode, "pers") != 0 : a.Skv > 2048;
#define KF_FWD(DD, PP)                                                          \
    do {                                                                        \
        if (ps) launch_fwd_tc<DD, PP, 3>(a);                                    \
        else if (k64) launch_fwd_tc<DD, PP, 5>(a);                              \
        else if (nh == 1 && !per_cta && !p_stale) launch_fwd_tc<DD, PP, 4>(a);  \
        else if (nh == 1) launch_fwd_tc<DD, PP, 1>(a);                          \
        else launch_fwd_tc<DD, PP, 2>(a);                                       \
    } while (0)
    const char *st_env = std::getenv("KF_ATTN_STALE");
    const bool p_stale = st_env && st_env[0] == '1';
    if (a.D == 64) {
        if (poly <= 0) KF_FWD(64, 0);
        else KF_FWD(64, 3);
    } else {
        if (poly <= 0) KF_FWD(128, 0);
        else if (poly == 2) KF_FWD(128, 2);
        else if (poly == 4) KF_FWD(128, 4);
        else KF_FWD(128, 3);
    }
#undef KF_FWD
    return true;
}

}  // namespace kf
