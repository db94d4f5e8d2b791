// Microbenchmark: tcgen05.mma issue->complete rate per SM for the shapes the attention kernels use (cta_group::1, kind::f16):
//   SS M128 N128 K16 (both operands in shared memory, 128B swizzle), SS N64, TS (A from tensor memory) N128, and the same with a
//   concurrent tcgen05.ld stream from 16 warps.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I kfunca_b200/csrc/kernels ...
#include <cstdio>
#include "tc_common.cuh"
using namespace kf::tc;

// mode: 0 = SS N128, 1 = SS N64, 2 = TS N128, 3 = SS N128 K-major A + MN-major B
template <int MODE, bool LDTM, int REPS>
__global__ void __launch_bounds__(576, 1) k(int groups, unsigned long long *out, int randomize, const unsigned char *gsrc, int commit_each) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __shared__ uint64_t tbar, dbar;
    if (randomize) {  // bf16 values in (-1, 1) with random mantissas: realistic toggling on the operand buses
        uint32_t x = threadIdx.x * 2654435761u + blockIdx.x * 40503u + 12345u;
        for (int i = threadIdx.x; i < 131072 / 4; i += blockDim.x) {
            x = x * 1664525u + 1013904223u;
            const uint32_t lo = 0x3f00u | ((x >> 8) & 0x80ffu) ^ ((x >> 3) & 0x8000u), hi = 0x3e80u | ((x >> 16) & 0x80ffu);
            reinterpret_cast<uint32_t *>(smem)[i] = lo | (hi << 16);
        }
        fence_proxy_async();
    }
    if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_init(&tbar, 1); mbar_init(&dbar, 1 << 20); fence_barrier_init(); }
    if (warp == 16) tmem_alloc(&slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tb = slot;
    if (warp == 16) {
        const bool leader = elect_one();
        const uint32_t a = smem_u32(smem), b = a + 65536;
        const uint32_t idesc = make_idesc_f16(1, 0, MODE == 3 ? 1 : 0, 128, MODE == 1 ? 64 : 128);
        const long long t0 = clock64();
        for (int g = 0; g < groups; ++g) {
            for (int rep = 0; rep < REPS; ++rep) {
            if (commit_each && rep) umma_commit_p(&dbar, leader);
#pragma unroll
            for (int kk = 0; kk < 8; ++kk) {
                const uint32_t off = (uint32_t)((kk >> 2) * 16384 + (kk & 3) * 32);
                if (MODE == 4 || MODE == 5) {
                    const uint32_t dreg = (MODE == 4) ? 0u : 128u;  // 4: the SS MMA overwrites the TS MMA's A columns; 5: disjoint columns
                    if (rep & 1) umma_f16_p(tb + dreg, make_sw128_desc(a + off, 0, 1024), make_sw128_desc(b + off, 0, 1024), make_idesc_f16(1, 0, 0, 128, 128), kk ? 1u : 0u, leader);
                    else umma_f16_ts_p(tb + 256, tb + (uint32_t)((kk >> 1) * 32 + (kk & 1) * 8), make_sw128_desc(b + kk * 2048, 16384, 1024), make_idesc_f16(1, 0, 1, 128, 128), kk ? 1u : 0u, leader);
                } else if (MODE == 2) umma_f16_ts_p(tb + 256, tb + (uint32_t)((kk >> 1) * 32 + (kk & 1) * 8), make_sw128_desc(b + kk * 2048, 16384, 1024), make_idesc_f16(1, 0, 1, 128, 128), kk ? 1u : 0u, leader);
                else if (MODE == 3) umma_f16_p(tb + 256, make_sw128_desc(a + off, 0, 1024), make_sw128_desc(b + kk * 2048, 16384, 1024), idesc, kk ? 1u : 0u, leader);
                else umma_f16_p(tb + 256, make_sw128_desc(a + off, 0, 1024), make_sw128_desc(b + off, 0, 1024), idesc, kk ? 1u : 0u, leader);
            }
            }
            umma_commit_p(&bar, leader);
            mbar_wait(&bar, (uint32_t)(g & 1));
        }
        const long long t1 = clock64();
        if (lane == 0) out[blockIdx.x] = (unsigned long long)(t1 - t0);
    } else if (warp == 17 && gsrc != nullptr) {
        if (lane == 0) {
            for (int i = 0; i < groups * REPS / 2; ++i) {  // ~32 KB per 8 MMAs (the attention kernels' ratio): 2 x 16 KB per 2 groups... one 32 KB copy per group pair
                mbar_arrive_expect_tx(&tbar, 32768);
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(smem_u32(smem + 131072)), "l"(reinterpret_cast<uint64_t>(gsrc + ((size_t)(blockIdx.x * 64 + (i & 63)) << 15))), "r"(32768), "r"(smem_u32(&tbar)) : "memory");
                mbar_wait(&tbar, (uint32_t)(i & 1));
            }
        }
    } else if (LDTM && warp < 16) {
        uint32_t r[32], sink = 0;
        const uint32_t base = tb + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * 32);
        for (int i = 0; i < groups * 4 * REPS; ++i) {
            tmem_ld32(base, r);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) sink ^= r[j];
        }
        if (sink == 0x1234567) out[200] = sink;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 16) { tc_fence_after(); tmem_dealloc(tb, 512); }
}

template <int MODE, bool LDTM, int REPS>
void run(const char *name, int randomize = 0, bool tma = false, int commit_each = 0) {
    unsigned long long *out;
    cudaMalloc(&out, 256 * 8);
    const int groups = 2000, smem = 131072 + 32768;
    static unsigned char *gsrc = nullptr;
    if (!gsrc) { cudaMalloc(&gsrc, (size_t)148 * 64 << 15); cudaMemset(gsrc, 1, (size_t)148 * 64 << 15); }
    cudaFuncSetAttribute(k<MODE, LDTM, REPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    k<MODE, LDTM, REPS><<<148, 576, smem>>>(groups, out, randomize, tma ? gsrc : nullptr, commit_each);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); return; }
    unsigned long long h[148];
    cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
    printf("%-66s %2d MMAs per commit: %7.1f clk per commit = %5.1f clk per MMA incl. the ~270 clk commit -> wake-up\n", name, 8 * REPS, (double)h[0] / groups, (double)h[0] / groups / (8 * REPS));
}

int main() {
    run<0, false, 1>("SS M128 N128 K16, 8 MMAs + commit + wait");
    run<0, false, 8>("SS M128 N128 K16, 64 MMAs per commit");
    run<1, false, 8>("SS M128 N64  K16");
    run<2, false, 8>("TS M128 N128 K16 (A in TMEM, B MN-major)");
    run<3, false, 8>("SS M128 N128 K16 (B MN-major)");
    run<0, true, 8>("SS N128 + 16 warps of tcgen05.ld");
    run<2, true, 8>("TS N128 + 16 warps of tcgen05.ld");
    run<0, false, 8>("SS N128, random operands", 1);
    run<0, false, 8>("SS N128, random operands + TMA 32 KB / 16 MMAs", 1, true);
    run<0, true, 8>("SS N128, random + TMA + 16 warps tcgen05.ld", 1, true);
    run<0, false, 8>("SS N128, commit after every 8 MMAs (nobody waits)", 0, false, 1);
    run<5, false, 8>("alternate TS(A = cols 0-63, D = 256) / SS(D = cols 128-255)");
    run<4, false, 8>("alternate TS(A = cols 0-63, D = 256) / SS(D = cols 0-127, over A)");
    return 0;
}
