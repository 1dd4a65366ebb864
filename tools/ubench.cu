// ubench.cu — issue-rate microbenchmarks for the integer pipes of one B200 SM sub-partition (SMSP).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench tools/ubench.cu ; run on a B200.
// Prints warp-instructions per cycle per SMSP for independent-chain loops of each instruction class, which is
// what bounds the multispin sweep kernel (LOP3-heavy with Philox IMAD.WIDEs).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int CH = 12;     // independent chains per thread
constexpr int IT = 4096;  // loop iterations

enum { K_LOP3, K_IMADW, K_IMAD, K_IADD3, K_SHF, K_SETSEL, K_MIX_LOP_IMADW_2_1, K_MIX_LOP_IMAD_1_1, K_MIX_LOP_IMADW_1_1, K_LDS128, K_PRMT, K_MIX_LOP_IMADW_4_1, K_IMADHI, K_MIX_LOP_IMAD_2_1, K_MIX_LOP_IMAD_3_1, K_MIX_LOP_VIADD_1_1, K_MIX_LOP_IMADW_3_1, K_LDS128_SEQ, K_MULW, K_MULHILO, K_MIX_LOP_MULW_4_1, K_MIX_LOP_MULHILO_4_1, K_SETP_PLOP, K_POPC, K_VOTE, K_N };
const char *NAMES[] = {"lop3", "imad.wide.u32", "imad.lo", "iadd3", "shf", "isetp+sel", "lop3:imad.wide 2:1", "lop3:imad 1:1", "lop3:imad.wide 1:1", "lds.128", "prmt", "lop3:imad.wide 4:1", "imad.hi.s32", "lop3:imad 2:1", "lop3:imad 3:1", "lop3:add 1:1", "lop3:imad.wide 3:1", "lds.128 conflict-free", "mul.wide.u32 (no addend)", "mul.hi + mul.lo", "lop3:mul.wide 4:1", "lop3:(mul.hi+mul.lo) 4:1", "setp + @p lop3", "popc", "vote.ballot"};

template <int KIND>
__global__ void __launch_bounds__(1024, 1) bench(uint32_t *out, uint32_t seed, unsigned long long *cycles) {
    __shared__ uint4 sm[1024];
    uint32_t x[CH];
    unsigned long long w[CH];
    const uint32_t a = seed | 1u, b = seed * 3u + 7u;
#pragma unroll
    for (int c = 0; c < CH; c++) { x[c] = seed + c * 977u + threadIdx.x; w[c] = x[c]; }
    sm[threadIdx.x] = make_uint4(x[0], x[1], x[2], x[3]);
    __syncthreads();
    const unsigned long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < IT; i++) {
#pragma unroll
        for (int c = 0; c < CH; c++) {
            if (KIND == K_LOP3) asm("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[c]) : "r"(a), "r"(b));
            if (KIND == K_IMADW) asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[c]) : "r"((uint32_t)w[c]), "r"(b));
            if (KIND == K_IMAD) asm("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[c]) : "r"(a), "r"(b));
            if (KIND == K_IMADHI) asm("mul.hi.s32 %0, %0, %1;" : "+r"(x[c]) : "r"(a));
            if (KIND == K_IADD3) asm("add.u32 %0, %0, %1;" : "+r"(x[c]) : "r"(a));
            if (KIND == K_SHF) asm("shf.l.wrap.b32 %0, %0, %1, 7;" : "+r"(x[c]) : "r"(a));
            if (KIND == K_PRMT) asm("prmt.b32 %0, %0, %1, 0x2103;" : "+r"(x[c]) : "r"(a));
            if (KIND == K_SETSEL) asm("{.reg .pred p; setp.lt.u32 p, %0, %1; selp.u32 %0, %2, %0, p;}" : "+r"(x[c]) : "r"(a), "r"(b));
            if (KIND == K_MIX_LOP_IMADW_2_1) {
                if (c % 3 == 2) asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[c]) : "r"((uint32_t)w[c]), "r"(b));
                else asm("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[c]) : "r"(a), "r"(b));
            }
            if (KIND == K_MIX_LOP_IMADW_4_1) {
                if (c % 5 == 4) asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[c]) : "r"((uint32_t)w[c]), "r"(b));
                else asm("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[c]) : "r"(a), "r"(b));
            }
            if (KIND == K_MIX_LOP_IMADW_1_1) {
                if (c % 2) asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[c]) : "r"((uint32_t)w[c]), "r"(b));
                else asm("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[c]) : "r"(a), "r"(b));
            }
            if (KIND == K_MIX_LOP_IMAD_1_1) {
                if (c % 2) asm("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[c]) : "r"(a), "r"(b));
                else asm("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[c]) : "r"(a), "r"(b));
            }
            if (KIND == K_MIX_LOP_IMAD_2_1) {
                if (c % 3 == 2) asm("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[c]) : "r"(a), "r"(b));
                else asm("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[c]) : "r"(a), "r"(b));
            }
            if (KIND == K_MIX_LOP_IMAD_3_1) {
                if (c % 4 == 3) asm("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[c]) : "r"(a), "r"(b));
                else asm("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[c]) : "r"(a), "r"(b));
            }
            if (KIND == K_MIX_LOP_IMADW_3_1) {
                if (c % 4 == 3) asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[c]) : "r"((uint32_t)w[c]), "r"(b));
                else asm("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[c]) : "r"(a), "r"(b));
            }
            if (KIND == K_MIX_LOP_VIADD_1_1) {
                if (c % 2) asm("add.u32 %0, %0, %1;" : "+r"(x[c]) : "r"(a));
                else asm("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[c]) : "r"(a), "r"(b));
            }
            if (KIND == K_MULW) asm("mul.wide.u32 %0, %1, %2;" : "=l"(w[c]) : "r"((uint32_t)w[c] ^ (uint32_t)(w[c] >> 32)), "r"(b));
            if (KIND == K_MULHILO) {
                uint32_t hi, lo;
                asm("mul.hi.u32 %0, %1, %2;" : "=r"(hi) : "r"(x[c]), "r"(b));
                asm("mul.lo.u32 %0, %1, %2;" : "=r"(lo) : "r"(x[c]), "r"(b));
                x[c] = hi ^ lo;
            }
            if (KIND == K_MIX_LOP_MULW_4_1) {
                if (c % 5 == 4) asm("mul.wide.u32 %0, %1, %2;" : "=l"(w[c]) : "r"((uint32_t)w[c] ^ (uint32_t)(w[c] >> 32)), "r"(b));
                else asm("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[c]) : "r"(a), "r"(b));
            }
            if (KIND == K_MIX_LOP_MULHILO_4_1) {
                if (c % 5 == 4) {
                    uint32_t hi, lo;
                    asm("mul.hi.u32 %0, %1, %2;" : "=r"(hi) : "r"(x[c]), "r"(b));
                    asm("mul.lo.u32 %0, %1, %2;" : "=r"(lo) : "r"(x[c]), "r"(b));
                    x[c] = hi ^ lo;
                } else asm("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[c]) : "r"(a), "r"(b));
            }
            if (KIND == K_SETP_PLOP) asm("{.reg .pred p; setp.lt.u32 p, %0, %1; @p or.b32 %0, %0, 0x10; add.u32 %0, %0, %2;}" : "+r"(x[c]) : "r"(a), "r"(b));
            if (KIND == K_POPC) asm("popc.b32 %0, %0;" : "+r"(x[c]));
            if (KIND == K_VOTE) asm("{.reg .pred p; setp.lt.u32 p, %0, %1; vote.sync.ballot.b32 %0, p, 0xffffffff;}" : "+r"(x[c]) : "r"(a));
            if (KIND == K_LDS128_SEQ) {
                uint4 v = sm[(threadIdx.x + (x[c] & 0x3e0)) & 1023];
                x[c] = (v.x ^ v.y) + (v.z ^ v.w);
            }
            if (KIND == K_LDS128) {
                uint4 v = sm[(x[c] + c * 33) & 1023];
                x[c] = (v.x ^ v.y) + (v.z ^ v.w);  // keeps all four words and the address chain alive
            }
        }
    }
    const unsigned long long t1 = clock64();
    uint32_t acc = 0;
#pragma unroll
    for (int c = 0; c < CH; c++) acc ^= x[c] ^ (uint32_t)w[c] ^ (uint32_t)(w[c] >> 32);
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int KIND>
void run(uint32_t *out, unsigned long long *cyc, int threads) {
    const int blocks = 148;
    bench<KIND><<<blocks, threads>>>(out, 12345u, cyc);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    bench<KIND><<<blocks, threads>>>(out, 12345u, cyc);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    unsigned long long h[148];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double mean = 0;
    for (int i = 0; i < blocks; i++) mean += (double)h[i];
    mean /= blocks;
    const double warps_per_smsp = threads / 32.0 / 4.0;
    const double inst = (double)IT * CH * warps_per_smsp;  // warp-instructions of the measured class per SMSP
    printf("%-22s threads=%4d  %.3f warp-inst/cycle/SMSP  (%.0f cycles, %.3f ms)\n", NAMES[KIND], threads, inst / mean, mean, ms);
}

int main() {
    uint32_t *out;
    unsigned long long *cyc;
    cudaMalloc(&out, 148 * 1024 * 4);
    cudaMalloc(&cyc, 148 * 8);
    for (int threads : {512, 1024}) {
        run<K_LOP3>(out, cyc, threads);
        run<K_IMADW>(out, cyc, threads);
        run<K_IMAD>(out, cyc, threads);
        run<K_IMADHI>(out, cyc, threads);
        run<K_IADD3>(out, cyc, threads);
        run<K_SHF>(out, cyc, threads);
        run<K_PRMT>(out, cyc, threads);
        run<K_SETSEL>(out, cyc, threads);
        run<K_MIX_LOP_IMADW_4_1>(out, cyc, threads);
        run<K_MIX_LOP_IMADW_2_1>(out, cyc, threads);
        run<K_MIX_LOP_IMADW_1_1>(out, cyc, threads);
        run<K_MIX_LOP_IMAD_1_1>(out, cyc, threads);
        run<K_MIX_LOP_IMAD_2_1>(out, cyc, threads);
        run<K_MIX_LOP_IMAD_3_1>(out, cyc, threads);
        run<K_MIX_LOP_IMADW_3_1>(out, cyc, threads);
        run<K_MIX_LOP_VIADD_1_1>(out, cyc, threads);
        run<K_MULW>(out, cyc, threads);
        run<K_MULHILO>(out, cyc, threads);
        run<K_MIX_LOP_MULW_4_1>(out, cyc, threads);
        run<K_MIX_LOP_MULHILO_4_1>(out, cyc, threads);
        run<K_SETP_PLOP>(out, cyc, threads);
        run<K_POPC>(out, cyc, threads);
        run<K_VOTE>(out, cyc, threads);
        run<K_LDS128>(out, cyc, threads);
        run<K_LDS128_SEQ>(out, cyc, threads);
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
