// Which integer instructions overlap on sm_100a?  Each thread runs NA independent chains of op A and NB independent
// chains of op B (no data flow between the two sets).  If A and B use different pipes, time(A+B) ~ max(time(A), time(B));
// if they share one, time(A+B) ~ time(A) + time(B).  Test infrastructure only.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
constexpr int ITER = 2048;
constexpr int CH = 6;
// ops: 0 none, 1 IMAD.HI with 64-bit addend, 2 LOP3, 3 IMAD (lo), 4 PRMT, 5 FFMA, 6 LDS.32 (conflict free), 7 IADD3, 8 SHF, 9 IMAD.HI 32-bit addend, 10 IMAD.WIDE
template <int OP> __device__ __forceinline__ void step(int &a, long long &w, float &f, int b, int c, unsigned saddr)
{
    if (OP == 1) { long long cc = ((long long)a << 32) | (unsigned)c; asm volatile("" : "+l"(cc)); a = (int)(((long long)b * (long long)c + cc) >> 32); }
    if (OP == 2) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a) : "r"(b), "r"(c));
    if (OP == 3) asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(a) : "r"(b), "r"(c));
    if (OP == 4) asm volatile("prmt.b32 %0, %0, %1, 0x3210;" : "+r"(a) : "r"(b));
    if (OP == 5) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f) : "f"(1.0001f), "f"(0.5f));
    if (OP == 6) { int v; asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(saddr + ((unsigned)a & 0xF80u))); a ^= v; }
    if (OP == 7) asm volatile("add.s32 %0, %0, %1;" : "+r"(a) : "r"(b));
    if (OP == 8) asm volatile("shf.r.wrap.b32 %0, %0, %1, 3;" : "+r"(a) : "r"(b));
    if (OP == 9) asm volatile("mad.hi.s32 %0, %1, %2, %0;" : "+r"(a) : "r"(b), "r"(c));
    if (OP == 10) asm volatile("mad.wide.s32 %0, %1, %2, %0;" : "+l"(w) : "r"(b), "r"(c));
    if (OP == 11) { int v; asm volatile("ld.shared.s16 %0, [%1];" : "=r"(v) : "r"(saddr + ((unsigned)a & 0xF80u))); a ^= v; }
}
template <int A, int NA, int B, int NB>
__global__ void __launch_bounds__(512) k(int *out, int b0, int c0)
{
    __shared__ int buf[2048];
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) buf[i] = 0;
    __syncthreads();
    const unsigned saddr = (unsigned)__cvta_generic_to_shared(buf) + (threadIdx.x & 31) * 4;
    int a[CH], a2[CH]; long long w[CH], w2[CH]; float f[CH], f2[CH];
    int b = b0 + threadIdx.x, c = c0;
#pragma unroll
    for (int i = 0; i < CH; ++i) { a[i] = threadIdx.x * (i + 1) + b0; a2[i] = a[i] * 3; w[i] = a[i]; w2[i] = a2[i]; f[i] = (float)i; f2[i] = (float)i + 0.5f; }
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int i = 0; i < CH; ++i) {
            if (i < NA) step<A>(a[i], w[i], f[i], b, c, saddr);
            if (i < NB) step<B>(a2[i], w2[i], f2[i], b, c, saddr);
        }
    }
    int r = 0;
#pragma unroll
    for (int i = 0; i < CH; ++i) r ^= a[i] ^ a2[i] ^ (int)w[i] ^ (int)w2[i] ^ __float_as_int(f[i]) ^ __float_as_int(f2[i]);
    if (r == 0x12345678) out[0] = r;
}
template <int A, int NA, int B, int NB> void run(const char *name, int *d, int sms, double mhz)
{
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int blocks = sms * 2;
    k<A, NA, B, NB><<<blocks, 512>>>(d, 3, 5);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(e0); k<A, NA, B, NB><<<blocks, 512>>>(d, 3, 5); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    // clocks per SMSP per loop iteration per warp: 32 warps per SM / 4 = 8 warps per SMSP
    const double clk = best * 1e-3 * mhz * 1e6 / ITER / 8.0;
    printf("{\"mix\": \"%s\", \"ms\": %.4f, \"smsp_clk_per_warp_iter\": %.2f, \"A_per_iter\": %d, \"B_per_iter\": %d, \"error\": \"%s\"}\n", name, best, clk, NA, NB, cudaGetErrorString(cudaGetLastError()));
}
int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    const double mhz = clk_khz / 1000.0; const int sms = p.multiProcessorCount;
    int *d; cudaMalloc(&d, 4);
    printf("{\"device\": \"%s\", \"sms\": %d, \"clock_mhz_nominal\": %.0f, \"layout\": \"2 CTAs x 512 threads per SM, 6 chains per op per thread\"}\n", p.name, sms, mhz);
    run<1, 6, 0, 0>("6 imad.hi64", d, sms, mhz);
    run<9, 6, 0, 0>("6 imad.hi32", d, sms, mhz);
    run<10, 6, 0, 0>("6 imad.wide", d, sms, mhz);
    run<2, 6, 0, 0>("6 lop3", d, sms, mhz);
    run<3, 6, 0, 0>("6 imad", d, sms, mhz);
    run<5, 6, 0, 0>("6 ffma", d, sms, mhz);
    run<6, 6, 0, 0>("6 lds.32", d, sms, mhz);
    run<11, 6, 0, 0>("6 lds.s16", d, sms, mhz);
    run<1, 6, 2, 6>("6 imad.hi64 + 6 lop3", d, sms, mhz);
    run<1, 6, 4, 6>("6 imad.hi64 + 6 prmt", d, sms, mhz);
    run<1, 6, 3, 6>("6 imad.hi64 + 6 imad", d, sms, mhz);
    run<1, 6, 5, 6>("6 imad.hi64 + 6 ffma", d, sms, mhz);
    run<1, 6, 7, 6>("6 imad.hi64 + 6 iadd", d, sms, mhz);
    run<1, 6, 8, 6>("6 imad.hi64 + 6 shf", d, sms, mhz);
    run<1, 6, 6, 6>("6 imad.hi64 + 6 lds.32", d, sms, mhz);
    run<1, 6, 11, 6>("6 imad.hi64 + 6 lds.s16", d, sms, mhz);
    run<9, 6, 2, 6>("6 imad.hi32 + 6 lop3", d, sms, mhz);
    run<2, 6, 3, 6>("6 lop3 + 6 imad", d, sms, mhz);
    run<2, 6, 5, 6>("6 lop3 + 6 ffma", d, sms, mhz);
    run<2, 6, 6, 6>("6 lop3 + 6 lds.32", d, sms, mhz);
    run<1, 6, 2, 3>("6 imad.hi64 + 3 lop3", d, sms, mhz);
    run<1, 3, 2, 6>("3 imad.hi64 + 6 lop3", d, sms, mhz);
    return 0;
}
