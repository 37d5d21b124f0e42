// Instruction-issue microbenchmark for sm_100a: measures warp-instructions per clock per SM
// for the integer ops the exact per-tap-truncating MAC needs (IMAD, IMAD.WIDE, IMAD.HI, LOP3,
// PRMT, SHF, IADD3) alone and in the mixes the resample kernels use.  Test infrastructure only.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

constexpr int ITER = 4096;
constexpr int CH = 8; // independent chains per thread

template <int OP>
__global__ void __launch_bounds__(256) k(int *out, int b0, int c0)
{
    int a[CH]; long long w[CH];
    int b = b0 + threadIdx.x, c = c0;
#pragma unroll
    for (int i = 0; i < CH; ++i) { a[i] = threadIdx.x * (i + 1) + b0; w[i] = (long long)a[i] * 77; }
    for (int it = 0; it < ITER / 8; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
#pragma unroll
            for (int i = 0; i < CH; ++i) {
                if (OP == 0) asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c));
                const int x = (int)w[i]; (void)x;
                if (OP == 1) asm volatile("mad.wide.s32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(x), "r"(b));
                if (OP == 2) asm volatile("mad.hi.s32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c));
                if (OP == 3) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b), "r"(c));
                if (OP == 4) asm volatile("prmt.b32 %0, %0, %1, 0xBB32;" : "+r"(a[i]) : "r"(b));
                if (OP == 5) asm volatile("shr.s32 %0, %0, 1;" : "+r"(a[i]));
                if (OP == 6) asm volatile("add.s32 %0, %0, %1;" : "+r"(a[i]) : "r"(b));
                if (OP == 7) { // MAC recipe A: bias (lop3) into low word + IMAD.HI with 64-bit addend (2 instr)
                    const int xm = a[(i + 1) % CH];
                    unsigned lo; asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(lo) : "r"(xm), "r"(b), "r"(c));
                    const long long cc = ((long long)a[i] << 32) | lo;
                    a[i] = (int)(((long long)xm * b + cc) >> 32);
                }
                if (OP == 8) { // MAC recipe B: shl + prmt + lop3 + IMAD.HI (4 instr)
                    const int xm = a[(i + 1) % CH];
                    unsigned lo, s2, m; 
                    asm volatile("shl.b32 %0, %1, 16;" : "=r"(s2) : "r"(xm));
                    asm volatile("prmt.b32 %0, %1, %1, 0x9999;" : "=r"(m) : "r"(xm));
                    asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(lo) : "r"(m), "r"(b), "r"(c));
                    const long long cc = ((long long)a[i] << 32) | lo;
                    a[i] = (int)(((long long)(int)s2 * b + cc) >> 32);
                }
                if (OP == 16) { // MAC recipe C (static-sign chain): shl + prmt(bias) + IMAD.HI (3 instr)
                    const int xm = a[(i + 1) % CH];
                    unsigned s2, m; 
                    asm volatile("shl.b32 %0, %1, 16;" : "=r"(s2) : "r"(xm));
                    asm volatile("prmt.b32 %0, %1, %1, 0x9999;" : "=r"(m) : "r"(xm));
                    const long long cc = ((long long)a[i] << 32) | m;
                    a[i] = (int)(((long long)(int)s2 * b + cc) >> 32);
                }
                if (OP == 17) { // IMAD.HI with full 64-bit addend alone
                    const int xm = a[(i + 1) % CH];
                    const long long cc = ((long long)a[i] << 32) | (unsigned)c;
                    a[i] = (int)(((long long)xm * b + cc) >> 32);
                }
                if (OP == 9) asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(b));
                if (OP == 10) { // imad.hi + lop3 + iadd (sign-magnitude recipe)
                    int m; asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(m) : "r"(x), "r"(b));
                    asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(m) : "r"(b), "r"(c));
                    long long t = w[i]; int lo2 = (int)t; lo2 += m; w[i] = (t & 0xffffffff00000000ll) | (unsigned)lo2;
                }
                if (OP == 11) { // imad + imad.wide alternating (two fma-pipe ops)
                    asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c));
                    asm volatile("mad.wide.s32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(x), "r"(b));
                }
                if (OP == 12) { // lop3 + shf alternating (two alu-pipe ops)
                    asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b), "r"(c));
                    asm volatile("shr.s32 %0, %0, 1;" : "+r"(a[i]));
                }
                if (OP == 13) { // imad + lop3 alternating (one per pipe)
                    asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c));
                    int t = (int)w[i]; asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(t) : "r"(b), "r"(c)); w[i] = t;
                }
                if (OP == 14) asm volatile("min.s32 %0, %0, %1;" : "+r"(a[i]) : "r"(b));
                if (OP == 18) asm volatile("cvt.pack.sat.s16.s32 %0, %0, %1;" : "+r"(a[i]) : "r"(b));
                if (OP == 19) { unsigned v = __vmaxs2((unsigned)a[i], (unsigned)b); a[i] = (int)v; }
                if (OP == 20) asm volatile("shf.r.wrap.b32 %0, %0, %1, 3;" : "+r"(a[i]) : "r"(b));
                if (OP == 15) { float f = __int_as_float(a[i]); asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f) : "f"(__int_as_float(b)), "f"(__int_as_float(c))); a[i] = __float_as_int(f); }
            }
        }
    }
    int r = 0;
#pragma unroll
    for (int i = 0; i < CH; ++i) r ^= a[i] ^ (int)w[i] ^ (int)(w[i] >> 32);
    if (r == 0x12345678) out[0] = r;
}


// MAC recipes fed from shared memory the way the resample kernels are: one LDS.128 brings 4 packed
// words; every word feeds REC-specific prep + one IMAD.HI with a 64-bit addend (acc:bias).
//  REC 0: IMAD.HI only (multiplicand = loaded word, bias = loaded word)         1 instr / MAC
//  REC 1: PRMT bias + IMAD.HI                                                    2 instr / MAC
//  REC 2: SHL + PRMT bias + IMAD.HI              (static-sign chain)             3 instr / MAC
//  REC 3: SHL + PRMT + LOP3 bias + IMAD.HI       (generic signed weight)         4 instr / MAC
//  REC 4: stereo packed word: SHL, LOP(and), 2x PRMT, 2x IMAD.HI                 6 instr / 2 MAC
template <int REC>
__global__ void __launch_bounds__(256) mac(int *out, int b0, int c0)
{
    __shared__ int4 buf[2048];      /* 32 KB: the two loads below reach 4080 + 7 * 1024 + 8192 + 16 bytes */
    for (int i = threadIdx.x; i < 2048; i += 256) buf[i] = make_int4(i * 2654435761u, i * 40503u, ~i * 977u, i * 31337u);
    __syncthreads();
    int acc[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) acc[i] = i;
    const int b = b0 + (threadIdx.x & 3), c = c0;
    unsigned addr = (unsigned)__cvta_generic_to_shared(buf) + threadIdx.x * 16;
    for (int it = 0; it < ITER / 8; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            int v[8];
            asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]) : "r"(addr + u * 1024));
            asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(addr + u * 1024 + 8192));
#pragma unroll
            for (int i = 0; i < CH; ++i) {
                const int x = v[i];
                if (REC == 0) { const long long cc = ((long long)acc[i] << 32) | (unsigned)x; acc[i] = (int)(((long long)x * b + cc) >> 32); }
                if (REC == 1) { unsigned m; asm volatile("prmt.b32 %0, %1, %1, 0x9999;" : "=r"(m) : "r"(x));
                                const long long cc = ((long long)acc[i] << 32) | m; acc[i] = (int)(((long long)x * b + cc) >> 32); }
                if (REC == 2) { unsigned m, s2; asm volatile("prmt.b32 %0, %1, %1, 0x9999;" : "=r"(m) : "r"(x)); asm volatile("shl.b32 %0, %1, 16;" : "=r"(s2) : "r"(x));
                                const long long cc = ((long long)acc[i] << 32) | m; acc[i] = (int)(((long long)(int)s2 * b + cc) >> 32); }
                if (REC == 3) { unsigned m, s2, lo; asm volatile("prmt.b32 %0, %1, %1, 0x9999;" : "=r"(m) : "r"(x)); asm volatile("shl.b32 %0, %1, 16;" : "=r"(s2) : "r"(x));
                                asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(lo) : "r"(m), "r"(b), "r"(c));
                                const long long cc = ((long long)acc[i] << 32) | lo; acc[i] = (int)(((long long)(int)s2 * b + cc) >> 32); }
                if (REC == 4 && (i & 1) == 0) {
                    unsigned ml, mr, sl, sr;
                    asm volatile("prmt.b32 %0, %1, %1, 0x9999;" : "=r"(ml) : "r"(x)); asm volatile("prmt.b32 %0, %1, %1, 0xBBBB;" : "=r"(mr) : "r"(x));
                    asm volatile("shl.b32 %0, %1, 16;" : "=r"(sl) : "r"(x)); asm volatile("and.b32 %0, %1, 0xFFFF0000;" : "=r"(sr) : "r"(x));
                    const long long c0l = ((long long)acc[i] << 32) | ml; acc[i] = (int)(((long long)(int)sl * b + c0l) >> 32);
                    const long long c1l = ((long long)acc[i + 1] << 32) | mr; acc[i + 1] = (int)(((long long)(int)sr * b + c1l) >> 32);
                }
            }
        }
        addr ^= (acc[0] & 16);
    }
    int r = 0;
#pragma unroll
    for (int i = 0; i < CH; ++i) r ^= acc[i];
    if (r == 0x12345678) out[0] = r;
}

template <int REC> int runmac(const char *name, double instr_per_mac, double macs_per_slot, int *d, int sms, double mhz)
{
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int blocks = sms * 8;
    mac<REC><<<blocks, 256>>>(d, 3, 5);
    CHECK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(e0); mac<REC><<<blocks, 256>>>(d, 3, 5); cudaEventRecord(e1); CHECK(cudaEventSynchronize(e1));
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    double wmacs = (double)blocks * 8 * ITER * CH * macs_per_slot;
    double per_clk_sm = wmacs / (best * 1e-3 * mhz * 1e6) / sms;
    printf("{\"op\": \"%s\", \"ms\": %.4f, \"warp_macs_per_clk_per_sm\": %.3f, \"lane_macs_per_clk_per_sm\": %.1f, \"alu_instr_per_mac\": %.2f, \"tmacs_per_s_at_nominal\": %.2f}\n",
           name, best, per_clk_sm, per_clk_sm * 32, instr_per_mac, per_clk_sm * 32 * sms * mhz * 1e6 / 1e12);
    return 0;
}

// shared-memory load throughput: LDS.32 / LDS.64 / LDS.128, conflict-free consecutive lanes
template <int W>
__global__ void __launch_bounds__(256) lds(int *out, int stride)
{
    __shared__ int4 buf[2048];
    for (int i = threadIdx.x; i < 2048; i += 256) buf[i] = make_int4(i, i, i, i);
    __syncthreads();
    int acc = 0; int idx = threadIdx.x * stride;
    const char *base = (const char *)buf;
    for (int it = 0; it < ITER / 8; ++it) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            unsigned addr = (unsigned)__cvta_generic_to_shared(base) + (((idx + u * 37) * (W * 4)) & 32767 & ~(W * 4 - 1));
            if (W == 1) { int v; asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr)); acc ^= v; }
            if (W == 2) { int v, v2; asm volatile("ld.shared.v2.b32 {%0,%1}, [%2];" : "=r"(v), "=r"(v2) : "r"(addr)); acc ^= v ^ v2; }
            if (W == 4) { int v, v2, v3, v4; asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v), "=r"(v2), "=r"(v3), "=r"(v4) : "r"(addr)); acc ^= v ^ v2 ^ v3 ^ v4; }
        }
        idx += acc & 1;
    }
    if (acc == 0x12345678) out[0] = acc;
}

template <int OP> int run(const char *name, int instr_per_iter, int *d, int sms, double mhz)
{
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int blocks = sms * 8;
    k<OP><<<blocks, 256>>>(d, 3, 5);
    CHECK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(e0); k<OP><<<blocks, 256>>>(d, 3, 5); cudaEventRecord(e1); CHECK(cudaEventSynchronize(e1));
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    double winstr = (double)blocks * 8 /*warps*/ * ITER * CH * instr_per_iter;
    double per_clk_sm = winstr / (best * 1e-3 * mhz * 1e6) / sms;
    printf("{\"op\": \"%s\", \"ms\": %.4f, \"warp_instr_per_clk_per_sm\": %.3f, \"units_per_clk_per_sm\": %.3f}\n", name, best, per_clk_sm, per_clk_sm / instr_per_iter);
    return 0;
}

template <int W> int runlds(const char *name, int stride, int *d, int sms, double mhz)
{
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int blocks = sms * 8;
    lds<W><<<blocks, 256>>>(d, stride);
    CHECK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(e0); lds<W><<<blocks, 256>>>(d, stride); cudaEventRecord(e1); CHECK(cudaEventSynchronize(e1));
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    double winstr = (double)blocks * 8 * (ITER / 8) * 16;
    double per_clk_sm = winstr / (best * 1e-3 * mhz * 1e6) / sms;
    printf("{\"op\": \"%s\", \"ms\": %.4f, \"warp_instr_per_clk_per_sm\": %.3f, \"bytes_per_clk_per_sm\": %.1f}\n", name, best, per_clk_sm, per_clk_sm * 32 * W * 4);
    return 0;
}

int main()
{
    cudaDeviceProp p; CHECK(cudaGetDeviceProperties(&p, 0));
    int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    double mhz = clk_khz / 1000.0;
    printf("{\"device\": \"%s\", \"sms\": %d, \"clock_mhz_nominal\": %.0f, \"note\": \"rates assume the nominal max clock; real clock may be lower\"}\n", p.name, p.multiProcessorCount, mhz);
    int *d; CHECK(cudaMalloc(&d, 4));
    int sms = p.multiProcessorCount;
    run<0>("imad", 1, d, sms, mhz);
    run<1>("imad.wide", 1, d, sms, mhz);
    run<2>("imad.hi.s32(mad)", 1, d, sms, mhz);
    run<9>("imad.hi.u32(mul)", 1, d, sms, mhz);
    run<3>("lop3", 1, d, sms, mhz);
    run<4>("prmt", 1, d, sms, mhz);
    run<5>("shf", 1, d, sms, mhz);
    run<6>("iadd", 1, d, sms, mhz);
    run<14>("imnmx", 1, d, sms, mhz);
    run<15>("ffma", 1, d, sms, mhz);
    run<18>("i2ip.s16.s32.sat (cvt.pack)", 1, d, sms, mhz);
    run<19>("vimnmx.s16x2", 1, d, sms, mhz);
    run<20>("shf (funnel)", 1, d, sms, mhz);
    run<11>("imad+imad.wide", 2, d, sms, mhz);
    run<12>("lop3+shf", 2, d, sms, mhz);
    run<13>("imad+lop3", 2, d, sms, mhz);
    run<7>("mac2: lop3+imad.hi64", 2, d, sms, mhz);
    run<8>("mac4: shl+prmt+lop3+imad.hi64", 4, d, sms, mhz);
    run<16>("mac3s: shl+prmt+imad.hi64", 3, d, sms, mhz);
    run<17>("imad.hi64 (64-bit addend)", 1, d, sms, mhz);
    run<10>("mac3: umulhi+lop3+iadd", 3, d, sms, mhz);
    runmac<0>("smem-fed imad.hi64", 1.25, 1, d, sms, mhz);
    runmac<1>("smem-fed prmt+imad.hi64", 2.25, 1, d, sms, mhz);
    runmac<2>("smem-fed shl+prmt+imad.hi64 (static-sign MAC)", 3.25, 1, d, sms, mhz);
    runmac<3>("smem-fed shl+prmt+lop3+imad.hi64 (generic MAC)", 4.25, 1, d, sms, mhz);
    runmac<4>("smem-fed stereo word: 4 prep + 2 imad.hi64", 3.25, 1, d, sms, mhz);
    runlds<1>("lds.32 stride1", 1, d, sms, mhz);
    runlds<2>("lds.64 stride1", 1, d, sms, mhz);
    runlds<4>("lds.128 stride1", 1, d, sms, mhz);
    runlds<1>("lds.32 random", 7919, d, sms, mhz);
    return 0;
}
