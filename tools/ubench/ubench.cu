// ubench.cu — instruction-throughput microbenchmarks behind the pair-kernel design.
// Each test runs N_ITER x 8 independent chains of one instruction per thread, 148 blocks of 1024 threads
// (8 warps per SMSP), and reports warp-instructions per clock per SM from the per-block clock64() span.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench ubench.cu && ./ubench
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

constexpr int N_ITER = 4096;

enum Op { FFMA, FMUL, FADD, I2F, MUFU_RCP, MUFU_RSQ, MUFU_EX2, SHFL, IADD, LOP, FSEL, FSETP_SEL, LDS128_BCAST, LDS128_8, LDS128_RAND, LDS64_RAND, LDS32_RAND,
          FFMA_I2F, FFMA_MUFU, FFMA_IADD, FFMA_SHFL, FFMA_LDS, F2I, IMAD, RED64, RED32F, FFMA2, FADD2, FFMA2_ALU, FFMA_ALU3, N_OPS };
const char* names[] = {"FFMA", "FMUL", "FADD", "I2FP.F32.S32", "MUFU.RCP", "MUFU.RSQ", "MUFU.EX2", "SHFL.BFLY", "IADD3", "LOP3", "FSEL", "FSETP+FSEL",
                       "LDS.128 one address", "LDS.128 8 addresses", "LDS.128 32 random", "LDS.64 32 random", "LDS.32 32 random",
                       "FFMA+I2FP (1:1)", "FFMA+MUFU.RCP (3:1)", "FFMA+IADD3 (1:1)", "FFMA+SHFL (3:1)", "FFMA+LDS.128 8addr (3:1)", "F2I.S32.F32", "IMAD",
                       "RED.ADD.U64 global (distinct)", "RED.ADD.F32 global (distinct)", "FFMA2 (fma.rn.f32x2), per instr", "FADD2 (add.rn.f32x2), per instr",
                       "FFMA2+IADD3+LOP3+FSEL (1:1:1:1)", "FFMA+IADD3+LOP3+FSEL (1:1:1:1)"};

template <int OP>
__global__ void __launch_bounds__(1024) k(float* out, long long* cyc, unsigned long long* red, int seed) {
    __shared__ float4 sm[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = make_float4(i * 1e-3f, 1.f, 2.f, 3.f);
    __syncthreads();
    float a[8], a2[8];
    int b[8];
#pragma unroll
    for (int u = 0; u < 8; u++) { a[u] = 1.0f + (threadIdx.x + u) * 1e-6f; a2[u] = a[u] + 0.5f; b[u] = threadIdx.x * 7 + u + seed; }
    const float m = 0.999f + seed * 1e-9f, c = 1e-3f, cc = 1e-3f + seed * 1e-9f;
    const int lane = threadIdx.x & 31;
    unsigned int r = (threadIdx.x * 2654435761u) >> 22;  // random 10-bit index
    const long long t0 = clock64();
    for (int it = 0; it < N_ITER; it++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
            if (OP == FFMA) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[u]) : "f"(m), "f"(cc));
            if (OP == FMUL) asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(a[u]) : "f"(m));
            if (OP == FADD) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[u]) : "f"(cc));
            if (OP == I2F) { asm volatile("cvt.rn.f32.s32 %0, %1;" : "=f"(a[u]) : "r"(b[u])); b[u] = __float_as_int(a[u]); }
            if (OP == F2I) { asm volatile("cvt.rni.s32.f32 %0, %1;" : "=r"(b[u]) : "f"(a[u])); a[u] = __int_as_float(b[u] | 0x3f800000); }
            if (OP == MUFU_RCP) { float t; asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(a[u])); a[u] = t; }
            if (OP == MUFU_RSQ) asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(a[u]));
            if (OP == MUFU_EX2) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[u]));
            if (OP == SHFL) a[u] = __shfl_xor_sync(0xffffffffu, a[u], 1 + (u & 3));
            if (OP == IADD) asm volatile("add.s32 %0, %0, %1;" : "+r"(b[u]) : "r"(b[(u + 1) & 7]));
            if (OP == IMAD) asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(b[u]) : "r"(seed + 3), "r"(u));
            if (OP == LOP) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(b[u]) : "r"(seed + u), "r"(it));
            if (OP == FSEL) a[u] = (b[u] > seed) ? a[u] : c;
            if (OP == FSETP_SEL) a[u] = (a[u] < m) ? a[u] + 0.f : c;
            if (OP == LDS128_BCAST) { float4 v = sm[(it + u) & 1023]; a[u] += v.x; }
            if (OP == LDS128_8) { float4 v = sm[((it + u) * 8 + (lane & 7)) & 1023]; a[u] += v.x; }
            if (OP == LDS128_RAND) { float4 v = sm[(r + it * 37 + u * 101) & 1023]; a[u] += v.x; }
            if (OP == LDS64_RAND) { float2 v = reinterpret_cast<float2*>(sm)[(r + it * 37 + u * 101) & 2047]; a[u] += v.x; }
            if (OP == LDS32_RAND) { float v = reinterpret_cast<float*>(sm)[(r + it * 37 + u * 101) & 4095]; a[u] += v; }
            if (OP == FFMA_I2F) { if (u & 1) a[u] = fmaf(a[u], m, c); else { asm volatile("cvt.rn.f32.s32 %0, %1;" : "=f"(a[u]) : "r"(b[u])); b[u] = __float_as_int(a[u]); } }
            if (OP == FFMA_MUFU) { if (u & 3) a[u] = fmaf(a[u], m, c); else asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(a[u])); }
            if (OP == FFMA_IADD) { if (u & 1) a[u] = fmaf(a[u], m, c); else asm volatile("add.s32 %0, %0, %1;" : "+r"(b[u]) : "r"(seed + u)); }
            if (OP == FFMA_SHFL) { if (u & 3) a[u] = fmaf(a[u], m, c); else a[u] = __shfl_xor_sync(0xffffffffu, a[u], 1 + (u >> 2)); }
            if (OP == FFMA_LDS) { if (u & 3) a[u] = fmaf(a[u], m, c); else { float4 v = sm[((it + u) * 8 + (lane & 7)) & 1023]; a[u] += v.x; } }
            if (OP == FFMA2) { float2 v = __ffma2_rn(make_float2(a[u], a2[u]), make_float2(m, m), make_float2(c, c)); a[u] = v.x; a2[u] = v.y; }
            if (OP == FADD2) { float2 v = __fadd2_rn(make_float2(a[u], a2[u]), make_float2(c, c)); a[u] = v.x; a2[u] = v.y; }
            if (OP == FFMA2_ALU) {
                if ((u & 3) == 0) { float2 v = __ffma2_rn(make_float2(a[u], a2[u]), make_float2(m, m), make_float2(c, c)); a[u] = v.x; a2[u] = v.y; }
                if ((u & 3) == 1) asm volatile("add.s32 %0, %0, %1;" : "+r"(b[u]) : "r"(seed + u));
                if ((u & 3) == 2) asm volatile("xor.b32 %0, %0, %1;" : "+r"(b[u]) : "r"(seed + u));
                if ((u & 3) == 3) a[u] = (b[u] > seed) ? a[u] : c;
            }
            if (OP == FFMA_ALU3) {
                if ((u & 3) == 0) a[u] = fmaf(a[u], m, c);
                if ((u & 3) == 1) asm volatile("add.s32 %0, %0, %1;" : "+r"(b[u]) : "r"(seed + u));
                if ((u & 3) == 2) asm volatile("xor.b32 %0, %0, %1;" : "+r"(b[u]) : "r"(seed + u));
                if ((u & 3) == 3) a[u] = (b[u] > seed) ? a[u] : c;
            }
            if (OP == RED64) atomicAdd(&red[(size_t)(blockIdx.x * blockDim.x + threadIdx.x) * 8 + u], (unsigned long long)b[u]);
            if (OP == RED32F) atomicAdd(reinterpret_cast<float*>(red) + (size_t)(blockIdx.x * blockDim.x + threadIdx.x) * 8 + u, a[u]);
        }
    }
    const long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int u = 0; u < 8; u++) s += a[u] + a2[u] + (float)b[u];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

static float g_ms = 0;
template <int OP>
void run(float* out, long long* cyc, unsigned long long* red, int nblk) {
    std::vector<long long> h(nblk);
    double best = 1e30;
    for (int rep = 0; rep < 3; rep++) {
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        k<OP><<<nblk, 1024>>>(out, cyc, red, rep);
        cudaEventRecord(e1);
        cudaDeviceSynchronize();
        float ms; cudaEventElapsedTime(&ms, e0, e1); g_ms = ms;
        cudaEventDestroy(e0); cudaEventDestroy(e1);
        cudaMemcpy(h.data(), cyc, nblk * sizeof(long long), cudaMemcpyDeviceToHost);
        double mx = 0;
        for (long long v : h) mx = v > mx ? v : mx;
        if (rep && mx < best) best = mx;
    }
    // per SM: 1 block x 32 warps x N_ITER x 8 instructions (mixed tests count every instruction)
    const double instr = 32.0 * N_ITER * 8;
    printf("%-34s %8.3f warp-instr/clk64/SM  (%.2f clk64 per warp-instr per SMSP)   kernel %.3f ms -> clock64 at %.0f MHz, %.3f warp-instr/ns/SM\n", names[OP], instr / best, best * 4 / instr, g_ms, best / (g_ms * 1e3), instr / (g_ms * 1e6));
    if (OP + 1 < N_OPS) run<(OP + 1 < N_OPS ? OP + 1 : OP)>(out, cyc, red, nblk);
}
template <> void run<N_OPS>(float*, long long*, unsigned long long*, int) {}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int nblk = p.multiProcessorCount;
    printf("%s, %d SMs, %d blocks x 1024 threads\n", p.name, p.multiProcessorCount, nblk);
    float* out; long long* cyc; unsigned long long* red;
    cudaMalloc(&out, (size_t)nblk * 1024 * 4);
    cudaMalloc(&cyc, nblk * 8);
    cudaMalloc(&red, (size_t)nblk * 1024 * 8 * 8);
    cudaMemset(red, 0, (size_t)nblk * 1024 * 8 * 8);
    run<0>(out, cyc, red, nblk);
    printf("cuda status: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
