// fp32_probe.cu -- developer micro-benchmarks behind DESIGN.md's scoring-kernel ceiling:
//   (1) FFMA and FFMA2 (packed f32x2) issue throughput, (2) the scoring inner loop (eval_pair) fed from
//   shared memory without TMA/barriers, for several warps-per-SM settings.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -I../orb-slam2-optimized_b200/csrc -o fp32_probe fp32_probe.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "score.cuh"

using namespace rsac;

__global__ void ffma_chain(float* out, int iters)
{
    float a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 1e-3f + i;
    const float m = 1.0001f, c = 1e-3f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = fmaf(a[i], m, c);
    }
    float s = 0;
    for (int i = 0; i < 8; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void ffma2_chain(float* out, int iters)
{
    float2 a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = make_float2(threadIdx.x * 1e-3f + i, threadIdx.x * 2e-3f + i);
    const float2 m = make_float2(1.0001f, 1.0002f), c = make_float2(1e-3f, 2e-3f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = __ffma2_rn(a[i], m, c);
    }
    float s = 0;
    for (int i = 0; i < 8; ++i) s += a[i].x + a[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// mixed: 15 FFMA2 + 2 FFMA + 2 FMNMX + 4 SHF per step, like eval_pair, register operands only
template <int HPL>
__global__ void eval_loop(const float4* __restrict__ pts, int words, int reps, uint32_t* out)
{
    extern __shared__ float4 sp[];
    for (int i = threadIdx.x; i < words * 64; i += blockDim.x) sp[i] = pts[i];
    __syncthreads();
    float2 C[HPL][12];
#pragma unroll
    for (int s = 0; s < HPL; ++s)
#pragma unroll
        for (int i = 0; i < 12; ++i) { const float v = 0.01f * (i + 1) + 1e-4f * threadIdx.x + s; C[s][i] = make_float2(v, v); }
    uint32_t acc = 0;
    for (int r = 0; r < reps; ++r)
        for (int w = 0; w < words; ++w) {
            uint32_t inl[HPL], cert[HPL];
#pragma unroll
            for (int s = 0; s < HPL; ++s) { inl[s] = 0; cert[s] = 0; }
            const float4* pp = sp + w * 64;
#pragma unroll
            for (int p = 15; p >= 0; --p) {
                const float4 q0 = pp[4 * p], q1 = pp[4 * p + 1], q2 = pp[4 * p + 2];
                const float2 bd = *reinterpret_cast<const float2*>(pp + 4 * p + 3);
#pragma unroll
                for (int s = 0; s < HPL; ++s) eval_pair(C[s], q0, q1, q2, bd, inl[s], cert[s]);
            }
#pragma unroll
            for (int s = 0; s < HPL; ++s) acc += __popc(inl[s] & cert[s]);
        }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <typename F>
static float time_ms(F f)
{
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    f();
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    f();
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    return ms;
}

int main()
{
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    const int sms = prop.multiProcessorCount;
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    float* out;
    cudaMalloc(&out, sizeof(float) * sms * 8 * 1024);
    const int iters = 4096;
    for (int warps : {4, 8, 16, 32}) {
        const int threads = 256, blocks = sms * warps * 32 / threads;
        float ms1 = time_ms([&] { ffma_chain<<<blocks, threads>>>(out, iters); });
        float ms2 = time_ms([&] { ffma2_chain<<<blocks, threads>>>(out, iters); });
        const double f1 = (double)blocks * threads * iters * 64 * 2, f2 = f1 * 2;
        printf("warps/SM %2d: FFMA %.2f TFLOP/s   FFMA2 %.2f TFLOP/s\n", warps, f1 / ms1 * 1e-9, f2 / ms2 * 1e-9);
    }
    // eval loop
    const int words = 8, reps = 400;
    std::vector<float4> h(words * 64);
    for (size_t i = 0; i < h.size(); ++i) h[i] = make_float4(0.1f * (i % 7), 0.2f * (i % 5), 1.f + 0.01f * (i % 11), 0.3f);
    float4* d;
    cudaMalloc(&d, sizeof(float4) * h.size());
    cudaMemcpy(d, h.data(), sizeof(float4) * h.size(), cudaMemcpyHostToDevice);
    uint32_t* o;
    cudaMalloc(&o, sizeof(uint32_t) * sms * 4 * 1024);
    for (int threads : {128, 256, 288, 512}) {
        for (int ctas : {1, 2}) {
            const int blocks = sms * ctas;
            float ms = time_ms([&] { eval_loop<2><<<blocks, threads, words * 1024>>>(d, words, reps, o); });
            const double evals = (double)blocks * threads * 2 * 32 * words * reps;
            printf("eval_loop<2> threads %3d ctas/SM %d: %.3f ms  %.3g evals/s  %.2f TFLOP/s (31 flop/eval)  %.2f clk/eval/lane\n",
                   threads, ctas, ms, evals / ms * 1e3, evals * 31 / ms * 1e-9,
                   ms * 1e-3 * khz * 1e3 * sms * 128 / evals);
        }
    }
    {
        const int threads = 256, blocks = sms * 2;
        float ms = time_ms([&] { eval_loop<1><<<blocks, threads, words * 1024>>>(d, words, reps, o); });
        const double evals = (double)blocks * threads * 1 * 32 * words * reps;
        printf("eval_loop<1> threads 256 ctas/SM 2: %.2f TFLOP/s\n", evals * 31 / ms * 1e-9);
        ms = time_ms([&] { eval_loop<3><<<blocks, threads, words * 1024>>>(d, words, reps, o); });
        const double evals3 = (double)blocks * threads * 3 * 32 * words * reps;
        printf("eval_loop<3> threads 256 ctas/SM 2: %.2f TFLOP/s\n", evals3 * 31 / ms * 1e-9);
    }
    printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
