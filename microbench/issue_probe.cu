// issue_probe.cu -- does a packed FFMA2 cost one or two issue slots, and do ALU ops (SHF/FMNMX) overlap
// with it?  Prints cycles per loop iteration per warp scheduler for several instruction mixes.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o issue_probe issue_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int MODE>
__global__ void mix(float* out, int iters, long long* cycles)
{
    float2 a[8];
    float b[8];
    uint32_t s[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        a[i] = make_float2(threadIdx.x * 1e-3f + i, threadIdx.x * 2e-3f + i);
        b[i] = threadIdx.x * 3e-3f + i;
        s[i] = threadIdx.x * 77u + i;
    }
    const float2 m = make_float2(1.0001f, 1.0002f), c = make_float2(1e-3f, 2e-3f);
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) { a[i] = __ffma2_rn(a[i], m, c); }                                        // 8 FFMA2
            if (MODE == 1) { a[i] = __ffma2_rn(a[i], m, c); s[i] = __funnelshift_l(s[(i + 1) & 7], s[i], 1); }   // 8 FFMA2 + 8 SHF
            if (MODE == 2) { a[i] = __ffma2_rn(a[i], m, c); b[i] = fmaf(b[i], m.x, c.x); }            // 8 FFMA2 + 8 FFMA
            if (MODE == 3) { a[i].x = fmaf(a[i].x, m.x, c.x); a[i].y = fmaf(a[i].y, m.y, c.y); s[i] = __funnelshift_l(s[(i + 1) & 7], s[i], 1); }  // 16 FFMA + 8 SHF
            if (MODE == 4) { a[i].x = fmaf(a[i].x, m.x, c.x); a[i].y = fmaf(a[i].y, m.y, c.y); }     // 16 FFMA
            if (MODE == 5) { s[i] = __funnelshift_l(s[(i + 1) & 7], s[i], 1); }                      // 8 SHF
            if (MODE == 6) { a[i] = __ffma2_rn(a[i], m, c); b[i] = fminf(b[i], fabsf(a[(i + 3) & 7].x)); }   // 8 FFMA2 + 8 FMNMX
            if (MODE == 7) { a[i] = __ffma2_rn(a[i], m, c); s[i] = __funnelshift_l(s[(i + 1) & 7], s[i], 1); b[i] = fminf(b[i], fabsf(a[(i + 3) & 7].x)); }   // 8 FFMA2 + 8 SHF + 8 FMNMX
        }
    }
    const long long t1 = clock64();
    float r = 0;
    for (int i = 0; i < 8; ++i) r += a[i].x + a[i].y + b[i] + (float)s[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

template <int MODE>
static void run(const char* name, int warps_per_sm, int sms)
{
    float* out;
    long long* cyc;
    cudaMalloc(&out, sizeof(float) * sms * 2048);
    cudaMalloc(&cyc, 8);
    const int iters = 2048;
    const int threads = warps_per_sm * 32 > 1024 ? 1024 : warps_per_sm * 32;
    mix<MODE><<<sms, threads>>>(out, iters, cyc);
    mix<MODE><<<sms, threads>>>(out, iters, cyc);
    long long h = 0;
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    // cycles per iteration per scheduler, normalised per warp on that scheduler
    const double per_iter = (double)h / iters;
    printf("%-28s warps/SM %2d: %.1f clk/iter/block  => %.2f clk per warp-iteration per scheduler\n", name, warps_per_sm, per_iter,
           per_iter / (warps_per_sm / 4.0));
    cudaFree(out); cudaFree(cyc);
}

int main()
{
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    const int sms = prop.multiProcessorCount;
    for (int w : {16, 32}) {
        run<0>("8 FFMA2", w, sms);
        run<4>("16 FFMA", w, sms);
        run<5>("8 SHF", w, sms);
        run<1>("8 FFMA2 + 8 SHF", w, sms);
        run<3>("16 FFMA + 8 SHF", w, sms);
        run<2>("8 FFMA2 + 8 FFMA", w, sms);
        run<6>("8 FFMA2 + 8 FMNMX", w, sms);
        run<7>("8 FFMA2 + 8 SHF + 8 FMNMX", w, sms);
    }
    printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
