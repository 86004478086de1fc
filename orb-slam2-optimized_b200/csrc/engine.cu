// engine.cu -- C ABI (include/ransac_b200.h) and host orchestration of the batched RANSAC
// engine.  One engine = one device + one stream + grow-only device buffers.  No CPU
// fallback: every compute entry point needs a CUDA device.
#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <tuple>
#include <vector>

#include "../../include/ransac_b200.h"
#include "common.cuh"
#include "engine_state.cuh"
#include "rng.cuh"

using namespace rsac;

// ------------------------------------------------------------------ helpers
int rsac_version(void) { return RSAC_VERSION; }

int rsac_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int rsac_create(int device, rsac_engine** out)
{
    if (!out) return RSAC_ERR_INVALID;
    *out = nullptr;
    int n = rsac_device_count();
    if (n <= 0 || device < 0 || device >= n) return RSAC_ERR_NO_DEVICE;
    if (cudaSetDevice(device) != cudaSuccess) { cudaGetLastError(); return RSAC_ERR_NO_DEVICE; }
    rsac_engine* e = new rsac_engine();
    e->device = device;
    { const char* g = getenv("RSAC_GRAPH"); if (g && *g == '0') e->graphs = false; }
    if (cudaStreamCreateWithFlags(&e->own_stream, cudaStreamNonBlocking) != cudaSuccess) { delete e; return RSAC_ERR_CUDA; }
    e->stream = e->own_stream;
    if (cudaStreamCreateWithFlags(&e->aux_stream, cudaStreamNonBlocking) != cudaSuccess) { cudaGetLastError(); e->aux_stream = nullptr; }
    cudaEventCreateWithFlags(&e->ev_fork, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&e->ev_join, cudaEventDisableTiming);
    cudaEventCreate(&e->t0);
    cudaEventCreate(&e->t1);
    cudaDeviceGetAttribute(&e->sm_count, cudaDevAttrMultiProcessorCount, device);
    *out = e;
    return RSAC_OK;
}

void rsac_destroy(rsac_engine* e)
{
    if (!e) return;
    cudaSetDevice(e->device);
    cudaStreamSynchronize(e->stream);
    rsac_nccl_destroy(e);          // an engine destroyed with a live communicator does not leak it
    e->free_all();
    for (auto& p : e->prof_events) { cudaEventDestroy(p.a); cudaEventDestroy(p.b); }
    for (auto& p : e->prof_pool) { cudaEventDestroy(p.a); cudaEventDestroy(p.b); }
    cudaEventDestroy(e->t0);
    cudaEventDestroy(e->t1);
    if (e->aux_stream) { cudaStreamSynchronize(e->aux_stream); cudaStreamDestroy(e->aux_stream); }
    cudaEventDestroy(e->ev_fork);
    cudaEventDestroy(e->ev_join);
    cudaStreamDestroy(e->own_stream);
    delete e;
}

const char* rsac_last_error(rsac_engine* e) { return e ? e->err.c_str() : "null engine"; }

int rsac_set_stream(rsac_engine* e, void* s)
{
    if (!e) return RSAC_ERR_INVALID;
    e->stream = s ? (cudaStream_t)s : e->own_stream;
    return RSAC_OK;
}

int rsac_set_problem_base(rsac_engine* e, int base)
{
    if (!e) return RSAC_ERR_INVALID;
    e->problem_base = base;
    return RSAC_OK;
}

int rsac_set_problem_ids(rsac_engine* e, const int32_t* ids, int C)
{
    if (!e || C < 0 || (C > 0 && !ids)) return RSAC_ERR_INVALID;
    e->n_problem_ids = 0;
    if (C == 0) return RSAC_OK;
    RSAC_CUDA(e, cudaSetDevice(e->device));
    RSAC_TRY(e->d_problem_ids.ensure(e, sizeof(int32_t) * (size_t)C));
    RSAC_CUDA(e, cudaMemcpyAsync(e->d_problem_ids.p, ids, sizeof(int32_t) * (size_t)C, cudaMemcpyHostToDevice, e->stream));
    RSAC_CUDA(e, cudaStreamSynchronize(e->stream));       // ids may live on the caller's stack
    e->n_problem_ids = C;
    return RSAC_OK;
}

int rsac_set_graphs(rsac_engine* e, int on)
{
    if (!e) return RSAC_ERR_INVALID;
    e->graphs = on != 0;
    return RSAC_OK;
}

int rsac_set_first_phase(rsac_engine* e, int hypotheses)
{
    if (!e || hypotheses < 0) return RSAC_ERR_INVALID;
    e->first_phase = hypotheses;
    e->second_phase = 0;
    e->stage_bounds.clear();
    return RSAC_OK;
}

int rsac_set_phases(rsac_engine* e, int first, int second)
{
    if (!e || first < 0 || second < 0 || (second > 0 && first > 0 && second <= first)) return RSAC_ERR_INVALID;
    e->first_phase = first;
    e->second_phase = second;
    e->stage_bounds.clear();
    return RSAC_OK;
}

int rsac_set_stages(rsac_engine* e, int n, const int32_t* bounds)
{
    if (!e || n < 0 || n > 7 || (n > 0 && !bounds)) return RSAC_ERR_INVALID;
    for (int i = 0; i < n; ++i)
        if (bounds[i] <= 0 || (i > 0 && bounds[i] <= bounds[i - 1])) return RSAC_ERR_INVALID;
    e->stage_bounds.assign(bounds, bounds + n);
    return RSAC_OK;
}

int rsac_sync(rsac_engine* e)
{
    if (!e) return RSAC_ERR_INVALID;
    RSAC_CUDA(e, cudaStreamSynchronize(e->stream));
    return RSAC_OK;
}

int rsac_host_alloc(void** ptr, uint64_t bytes)
{
    if (!ptr) return RSAC_ERR_INVALID;
    if (cudaHostAlloc(ptr, bytes, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); return RSAC_ERR_ALLOC; }
    return RSAC_OK;
}

int rsac_host_free(void* ptr) { return cudaFreeHost(ptr) == cudaSuccess ? RSAC_OK : RSAC_ERR_CUDA; }

int rsac_get_device_info(rsac_engine* e, rsac_device_info* info)
{
    if (!e || !info) return RSAC_ERR_INVALID;
    cudaDeviceProp p;
    RSAC_CUDA(e, cudaGetDeviceProperties(&p, e->device));
    memset(info, 0, sizeof(*info));
    info->sm_count = p.multiProcessorCount;
    cudaDeviceGetAttribute(&info->sm_clock_khz, cudaDevAttrClockRate, e->device);
    cudaDeviceGetAttribute(&info->mem_clock_khz, cudaDevAttrMemoryClockRate, e->device);
    info->cc_major = p.major;
    info->cc_minor = p.minor;
    info->total_mem = p.totalGlobalMem;
    strncpy(info->name, p.name, sizeof(info->name) - 1);
    return RSAC_OK;
}

int rsac_timer_begin(rsac_engine* e)
{
    if (!e) return RSAC_ERR_INVALID;
    RSAC_CUDA(e, cudaEventRecord(e->t0, e->stream));
    return RSAC_OK;
}

int rsac_timer_end(rsac_engine* e, float* ms)
{
    if (!e || !ms) return RSAC_ERR_INVALID;
    RSAC_CUDA(e, cudaEventRecord(e->t1, e->stream));
    RSAC_CUDA(e, cudaEventSynchronize(e->t1));
    RSAC_CUDA(e, cudaEventElapsedTime(ms, e->t0, e->t1));
    return RSAC_OK;
}

int rsac_profile_enable(rsac_engine* e, int on)
{
    if (!e) return RSAC_ERR_INVALID;
    e->profile = on != 0;
    return RSAC_OK;
}

static int profile_drain(rsac_engine* e)
{
    if (e->prof_events.empty()) return RSAC_OK;
    RSAC_CUDA(e, cudaStreamSynchronize(e->stream));
    for (auto& p : e->prof_events) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, p.a, p.b) == cudaSuccess) {
            e->stage_ms[p.stage] += ms;
            e->stage_launches[p.stage] += 1;
            if (e->prof_trace.size() < 4096) e->prof_trace.push_back({p.stage, ms});
        } else {
            cudaGetLastError();
        }
        e->prof_pool.push_back(p);
    }
    e->prof_events.clear();
    return RSAC_OK;
}

int rsac_profile_reset(rsac_engine* e)
{
    if (!e) return RSAC_ERR_INVALID;
    int rc = profile_drain(e);
    for (int i = 0; i < RSAC_STAGE_COUNT; ++i) { e->stage_ms[i] = 0.0; e->stage_launches[i] = 0; }
    e->launches = 0;
    e->prof_trace.clear();
    return rc;
}

int rsac_profile_trace(rsac_engine* e, int max_entries, int32_t* stages, float* ms)
{
    if (!e || max_entries < 0) return -1;
    if (profile_drain(e) != RSAC_OK) return -1;
    const int n = (int)std::min<size_t>(e->prof_trace.size(), (size_t)max_entries);
    for (int i = 0; i < n; ++i) {
        if (stages) stages[i] = e->prof_trace[i].first;
        if (ms) ms[i] = e->prof_trace[i].second;
    }
    return n;
}

int rsac_profile_get(rsac_engine* e, int stage, double* total_ms, int64_t* launches)
{
    if (!e || stage < 0 || stage >= RSAC_STAGE_COUNT) return RSAC_ERR_INVALID;
    int rc = profile_drain(e);
    if (total_ms) *total_ms = e->stage_ms[stage];
    if (launches) *launches = e->stage_launches[stage];
    return rc;
}

int64_t rsac_launch_count(rsac_engine* e) { return e ? e->launches : 0; }

// ------------------------------------------------------------ host helpers
int rsac_pnp_ransac_setup(int N, const rsac_ransac_params* p, int* min_inl, int* max_its)
{
    if (!p) return RSAC_ERR_INVALID;
    // PnPsolver::SetRansacParameters (PnPsolver.cpp:58-94); same arithmetic in MLPnPsolver.cpp:185-220
    float eps = p->eps;
    int nMinInliers = (int)((float)N * eps);
    if (nMinInliers < p->min_inliers) nMinInliers = p->min_inliers;
    if (nMinInliers < p->min_set) nMinInliers = p->min_set;
    if (eps < (float)nMinInliers / N) eps = (float)nMinInliers / N;
    int nIterations;
    if (nMinInliers == N)
        nIterations = 1;
    else
        nIterations = (int)std::ceil(std::log(1 - p->prob) / std::log(1 - std::pow(eps, 3)));
    int its = std::min(nIterations, p->max_its);
    if (max_its) *max_its = std::max(1, its);
    if (min_inl) *min_inl = nMinInliers;
    return RSAC_OK;
}

int rsac_sim3_ransac_setup(int N, const rsac_sim3_params* p, int* max_its)
{
    if (!p) return RSAC_ERR_INVALID;
    // Sim3Solver::SetRansacParameters (Sim3Solver.cpp:87-111)
    const float epsilon = (float)p->min_inliers / N;
    int nIterations;
    if (p->min_inliers == N)
        nIterations = 1;
    else
        nIterations = (int)std::ceil(std::log(1 - p->prob) / std::log(1 - std::pow(epsilon, 3)));
    int its = std::min(nIterations, p->max_its);
    if (max_its) *max_its = std::max(1, its);
    return RSAC_OK;
}

int rsac_index_table(uint32_t seed, int n, int k, int H, uint32_t* out)
{
    if (!out || k < 1 || k > 8 || n < k || H < 0) return RSAC_ERR_INVALID;
    GlibcRand g;
    g.seed(seed);
    for (int h = 0; h < H; ++h) draw_minimal_set<8>(g, n, k, out + (size_t)h * k);
    return RSAC_OK;
}

int rsac_rand_stream(uint32_t seed, int count, int32_t* out)
{
    if (!out || count < 0) return RSAC_ERR_INVALID;
    GlibcRand g;
    g.seed(seed);
    for (int i = 0; i < count; ++i) out[i] = g.next();
    return RSAC_OK;
}

int rsac_shard_range(int C, int rank, int world, int* first, int* count)
{
    if (world < 1 || rank < 0 || rank >= world || C < 0) return RSAC_ERR_INVALID;
    const int per = (C + world - 1) / world;           // contiguous blocks of ceil(C/world)
    const int f = std::min(C, rank * per);
    const int l = std::min(C, f + per);
    if (first) *first = f;
    if (count) *count = l - f;
    return RSAC_OK;
}

// ------------------------------------------------------------ peak probes
template <typename T>
__global__ void fma_peak_kernel(T* out, int iters)
{
    T a0 = (T)threadIdx.x * (T)1e-3, a1 = a0 + (T)1, a2 = a0 + (T)2, a3 = a0 + (T)3;
    T a4 = a0 + (T)4, a5 = a0 + (T)5, a6 = a0 + (T)6, a7 = a0 + (T)7;
    const T b = (T)0.999, c = (T)1e-4;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
            a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

int rsac_measure_peaks(rsac_engine* e, double* fp32_tflops, double* fp64_tflops)
{
    if (!e) return RSAC_ERR_INVALID;
    RSAC_CUDA(e, cudaSetDevice(e->device));
    const int blocks = e->sm_count * 8, threads = 256;
    RSAC_TRY(e->d_scratch.ensure(e, sizeof(double) * (size_t)blocks * threads));
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    auto run = [&](bool dbl, int iters) -> double {
        float best = 1e30f;
        for (int rep = 0; rep < 4; ++rep) {
            cudaEventRecord(a, e->stream);
            if (dbl) fma_peak_kernel<double><<<blocks, threads, 0, e->stream>>>((double*)e->d_scratch.p, iters);
            else fma_peak_kernel<float><<<blocks, threads, 0, e->stream>>>((float*)e->d_scratch.p, iters);
            cudaEventRecord(b, e->stream);
            cudaEventSynchronize(b);
            float ms = 0.f;
            cudaEventElapsedTime(&ms, a, b);
            if (rep > 0) best = std::min(best, ms);
        }
        e->launches += 4;
        const double flops = 2.0 * 8 * 16 * (double)iters * blocks * threads;
        return flops / (best * 1e-3) / 1e12;
    };
    if (fp32_tflops) *fp32_tflops = run(false, 4096);
    if (fp64_tflops) *fp64_tflops = run(true, 1024);
    cudaEventDestroy(a); cudaEventDestroy(b);
    RSAC_CUDA(e, cudaGetLastError());
    return RSAC_OK;
}
