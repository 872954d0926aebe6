/* tools/ubench.cu -- micro-benchmarks that decide the kernel design (measurement aid, not part of the product):
 *   1. issue rate of scalar FFMA / FADD against the packed FFMA2 / FADD2 (fma.rn.f32x2, add.rn.f32x2) of sm_100a
 *   2. distributed-shared-memory bandwidth inside a thread-block cluster (remote loads and remote stores)
 * build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench.bin ubench.cu ; run on a B200. */
#include <cuda_runtime.h>
#include <cooperative_groups.h>
#include <cstdio>
#include <cstdlib>
namespace cg = cooperative_groups;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ float2 add2(float2 a, float2 b)
{
    float2 r;
    asm("{ .reg .b64 ra, rb, rc; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; add.rn.f32x2 rc, ra, rb; mov.b64 {%0,%1}, rc; }"
        : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return r;
}
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c)
{
    float2 r;
    asm("{ .reg .b64 ra, rb, rc, rd; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; mov.b64 rc, {%6,%7}; fma.rn.f32x2 rd, ra, rb, rc; mov.b64 {%0,%1}, rd; }"
        : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
    return r;
}

/* MODE 0: scalar FFMA, 1: scalar FADD, 2: FFMA2, 3: FADD2, 4: FFMA2 with swapped operand (complex-multiply form) */
template <int MODE, int CH>
__global__ void __launch_bounds__(256) k_fp(float2* out, int iters, float2 seed)
{
    float2 v[CH];
#pragma unroll
    for (int i = 0; i < CH; i++) v[i] = make_float2(threadIdx.x * 0.001f + i, i * 0.5f + seed.x);
    const float2 w = seed;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < CH; i++) {
            if (MODE == 0) { v[i].x = fmaf(v[i].x, w.x, w.y); v[i].y = fmaf(v[i].y, w.x, w.y); }
            if (MODE == 1) { v[i].x = v[i].x + w.x; v[i].y = v[i].y + w.y; }
            if (MODE == 2) v[i] = fma2(v[i], w, w);
            if (MODE == 3) v[i] = add2(v[i], w);
            if (MODE == 4) v[i] = fma2(make_float2(-v[i].y, v[i].x), make_float2(w.y, w.y), v[i]);
        }
    }
    float2 s = make_float2(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < CH; i++) { s.x += v[i].x; s.y += v[i].y; }
    if (s.x == 12345.678f) out[threadIdx.x] = s;
}

template <int MODE, int CH> static void run_fp(const char* name, float2* d_out, int ctas_per_sm, int nsm)
{
    const int iters = 4096;
    cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    k_fp<MODE, CH><<<nsm * ctas_per_sm, 256>>>(d_out, 16, make_float2(1.0001f, 0.0001f));
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(a));
    k_fp<MODE, CH><<<nsm * ctas_per_sm, 256>>>(d_out, iters, make_float2(1.0001f, 0.0001f));
    CK(cudaEventRecord(b)); CK(cudaDeviceSynchronize());
    float ms = 0; CK(cudaEventElapsedTime(&ms, a, b));
    /* "instructions" = machine instructions: scalar modes issue two per complex value, packed modes one */
    const double per_thread = (double)iters * CH * ((MODE == 0 || MODE == 1) ? 2 : 1);
    const double warp_instr = per_thread * 8.0 * ctas_per_sm * nsm;      /* 8 warps per CTA */
    const double flop_lanes = (double)iters * CH * 2 * 256.0 * ctas_per_sm * nsm;   /* fp32 lane-operations */
    printf("%-28s ctas/SM %d: %.3f ms  %.2f warp-instr/ns/SM  %.1f G lane-ops/s/SM (x%d SMs = %.1f T lane-ops/s)\n", name, ctas_per_sm, ms,
           warp_instr / (ms * 1e6) / nsm, flop_lanes / (ms * 1e-3) / nsm * 1e-9, nsm, flop_lanes / (ms * 1e-3) * 1e-12);
    CK(cudaEventDestroy(a)); CK(cudaEventDestroy(b));
}

/* ---- DSMEM: every CTA of a cluster streams through the shared memory of the next CTA (loads) or writes it (stores) */
template <int CL, int VEC, bool STORE>
__global__ void __launch_bounds__(256) k_dsmem(float* out, int iters, int words)
{
    extern __shared__ __align__(16) float sm[];
    cg::cluster_group cluster = cg::this_cluster();
    const unsigned rank = cluster.block_rank();
    for (int i = threadIdx.x; i < words; i += blockDim.x) sm[i] = (float)i;
    cluster.sync();
    float acc = 0.f;
    for (int it = 0; it < iters; it++) {
        const unsigned peer = (rank + 1 + (it % (CL - 1))) % CL;
        float* remote = cluster.map_shared_rank(sm, peer);
        if (VEC == 4) {
            float4* r4 = reinterpret_cast<float4*>(remote);
            for (int i = threadIdx.x; i < words / 4; i += blockDim.x) {
                if (STORE) r4[i] = make_float4(acc, acc, acc, (float)it);
                else { const float4 v = r4[i]; acc += v.x + v.y + v.z + v.w; }
            }
        } else {
            float2* r2 = reinterpret_cast<float2*>(remote);
            for (int i = threadIdx.x; i < words / 2; i += blockDim.x) {
                if (STORE) r2[i] = make_float2(acc, (float)it);
                else { const float2 v = r2[i]; acc += v.x + v.y; }
            }
        }
    }
    cluster.sync();
    if (acc == 1234.5f) out[threadIdx.x] = acc + sm[threadIdx.x];
}
/* the same loop on the CTA's own shared memory, for comparison */
template <int VEC>
__global__ void __launch_bounds__(256) k_smem_local(float* out, int iters, int words)
{
    extern __shared__ __align__(16) float sm[];
    for (int i = threadIdx.x; i < words; i += blockDim.x) sm[i] = (float)i;
    __syncthreads();
    float acc = 0.f;
    for (int it = 0; it < iters; it++) {
        if (VEC == 4) {
            const float4* r4 = reinterpret_cast<const float4*>(sm);
            for (int i = threadIdx.x; i < words / 4; i += blockDim.x) { const float4 v = r4[i]; acc += v.x + v.y + v.z + v.w; }
        } else {
            const float2* r2 = reinterpret_cast<const float2*>(sm);
            for (int i = threadIdx.x; i < words / 2; i += blockDim.x) { const float2 v = r2[i]; acc += v.x + v.y; }
        }
        __syncthreads();
    }
    if (acc == 1234.5f) out[threadIdx.x] = acc;
}

template <int CL, int VEC, bool STORE> static void run_dsmem(const char* name, float* d_out, int nsm, double ghz)
{
    const int words = 16384;                 /* 64 KiB per CTA */
    const int iters = 200;
    const size_t smem = sizeof(float) * words;
    CK(cudaFuncSetAttribute(k_dsmem<CL, VEC, STORE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (CL > 8) CK(cudaFuncSetAttribute(k_dsmem<CL, VEC, STORE>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    cudaLaunchConfig_t cfg = {};
    cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = CL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int maxcl = 0;
    cfg.gridDim = dim3(CL);
    CK(cudaOccupancyMaxActiveClusters(&maxcl, k_dsmem<CL, VEC, STORE>, &cfg));
    const int ncl = maxcl;                   /* one wave of clusters */
    cfg.gridDim = dim3(ncl * CL);
    cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    CK(cudaLaunchKernelEx(&cfg, k_dsmem<CL, VEC, STORE>, d_out, 2, words));
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(a));
    CK(cudaLaunchKernelEx(&cfg, k_dsmem<CL, VEC, STORE>, d_out, iters, words));
    CK(cudaEventRecord(b)); CK(cudaDeviceSynchronize());
    float ms = 0; CK(cudaEventElapsedTime(&ms, a, b));
    const double bytes = (double)iters * words * 4.0 * ncl * CL;
    printf("%-34s max active clusters %3d (%3d CTAs, 64 KiB smem each): %.3f ms  %.2f TB/s chip  %.1f GB/s per CTA = %.1f B/clk/CTA at %.2f GHz\n", name, maxcl,
           ncl * CL, ms, bytes / (ms * 1e-3) * 1e-12, bytes / (ms * 1e-3) / (ncl * CL) * 1e-9, bytes / (ms * 1e-3) / (ncl * CL) / (ghz * 1e9), ghz);
    CK(cudaEventDestroy(a)); CK(cudaEventDestroy(b));
    (void)nsm;
}
template <int VEC> static void run_local(const char* name, float* d_out, int nsm, double ghz)
{
    const int words = 16384, iters = 200;
    const size_t smem = sizeof(float) * words;
    CK(cudaFuncSetAttribute(k_smem_local<VEC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    k_smem_local<VEC><<<nsm * 2, 256, smem>>>(d_out, 2, words);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(a));
    k_smem_local<VEC><<<nsm * 2, 256, smem>>>(d_out, iters, words);
    CK(cudaEventRecord(b)); CK(cudaDeviceSynchronize());
    float ms = 0; CK(cudaEventElapsedTime(&ms, a, b));
    const double bytes = (double)iters * words * 4.0 * nsm * 2;
    printf("%-34s 2 CTAs/SM: %.3f ms  %.2f TB/s chip  %.1f B/clk/SM at %.2f GHz\n", name, ms, bytes / (ms * 1e-3) * 1e-12,
           bytes / (ms * 1e-3) / nsm / (ghz * 1e9), ghz);
    CK(cudaEventDestroy(a)); CK(cudaEventDestroy(b));
}

int main()
{
    cudaDeviceProp pr; CK(cudaGetDeviceProperties(&pr, 0));
    int khz = 0; CK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0));
    const int nsm = pr.multiProcessorCount; const double ghz = khz * 1e-6;
    printf("%s, %d SMs, %.3f GHz nominal\n", pr.name, nsm, ghz);
    float2* d_out; CK(cudaMalloc(&d_out, 1 << 20));
    for (int c = 1; c <= 4; c *= 2) {
        if (c == 1) { run_fp<0, 8>("FFMA  (scalar, 8 chains x2)", d_out, 2, nsm); run_fp<1, 8>("FADD  (scalar, 8 chains x2)", d_out, 2, nsm);
                      run_fp<2, 8>("FFMA2 (packed, 8 chains)", d_out, 2, nsm); run_fp<3, 8>("FADD2 (packed, 8 chains)", d_out, 2, nsm);
                      run_fp<4, 8>("FFMA2 swapped operand", d_out, 2, nsm); }
        else { run_fp<0, 8>("FFMA  (scalar, 8 chains x2)", d_out, 2 * c, nsm); run_fp<1, 8>("FADD  (scalar, 8 chains x2)", d_out, 2 * c, nsm);
               run_fp<2, 8>("FFMA2 (packed, 8 chains)", d_out, 2 * c, nsm); run_fp<3, 8>("FADD2 (packed, 8 chains)", d_out, 2 * c, nsm);
               run_fp<4, 8>("FFMA2 swapped operand", d_out, 2 * c, nsm); }
    }
    run_fp<2, 2>("FFMA2 2 chains (latency)", d_out, 1, nsm);
    run_fp<0, 1>("FFMA 1x2 chains (latency)", d_out, 1, nsm);
    run_fp<2, 1>("FFMA2 1 chain (latency)", d_out, 1, nsm);
    run_local<2>("local smem LDS.64", (float*)d_out, nsm, ghz);
    run_local<4>("local smem LDS.128", (float*)d_out, nsm, ghz);
    run_dsmem<2, 2, false>("DSMEM load  64-bit cluster 2", (float*)d_out, nsm, ghz);
    run_dsmem<4, 2, false>("DSMEM load  64-bit cluster 4", (float*)d_out, nsm, ghz);
    run_dsmem<4, 4, false>("DSMEM load 128-bit cluster 4", (float*)d_out, nsm, ghz);
    run_dsmem<8, 2, false>("DSMEM load  64-bit cluster 8", (float*)d_out, nsm, ghz);
    run_dsmem<8, 4, false>("DSMEM load 128-bit cluster 8", (float*)d_out, nsm, ghz);
    run_dsmem<4, 2, true>("DSMEM store  64-bit cluster 4", (float*)d_out, nsm, ghz);
    run_dsmem<4, 4, true>("DSMEM store 128-bit cluster 4", (float*)d_out, nsm, ghz);
    run_dsmem<8, 2, true>("DSMEM store  64-bit cluster 8", (float*)d_out, nsm, ghz);
    run_dsmem<8, 4, true>("DSMEM store 128-bit cluster 8", (float*)d_out, nsm, ghz);
    run_dsmem<16, 4, false>("DSMEM load 128-bit cluster 16", (float*)d_out, nsm, ghz);
    return 0;
}
