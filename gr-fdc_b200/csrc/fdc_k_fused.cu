/* fdc_k_fused.cu -- the forward transform of one long block (N = N1 * N2 >= 32768) inside ONE thread-block cluster.
 *
 * The four-step scheme needs a transpose between its column and its row transforms.  fdc_k_fwd.cu does that
 * through global memory (two kernels, the N-point intermediate written and read back: 2 x 8 B per point, and with the
 * chunk sizes that keep the SMs busy it is DRAM traffic, profiles/r2_ncu_warm_step.txt).  Here the CL CTAs of a cluster
 * hold the block in their shared memories and the transpose is a scatter over distributed shared memory:
 *
 *   CTA r, columns [r*BC, (r+1)*BC) (BC = N2 / CL):  loads its column tile straight from the sample stream (overlap-save
 *       addressing, lib/overlap_save_impl.cc:70-78), N1-point transforms, four-step twiddle, and stores A[k1][n2] into the
 *       shared memory of the CTA that owns row k1 (st.shared::cluster) -- nothing of the intermediate touches L2 or HBM
 *   cluster barrier
 *   CTA r, rows [r*BR, (r+1)*BR) (BR = N1 / CL):  N2-point transforms out of its own shared memory, fft-shift and 1/N,
 *       spectrum rows to global memory (python/FrequencyDomainChannelizer.py:206,216)
 *
 * The kernel is persistent: a cluster walks blocks cluster_id, cluster_id + nclusters, ...; the samples of its next block
 * are fetched into registers right after the scatter, so that the loads are in flight during the barrier and the row
 * transforms.  Cluster barriers are split (arrive / wait) around work that does not depend on them. */
#include "fdc_kcommon.cuh"
#include <map>
#include <mutex>
#include <vector>
#include <cstdio>
#include <cstdlib>

namespace fdc {

__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cluster_idx() { uint32_t r; asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cluster_count() { uint32_t r; asm volatile("mov.u32 %0, %%nclusterid.x;" : "=r"(r)); return r; }
/* release / acquire at cluster scope: shared-memory stores into other CTAs issued before the arrive are visible after the wait */
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
/* address of the same shared-memory location in CTA `rank` of the cluster (shared::cluster window) */
__device__ __forceinline__ uint32_t map_to_rank(uint32_t smem_addr, uint32_t rank)
{
    uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank)); return r;
}
__device__ __forceinline__ void st_cluster(uint32_t addr, float2 v)
{
    asm volatile("st.shared::cluster.v2.f32 [%0], {%1, %2};" ::"r"(addr), "f"(v.x), "f"(v.y) : "memory");
}

/* ---- column side: the last pass of the column transforms scatters A[k1][n2] * W_N^{n2 k1} over the cluster.
 * Butterfly o of column `batch` produces k1 = o + NS t, t < R.  Rows are dealt out in runs of BR = N1 / CL: with NS | BR
 * the owner of k1 is (NS t) / BR, a compile-time function of t, and the row inside the owner is o + (NS t) % BR.
 * W_N^{n2 k1} = W_N^{n2 o} * W_N^{n2 NS t}: two small per-CTA tables (T1[o][column], T2[t][column]) instead of the N1 x BC
 * slice of the four-step table (64 KiB for 256 x 32) that fdc_k_fwd.cu keeps per CTA. */
template <int N1, int N2, int CL, int CB, int R, int NS> struct ClusterColStorer {
    static constexpr int BC = N2 / CL, BR = N1 / CL;       /* columns / rows per CTA; CB columns per sub-tile */
    static_assert(NS <= BR && BR % NS == 0, "row ownership must be a compile-time function of the butterfly output");
    struct Ctx { uint32_t dst[CL]; float2 t1; const float2* t2; };
    uint32_t a_base;            /* shared::cta address of A */
    int rank, col0;             /* col0: first column of the sub-tile inside the CTA's BC columns */
    const float2* t1; const float2* t2;
    __device__ __forceinline__ Ctx begin(int batch, int o) const
    {
        Ctx c;
        const uint32_t local = a_base + (uint32_t)sizeof(float2) * (uint32_t)(o * N2 + rank * BC + col0 + batch);
#pragma unroll
        for (int q = 0; q < CL; q++) c.dst[q] = map_to_rank(local, (uint32_t)q);
        c.t1 = t1[o * BC + col0 + batch]; c.t2 = t2 + (col0 + batch);
        return c;
    }
    template <int RR, int NNS> __device__ __forceinline__ void put(const Ctx& c, int t, float2 v) const
    {
        static_assert(RR == R && NNS == NS, "storer built for another last pass");
        const float2 w = t == 0 ? c.t1 : cmul(c.t1, c.t2[t * BC]);
        st_cluster(c.dst[(NS * t) / BR] + (uint32_t)(sizeof(float2) * (size_t)(((NS * t) % BR) * N2)), cmul(v, w));
    }
};
/* ---- row side: operands come from the CTA's own shared memory */
template <int N2> struct SmemRowLoader {
    typedef const float2* Ctx;
    static constexpr bool HAS_FINISH = false;
    const float2* a;
    __device__ __forceinline__ Ctx begin(int batch, int j) const { return a + (batch * N2 + j); }
    template <int R, int STRIDE> __device__ __forceinline__ float2 fetch(const Ctx& row, int t) const { return row[t * STRIDE]; }
    template <int R, int STRIDE> __device__ __forceinline__ float2 finish(const Ctx&, int, float2 raw) const { return raw; }
};

/* SUB sub-tiles per phase: a CTA of T = N / (16 CL SUB) threads walks its BC columns (and then its BR rows) in SUB steps.
 * SUB = 2 halves the CTA (256 threads, 112 KiB for 65536 points on 8 CTAs) so that TWO clusters are resident on every SM:
 * while one drains its scatter or waits at the cluster barrier the other one computes. */
template <int N1, int N2, int CL, int SUB> struct ClusterFwd {
    static constexpr int BC = N2 / CL, BR = N1 / CL, CB = BC / SUB, RB = BR / SUB;
    typedef TileFFT<N1, CB, 1, true, true, 16> CE;               /* CB columns of N1 points */
    typedef TileFFT<N2, RB, 1, false, true, 16> RE;              /* RB rows of N2 points */
    static_assert(CE::T == RE::T, "column and row tiles use the same CTA");
    static constexpr int T = CE::T;
    static constexpr int RC = fft_radix(N1, CE::NP - 1, 16), NSC = fft_ns(N1, CE::NP - 1, 16);      /* last pass of the column transform */
    static constexpr int A_ELEMS = BR * N2;
    static constexpr int X_ELEMS = CE::SMEM_ELEMS > RE::SMEM_ELEMS ? CE::SMEM_ELEMS : RE::SMEM_ELEMS;
    static constexpr int TW_ELEMS = CE::TWSIZE + RE::TWSIZE + NSC * BC + RC * BC;
    static constexpr size_t SMEM_BYTES = sizeof(float2) * (size_t)(A_ELEMS + X_ELEMS + TW_ELEMS);
};

template <int N1, int N2, int CL, int SUB>
__global__ void __launch_bounds__((ClusterFwd<N1, N2, CL, SUB>::T), SUB)
k_fwd_cluster(const BigParams p, const float2* __restrict__ twc_g, const float2* __restrict__ twr_g, long long* prof)
{
    long long tp[6] = {0, 0, 0, 0, 0, 0}, tc = 0;
#define FDC_TICK(i) do { if (prof) { const long long now_ = clock64(); tp[i] += now_ - tc; tc = now_; } } while (0)
    typedef ClusterFwd<N1, N2, CL, SUB> K;
    typedef typename K::CE CE; typedef typename K::RE RE;
    constexpr int BC = K::BC, CB = K::CB, RB = K::RB;
    float2* A = reinterpret_cast<float2*>(fdc_smem_raw);
    float2* X = A + K::A_ELEMS;
    float2* twc = X + K::X_ELEMS;
    float2* twr = twc + CE::TWSIZE;
    float2* t1 = twr + RE::TWSIZE;
    float2* t2 = t1 + K::NSC * BC;
    const int tid = (int)threadIdx.x, rank = (int)cluster_ctarank();
    for (int i = tid; i < CE::TWSIZE; i += K::T) twc[i] = twc_g[i];
    for (int i = tid; i < RE::TWSIZE; i += K::T) twr[i] = twr_g[i];
    /* p.tw4 is [k1][n2] = W_N^{n2 k1} */
    for (int i = tid; i < K::NSC * BC; i += K::T) t1[i] = p.tw4[(long)(i / BC) * N2 + rank * BC + (i % BC)];
    for (int i = tid; i < K::RC * BC; i += K::T) t2[i] = p.tw4[(long)((i / BC) * K::NSC) * N2 + rank * BC + (i % BC)];
    __syncthreads();
    /* every CTA of the cluster is resident from here on: its shared memory may be written by the others */
    cluster_arrive(); cluster_wait();
    const long stride = (long)cluster_count();
    long blk = (long)cluster_idx();
    float2 v[16];
    if (blk < p.nblocks) CE::fetch(tid, v, ColLoader<N1, N2, CB>{p, rank * SUB, blk});
    if (prof) tc = clock64();
    for (; blk < p.nblocks; blk += stride) {
        const long next = blk + stride;
        /* columns of this block -> rows in the shared memories of the cluster; the samples of the next sub-tile (or of the
         * next block's first one) are fetched before the current one is transformed */
#pragma unroll
        for (int sub = 0; sub < SUB; sub++) {
            float2 nx[16];
            asm volatile("" ::: "memory");
            if (sub + 1 < SUB) CE::fetch(tid, nx, ColLoader<N1, N2, CB>{p, rank * SUB + sub + 1, blk});
            else if (next < p.nblocks) CE::fetch(tid, nx, ColLoader<N1, N2, CB>{p, rank * SUB, next});
            asm volatile("" ::: "memory");
            const ClusterColStorer<N1, N2, CL, CB, K::RC, K::NSC> scatter{smem_u32(A), rank, sub * CB, t1, t2};
            tile_fft_from<CE, 0, true>(v, X, twc, scatter);
#pragma unroll
            for (int i = 0; i < 16; i++) v[i] = nx[i];
            if (sub + 1 < SUB) __syncthreads();             /* the exchange buffer is reused */
        }
        FDC_TICK(0);
        cluster_arrive(); cluster_wait();                   /* the intermediate is complete in every CTA */
        FDC_TICK(1);
#pragma unroll
        for (int sub = 0; sub < SUB; sub++) {
            float2 r[16];
            RE::fetch(tid, r, SmemRowLoader<N2>{A + sub * RB * N2});
            if (sub + 1 == SUB) cluster_arrive();           /* this CTA has taken its rows out of A ... */
            tile_fft_from<RE, 0, true>(r, X, twr, RowStorer<N1, N2, RB>{p, rank * SUB + sub, blk});
            __syncthreads();                                /* the exchange buffer is reused */
        }
        FDC_TICK(2);
        cluster_wait();                                     /* ... and so has every other CTA: A may be overwritten */
        FDC_TICK(3);
    }
    if (prof && (tid & 31) == 0) {
#pragma unroll
        for (int i = 0; i < 5; i++) prof[((long)blockIdx.x * (K::T / 32) + (tid >> 5)) * 8 + i] = tp[i];
    }
#undef FDC_TICK
}

template <int N1, int N2, int CL, int SUB> static cudaError_t go_cluster(const BigParams& p, cudaStream_t s)
{
    typedef ClusterFwd<N1, N2, CL, SUB> K;
    auto kernel = k_fwd_cluster<N1, N2, CL, SUB>;
    static std::mutex m;
    static std::map<int, int> max_clusters;          /* per device */
    int dev = 0;
    FDC_CHECK(cudaGetDevice(&dev));
    cudaLaunchConfig_t cfg = {};
    cfg.blockDim = dim3((unsigned)K::T); cfg.dynamicSmemBytes = K::SMEM_BYTES; cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    int ncl = 0;
    {
        std::lock_guard<std::mutex> g(m);
        std::map<int, int>::iterator it = max_clusters.find(dev);
        if (it == max_clusters.end()) {
            FDC_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)K::SMEM_BYTES));
            cfg.gridDim = dim3(CL);
            FDC_CHECK(cudaOccupancyMaxActiveClusters(&ncl, kernel, &cfg));
            if (ncl < 1) return cudaErrorLaunchOutOfResources;
            max_clusters[dev] = ncl;
        } else ncl = it->second;
    }
    if ((long)ncl > p.nblocks) ncl = (int)p.nblocks;
    cfg.gridDim = dim3((unsigned)(ncl * CL));
    count_launch();
    /* FDC_CLUSTER_PROF=1: per-warp clock64 totals of the phases (measurement aid) */
    static long long* d_prof = 0; static int prof_calls = 0;
    const bool want_prof = getenv("FDC_CLUSTER_PROF") != 0;
    if (want_prof && !d_prof) cudaMalloc(&d_prof, sizeof(long long) * 8 * 4096);
    if (want_prof) cudaMemsetAsync(d_prof, 0, sizeof(long long) * 8 * 4096, s);
    const cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, p, twiddle_table(N1), twiddle_table(N2), want_prof ? d_prof : (long long*)0);
    if (want_prof && e == cudaSuccess && ++prof_calls % 8 == 0) {
        std::vector<long long> h(8 * 4096);
        cudaStreamSynchronize(s);
        cudaMemcpy(h.data(), d_prof, sizeof(long long) * h.size(), cudaMemcpyDeviceToHost);
        const int nw = ncl * CL * (K::T / 32);
        double sum[5] = {0, 0, 0, 0, 0};
        for (int w = 0; w < nw && w < 4096; w++) for (int i = 0; i < 5; i++) sum[i] += (double)h[(size_t)w * 8 + i];
        const double per = (double)((p.nblocks + ncl - 1) / ncl) * (nw < 4096 ? nw : 4096);
        fprintf(stderr, "cluster fwd: %ld blocks on %d clusters; cycles per block per warp: cols+scatter %.0f, barrier A %.0f, rows+store %.0f, barrier B %.0f\n",
                p.nblocks, ncl, sum[0] / per, sum[1] / per, sum[2] / per, sum[3] / per);
    }
    return e;
}

bool fwd_cluster_supported(int N) { return N == 65536 || N == 32768; }
cudaError_t launch_fwd_cluster(const BigParams& p, int N, cudaStream_t s)
{
    if (p.nblocks <= 0) return cudaSuccess;
    switch (N) {
    case 32768: return go_cluster<128, 256, 8, 1>(p, s);
    case 65536: return tuning().fused > 1 ? go_cluster<256, 256, 8, 1>(p, s) : go_cluster<256, 256, 8, 2>(p, s);
    }
    return cudaErrorInvalidValue;
}

}  // namespace fdc
