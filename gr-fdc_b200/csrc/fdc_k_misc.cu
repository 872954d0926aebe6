/* fdc_k_misc.cu -- twiddle tables, the copy-type blocks (overlap_save, vector_cut_vxx), the stand-alone
 * phase_shifting_windowing_vcc multiply, and K3: power / threshold / edge kernels of the activity-gated blocks. */
#include "fdc_kcommon.cuh"
#include <atomic>
#include <map>
#include <mutex>
#include <vector>
#include <cmath>
#include <cfloat>
#include <cstdlib>

namespace fdc {

static std::atomic<unsigned long long> g_launches(0);
void count_launch(int n) { g_launches += (unsigned long long)n; }
unsigned long long launch_count() { return g_launches.load(); }

/* ---- run-time switches ------------------------------------------------------------------------ */
static int env_int(const char* name, int dflt)
{
    const char* v = getenv(name);
    return (v && *v) ? atoi(v) : dflt;
}
const Tuning& tuning()
{
    static const Tuning t = { env_int("FDC_PREFETCH", 1), env_int("FDC_CTAS_PER_SM", 0), env_int("FDC_STREAMS", 3), env_int("FDC_EXTRACT_E8", -1), env_int("FDC_CTAS_FWD", 0), env_int("FDC_CTAS_EXT", 0), env_int("FDC_PDL", 1), env_int("FDC_FWD_SPLIT", 32768), env_int("FDC_HOST_CHUNK_MB", 8), env_int("FDC_EXTRACT_E32", 1), env_int("FDC_FWD_E32", 1), env_int("FDC_L2PF", 1), env_int("FDC_PACK", 1), env_int("FDC_FUSED", 0), env_int("FDC_HOST_STAGING", 1), env_int("FDC_SINK_DMA", 0), env_int("FDC_FUSE_SMALL", 0) };
    return t;
}

/* ---- launch geometry cache --------------------------------------------------------------------- */
cudaError_t kernel_capacity(const void* kernel, int threads, size_t smem, int* capacity, int limit)
{
    static std::mutex m;
    static std::map<std::pair<int, const void*>, int> cache;
    int dev = 0;
    FDC_CHECK(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> g(m);
    std::map<std::pair<int, const void*>, int>::iterator it = cache.find(std::make_pair(dev, kernel));
    if (it != cache.end()) { *capacity = it->second; return cudaSuccess; }
    int sms = 0, per_sm = 0;
    if (smem > 48 * 1024) FDC_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    FDC_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    FDC_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem));
    if (per_sm < 1) per_sm = 1;
    int lim = tuning().ctas_per_sm;
    if (limit > 0 && (lim <= 0 || limit < lim)) lim = limit;
    if (lim > 0 && per_sm > lim) per_sm = lim;
    *capacity = sms * per_sm;
    cache[std::make_pair(dev, kernel)] = *capacity;
    return cudaSuccess;
}

/* ---- twiddle tables (per device, per length) ------------------------------------------------ */
static std::mutex g_tw_mutex;
static std::map<std::pair<int, long>, float2*> g_tw;      /* (device, key) -> device pointer */

static float2 root(long m, long N)                          /* exp(-2 pi i m / N), computed in long double */
{
    const long double a = -2.0L * 3.141592653589793238462643383279502884L * (long double)(m % N) / (long double)N;
    return make_float2((float)cosl(a), (float)sinl(a));
}
static float2* upload(const std::vector<float2>& h)
{
    float2* d = 0;
    if (cudaMalloc(&d, sizeof(float2) * h.size()) != cudaSuccess) return 0;
    if (cudaMemcpy(d, h.data(), sizeof(float2) * h.size(), cudaMemcpyHostToDevice) != cudaSuccess) { cudaFree(d); return 0; }
    return d;
}
static void host_pass_twiddles(int L, int E, std::vector<float2>& h)
{
    h.assign((size_t)fft_twsize(L, E), make_float2(1.f, 0.f));
    const int np = fft_npasses(L, E);
    for (int p = 1; p < np; p++) {
        const int R = fft_radix(L, p, E), NS = fft_ns(L, p, E), off = fft_twoff(L, p, E);
        for (int t = 1; t < R; t++) {
            if (!fft_twstored(t, R)) continue;                /* FDC_TWGEN: the kernels generate the powers in between */
            for (int k = 0; k < NS; k++) h[(size_t)(off + fft_twrow(t, R) * NS + k)] = root((long)k * t, (long)NS * R);
        }
    }
}
const float2* twiddle_table(int L, int E)
{
    int dev = 0; cudaGetDevice(&dev);
    std::lock_guard<std::mutex> g(g_tw_mutex);
    float2*& p = g_tw[std::make_pair(dev, (long)L * 64 + E)];
    if (!p) { std::vector<float2> h; host_pass_twiddles(L < 1 ? 1 : L, E, h); p = upload(h); }
    return p;
}
const float2* fourstep_table(int N1, int N2)
{
    int dev = 0; cudaGetDevice(&dev);
    const long N = (long)N1 * N2;
    std::lock_guard<std::mutex> g(g_tw_mutex);
    float2*& p = g_tw[std::make_pair(dev, -(N + ((long)N1 << 32)))];
    if (!p) {
        std::vector<float2> h((size_t)N);
        for (long k1 = 0; k1 < N1; k1++)
            for (long n2 = 0; n2 < N2; n2++) h[(size_t)(k1 * N2 + n2)] = root(k1 * n2, N);
        p = upload(h);
    }
    return p;
}

/* ---- row copy: overlap_save (lib/overlap_save_impl.cc:70-78) and vector_cut_vxx (lib/vector_cut_vxx_impl.cc:67-68)
 * both are "row b of the output = a window of the input byte stream"; bytes before the start of the input come
 * from the history saved by the previous call. */
template <class U>
__global__ void __launch_bounds__(256) k_rowcopy(const U* __restrict__ src, const U* __restrict__ hist, long hist_units,
                                                 U* __restrict__ dst, long nrows, long row_units, long src_stride,
                                                 long src_off)
{
    const long total = nrows * row_units;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const long b = i / row_units, k = i - b * row_units;
        const long g = b * src_stride + src_off + k;
        dst[i] = g < 0 ? hist[hist_units + g] : src[g];
    }
}
cudaError_t launch_rowcopy(const void* src, const void* hist, long hist_bytes, void* dst, long nrows, long row_bytes,
                           long src_stride, long src_off, cudaStream_t s)
{
    if (nrows <= 0 || row_bytes <= 0) return cudaSuccess;
    const uintptr_t al = (uintptr_t)src | (uintptr_t)hist | (uintptr_t)dst | (uintptr_t)hist_bytes | (uintptr_t)row_bytes |
                         (uintptr_t)src_stride | (uintptr_t)(src_off < 0 ? -src_off : src_off);
    const long total = nrows * row_bytes;
    int unit = 1;
    if ((al & 15) == 0) unit = 16; else if ((al & 7) == 0) unit = 8; else if ((al & 3) == 0) unit = 4; else if ((al & 1) == 0) unit = 2;
    long blocks = (total / unit + 255) / 256; if (blocks > 148L * 16) blocks = 148L * 16; if (blocks < 1) blocks = 1;
#define GO(TT) k_rowcopy<TT><<<(unsigned)blocks, 256, 0, s>>>((const TT*)src, (const TT*)hist, hist_bytes / unit, (TT*)dst, nrows, \
                                                              row_bytes / unit, src_stride / unit, src_off / unit)
    switch (unit) {
    case 16: GO(uint4); break;
    case 8: GO(uint2); break;
    case 4: GO(unsigned); break;
    case 2: GO(unsigned short); break;
    default: GO(unsigned char); break;
    }
#undef GO
    count_launch();
    return cudaGetLastError();
}

/* ---- phase_shifting_windowing_vcc::work (lib/phase_shifting_windowing_vcc_impl.cc:72-86) ------- */
__global__ void __launch_bounds__(256) k_psw(const float2* __restrict__ in, float2* __restrict__ out,
                                             const float2* __restrict__ table, long nblocks, int l, int nphase, int counter,
                                             int shift)
{
    const long total = nblocks * l;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const long b = i / l; const int k = (int)(i - b * l);
        const int ph = (int)((counter + (b % nphase) * shift) % nphase);
        out[i] = cmul_exact(in[i], __ldg(table + (long)ph * l + k));
    }
}
cudaError_t launch_psw(const float2* in, float2* out, const float2* table, long nblocks, int l, int nphase, int counter,
                       int shift, cudaStream_t s)
{
    if (nblocks <= 0) return cudaSuccess;
    long blocks = (nblocks * l + 255) / 256; if (blocks > 148L * 16) blocks = 148L * 16;
    k_psw<<<(unsigned)blocks, 256, 0, s>>>(in, out, table, nblocks, l, nphase, counter, shift);
    count_launch();
    return cudaGetLastError();
}

/* ---- blocks.multiply_const_cc(k) with a real constant (the hier block's normalize_input, python/FrequencyDomainChannelizer.py:216),
 * for spectra that arrive already transformed (inpveclen > 1, :284-290) ---- */
__global__ void __launch_bounds__(256) k_scale(const float4* __restrict__ in, float4* __restrict__ out, long n4, float k)
{
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long)gridDim.x * blockDim.x) {
        const float4 v = in[i];
        out[i] = make_float4(fdc_mul(v.x, k), fdc_mul(v.y, k), fdc_mul(v.z, k), fdc_mul(v.w, k));
    }
}
cudaError_t launch_scale(const float2* in, float2* out, long n, float k, cudaStream_t s)
{
    if (n <= 0) return cudaSuccess;
    if ((n & 1) || ((uintptr_t)in & 15) || ((uintptr_t)out & 15)) return cudaErrorInvalidValue;
    long blocks = (n / 2 + 255) / 256; if (blocks > 148L * 16) blocks = 148L * 16;
    k_scale<<<(unsigned)blocks, 256, 0, s>>>((const float4*)in, (float4*)out, n / 2, k);
    count_launch();
    return cudaGetLastError();
}

/* ---- K3 ------------------------------------------------------------------------------------------ */
/* P[b][i] = sum_{k < D} |X_b[start + i D + k]|^2 in the order of the generic VOLK accumulator (k ascending, every product and
 * sum rounded on its own).  A warp owns 32 adjacent power bins = 32 D contiguous spectrum bins and walks them in 32 x 32 tiles:
 * for every bin row the 32 lanes load 32 consecutive spectrum bins (one coalesced 256-byte request), square them and park the
 * 32 x 32 powers in shared memory; then lane r adds row r's 32 values sequentially.  |x|^2 of an element does not depend on
 * the order, the running sum keeps it: bit-identical to one thread walking its D bins, without its stride-D loads. */
__global__ void __launch_bounds__(256) k_group_power(const float2* __restrict__ spec, long spec_stride, int start, int D,
                                                     int M, int mean, float* __restrict__ P)
{
    __shared__ float tile[8][32][33];
    const long b = blockIdx.y;
    const float2* row = spec + b * spec_stride + start;
    const float norm = 1.0f / (float)D;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i0 = (blockIdx.x * 8 + warp) * 32; i0 < M; i0 += gridDim.x * 256) {
        const int rows = M - i0 < 32 ? M - i0 : 32;
        float acc = 0.0f;
        for (int k0 = 0; k0 < D; k0 += 32) {
            const int kn = D - k0 < 32 ? D - k0 : 32;
            for (int r = 0; r < rows; r++) {
                if (lane < kn) {
                    const float2 v = __ldg(row + ((long)(i0 + r) * D + k0 + lane));
                    tile[warp][r][lane] = fdc_add(fdc_mul(v.x, v.x), fdc_mul(v.y, v.y));
                }
            }
            __syncwarp();
            if (lane < rows)
                for (int k = 0; k < kn; k++) acc = fdc_add(acc, tile[warp][lane][k]);
            __syncwarp();
        }
        if (lane < rows) P[b * M + i0 + lane] = mean ? fdc_mul(acc, norm) : acc;
    }
}
cudaError_t launch_group_power(const float2* spec, long spec_stride, long nblocks, int start, int D, int M, int mean,
                               float* P, cudaStream_t s)
{
    if (nblocks <= 0 || M <= 0) return cudaSuccess;
    for (long b0 = 0; b0 < nblocks; b0 += 65535) {
        const long nb = nblocks - b0 < 65535 ? nblocks - b0 : 65535;
        int gx = (M + 255) / 256; if (gx > 1024) gx = 1024;      /* a CTA takes 8 x 32 power bins per round */
        k_group_power<<<dim3((unsigned)gx, (unsigned)nb), 256, 0, s>>>(spec + b0 * spec_stride, spec_stride, start, D, M, mean,
                                                                      P + b0 * M);
        count_launch();
    }
    return cudaGetLastError();
}

/* one CTA per block: ratios of adjacent power bins, threshold classification, ordered compaction of the two edge
 * lists with warp ballots + a CTA-level running offset (ascending bin order is what the reference's deque holds
 * before std::sort, lib/SegmentDetection_impl.cc:206-217). */
__global__ void __launch_bounds__(256) k_edges(const float* __restrict__ P, int M, float T, float invT, int guard, int cap,
                                               int* __restrict__ counts, float* __restrict__ rise_ratio,
                                               int* __restrict__ rise_idx, int* __restrict__ fall_idx)
{
    __shared__ int s_warp[2][8];
    __shared__ int s_base[2];
    const long b = blockIdx.x;
    const float* p = P + b * M;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x < 2) s_base[threadIdx.x] = 0;
    __syncthreads();
    const int n = M - 1;
    for (int i0 = 0; i0 < n; i0 += 256) {
        const int i = i0 + threadIdx.x;
        bool rise = false, fall = false; float r = 0.0f;
        if (i < n) {
            float den = p[i];
            if (guard && den == 0.0f) den = FLT_MIN;
            r = fdc_div(p[i + 1], den);
            rise = r > T;
            fall = !rise && r < invT;
        }
        const unsigned mr = __ballot_sync(0xffffffffu, rise), mf = __ballot_sync(0xffffffffu, fall);
        if (lane == 0) { s_warp[0][warp] = __popc(mr); s_warp[1][warp] = __popc(mf); }
        __syncthreads();
        int offr = s_base[0], offf = s_base[1];
        for (int w = 0; w < warp; w++) { offr += s_warp[0][w]; offf += s_warp[1][w]; }
        const unsigned lower = (1u << lane) - 1u;
        if (rise) { const int o = offr + __popc(mr & lower); if (o < cap) { rise_ratio[b * cap + o] = r; rise_idx[b * cap + o] = i; } }
        if (fall) { const int o = offf + __popc(mf & lower); if (o < cap) fall_idx[b * cap + o] = i; }
        __syncthreads();
        if (threadIdx.x == 0) {
            int tr = 0, tf = 0;
            for (int w = 0; w < 8; w++) { tr += s_warp[0][w]; tf += s_warp[1][w]; }
            s_base[0] += tr; s_base[1] += tf;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) { counts[b * 2] = s_base[0]; counts[b * 2 + 1] = s_base[1]; }
}
cudaError_t launch_edges(const float* P, long nblocks, int M, float T, float invT, int guard, int cap, int* counts,
                         float* rise_ratio, int* rise_idx, int* fall_idx, cudaStream_t s)
{
    if (nblocks <= 0) return cudaSuccess;
    k_edges<<<(unsigned)nblocks, 256, 0, s>>>(P, M, T, invT, guard, cap, counts, rise_ratio, rise_idx, fall_idx);
    count_launch();
    return cudaGetLastError();
}

/* pwr[b] = sum over the measurement band [m0, m1), strictly sequential (lib/PowerActivationChannel_impl.cc:289-291);
 * std::real(x * conj(x)) is a*a - b*(-b): two rounded products and one rounded sum per element.  One warp per block: the lanes
 * load 32 consecutive bins at a time (coalesced, the next chunk in flight while the current one is added) and the running sum
 * takes the 32 element powers in bin order through shuffles -- the same additions in the same order as a single thread, without
 * its chain of dependent, uncoalesced loads (a scheduler hands over 1 - 8 blocks per call: latency is what counts). */
__global__ void __launch_bounds__(128) k_band_power(const float2* __restrict__ spec, long spec_stride, long nblocks, int m0,
                                                    int m1, float* __restrict__ pwr)
{
    const long b = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (b >= nblocks) return;
    const int lane = threadIdx.x & 31;
    const float2* x = spec + b * spec_stride;
    float acc = 0.0f;
    float2 nx = make_float2(0.f, 0.f);
    if (m0 + lane < m1) nx = __ldg(x + m0 + lane);
    for (int i0 = m0; i0 < m1; i0 += 32) {
        const float2 v = nx;
        if (i0 + 32 + lane < m1) nx = __ldg(x + i0 + 32 + lane);
        const float p = fdc_sub(fdc_mul(v.x, v.x), fdc_mul(v.y, -v.y));
        const int n = m1 - i0 < 32 ? m1 - i0 : 32;
        for (int j = 0; j < n; j++) acc = fdc_add(acc, __shfl_sync(0xffffffffu, p, j));
    }
    if (lane == 0) pwr[b] = acc;
}
cudaError_t launch_band_power(const float2* spec, long spec_stride, long nblocks, int m0, int m1, float* pwr, cudaStream_t s)
{
    if (nblocks <= 0) return cudaSuccess;
    k_band_power<<<(unsigned)((nblocks + 3) / 4), 128, 0, s>>>(spec, spec_stride, nblocks, m0, m1, pwr);       /* 4 warps = 4 blocks per CTA */
    count_launch();
    return cudaGetLastError();
}

/* Decimated power rows for a waterfall display (python/WaterfallMsgTagging.py:272-277 reduces every input vector to 1024
 * columns by a mean; upstream of it a flowgraph has complex_to_mag_squared and, for a dB display, nlog10_ff).  All three on
 * the device: W columns per block leave the GPU instead of N.  red = N / W bins per column (N >= W) or every bin repeated
 * rep = W / N times (N < W).  A warp takes 32 consecutive bins at a time, so the reads are coalesced. */
__global__ void __launch_bounds__(256) k_waterfall_rows(const float2* __restrict__ spec, long spec_stride, long nblocks, int W, int red,
                                                         int rep, int logmode, float* __restrict__ out)
{
    const int lane = threadIdx.x & 31;
    const long warp = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int cpw = red >= 32 ? 1 : 32 / red;                  /* columns per warp */
    const int src_w = W / rep;                                 /* distinct columns */
    const long wpb = (src_w + cpw - 1) / cpw;                  /* warps per block */
    const long b = warp / wpb;
    if (b >= nblocks) return;
    const int c0 = (int)(warp - b * wpb) * cpw;
    const float2* x = spec + b * spec_stride + (long)c0 * red;
    float acc = 0.0f;
    if (red >= 32) {
        for (int i = lane; i < red; i += 32) {
            const float2 v = __ldg(x + i);
            const float pw = v.x * v.x + v.y * v.y;
            acc += logmode ? 10.0f * log10f(pw) : pw;
        }
        for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    } else {
        if (c0 + lane / red < src_w) {
            const float2 v = __ldg(x + lane);
            const float pw = v.x * v.x + v.y * v.y;
            acc = logmode ? 10.0f * log10f(pw) : pw;
        }
        for (int o = red >> 1; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    }
    const int col = c0 + (red >= 32 ? 0 : lane / red);
    const bool leader = red >= 32 ? lane == 0 : (lane % red) == 0;
    if (leader && col < src_w) {
        const float m = acc / (float)red;
        for (int r = 0; r < rep; r++) out[b * W + (long)col * rep + r] = m;
    }
}
cudaError_t launch_waterfall_rows(const float2* spec, long spec_stride, long nblocks, int N, int W, int logmode, float* out, cudaStream_t s)
{
    if (nblocks <= 0) return cudaSuccess;
    const int red = N >= W ? N / W : 1, rep = N >= W ? 1 : W / N;
    const int cpw = red >= 32 ? 1 : 32 / red;
    const long wpb = (W / rep + cpw - 1) / cpw;
    const long warps = nblocks * wpb;
    k_waterfall_rows<<<(unsigned)((warps + 7) / 8), 256, 0, s>>>(spec, spec_stride, nblocks, W, red, rep, logmode, out);
    count_launch();
    return cudaGetLastError();
}

}  // namespace fdc
