/* fdc_launch.h -- host-callable launchers of the sm_100a kernels (internal to libfdc_b200.so). */
#ifndef FDC_LAUNCH_H
#define FDC_LAUNCH_H
#include <cuda_runtime.h>
#include "fdc_functors.cuh"

namespace fdc {

/* every launcher counts its launches here (fdc_launch_count) */
void count_launch(int n = 1);
unsigned long long launch_count();

/* run-time switches for measurements (environment: FDC_PREFETCH=bit 0 forward kernels | bit 1 extract kernels, FDC_CTAS_FWD / FDC_CTAS_EXT=n, FDC_CTAS_PER_SM=n, FDC_STREAMS=1..4, FDC_EXTRACT_E8=-1 (auto: slices <= 64) | 0 | 1, FDC_PDL=0|1, FDC_FWD_SPLIT=N, FDC_EXTRACT_E32=0|1, FDC_FWD_E32=0|1|2 (2: also 512/1024-point transforms), FDC_HOST_CHUNK_MB=n, FDC_L2PF=0|1 (bulk L2 prefetch of the extract's next tile; measured: +1 %, default on), FDC_PACK=0|1 (tiles spanning several blocks when a launch has few channels; default on), FDC_HOST_STAGING=0|1|2 (host path: never / for pageable caller memory (default) / always stage through the library's pinned slots), FDC_COPY_THREADS=n (threads of the staging copy pool), FDC_SINK_DMA=0|1 (channel-sharded sinks: 0 (default) the extract kernel stores into the owners' memory itself, 1 rows of remote owners go to a local staging slab and the copy engines forward them behind the kernels; measured on 2 GPUs: 111 against 97 Gsample/s), FDC_FUSE_SMALL=0|1|2 (N <= 16384, one slice length: forward transform and all channels in ONE kernel, the spectrum stays in shared memory, fdc_k_chanfused.cu; 1: one thread group does both halves in turn, 2: two groups of one CTA, forward and extract, hand spectra over through named barriers; measured on cfg2 107 / 100 against 112 Gsample/s of the two kernels, cfg1 100 / 86 against 97, default 0), FDC_FUSED=0|1 (N >= 32768: forward transform inside a thread-block cluster, the four-step transpose over distributed shared memory, fdc_k_fused.cu)) */
struct Tuning { int prefetch; int ctas_per_sm; int streams; int extract_e8; int ctas_fwd; int ctas_ext; int pdl; int fwd_split; int host_chunk_mb; int extract_e32; int fwd_e32; int l2pf; int pack; int fused; int host_staging; int sink_dma; int fuse_small; };
const Tuning& tuning();

/* per-pass Stockham twiddles of a length-L tile FFT (layout of fdc_tile_fft.cuh: pass p at fft_twoff(L, p), entry
 * [(t-1)*Ns + k] = exp(-2 pi i k t / (Ns R))), on the current device (cached per device and L) */
const float2* twiddle_table(int L, int E = 16);
/* four-step twiddles W_N^{n2 k1} laid out [k1][n2] (N = N1*N2 entries) */
const float2* fourstep_table(int N1, int N2);

bool fwd_small_supported(int N);      /* N handled by one CTA (16 .. 16384) */
bool fwd_big_supported(int N, int* N1, int* N2);
bool tile_len_supported(int L);       /* inverse / plain tile lengths (2 .. 16384) */

cudaError_t launch_fwd_small(const FwdParams& p, cudaStream_t s);
cudaError_t launch_fwd_big(const BigParams& p, int N, cudaStream_t s);
bool fwd_cluster_supported(int N);    /* N whose block fits the shared memory of one thread-block cluster */
cudaError_t launch_fwd_cluster(const BigParams& p, int N, cudaStream_t s);
/* nsel channels sharing slice length l */
cudaError_t launch_extract(const ExtractParams& p, int l, cudaStream_t s);
cudaError_t launch_jobs(const JobParams& p, int l, cudaStream_t s);
/* K1 + K2 in one kernel (N <= 16384, every channel of slice length l): the spectrum never leaves shared memory */
bool chan_fused_supported(int N, int l);
cudaError_t launch_chan_fused(const FwdParams& fp, const ExtractParams& xp, int l, cudaStream_t s);
cudaError_t launch_plain_fft(const PlainParams& p, int L, int forward, cudaStream_t s);

/* out row b (row_bytes) <- src[b*src_stride + src_off ...); source bytes before 0 come from hist
 * (hist holds hist_bytes bytes that logically precede src[0]).  overlap_save and vector_cut_vxx. */
cudaError_t launch_rowcopy(const void* src, const void* hist, long hist_bytes, void* dst, long nrows,
                           long row_bytes, long src_stride, long src_off, cudaStream_t s);
/* out[b][k] = in[b][k] * table[(counter + b*shift) % nphase][k]  (VOLK-exact complex multiply) */
cudaError_t launch_psw(const float2* in, float2* out, const float2* table, long nblocks, int l, int nphase,
                       int counter, int shift, cudaStream_t s);

/* out = in * k (real constant), n complex items (n even, 16-byte aligned pointers) */
cudaError_t launch_scale(const float2* in, float2* out, long n, float k, cudaStream_t s);

/* ---- K3: power / threshold / edges ---------------------------------------------------------- */
/* P[b*M + i] = sum_{k<D} |X_b[start + i*D + k]|^2, strictly sequential fp32 (generic VOLK order,
 * lib/SegmentDetection_impl.cc:185-190); mean != 0 multiplies by 1/D afterwards
 * (lib/activity_detection_channelizer_vcm_impl.cc:633-648). */
cudaError_t launch_group_power(const float2* spec, long spec_stride, long nblocks, int start, int D, int M,
                               int mean, float* P, cudaStream_t s);
/* per block: rising edges (ratio > T) as (ratio, index i) and falling edges (ratio < 1/T) as index i, in
 * ascending i, compacted with warp ballots.  guard != 0: a zero denominator is replaced by FLT_MIN
 * (activity_detection_channelizer_vcm_impl.cc:703-704).  Layout per block: counts[b*2+{0,1}],
 * rise_ratio[b*cap + n], rise_idx[b*cap + n], fall_idx[b*cap + n]. */
cudaError_t launch_edges(const float* P, long nblocks, int M, float T, float invT, int guard, int cap,
                         int* counts, float* rise_ratio, int* rise_idx, int* fall_idx, cudaStream_t s);
/* pwr[b] = sum_{i in [m0, m1)} re(x * conj x), sequential (lib/PowerActivationChannel_impl.cc:289-291) */
cudaError_t launch_waterfall_rows(const float2* spec, long spec_stride, long nblocks, int N, int W, int logmode, float* out, cudaStream_t s);
cudaError_t launch_band_power(const float2* spec, long spec_stride, long nblocks, int m0, int m1, float* pwr,
                              cudaStream_t s);

}  // namespace fdc
#endif
