/* fdc_cabi.h -- C ABI of the B200-native gr-FDC hot path (libfdc_b200.so).
 *
 * gr-FDC itself has no FFI: its blocks are C++ gr::sync_block subclasses created through
 * `static sptr make(...)` (include/FDC/<block>.h:49 in the reference) whose work() bodies live in
 * lib/<block>_impl.cc.  This header is the boundary a maintainer binds instead of those bodies:
 * every entry point below names the reference interface it replaces.  Plain pointers and sizes
 * only, no C++/torch types, no exceptions across the boundary: constructors return NULL and every
 * other call a negative status on failure, with the text in fdc_last_error() (the block wrappers
 * re-throw it as std::invalid_argument, which is what the reference constructors throw).
 *
 * All sample buffers are gr_complex = interleaved float32 (re, im), 8 bytes per item.
 * "host" entry points take ordinary (preferably pinned, see fdc_host_alloc) host memory and
 * include the host<->device copies; "device" entry points take CUDA device pointers and a
 * cudaStream_t passed as void* (NULL = the context's own stream) and only enqueue work.
 * A context is used by one caller thread at a time (GNU Radio never re-enters work() of one
 * block); different contexts may be used concurrently.  There is NO CPU fallback: without a
 * usable CUDA device every constructor fails. */
#ifndef FDC_CABI_H
#define FDC_CABI_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default)
#endif

#define FDC_API_VERSION 1

/* window types, lib/windows.h:29-33 */
enum { FDC_WIN_RECTANGULAR = 0, FDC_WIN_HANN = 1, FDC_WIN_RAMP = 2 };

/* ---- library / device -------------------------------------------------------------------- */
int fdc_api_version(void);
const char* fdc_last_error(void);                 /* thread local, never NULL */
int fdc_device_count(void);                        /* number of CUDA devices, <= 0 if none */
int fdc_set_device(int device);                    /* device used by contexts created afterwards on this thread */
/* Host buffers of the *_host entry points.  GNU Radio's work() hands over scheduler-owned, pageable buffers
 * (lib/overlap_save_impl.cc:62-81): the library stages them through pinned slots it owns (a ring of 4 per context) with a
 * small pool of copy threads, so the caller needs nothing special.  Memory from fdc_host_alloc (or any page-locked
 * memory) is recognised and used in place: no staging copy. */
void* fdc_host_alloc(size_t bytes);                /* pinned host memory */
void fdc_host_free(void* p);
int fdc_copy_threads(void);                        /* worker threads of the staging copy pool (FDC_COPY_THREADS) */
/* Page-lock memory the caller owns for as long as it stays registered (e.g. the scheduler's stream buffers, once, when the
 * flowgraph starts): the *_host entry points then move it by DMA in place, at the speed of fdc_host_alloc memory.  The
 * caller must unregister before freeing or remapping the range. */
int fdc_host_register(void* p, size_t bytes);
int fdc_host_unregister(void* p);
/* A consumer that reads a PDU payload IN PLACE (fdc_msg.data may point into the block's pinned result buffer, which the GPU
 * writes again at the next work() call) should call this on the range when it is done: it drops the lines from the CPU caches
 * (clflushopt).  A DMA write into lines that several cores still hold is up to 10x slower than into lines nobody caches
 * (measured: 10 MB in 1.1-2.0 ms against 0.185 ms, profiles/r2_d2h_after_cpu_reads.txt).  The library's own copies
 * (*_msg_copy_data, the staging slots of the *_work_host calls) do it themselves.  Never required for correctness. */
void fdc_host_evict(const void* p, size_t bytes);
/* plain device memory + blocking copies, so that a host language without a CUDA binding can keep a call's
 * spectrum on the GPU between fdc_chan_work_device and the activity-gated blocks' *_work_device */
void* fdc_dev_alloc(size_t bytes);
void fdc_dev_free(void* p);
int fdc_memcpy_h2d(void* dst_device, const void* src_host, size_t bytes);
int fdc_memcpy_d2h(void* dst_host, const void* src_device, size_t bytes);
int fdc_device_synchronize(void);
/* number of kernel launches this library has enqueued so far (all contexts, this process) */
unsigned long long fdc_launch_count(void);

/* ---- geometry and window tables (host side, bit exact restatements) ----------------------- */
/* FrequencyDomainChannelizer.get_opt_channelparams, python/FrequencyDomainChannelizer.py:322-345
 * (freq and bw already converted with get_freq/get_bw, i.e. DC at 0.5). */
int fdc_opt_channelparams(int blocksize, int relinvovl, double freq, double bw,
                          int* f, int* l, int* lout, double* passband, double* stopband);
/* cr_win(wintype, blocksize, passbw, stopbw, w, relinvovl, 1, false), lib/windows.h:41-78, as called by
 * phase_shifting_windowing_vcc_impl (lib/phase_shifting_windowing_vcc_impl.cc:62).
 * out: relinvovl * blocksize complex floats. */
int fdc_psw_build_tables(int blocklen, int numphasestates, float passbw, float stopbw, int windowtype, float* out);

/* ---- the fused throughput channelizer ------------------------------------------------------
 * Replaces the chain the hier block wires (python/FrequencyDomainChannelizer.py:201-231, 284-315):
 *   stream_to_vector -> FDC.overlap_save -> fft_vcc(N, fwd, shift) -> multiply_const(1/N)
 *   -> per channel: vector_cut_vxx(f,l) -> phase_shifting_windowing_vcc -> fft_vcc(l, inverse, shift)
 *                   -> vector_cut_vxx(l-lout, lout) -> vector_to_stream -> multiply_const(l)
 * One forward-FFT kernel (overlap staging, fft-shift and 1/N fused) and one batched channel-extract
 * kernel per distinct l. */
typedef struct fdc_chan fdc_chan;
typedef struct {
    int f;               /* first bin of the slice in the fft-shifted spectrum (vector_cut offset) */
    int l;               /* slice / inverse FFT length, power of two */
    int lout;            /* samples kept per block: the last lout of the l IFFT outputs */
    int shift;           /* phase table advance per block (phase_shifting_windowing_vcc `shifts`) */
    float gain;          /* output scale (multiply_const_cc, the hier block uses l) */
    const float* table;  /* nphase * l complex floats: table[p][k] multiplies bin f+k on blocks with phase p */
} fdc_chan_desc;

/* N: forward FFT length (power of two), ovl: samples re-used from the previous block (0 <= ovl < N;
 * the reference's overlap_save is only defined for ovl <= N/2, lib/overlap_save_impl.cc:74-78),
 * nphase: number of phase states of every table. */
fdc_chan* fdc_chan_create(int N, int ovl, int nphase, int nchan, const fdc_chan_desc* ch);
void fdc_chan_destroy(fdc_chan* c);
int fdc_chan_hop(const fdc_chan* c);              /* N - ovl new samples per block */
long fdc_chan_blockcount(const fdc_chan* c);      /* blocks consumed so far (phase counter origin) */
int fdc_chan_reset(fdc_chan* c);                  /* zero history, block counter 0 */
/* Position a fresh/reset context in the middle of a stream (time sharding): the next block is
 * global block `first_block`; history is NOT touched (feed the halo with fdc_chan_set_history). */
int fdc_chan_seek(fdc_chan* c, long first_block);
int fdc_chan_set_history(fdc_chan* c, const void* host_last_ovl_samples);
/* in: nblocks*hop new samples.  outs[i]: nblocks*lout_i samples for channel i (NULL entries are skipped).
 * spectrum: optional nblocks*N normalised fft-shifted spectrum (the hier block's debug port). */
int fdc_chan_work_host(fdc_chan* c, const void* in, long nblocks, void* const* outs, void* spectrum);
/* Device-resident variant.  d_out: one buffer, channel i starts at item offset
 * sum_{j<i} nblocks*lout_j ("channel-major slabs").  d_spectrum may be NULL (an internal L2-sized
 * ring is used then).  Only enqueues on `stream`; history/counters are advanced. */
int fdc_chan_work_device(fdc_chan* c, const void* d_in, long nblocks, void* d_out, void* d_spectrum, void* stream);
/* The same with an explicit placement of the outputs: channel i occupies slab_blocks * lout_i items starting at item offset
 * slab_blocks * sum_{j<i} lout_j, and this call's blocks go to rows [slab_first_block, slab_first_block + nblocks) of every
 * channel's slab.  With d_out in ANOTHER GPU's memory (fdc_ipc_open) the extract kernel's stores travel over NVLink: the
 * time-sharded ranks write one stream-ordered buffer on the sink rank and no separate gather is needed. */
int fdc_chan_work_device_slab(fdc_chan* c, const void* d_in, long nblocks, void* d_out, long slab_blocks, long slab_first_block,
                              void* d_spectrum, void* stream);
/* CUDA IPC plumbing for that (one process per GPU): export a buffer made by fdc_dev_alloc as a 64-byte handle, map it in
 * another process (peer access is enabled on demand), unmap. */
int fdc_ipc_export(const void* d_ptr, void* handle64);
void* fdc_ipc_open(const void* handle64);
int fdc_ipc_close(void* d_ptr);
/* Channel-sharded sinks for a time-sharded stream on several GPUs (the reference has one flowgraph and one sink
 * per channel, python/FrequencyDomainChannelizer.py:312 -- with N GPUs the channels are dealt out over the ranks, each rank
 * feeding the sinks of its own channels).  d_bases[k] is sink k's buffer as seen from THIS GPU (local, or fdc_ipc_open of the
 * owner's export); owner[i] in [0, nsinks) names the sink of channel i.  Sink k holds only its channels, channel-major:
 * slab_blocks * lout_i items per channel in channel order, this call's blocks in rows [slab_first_block, +nblocks).  The
 * extract kernel stores every channel's rows straight into its owner's memory (peer stores over NVLink for remote owners):
 * an all-to-all that overlaps the transforms, every GPU receiving 1/N of each rank's output.  local_sink: the index of the
 * sink that lies in this GPU's own memory (-1: none; with FDC_SINK_DMA=1 the rows of all OTHER sinks are written to a local
 * staging slab and forwarded by the copy engines behind the kernels instead of being stored by the SMs). */
int fdc_chan_set_sinks(fdc_chan* c, int nsinks, void* const* d_bases, const int* owner, int local_sink);
int fdc_chan_work_device_sinks(fdc_chan* c, const void* d_in, long nblocks, long slab_blocks, long slab_first_block, void* stream);
/* The hier block's inpveclen > 1 mode (python/FrequencyDomainChannelizer.py:284-290): the input items are vectors that are
 * already transformed (fft-shifted, unnormalised spectra of N bins).  Runs normalize_input (x 1/N, :216) and the
 * per-channel chains; overlap-save and the forward FFT are skipped.  d_spectrum (optional) receives the normalised
 * spectra for the activity-gated blocks / the debug port. */
int fdc_chan_work_spectrum_device(fdc_chan* c, const void* d_spectra, long nblocks, void* d_out, void* d_spectrum, void* stream);
/* 1 when calls that do not ask for the spectrum run overlap-save, forward FFT and all channels as ONE kernel (N <= 16384, every
 * channel of the same slice length; the spectrum stays in shared memory) */
int fdc_chan_is_fused(const fdc_chan* c);
int fdc_chan_sync(fdc_chan* c);
/* Measurement hooks.  Profiling records CUDA events around the forward-FFT and the channel-extract kernels of every
 * chunk; get_profile synchronises and returns the summed kernel times (ms) since the last call.  chunk_blocks is
 * the number of blocks per K1 -> K2 round trip (the spectrum ring that is meant to stay resident in L2). */
int fdc_chan_set_profiling(fdc_chan* c, int enable);
int fdc_chan_get_profile(fdc_chan* c, double* ms_fwd, double* ms_extract, long* chunks);
int fdc_chan_chunk_blocks(const fdc_chan* c);
int fdc_chan_set_chunk_blocks(fdc_chan* c, int blocks);

/* ---- individual block replacements (host buffers in, host buffers out) ---------------------- */
/* FDC.overlap_save, include/FDC/overlap_save.h:49, lib/overlap_save_impl.cc:62-81 */
typedef struct fdc_overlap_save fdc_overlap_save;
fdc_overlap_save* fdc_overlap_save_create(int itemsize, int outputlen, int overlaplen);
int fdc_overlap_save_work(fdc_overlap_save* b, int noutput_items, const void* in, void* out);
void fdc_overlap_save_destroy(fdc_overlap_save* b);

/* FDC.vector_cut_vxx, include/FDC/vector_cut_vxx.h:49, lib/vector_cut_vxx_impl.cc:59-72 */
typedef struct fdc_vector_cut fdc_vector_cut;
fdc_vector_cut* fdc_vector_cut_create(int itemsize, int veclen, int offset, int blocklen);
int fdc_vector_cut_work(fdc_vector_cut* b, int noutput_items, const void* in, void* out);
void fdc_vector_cut_destroy(fdc_vector_cut* b);

/* FDC.phase_shifting_windowing_vcc, include/FDC/phase_shifting_windowing_vcc.h:49,
 * lib/phase_shifting_windowing_vcc_impl.cc:41-86 */
typedef struct fdc_psw fdc_psw;
fdc_psw* fdc_psw_create(int blocklen, int numphasestates, int shifts, float passbw, float stopbw, int windowtype);
int fdc_psw_work(fdc_psw* b, int noutput_items, const void* in, void* out);
int fdc_psw_state(const fdc_psw* b, int* blocksize, int* relinvovl, int* counter, int* shift);
int fdc_psw_tables(const fdc_psw* b, float* out);   /* relinvovl*blocklen complex floats */
void fdc_psw_destroy(fdc_psw* b);

/* gr::fft::fft_vcc(n, forward, rectangular window, shift) as the hier block uses it
 * (python/FrequencyDomainChannelizer.py:206, 228); third-party stage, provided so that a flowgraph
 * can stay on the GPU library end to end. */
typedef struct fdc_fft fdc_fft;
fdc_fft* fdc_fft_create(int n, int forward, int shift);
int fdc_fft_work(fdc_fft* b, long nvec, const void* in, void* out);
void fdc_fft_destroy(fdc_fft* b);

/* ---- decimated power rows for a waterfall display --------------------------------------------
 * Replaces complex_to_mag_squared (+ nlog10_ff for a dB display) and the column reduction at the input of the reference's
 * waterfall consumer (python/WaterfallMsgTagging.py:272-277: every blocklen vector is reduced to 1024 columns by a mean, or
 * repeated when blocklen < 1024).  Input: spectrum rows (blocklen complex floats each); output: `width` floats per block:
 * logmode 0 -> mean of |X|^2 over blocklen / width bins, logmode 1 -> mean of 10 log10 |X|^2. */
typedef struct fdc_waterfall fdc_waterfall;
fdc_waterfall* fdc_waterfall_create(int blocklen, int width, int logmode);
int fdc_waterfall_work_host(fdc_waterfall* b, int nblocks, const void* spectrum, float* out_host);
int fdc_waterfall_work_device(fdc_waterfall* b, int nblocks, const void* d_spectrum, float* out_host, void* stream);
void fdc_waterfall_destroy(fdc_waterfall* b);

/* ---- activity-gated channels: PDUs ---------------------------------------------------------- */
/* One published message: the pmt dict of lib/SegmentDetection_impl.cc:446-460,502-515 /
 * lib/PowerActivationChannel_impl.cc:222-232 plus the c32vector payload. */
typedef struct {
    char id[160];          /* "<time>.PowActChan.<ID>.<n>.fin|.part" or "<time>.DETECTED.<seg>.<chan>" */
    int finalized;
    long part;             /* -1: key absent */
    double rel_cfreq, rel_bw;
    long blockstart, blockend;
    long vectorstart, vectorend;   /* -1: key absent (PowerActivationChannel) */
    long nsamples;
    const float* data;     /* nsamples complex floats, owned by the block until the next msg_clear / destroy.  A burst extracted and
                            * published in one call is a view of that call's pinned result buffer (no copy); re-fetch the messages
                            * (msg_get) after a later work() call if they were not cleared: uncollected views are moved then */
} fdc_msg;

/* FDC.PowerActivationChannel, include/FDC/PowerActivationChannel.h:49, lib/PowerActivationChannel_impl.cc */
typedef struct fdc_pac fdc_pac;
fdc_pac* fdc_pac_create(int blocklen, float cfreq, float bw, int relinvovl, float thresh, int maxblocks,
                        int deactivation_delay, int msg, int fileoutput, const char* path, int verbose, int ID);
int fdc_pac_work_host(fdc_pac* b, int noutput_items, const void* in);
int fdc_pac_work_device(fdc_pac* b, int noutput_items, const void* d_in, void* stream);
/* geo[12]: extract_start, extract_stop, extract_width, measure_start, measure_stop, deltaphase, output_len,
 * output_ovl_offset, active, count, phase, blockcount; f[2]: thresh (linear), lastpower */
int fdc_pac_state(const fdc_pac* b, int* geo, float* f);
int fdc_pac_tables(const fdc_pac* b, float* out);    /* relinvovl * blocklen complex floats */
int fdc_pac_msg_count(const fdc_pac* b);
int fdc_pac_msg_get(const fdc_pac* b, int i, fdc_msg* out);
/* all pending messages at once: out[fdc_pac_msg_count()] records; payloads of all messages concatenated (NULL: only count) */
int fdc_pac_msg_get_all(const fdc_pac* b, fdc_msg* out);
long fdc_pac_msg_copy_data(const fdc_pac* b, float* out);
void fdc_pac_msg_clear(fdc_pac* b);
void fdc_pac_destroy(fdc_pac* b);

/* FDC.SegmentDetection, include/FDC/SegmentDetection.h:49, lib/SegmentDetection_impl.cc */
typedef struct fdc_segdet fdc_segdet;
fdc_segdet* fdc_segdet_create(int ID, int blocklen, int relinvovl, float seg_start, float seg_stop, float thresh,
                              float minchandist, float window_flank_puffer, int maxblocks_to_emit,
                              int channel_deactivation_delay, int messageoutput, int fileoutput, const char* path,
                              int threads, int verbose);
int fdc_segdet_work_host(fdc_segdet* b, int noutput_items, const void* in);
int fdc_segdet_work_device(fdc_segdet* b, int noutput_items, const void* d_in, void* stream);
/* geo[8]: d_start, d_stop, d_width, D, M, blockcount, n_active, active_channels_counter; f[1]: thresh */
int fdc_segdet_state(const fdc_segdet* b, long* geo, float* f);
int fdc_segdet_window(const fdc_segdet* b, int log2w, int phase, float* out);
int fdc_segdet_power(const fdc_segdet* b, float* out);          /* decimated power of the last block, M floats */
/* active channel i, out[14]: ID, detect_start, detect_stop, extract_start, extract_stop, extract_width, ovlskip,
 * outputsamples, count, phase, phaseincrement, inactive, part, buffered blocks */
int fdc_segdet_active(const fdc_segdet* b, int i, int* out);
int fdc_segdet_msg_count(const fdc_segdet* b);
int fdc_segdet_msg_get(const fdc_segdet* b, int i, fdc_msg* out);
/* all pending messages at once: out[fdc_segdet_msg_count()] records; payloads of all messages concatenated (NULL: only count) */
int fdc_segdet_msg_get_all(const fdc_segdet* b, fdc_msg* out);
long fdc_segdet_msg_copy_data(const fdc_segdet* b, float* out);
void fdc_segdet_msg_clear(fdc_segdet* b);
void fdc_segdet_destroy(fdc_segdet* b);

/* FDC.activity_detection_channelizer_vcm, include/FDC/activity_detection_channelizer_vcm.h:49,
 * lib/activity_detection_channelizer_vcm_impl.cc.  segments: nsegs pairs (start, stop). */
typedef struct fdc_actdet fdc_actdet;
fdc_actdet* fdc_actdet_create(int blocklen, const float* segments, int nsegs, float thresh, int relinvovl,
                              int maxblocks, int message, int fileoutput, const char* path, int threads,
                              float minchandist, int channel_deactivation_delay, double window_flank_puffer,
                              int verbose);
int fdc_actdet_work_host(fdc_actdet* b, int noutput_items, const void* in);
int fdc_actdet_work_device(fdc_actdet* b, int noutput_items, const void* d_in, void* stream);
int fdc_actdet_nsegments(const fdc_actdet* b);
/* out[7]: ID, start, stop, width, D, M, n_active */
int fdc_actdet_segment(const fdc_actdet* b, int i, int* out);
int fdc_actdet_power(const fdc_actdet* b, int seg, float* out);
int fdc_actdet_msg_count(const fdc_actdet* b);
int fdc_actdet_msg_get(const fdc_actdet* b, int i, fdc_msg* out);
/* all pending messages at once: out[fdc_actdet_msg_count()] records; payloads of all messages concatenated (NULL: only count) */
int fdc_actdet_msg_get_all(const fdc_actdet* b, fdc_msg* out);
long fdc_actdet_msg_copy_data(const fdc_actdet* b, float* out);
void fdc_actdet_msg_clear(fdc_actdet* b);
void fdc_actdet_destroy(fdc_actdet* b);

/* ---- host-logic hooks -------------------------------------------------------------------------
 * Contexts made by *_create_logic run ONLY the host-side bookkeeping of the activity-gated blocks (argument checks,
 * geometry, window tables, edge pairing, channel matching, PDU metadata) and need no CUDA device.  They exist so that the
 * CPU test-suite can compare that logic with the reference; *_work_host / *_work_device refuse them, and the messages they
 * produce carry metadata and sample counts but no samples (data == NULL).  Input is what K3 would have measured:
 * one band power per block (PowerActivationChannel) or the decimated power rows (M floats per block and segment). */
fdc_pac* fdc_pac_create_logic(int blocklen, float cfreq, float bw, int relinvovl, float thresh, int maxblocks,
                              int deactivation_delay, int msg, int fileoutput, const char* path, int verbose, int ID);
int fdc_pac_logic_work(fdc_pac* b, int nblocks, const float* band_power);
fdc_segdet* fdc_segdet_create_logic(int ID, int blocklen, int relinvovl, float seg_start, float seg_stop, float thresh,
                                    float minchandist, float window_flank_puffer, int maxblocks_to_emit,
                                    int channel_deactivation_delay, int messageoutput, int fileoutput, const char* path,
                                    int threads, int verbose);
int fdc_segdet_logic_work(fdc_segdet* b, int nblocks, const float* power_rows);
fdc_actdet* fdc_actdet_create_logic(int blocklen, const float* segments, int nsegs, float thresh, int relinvovl,
                                    int maxblocks, int message, int fileoutput, const char* path, int threads,
                                    float minchandist, int channel_deactivation_delay, double window_flank_puffer,
                                    int verbose);
int fdc_actdet_logic_work(fdc_actdet* b, int nblocks, const float* power_rows);

/* ---- time-sharded calls of the activity-gated blocks (one process per GPU) ----------------------
 * The reference blocks carry state from block to block (active flag and last power, PowerActivationChannel_impl.cc:137-177;
 * the active-channel list and its inactivity counters, SegmentDetection_impl.cc:131-161,163-300), their measurements do not.
 * A call over nblocks_total blocks of the stream, cut into contiguous runs of rows, one run per rank:
 *   1. every rank:  n = *_shard_measure(b, nrows, d_rows, stream); *_shard_blob(b, buf)   compact record of ITS rows (bytes)
 *   2. the records of all ranks are concatenated in row order (an all-gather of a few bytes per block, host plumbing)
 *   3. every rank:  *_shard_decide(b, nblocks_total, records, bytes)   the sequential bookkeeping, replicated: every rank derives
 *                   the same extraction jobs and message list; returns the number of jobs
 *   4. every rank:  *_shard_extract(b, first_row, nrows, d_rows, d_prev, stream, out)   runs the jobs its rows emitted (a
 *                   contiguous range of the job list) and writes *_shard_samples(b, first_row, nrows) complex floats to the
 *                   host buffer `out`, in job order.  d_prev: device pointer to the spectrum row before first_row (a freshly
 *                   activated channel also takes the previous block); NULL = all-zero row (start of the stream)
 *   5. sink rank:   *_shard_assemble(b, results, nsamples) with the ranks' outputs concatenated in rank order -> messages
 *                   (fdc_*_msg_*); every other rank passes NULL, which only closes the call.
 * Device form of steps 4 and 5 (the gather fused into the extract kernel): *_shard_extract_device stores the samples at d_dst,
 * which may be the sink rank's buffer mapped over CUDA IPC (fdc_ipc_open) -- the kernel's stores then cross NVLink -- and returns
 * after its stream has drained; after a barrier the sink calls *_shard_assemble_device on its buffer.
 * *_shard_layout(b, 1) between steps 3 and 4 switches the device form to ONE run per block instance for all ranks, laid out
 * channel by channel (every rank derives the same offsets from the same job list): d_dst / d_results are then the start of that
 * run on the sink, and bursts that begin and end inside the call are published as views of the assembled buffer instead of
 * being copied block by block.  Returns the samples of the whole call.
 * All ranks must make the same sequence of calls on contexts built with the same arguments.  *_shard_measure_logic is the
 * host-logic form of step 1 (contexts from *_create_logic; input as for *_logic_work), for CPU tests of the plumbing. */
long fdc_pac_shard_measure(fdc_pac* b, int nrows, const void* d_rows, void* stream);
long fdc_pac_shard_measure_logic(fdc_pac* b, int nrows, const float* power);
int fdc_pac_shard_blob(const fdc_pac* b, void* out);
long fdc_pac_shard_decide(fdc_pac* b, int nblocks_total, const void* records, long bytes);
long fdc_pac_shard_samples(const fdc_pac* b, int first_row, int nrows);
long fdc_pac_shard_layout(fdc_pac* b, int by_channel);
long fdc_pac_shard_extract(fdc_pac* b, int first_row, int nrows, const void* d_rows, const void* d_prev, void* stream, void* out_host);
int fdc_pac_shard_assemble(fdc_pac* b, const void* results, long nsamples);
long fdc_pac_shard_extract_device(fdc_pac* b, int first_row, int nrows, const void* d_rows, const void* d_prev, void* stream, void* d_dst);
int fdc_pac_shard_assemble_device(fdc_pac* b, const void* d_results, long nsamples, void* stream);
long fdc_segdet_shard_measure(fdc_segdet* b, int nrows, const void* d_rows, void* stream);
long fdc_segdet_shard_measure_logic(fdc_segdet* b, int nrows, const float* power);
int fdc_segdet_shard_blob(const fdc_segdet* b, void* out);
long fdc_segdet_shard_decide(fdc_segdet* b, int nblocks_total, const void* records, long bytes);
long fdc_segdet_shard_samples(const fdc_segdet* b, int first_row, int nrows);
long fdc_segdet_shard_layout(fdc_segdet* b, int by_channel);
long fdc_segdet_shard_extract(fdc_segdet* b, int first_row, int nrows, const void* d_rows, const void* d_prev, void* stream, void* out_host);
int fdc_segdet_shard_assemble(fdc_segdet* b, const void* results, long nsamples);
long fdc_segdet_shard_extract_device(fdc_segdet* b, int first_row, int nrows, const void* d_rows, const void* d_prev, void* stream, void* d_dst);
int fdc_segdet_shard_assemble_device(fdc_segdet* b, const void* d_results, long nsamples, void* stream);
long fdc_actdet_shard_measure(fdc_actdet* b, int nrows, const void* d_rows, void* stream);
long fdc_actdet_shard_measure_logic(fdc_actdet* b, int nrows, const float* power);
int fdc_actdet_shard_blob(const fdc_actdet* b, void* out);
long fdc_actdet_shard_decide(fdc_actdet* b, int nblocks_total, const void* records, long bytes);
long fdc_actdet_shard_samples(const fdc_actdet* b, int first_row, int nrows);
long fdc_actdet_shard_layout(fdc_actdet* b, int by_channel);
long fdc_actdet_shard_extract(fdc_actdet* b, int first_row, int nrows, const void* d_rows, const void* d_prev, void* stream, void* out_host);
int fdc_actdet_shard_assemble(fdc_actdet* b, const void* results, long nsamples);
long fdc_actdet_shard_extract_device(fdc_actdet* b, int first_row, int nrows, const void* d_rows, const void* d_prev, void* stream, void* d_dst);
int fdc_actdet_shard_assemble_device(fdc_actdet* b, const void* d_results, long nsamples, void* stream);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif
