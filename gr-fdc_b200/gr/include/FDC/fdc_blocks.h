/* FDC/fdc_blocks.h -- public block classes of the B200 drop-in, namespace gr::FDC.
 *
 * Same class names, base class (virtual gr::sync_block), sptr typedef and `make` parameter lists as the
 * reference's include/FDC/<block>.h:49, so that existing C++ flowgraphs and the SWIG interface
 * (swig/FDC_swig.i) bind unchanged.  The per-block headers FDC/<block>.h only include this file.
 * The implementations (gr/lib/fdc_gr_blocks.cc) hold a context of libfdc_b200.so (include/fdc_cabi.h)
 * and forward work() to it; constructors re-throw the library's error text as std::invalid_argument,
 * which is what the reference constructors throw. */
#ifndef INCLUDED_FDC_BLOCKS_H
#define INCLUDED_FDC_BLOCKS_H
#include <FDC/api.h>
#include <gnuradio/sync_block.h>
#include <string>
#include <vector>

namespace gr {
namespace FDC {

#define FDC_DECLARE_BLOCK(NAME, ...)                                   \
    class FDC_API NAME : virtual public gr::sync_block {               \
    public:                                                            \
        typedef boost::shared_ptr<NAME> sptr;                          \
        static sptr make(__VA_ARGS__);                                 \
    };

/* include/FDC/overlap_save.h:49 */
FDC_DECLARE_BLOCK(overlap_save, int itemsize, int outputlen, int overlaplen)
/* include/FDC/vector_cut_vxx.h:49 */
FDC_DECLARE_BLOCK(vector_cut_vxx, int itemsize, int veclen, int offset, int blocklen)
/* include/FDC/phase_shifting_windowing_vcc.h:49 */
FDC_DECLARE_BLOCK(phase_shifting_windowing_vcc, int blocklen, int numphasestates, int shifts, float passbw, float stopbw, int windowtype)
/* include/FDC/PowerActivationChannel.h:49 */
FDC_DECLARE_BLOCK(PowerActivationChannel, int v_blocklen, float v_cfreq, float v_bw, int v_relinvovl, float v_thresh, int v_maxblocks,
                  int v_deactivation_delay, bool v_msg, bool v_fileoutput, std::string v_path, int verbose, int v_ID)
/* include/FDC/SegmentDetection.h:49 */
FDC_DECLARE_BLOCK(SegmentDetection, int ID, int blocklen, int relinvovl, float seg_start, float seg_stop, float thresh, float minchandist,
                  float window_flank_puffer, int maxblocks_to_emit, int channel_deactivation_delay, bool messageoutput, bool fileoutput,
                  std::string path, bool threads, int verbose)
/* include/FDC/activity_detection_channelizer_vcm.h:49 */
FDC_DECLARE_BLOCK(activity_detection_channelizer_vcm, int v_blocklen, std::vector<std::vector<float> > v_segments, float v_thresh,
                  int v_relinvovl, int v_maxblocks, bool v_message, bool v_fileoutput, std::string v_path, bool v_threads,
                  float v_minchandist, int v_channel_deactivation_delay, double v_window_flank_puffer, int verbose)

#undef FDC_DECLARE_BLOCK

}  // namespace FDC
}  // namespace gr
#endif
