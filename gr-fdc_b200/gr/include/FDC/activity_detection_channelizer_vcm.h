/* FDC/activity_detection_channelizer_vcm.h -- gr::FDC::activity_detection_channelizer_vcm, see FDC/fdc_blocks.h */
#include <FDC/fdc_blocks.h>
