/* FDC/SegmentDetection.h -- gr::FDC::SegmentDetection, see FDC/fdc_blocks.h */
#include <FDC/fdc_blocks.h>
