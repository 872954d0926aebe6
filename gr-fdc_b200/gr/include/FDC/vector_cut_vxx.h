/* FDC/vector_cut_vxx.h -- gr::FDC::vector_cut_vxx, see FDC/fdc_blocks.h */
#include <FDC/fdc_blocks.h>
