/* FDC/overlap_save.h -- gr::FDC::overlap_save, see FDC/fdc_blocks.h */
#include <FDC/fdc_blocks.h>
