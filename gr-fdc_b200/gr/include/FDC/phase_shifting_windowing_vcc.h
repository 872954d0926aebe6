/* FDC/phase_shifting_windowing_vcc.h -- gr::FDC::phase_shifting_windowing_vcc, see FDC/fdc_blocks.h */
#include <FDC/fdc_blocks.h>
