/* FDC/PowerActivationChannel.h -- gr::FDC::PowerActivationChannel, see FDC/fdc_blocks.h */
#include <FDC/fdc_blocks.h>
