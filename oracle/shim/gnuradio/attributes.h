/* Oracle shim (TEST INFRASTRUCTURE ONLY): stands in for <gnuradio/attributes.h>
 * so the unmodified gr-FDC sources under /root/reference compile without GNU Radio. */
#ifndef FDC_SHIM_GR_ATTRIBUTES_H
#define FDC_SHIM_GR_ATTRIBUTES_H
#define __GR_ATTR_EXPORT __attribute__((visibility("default")))
#define __GR_ATTR_IMPORT __attribute__((visibility("default")))
#endif
