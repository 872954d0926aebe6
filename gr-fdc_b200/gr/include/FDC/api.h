/* FDC/api.h -- symbol visibility of the B200 drop-in for gr-FDC's C++ block library. */
#ifndef INCLUDED_FDC_API_H
#define INCLUDED_FDC_API_H
#include <gnuradio/attributes.h>
#ifdef gnuradio_FDC_EXPORTS
#define FDC_API __GR_ATTR_EXPORT
#else
#define FDC_API __GR_ATTR_IMPORT
#endif
#endif
