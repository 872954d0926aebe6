/* Oracle TU: unmodified /root/reference/lib/vector_cut_vxx_impl.cc */
#include "ref_common.h"
#define private public
#include "vector_cut_vxx_impl.cc"
#undef private
extern "C" gr::sync_block* ref_vector_cut_make(int itemsize, int veclen, int offset, int blocklen)
{ REF_TRY return new gr::FDC::vector_cut_vxx_impl(itemsize, veclen, offset, blocklen); REF_CATCH(0) }
