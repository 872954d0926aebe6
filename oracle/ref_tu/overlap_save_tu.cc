/* Oracle TU: unmodified /root/reference/lib/overlap_save_impl.cc */
#include "ref_common.h"
#define private public
#include "overlap_save_impl.cc"
#undef private
extern "C" gr::sync_block* ref_overlap_save_make(int itemsize, int outputlen, int overlaplen)
{ REF_TRY return new gr::FDC::overlap_save_impl(itemsize, outputlen, overlaplen); REF_CATCH(0) }
