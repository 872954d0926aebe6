/* Oracle shim (TEST INFRASTRUCTURE ONLY): SegmentDetection_impl.h includes this but never uses it. */
#ifndef FDC_SHIM_BOOST_LEXICAL_CAST_HPP
#define FDC_SHIM_BOOST_LEXICAL_CAST_HPP
#endif
