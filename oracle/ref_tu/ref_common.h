/* Oracle (TEST INFRASTRUCTURE ONLY): shared declarations of the C driver around the
 * UNMODIFIED gr-FDC sources.  Each reference lib/<x>_impl.cc is unity-included by its own
 * translation unit (lib/windows.h defines non-inline functions, so two TUs must not both
 * include phase_shifting_windowing_vcc_impl.h) with `private` opened for state inspection. */
#ifndef FDC_REF_COMMON_H
#define FDC_REF_COMMON_H
/* every std / shim header the reference pulls in, BEFORE private is opened */
#include <algorithm>
#include <array>
#include <complex>
#include <ctime>
#include <deque>
#include <exception>
#include <iostream>
#include <limits>
#include <map>
#include <memory>
#include <mutex>
#include <sstream>
#include <stdexcept>
#include <string>
#include <thread>
#include <utility>
#include <vector>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <gnuradio/io_signature.h>
#include <gnuradio/sync_block.h>
#include <gnuradio/fft/fft.h>
#include <volk/volk.h>
#include <pmt/pmt.h>
#include <boost/lexical_cast.hpp>

extern "C" void ref_set_error(const char* msg);
#define REF_TRY try {
#define REF_CATCH(ret) } catch (const std::exception& e) { ref_set_error(e.what()); return ret; } catch (...) { ref_set_error("unknown exception"); return ret; }
#endif
