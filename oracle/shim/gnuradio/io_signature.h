/* Oracle shim (TEST INFRASTRUCTURE ONLY): minimal gr::io_signature. */
#ifndef FDC_SHIM_GR_IO_SIGNATURE_H
#define FDC_SHIM_GR_IO_SIGNATURE_H
#include <memory>
#include <cstring>
#include <string>
#include <vector>
#include <complex>
#include <stdexcept>
namespace boost { using std::shared_ptr; }
namespace gr {
class io_signature {
public:
    typedef std::shared_ptr<io_signature> sptr;
    int d_min, d_max, d_itemsize;
    io_signature(int mn, int mx, int sz) : d_min(mn), d_max(mx), d_itemsize(sz) {}
    static sptr make(int min_streams, int max_streams, int sizeof_stream_item)
    { return sptr(new io_signature(min_streams, max_streams, sizeof_stream_item)); }
    int sizeof_stream_item(int) const { return d_itemsize; }
};
}
#endif
