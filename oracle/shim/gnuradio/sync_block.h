/* Oracle shim (TEST INFRASTRUCTURE ONLY): minimal gr::sync_block.  The real class
 * lives in gnuradio-runtime (third-party, absent).  Only what lib/*_impl.cc touches:
 * ctor(name, in_sig, out_sig), message_port_register_out, message_port_pub, work(). */
#ifndef FDC_SHIM_GR_SYNC_BLOCK_H
#define FDC_SHIM_GR_SYNC_BLOCK_H
#include <gnuradio/io_signature.h>
#include <pmt/pmt.h>
#include <mutex>
typedef std::complex<float> gr_complex;
typedef std::vector<const void*> gr_vector_const_void_star;
typedef std::vector<void*> gr_vector_void_star;
namespace gr {
class sync_block {
public:
    std::string d_name;
    io_signature::sptr d_in_sig, d_out_sig;
    std::vector<pmt::pmt_t> d_published;   /* captured PDUs, in publication order */
    std::mutex d_pub_mutex;                /* threaded reference modes publish from workers */
    sync_block() {}
    sync_block(const std::string& name, io_signature::sptr in, io_signature::sptr out)
        : d_name(name), d_in_sig(in), d_out_sig(out) {}
    virtual ~sync_block() {}
    void message_port_register_out(pmt::pmt_t) {}
    void message_port_pub(pmt::pmt_t, pmt::pmt_t msg)
    { std::lock_guard<std::mutex> g(d_pub_mutex); d_published.push_back(msg); }
    virtual int work(int noutput_items, gr_vector_const_void_star& input_items,
                     gr_vector_void_star& output_items) = 0;
};
}
namespace gnuradio {
template <class T> std::shared_ptr<T> get_initial_sptr(T* p) { return std::shared_ptr<T>(p); }
}
#endif
