/* Oracle shim (TEST INFRASTRUCTURE ONLY): gr::fft::fft_complex stand-in.
 * The real class (gnuradio-fft, third-party, absent) wraps an FFTW3f plan:
 * fftwf_plan_dft_1d(size, inbuf, outbuf, forward ? FFTW_FORWARD : FFTW_BACKWARD, FFTW_MEASURE),
 * unnormalised in both directions.  FFTW is absent too, so execute() runs the
 * transforms in oracle/shim/fft_shim.cc:
 *   mode 0 (default, parity anchor): fp64 arithmetic, result rounded once to fp32
 *   mode 1 (timing):                 fp32 Stockham radix-4, cached twiddles
 * Selected process-wide with fdc_shim_set_fft_mode(). */
#ifndef FDC_SHIM_GR_FFT_H
#define FDC_SHIM_GR_FFT_H
#include <complex>
typedef std::complex<float> gr_complex;
extern "C" void fdc_shim_set_fft_mode(int mode);
extern "C" int fdc_shim_get_fft_mode(void);
namespace gr { namespace fft {
class fft_complex {
    int d_size; bool d_forward;
    gr_complex *d_in, *d_out;
public:
    fft_complex(int fft_size, bool forward = true, int nthreads = 1);
    ~fft_complex();
    gr_complex* get_inbuf() const { return d_in; }
    gr_complex* get_outbuf() const { return d_out; }
    int inbuf_length() const { return d_size; }
    int outbuf_length() const { return d_size; }
    void execute();
};
/* free-function form used by the third-party-stage restatements in ref_driver.cc */
void fft_exec(int n, bool forward, const gr_complex* in, gr_complex* out);
}}
#endif
