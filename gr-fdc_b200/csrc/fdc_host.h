/* fdc_host.h -- host-side geometry / table builders (see fdc_host.cc). */
#ifndef FDC_HOST_H
#define FDC_HOST_H
#include <complex>
#include <vector>
namespace fdc {
void opt_channelparams(int blocksize, int relinvovl, double freq, double bw, int* f, int* l, int* lout, double* passband,
                       double* stopband);
void psw_tables(int blocksize, int relinvovl, float passbw, float stopbw, int wintype, std::vector<std::complex<float> >& out);
void psw_check_args(float passbw, float stopbw);
int nextpow2_int(double v);
float db_to_ratio(float db);
}
#endif
