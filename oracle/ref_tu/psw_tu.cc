/* Oracle TU: unmodified /root/reference/lib/phase_shifting_windowing_vcc_impl.cc (+ lib/windows.h) */
#include "ref_common.h"
#define private public
#include "phase_shifting_windowing_vcc_impl.cc"
#undef private
using gr::FDC::phase_shifting_windowing_vcc_impl;
extern "C" gr::sync_block* ref_psw_make(int blocklen, int numphasestates, int shifts, float passbw, float stopbw, int windowtype)
{ REF_TRY return new phase_shifting_windowing_vcc_impl(blocklen, numphasestates, shifts, passbw, stopbw, windowtype); REF_CATCH(0) }
/* table[i][k] as interleaved floats, R*blocklen*2 */
extern "C" int ref_psw_tables(gr::sync_block* b, float* out)
{
    phase_shifting_windowing_vcc_impl* p = dynamic_cast<phase_shifting_windowing_vcc_impl*>(b);
    if (!p) return -1;
    for (int i = 0; i < p->relinvovl; i++)
        memcpy(out + (size_t)2 * i * p->blocksize, p->windows[i].data(), sizeof(gr_complex) * p->blocksize);
    return 0;
}
/* state[0..3] = blocksize, relinvovl, counter, shift */
extern "C" int ref_psw_state(gr::sync_block* b, int* st)
{
    phase_shifting_windowing_vcc_impl* p = dynamic_cast<phase_shifting_windowing_vcc_impl*>(b);
    if (!p) return -1;
    st[0] = p->blocksize; st[1] = p->relinvovl; st[2] = p->counter; st[3] = p->shift; return 0;
}
