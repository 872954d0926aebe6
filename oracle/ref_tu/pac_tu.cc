/* Oracle TU: unmodified /root/reference/lib/PowerActivationChannel_impl.cc */
#include "ref_common.h"
#define private public
#include "PowerActivationChannel_impl.cc"
#undef private
using gr::FDC::PowerActivationChannel_impl;
extern "C" gr::sync_block* ref_pac_make(int blocklen, float cfreq, float bw, int relinvovl, float thresh, int maxblocks,
                                        int deactivation_delay, int msg, int fileoutput, const char* path, int verbose, int ID)
{ REF_TRY return new PowerActivationChannel_impl(blocklen, cfreq, bw, relinvovl, thresh, maxblocks, deactivation_delay,
                                                  msg != 0, fileoutput != 0, std::string(path ? path : ""), verbose, ID); REF_CATCH(0) }
/* geo: extract_start, extract_stop, extract_width, measure_start, measure_stop, deltaphase, output_len,
 *      output_ovl_offset, active, count, phase, blockcount ; f: thresh, lastpower */
extern "C" int ref_pac_state(gr::sync_block* b, int* geo, float* f)
{
    PowerActivationChannel_impl* p = dynamic_cast<PowerActivationChannel_impl*>(b);
    if (!p) return -1;
    geo[0] = p->extract_start; geo[1] = p->extract_stop; geo[2] = p->extract_width; geo[3] = p->measure_start;
    geo[4] = p->measure_stop; geo[5] = p->deltaphase; geo[6] = p->output_len; geo[7] = p->output_ovl_offset;
    geo[8] = p->active ? 1 : 0; geo[9] = p->count; geo[10] = p->phase; geo[11] = p->blockcount;
    f[0] = p->thresh; f[1] = p->lastpower; return 0;
}
/* windows[i] has length blocklen; out: R*blocklen*2 floats */
extern "C" int ref_pac_tables(gr::sync_block* b, float* out)
{
    PowerActivationChannel_impl* p = dynamic_cast<PowerActivationChannel_impl*>(b);
    if (!p) return -1;
    for (int i = 0; i < p->relinvovl; i++)
        memcpy(out + (size_t)2 * i * p->blocklen, p->windows[i].data(), sizeof(gr_complex) * p->blocklen);
    return 0;
}
