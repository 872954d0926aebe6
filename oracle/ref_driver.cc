/* Oracle (TEST INFRASTRUCTURE ONLY): C driver around the UNMODIFIED gr-FDC blocks built
 * from /root/reference/lib by oracle/Makefile, plus restatements of the third-party
 * GNU Radio 3.7 stages the hier block wires between them
 * (/root/reference/python/FrequencyDomainChannelizer.py:201-231,284-315):
 *   blocks.stream_to_vector / vector_to_stream : reinterpretation of the same bytes
 *   fft.fft_vcc(N, fwd, rect window, shift, nthreads) : window multiply (x1.0f), for the
 *        inverse+shift case the input halves are swapped before the transform, for the
 *        forward+shift case the output halves are swapped after it (gr-fft fft_vcc_fftw::work)
 *   blocks.multiply_const_cc(k, vlen) : complex scale by the real constant k
 * Nothing here is product code; the product never links or loads this library. */
#include "ref_tu/ref_common.h"

static thread_local std::string g_err;
extern "C" void ref_set_error(const char* msg) { g_err = msg ? msg : ""; }
extern "C" const char* ref_last_error(void) { return g_err.c_str(); }

extern "C" {
gr::sync_block* ref_overlap_save_make(int, int, int);
gr::sync_block* ref_vector_cut_make(int, int, int, int);
gr::sync_block* ref_psw_make(int, int, int, float, float, int);
}

/* ---- generic sync_block handling ------------------------------------------------------- */
extern "C" int ref_work(gr::sync_block* b, int nitems, const void* in, void* out)
{
    REF_TRY
    gr_vector_const_void_star iv(1, in); gr_vector_void_star ov(1, out);
    return b->work(nitems, iv, ov);
    REF_CATCH(-1)
}
extern "C" void ref_free(gr::sync_block* b) { delete b; }
extern "C" const char* ref_name(gr::sync_block* b) { return b->d_name.c_str(); }
extern "C" int ref_itemsizes(gr::sync_block* b, int* in_sz, int* out_sz)
{ *in_sz = b->d_in_sig->d_itemsize; *out_sz = b->d_out_sig->d_itemsize; return 0; }

/* ---- captured PDUs ---------------------------------------------------------------------- */
extern "C" int ref_msg_count(gr::sync_block* b) { return (int)b->d_published.size(); }
extern "C" void ref_msg_clear(gr::sync_block* b) { b->d_published.clear(); }
/* ints: finalized, part(-1 = key absent), blockstart, blockend, vectorstart(-1 absent), vectorend(-1), nsamples
 * dbl : rel_bw, rel_cfreq */
extern "C" int ref_msg_meta(gr::sync_block* b, int i, char* id, int idcap, long* ints, double* dbl)
{
    if (i < 0 || i >= (int)b->d_published.size()) return -1;
    const pmt::pmt_t& m = b->d_published[i];
    ints[0] = 0; ints[1] = -1; ints[2] = 0; ints[3] = 0; ints[4] = -1; ints[5] = -1; dbl[0] = dbl[1] = 0.0;
    if (idcap > 0) id[0] = 0;
    for (size_t k = 0; k < m->car->dict.size(); k++) {
        const std::string& key = m->car->dict[k].first; const pmt::pmt_t& v = m->car->dict[k].second;
        if (key == "ID") { strncpy(id, v->sym.c_str(), idcap - 1); id[idcap - 1] = 0; }
        else if (key == "finalized") ints[0] = v->b ? 1 : 0;
        else if (key == "part") ints[1] = v->l;
        else if (key == "blockstart") ints[2] = v->l;
        else if (key == "blockend") ints[3] = v->l;
        else if (key == "vectorstart") ints[4] = v->l;
        else if (key == "vectorend") ints[5] = v->l;
        else if (key == "rel_bw") dbl[0] = v->d;
        else if (key == "rel_cfreq") dbl[1] = v->d;
    }
    ints[6] = (long)m->cdr->c32.size();
    return 0;
}
/* the dict keys of PDU i in insertion order, comma separated (a GNU Radio 3.7 pmt dict is an association list: the order is
 * visible to whoever prints or serialises the message) */
extern "C" int ref_msg_keys(gr::sync_block* b, int i, char* out, int cap)
{
    if (i < 0 || i >= (int)b->d_published.size() || cap < 1) return -1;
    std::string all;
    const pmt::pmt_t& m = b->d_published[i];
    for (size_t k = 0; k < m->car->dict.size(); k++) { if (k) all += ","; all += m->car->dict[k].first; }
    strncpy(out, all.c_str(), cap - 1); out[cap - 1] = 0;
    return 0;
}
extern "C" int ref_msg_data(gr::sync_block* b, int i, float* out)
{
    if (i < 0 || i >= (int)b->d_published.size()) return -1;
    const std::vector<gr_complex>& d = b->d_published[i]->cdr->c32;
    memcpy(out, d.data(), sizeof(gr_complex) * d.size());
    return 0;
}

/* ---- third-party stage restatements ------------------------------------------------------ */
/* fft.fft_vcc with an all-ones window: nvec vectors of length n. */
extern "C" int ref_fft_vcc(int n, int forward, int shift, long nvec, const float* in_f, float* out_f)
{
    REF_TRY
    const gr_complex* in = (const gr_complex*)in_f; gr_complex* out = (gr_complex*)out_f;
    gr::fft::fft_complex f(n, forward != 0, 1);
    const int h = n / 2;
    for (long v = 0; v < nvec; v++, in += n, out += n) {
        gr_complex* dst = f.get_inbuf();
        if (!forward && shift) {
            for (int i = 0; i < h; i++) dst[i + (n - h)] = in[i] * 1.0f;
            for (int i = h; i < n; i++) dst[i - h] = in[i] * 1.0f;
        } else {
            for (int i = 0; i < n; i++) dst[i] = in[i] * 1.0f;
        }
        f.execute();
        if (forward && shift) {
            memcpy(out, f.get_outbuf() + h, sizeof(gr_complex) * (n - h));
            memcpy(out + (n - h), f.get_outbuf(), sizeof(gr_complex) * h);
        } else {
            memcpy(out, f.get_outbuf(), sizeof(gr_complex) * n);
        }
    }
    return 0;
    REF_CATCH(-1)
}
/* blocks.multiply_const_cc(k) : VOLK 32fc_s32fc_multiply with a real-valued complex constant */
static inline void mul_const_cc(gr_complex* out, const gr_complex* in, float k, size_t n)
{
    for (size_t i = 0; i < n; i++) {
        const float a = in[i].real(), b = in[i].imag();
        out[i] = gr_complex(a * k - b * 0.0f, a * 0.0f + b * k);
    }
}
extern "C" int ref_multiply_const_cc(float k, long n, const float* in, float* out)
{ mul_const_cc((gr_complex*)out, (const gr_complex*)in, k, (size_t)n); return 0; }

/* ---- the hier block's throughput flowgraph, same topology, mini scheduler ----------------- */
struct ref_chain {
    int N, R, ovl, hop, nchan;
    gr::sync_block* os;
    std::vector<int> f, l, lout;
    std::vector<gr::sync_block*> cut0, psw, cut3;
    std::vector<gr_complex> spectrum;     /* last run's normalised spectrum (debug port) */
    ~ref_chain() { delete os; for (size_t i = 0; i < cut0.size(); i++) { delete cut0[i]; delete psw[i]; delete cut3[i]; } }
};
extern "C" ref_chain* ref_chain_make(int N, int R, int nchan, const int* f, const int* l, const int* lout,
                                     const float* pbw, const float* sbw, int windowtype)
{
    REF_TRY
    ref_chain* c = new ref_chain; c->N = N; c->R = R; c->ovl = N / R; c->hop = N - N / R; c->nchan = nchan;
    c->os = ref_overlap_save_make(8, N, c->ovl);                                       /* FrequencyDomainChannelizer.py:203 */
    for (int i = 0; i < nchan; i++) {
        c->f.push_back(f[i]); c->l.push_back(l[i]); c->lout.push_back(lout[i]);
        gr::sync_block* a = ref_vector_cut_make(8, N, f[i], l[i]);                      /* :226 */
        gr::sync_block* b = ref_psw_make(l[i], R, f[i], pbw[i], sbw[i], windowtype);    /* :227 */
        gr::sync_block* d = ref_vector_cut_make(8, l[i], l[i] - lout[i], lout[i]);      /* :229 */
        if (!a || !b || !d) { delete c; return (ref_chain*)0; }
        c->cut0.push_back(a); c->psw.push_back(b); c->cut3.push_back(d);
    }
    return c;
    REF_CATCH(0)
}
extern "C" void ref_chain_free(ref_chain* c) { delete c; }

template <class F> static void parallel_for(long n, int nthreads, F fn)
{
    if (nthreads <= 1 || n <= 1) { for (long i = 0; i < n; i++) fn(i); return; }
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; t++)
        th.push_back(std::thread([=]() { for (long i = t; i < n; i += nthreads) fn(i); }));
    for (size_t t = 0; t < th.size(); t++) th[t].join();
}
/* in: nblocks*hop samples; outs[c]: nblocks*lout[c] samples (may be NULL to drop); spectrum_out optional nblocks*N */
extern "C" int ref_chain_run(ref_chain* c, const float* in_f, long nblocks, float** outs, float* spectrum_out, int nthreads)
{
    REF_TRY
    const int N = c->N;
    std::vector<gr_complex> staged((size_t)nblocks * N);
    if (ref_work(c->os, (int)nblocks, in_f, staged.data()) != nblocks) return -1;
    c->spectrum.resize((size_t)nblocks * N);
    gr_complex* S = c->spectrum.data();
    const float invN = (float)(1.0 / (double)N);                                       /* :216 multiply_const_cc(1.0/float(N)) */
    parallel_for(nblocks, nthreads, [&](long b) {
        ref_fft_vcc(N, 1, 1, 1, (const float*)(staged.data() + (size_t)b * N), (float*)(S + (size_t)b * N));
        mul_const_cc(S + (size_t)b * N, S + (size_t)b * N, invN, N);
    });
    if (spectrum_out) memcpy(spectrum_out, S, sizeof(gr_complex) * (size_t)nblocks * N);
    std::vector<int> rc(c->nchan, 0);
    parallel_for(c->nchan, nthreads, [&](long i) {
        const int l = c->l[i], lout = c->lout[i];
        std::vector<gr_complex> t1((size_t)nblocks * l), t2((size_t)nblocks * l);
        if (ref_work(c->cut0[i], (int)nblocks, S, t1.data()) != nblocks) rc[i] = -1;
        if (ref_work(c->psw[i], (int)nblocks, t1.data(), t2.data()) != nblocks) rc[i] = -1;
        ref_fft_vcc(l, 0, 1, nblocks, (const float*)t2.data(), (float*)t1.data());       /* :228 */
        if (outs && outs[i]) {
            gr_complex* o = (gr_complex*)outs[i];
            if (ref_work(c->cut3[i], (int)nblocks, t1.data(), o) != nblocks) rc[i] = -1;
            mul_const_cc(o, o, (float)l, (size_t)nblocks * lout);                        /* :231 blocksize/dec == l */
        }
    });
    for (int i = 0; i < c->nchan; i++) if (rc[i]) return -1;
    return 0;
    REF_CATCH(-1)
}
