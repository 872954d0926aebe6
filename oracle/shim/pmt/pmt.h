/* Oracle shim (TEST INFRASTRUCTURE ONLY): the subset of GNU Radio's pmt that the
 * gr-FDC emit_* functions use (intern, dict, bool/long/double, cons, c32vector). */
#ifndef FDC_SHIM_PMT_H
#define FDC_SHIM_PMT_H
#include <memory>
#include <string>
#include <vector>
#include <complex>
#include <utility>
namespace pmt {
struct pmt_base;
typedef std::shared_ptr<pmt_base> pmt_t;
struct pmt_base {
    enum kind_t { SYMBOL, BOOL, LONG, DOUBLE, DICT, PAIR, C32VEC } kind;
    std::string sym; bool b; long l; double d;
    std::vector<std::pair<std::string, pmt_t> > dict;   /* insertion ordered */
    pmt_t car, cdr;
    std::vector<std::complex<float> > c32;
    pmt_base(kind_t k) : kind(k), b(false), l(0), d(0.0) {}
};
inline pmt_t intern(const std::string& s) { pmt_t p(new pmt_base(pmt_base::SYMBOL)); p->sym = s; return p; }
inline pmt_t from_bool(bool v) { pmt_t p(new pmt_base(pmt_base::BOOL)); p->b = v; return p; }
inline pmt_t from_long(long v) { pmt_t p(new pmt_base(pmt_base::LONG)); p->l = v; return p; }
inline pmt_t from_double(double v) { pmt_t p(new pmt_base(pmt_base::DOUBLE)); p->d = v; return p; }
inline pmt_t make_dict() { return pmt_t(new pmt_base(pmt_base::DICT)); }
inline pmt_t dict_add(const pmt_t& dict, const pmt_t& key, const pmt_t& val)
{ pmt_t p(new pmt_base(*dict)); p->dict.push_back(std::make_pair(key->sym, val)); return p; }
inline pmt_t cons(const pmt_t& a, const pmt_t& b) { pmt_t p(new pmt_base(pmt_base::PAIR)); p->car = a; p->cdr = b; return p; }
inline pmt_t init_c32vector(size_t n, const std::vector<std::complex<float> >& v)
{ pmt_t p(new pmt_base(pmt_base::C32VEC)); p->c32.assign(v.begin(), v.begin() + n); return p; }
inline pmt_t init_c32vector(size_t n, const std::complex<float>* v)
{ pmt_t p(new pmt_base(pmt_base::C32VEC)); p->c32.assign(v, v + n); return p; }
}
#endif
