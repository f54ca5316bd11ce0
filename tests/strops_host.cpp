// Host shim over tuplex_b200/csrc/strops.cuh: the very functions the CUDA VM calls, compiled for the CPU so
// that tests/test_strops_host.py can fuzz them against CPython. Test infrastructure only.
#include "../tuplex_b200/csrc/strops.cuh"
#include <string.h>
using namespace tplx;
static StrV mk(const uint8_t *base, uint32_t off, uint32_t len, uint32_t flags) { StrV s; s.p = base + off; s.len = len; s.flags = flags; return s; }
extern "C" {
long long h_find(const uint8_t *hb, unsigned ho, unsigned hl, unsigned hf, const uint8_t *nb, unsigned no, unsigned nl, unsigned nf) {
    return str_find(mk(hb, ho, hl, hf), mk(nb, no, nl, nf)); }
long long h_rfind(const uint8_t *hb, unsigned ho, unsigned hl, unsigned hf, const uint8_t *nb, unsigned no, unsigned nl, unsigned nf) {
    return str_rfind(mk(hb, ho, hl, hf), mk(nb, no, nl, nf)); }
int h_eq(const uint8_t *hb, unsigned ho, unsigned hl, unsigned hf, const uint8_t *nb, unsigned no, unsigned nl, unsigned nf) {
    return str_eq(mk(hb, ho, hl, hf), mk(nb, no, nl, nf)); }
int h_atoi(const uint8_t *b, unsigned o, unsigned l, long long *out) { int64_t v = 0; bool ok = str_to_i64(mk(b, o, l, 0), &v); *out = v; return ok; }
void h_copy(uint8_t *dst, unsigned dofs, const uint8_t *sb, unsigned so, unsigned sl, unsigned sf) { str_copy(dst + dofs, mk(sb, so, sl, sf)); }
unsigned h_lower4(unsigned w) { return lower4(w); }
unsigned h_upper4(unsigned w) { return upper4(w); }
long long h_slice_index(long long i, long long n) { return slice_index(i, n); }
long long h_floordiv(long long a, long long b) { return floordiv_i64(a, b); }
long long h_floormod(long long a, long long b) { return floormod_i64(a, b); }
}
