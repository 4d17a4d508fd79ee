// Host-side NNC entropy coder and dependent quantiser behind the C ABI of include/nncabac.h (libnncabac.so).
//
// Stands in for the `deepCABAC` extension module the reference binds (nnc_core/approximator/baseline.py:24-57,89-98;
// nnc_core/coder/baseline.py:5-57; nnc_core/coder/__init__.py:118-140,439-483).  The real module (fraunhoferhhi/nncodec,
// extensions/deepCABAC) is not in /root/reference and cannot be installed here: PARITY UNPINNED.  What is restated, from
// ISO/IEC 15938-17 (NNC) as published:
//   * binarisation of an integer level: sig_flag, sign_flag, a unary run of abs_level_greater_x flags
//     (cabac_unary_length_minus1 + 1 of them), Exp-Golomb remainder with context-coded prefix and bypass suffix;
//   * context selection from the previously coded level (zero / negative / positive) and, under dependent quantisation,
//     from the trellis state;
//   * a two-rate adaptive probability estimator per context with optional per-context rate selection signalled in the
//     stream (param_opt_flag), binary arithmetic coding with 9-bit range, bypass and terminating bins;
//   * dependent quantisation: two scalar quantisers (even multiples / odd multiples and zero of the step size) switched by
//     an 8-state machine driven by the parity of the transmitted index; the encoder searches the trellis (Viterbi).
// Every decision the standard leaves to the encoder (trellis cost, rate selection) only has to be decodable, and every
// syntax choice that could not be checked against the real module is written so that encoder and decoder here agree:
// tests/test_cpu_codec.py drives the UNMODIFIED reference compress -> decompress through this code.
#include "../../include/nncabac.h"

#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <vector>

namespace {

// ------------------------------------------------------------------------------------------------
// step size and uniform quantisation (same arithmetic as csrc/quant.cu)
// ------------------------------------------------------------------------------------------------
inline float stepsize(int qp, int qp_density) {
    const int k = 1 << qp_density;
    const int mul = k + (qp & (k - 1));
    const int shift = (qp >> qp_density) - qp_density;
    return std::ldexp((float)mul, shift);
}

// smallest qp' >= qp whose largest level (+ `slack` union-grid steps) fits int32; non-finite maximum: qp as requested
inline int clip_qp(float max_abs, int qp, int qp_density, float slack) {
    if (!(max_abs <= std::numeric_limits<float>::max())) return qp;
    for (int it = 0; it < 1024; ++it) {
        const float d = stepsize(qp, qp_density);
        volatile float t = max_abs / d;
        const float q = t + 0.5f + slack;
        if (q < 2147483648.0f) return qp;
        ++qp;
    }
    return qp;
}

inline int32_t urq_level(float x, float d) {
    const float a = std::fabs(x);
    if (!(a <= std::numeric_limits<float>::max())) return 0;
    volatile float q = a / d;
    volatile float r = q + 0.5f;
    const int32_t m = r >= 2147483648.0f ? 2147483647 : (int32_t)r;
    return x < 0.0f ? -m : m;
}

// ------------------------------------------------------------------------------------------------
// context model: two exponentially decaying estimates of P(bin = 1), 15-bit, with different adaptation rates
// ------------------------------------------------------------------------------------------------
constexpr int kNumRateSets = 4;
constexpr uint8_t kRate0[kNumRateSets] = {4, 3, 5, 2};
constexpr uint8_t kRate1[kNumRateSets] = {7, 6, 8, 5};
constexpr uint16_t kMask0 = (uint16_t)(~((1u << 5) - 1)) & 0x7fff;    // 10 significant bits
constexpr uint16_t kMask1 = (uint16_t)(~((1u << 1) - 1)) & 0x7fff;    // 14 significant bits

struct Ctx {
    uint16_t s0 = 1 << 14, s1 = 1 << 14;
    uint8_t r0 = kRate0[0], r1 = kRate1[0];
    void reset(int rate_set = 0) { s0 = s1 = 1 << 14; r0 = kRate0[rate_set]; r1 = kRate1[rate_set]; }
    inline unsigned state() const { return (unsigned)(s0 + s1) >> 8; }         // 0..255, P(1) ~ state / 256
    inline unsigned mps() const { return state() >> 7; }
    inline unsigned lps(unsigned range) const {
        unsigned q = state();
        if (q & 0x80) q ^= 0xff;
        return ((q >> 2) * (range >> 5) >> 1) + 4;
    }
    inline void update(unsigned bin) {
        s0 -= (s0 >> r0) & kMask0;
        s1 -= (s1 >> r1) & kMask1;
        if (bin) {
            s0 += (0x7fffu >> r0) & kMask0;
            s1 += (0x7fffu >> r1) & kMask1;
        }
    }
};

struct BitCost {
    float bits[256][2];
    BitCost() {
        for (int s = 0; s < 256; ++s) {
            const double p1 = (s + 0.5) / 256.0;
            bits[s][1] = (float)-std::log2(p1);
            bits[s][0] = (float)-std::log2(1.0 - p1);
        }
    }
};
const BitCost g_cost;

constexpr uint8_t kRenorm[32] = {6, 5, 4, 4, 3, 3, 3, 3, 2, 2, 2, 2, 2, 2, 2, 2, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1};

// ------------------------------------------------------------------------------------------------
// context layout of one layer
// ------------------------------------------------------------------------------------------------
constexpr int kSigCtx = 0;                 // 3 neighbour classes x 8 trellis states
constexpr int kSignCtx = kSigCtx + 24;     // 3 neighbour classes
constexpr int kGtxCtx = kSignCtx + 3;      // 2 (sign) x up to 32 flags
constexpr int kEgkCtx = kGtxCtx + 64;      // Exp-Golomb prefix positions
constexpr int kNumCtx = kEgkCtx + 32;
constexpr int kMaxUnary = 32;

// 8-state machine of the dependent quantiser: next state from (state, parity of the transmitted index);
// states with an odd number use quantiser Q1 (odd multiples of the step size and zero), the others Q0 (even multiples)
constexpr uint8_t kNextState[8][2] = {{0, 2}, {7, 5}, {1, 3}, {6, 4}, {2, 0}, {5, 7}, {3, 1}, {4, 6}};
inline int sgn(int64_t v) { return (v > 0) - (v < 0); }
// union-grid integer n of index k in quantiser q, and back
inline int64_t grid_of(int64_t k, int q) { return q ? 2 * k - sgn(k) : 2 * k; }

// One level as a sequence of (context index | bypass, bin) pairs -- shared by encoder, decoder and the rate estimate
struct Binariser {
    int unary = 11;                        // cabac_unary_length_minus1 + 1
    int32_t prev = 0;                      // previously coded level (context neighbour)
    void start_layer() { prev = 0; }
    inline int nb() const { return prev == 0 ? 0 : (prev < 0 ? 1 : 2); }
};

// ------------------------------------------------------------------------------------------------
// arithmetic encoder
// ------------------------------------------------------------------------------------------------
struct ArithEncoder {
    std::vector<uint8_t> out;
    uint32_t low = 0, range = 510;
    int bits_left = 23;
    uint32_t buffered_byte = 0xff;
    int num_buffered = 0;
    void reset() { out.clear(); low = 0; range = 510; bits_left = 23; buffered_byte = 0xff; num_buffered = 0; }
    void write_out() {
        const uint32_t lead = low >> (24 - bits_left);
        bits_left += 8;
        low &= 0xffffffffu >> bits_left;
        if (lead == 0xff) {
            ++num_buffered;
        } else if (num_buffered > 0) {
            const uint32_t carry = lead >> 8;
            uint32_t byte = buffered_byte + carry;
            buffered_byte = lead & 0xff;
            out.push_back((uint8_t)byte);
            byte = (0xff + carry) & 0xff;
            while (num_buffered > 1) { out.push_back((uint8_t)byte); --num_buffered; }
        } else {
            num_buffered = 1;
            buffered_byte = lead;
        }
    }
    inline void test_write() { if (bits_left < 12) write_out(); }
    void encode(unsigned bin, Ctx& c) {
        const uint32_t lps = c.lps(range);
        range -= lps;
        if (bin != c.mps()) {
            const int nb = kRenorm[lps >> 3];
            low = (low + range) << nb;
            range = lps << nb;
            bits_left -= nb;
        } else if (range < 256) {
            low <<= 1;
            range <<= 1;
            bits_left -= 1;
        }
        c.update(bin);
        test_write();
    }
    void encode_bypass(unsigned bin) {
        low <<= 1;
        if (bin) low += range;
        --bits_left;
        test_write();
    }
    void encode_bypass_bits(uint32_t value, int n) { for (int i = n - 1; i >= 0; --i) encode_bypass((value >> i) & 1u); }
    void encode_terminate(unsigned bin) {
        range -= 2;
        if (bin) {
            low += range;
            low <<= 7;
            range = 2 << 7;
            bits_left -= 7;
        } else if (range < 256) {
            low <<= 1;
            range <<= 1;
            --bits_left;
        }
        test_write();
    }
    // flush the codeword, then a stop bit and zero bits up to the byte boundary
    void finish() {
        if (low >> (32 - bits_left)) {
            out.push_back((uint8_t)(buffered_byte + 1));
            while (num_buffered > 1) { out.push_back(0x00); --num_buffered; }
            low -= 1u << (32 - bits_left);
        } else {
            if (num_buffered > 0) out.push_back((uint8_t)buffered_byte);
            while (num_buffered > 1) { out.push_back(0xff); --num_buffered; }
        }
        // remaining (24 - bits_left) bits of low >> 8, then '1', then alignment zeros
        const int nbits = 24 - bits_left;
        uint64_t acc = nbits > 0 ? ((uint64_t)(low >> 8) & ((1ull << nbits) - 1)) : 0;
        acc = (acc << 1) | 1u;
        int total = nbits + 1;
        const int pad = (8 - total % 8) % 8;
        acc <<= pad;
        total += pad;
        for (int sh = total - 8; sh >= 0; sh -= 8) out.push_back((uint8_t)(acc >> sh));
    }
};

// ------------------------------------------------------------------------------------------------
// arithmetic decoder
// ------------------------------------------------------------------------------------------------
struct ArithDecoder {
    std::vector<uint8_t> buf;
    size_t pos = 0;
    uint32_t range = 510, value = 0;
    int bits_needed = -8;
    bool overrun = false;
    inline uint32_t read_byte() {
        if (pos < buf.size()) return buf[pos++];
        overrun = true;
        ++pos;
        return 0;
    }
    void start() {
        pos = 0; overrun = false;
        range = 510;
        bits_needed = -8;
        value = (read_byte() << 8);
        value |= read_byte();
    }
    unsigned decode(Ctx& c) {
        const uint32_t lps = c.lps(range);
        unsigned bin = c.mps();
        range -= lps;
        const uint32_t scaled = range << 7;
        if (value < scaled) {
            if (scaled < (256u << 7)) {
                range = scaled >> 6;
                value += value;
                if (++bits_needed == 0) {
                    bits_needed = -8;
                    value += read_byte();
                }
            }
        } else {
            bin = 1 - bin;
            const int nb = kRenorm[lps >> 3];
            value = (value - scaled) << nb;
            range = lps << nb;
            bits_needed += nb;
            if (bits_needed >= 0) {
                value += read_byte() << bits_needed;
                bits_needed -= 8;
            }
        }
        c.update(bin);
        return bin;
    }
    unsigned decode_bypass() {
        value += value;
        if (++bits_needed >= 0) {
            bits_needed = -8;
            value += read_byte();
        }
        const uint32_t scaled = range << 7;
        if (value >= scaled) {
            value -= scaled;
            return 1;
        }
        return 0;
    }
    uint32_t decode_bypass_bits(int n) {
        uint32_t v = 0;
        for (int i = 0; i < n; ++i) v = (v << 1) | decode_bypass();
        return v;
    }
    unsigned decode_terminate() {
        range -= 2;
        const uint32_t scaled = range << 7;
        if (value >= scaled) return 1;
        if (scaled < (256u << 7)) {
            range = scaled >> 6;
            value += value;
            if (++bits_needed == 0) {
                bits_needed = -8;
                value += read_byte();
            }
        }
        return 0;
    }
};

// ------------------------------------------------------------------------------------------------
// level coding on top of the engines
// ------------------------------------------------------------------------------------------------
template <class Sink>   // Sink::bin(ctx_index, bin) / Sink::bypass(bits, n)
inline void binarise_level(Sink& s, const Binariser& b, int32_t k, int state) {
    const int nbc = b.nb();
    s.bin(kSigCtx + 3 * state + nbc, k != 0);
    if (k == 0) return;
    const unsigned neg = k < 0;
    s.bin(kSignCtx + nbc, neg);
    uint32_t a = (uint32_t)(neg ? -(int64_t)k : (int64_t)k) - 1;            // abs level - 1
    int i = 0;
    for (; i < b.unary; ++i) {
        const unsigned gt = a > 0;
        s.bin(kGtxCtx + 2 * (i < kMaxUnary ? i : kMaxUnary - 1) + neg, gt);
        if (!gt) return;
        --a;
    }
    // remainder a >= 0: Exp-Golomb order 0, prefix bins context coded by position
    const uint32_t v = a + 1;                  // >= 1; a = 2^31 - 1 - unary at most, no overflow
    int n = 0;
    while ((v >> (n + 1)) != 0) ++n;           // n = floor(log2 v)
    for (int j = 0; j < n; ++j) s.bin(kEgkCtx + (j < 31 ? j : 31), 1);
    s.bin(kEgkCtx + (n < 31 ? n : 31), 0);
    if (n > 0) s.bypass(v - (1u << n), n);
}

struct EncSink {
    ArithEncoder& ae;
    Ctx* ctx;
    inline void bin(int c, unsigned b) { ae.encode(b, ctx[c]); }
    inline void bypass(uint32_t v, int n) { ae.encode_bypass_bits(v, n); }
};

// records the bins per context (for the rate-set choice) without coding anything
struct TraceSink {
    std::vector<uint8_t>* per_ctx;
    inline void bin(int c, unsigned b) { per_ctx[c].push_back((uint8_t)b); }
    inline void bypass(uint32_t, int) {}
};

inline int32_t decode_level(ArithDecoder& ad, Ctx* ctx, const Binariser& b, int state) {
    const int nbc = b.nb();
    if (!ad.decode(ctx[kSigCtx + 3 * state + nbc])) return 0;
    const unsigned neg = ad.decode(ctx[kSignCtx + nbc]);
    uint32_t a = 1;
    int i = 0;
    for (; i < b.unary; ++i) {
        if (!ad.decode(ctx[kGtxCtx + 2 * (i < kMaxUnary ? i : kMaxUnary - 1) + neg])) break;
        ++a;
    }
    if (i == b.unary) {
        int n = 0;
        while (n < 32 && ad.decode(ctx[kEgkCtx + (n < 31 ? n : 31)])) ++n;
        if (n >= 32) { ad.overrun = true; return 0; }
        uint32_t v = 1u << n;
        if (n > 0) v += ad.decode_bypass_bits(n);
        a += v - 1;
    }
    return neg ? -(int32_t)a : (int32_t)a;
}

// index transmitted for union-grid integer n in trellis state `state` (false: n is not on that quantiser's grid)
inline bool index_of(int64_t n, int state, int64_t* k) {
    if (state & 1) {
        if (n != 0 && !(n & 1)) return false;
        *k = (n + sgn(n)) / 2;
    } else {
        if (n & 1) return false;
        *k = n / 2;
    }
    return true;
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// handles
// ------------------------------------------------------------------------------------------------
struct nncabac_encoder {
    ArithEncoder ae;
    Ctx ctx[kNumCtx];
    Binariser bin;
    int param_opt = 0;
    bool finished = false;
};
struct nncabac_decoder {
    ArithDecoder ad;
    Ctx ctx[kNumCtx];
    Binariser bin;
    bool started = false;
};

extern "C" {

float nncabac_stepsize(int qp, int qp_density) { return stepsize(qp, qp_density); }

nncabac_encoder* nncabac_encoder_new(void) { return new (std::nothrow) nncabac_encoder(); }
void nncabac_encoder_free(nncabac_encoder* e) { delete e; }

int nncabac_encoder_init_ctx(nncabac_encoder* e, int cabac_unary_length_minus1, int param_opt_flag) {
    if (!e || cabac_unary_length_minus1 < 0 || cabac_unary_length_minus1 >= kMaxUnary) return -1;
    for (auto& c : e->ctx) c.reset();
    e->bin.unary = cabac_unary_length_minus1 + 1;
    e->bin.start_layer();
    e->param_opt = param_opt_flag ? 1 : 0;
    return 0;
}

int nncabac_encoder_iae_v(nncabac_encoder* e, int n_bits, int value) {
    if (!e || n_bits < 1 || n_bits > 32) return -1;
    if (e->finished) { e->ae.reset(); e->finished = false; }
    e->ae.encode_bypass_bits((uint32_t)value & (n_bits == 32 ? 0xffffffffu : ((1u << n_bits) - 1)), n_bits);
    return 0;
}

int nncabac_quant_layer(nncabac_encoder* e, const float* w, int32_t* lvl, int64_t n, int dq_flag, int qp_density, int qp,
                        float lambda_scale, int cabac_unary_length_minus1, int scan_order, int* qp_used) {
    (void)e; (void)cabac_unary_length_minus1;
    if (n < 0 || (n > 0 && (!w || !lvl)) || qp_density < 0 || qp_density > 8) return -1;
    if (scan_order != 0 && dq_flag) return -2;            // the trellis follows the coding order; block scans are not implemented
    float max_abs = 0.0f;
    for (int64_t i = 0; i < n; ++i) {
        const float a = std::fabs(w[i]);
        if (a > max_abs || a != a) max_abs = a;
        if (max_abs != max_abs) break;
    }
    const int q = clip_qp(max_abs, qp, qp_density, dq_flag ? 2.0f : 0.0f);
    if (qp_used) *qp_used = q;
    const float d = stepsize(q, qp_density);
    if (!dq_flag) {
        for (int64_t i = 0; i < n; ++i) lvl[i] = urq_level(w[i], d);
        return 0;
    }
    // ---- dependent quantisation: Viterbi over the 8-state trellis ----
    // cost = squared error in units of delta^2 (+ lambda_scale * a static bit estimate); per weight and state the two
    // indices whose reconstruction points bracket the weight are tried (they differ in parity, so both transitions exist)
    const double inf = std::numeric_limits<double>::infinity();
    double cost[8], next_cost[8];
    for (int s = 0; s < 8; ++s) cost[s] = s == 0 ? 0.0 : inf;
    struct Back { uint8_t prev; int32_t k; };
    std::vector<Back> back((size_t)n * 8);
    auto bits_of = [](int64_t k) { const double a = (double)(k < 0 ? -k : k); return a == 0 ? 1.0 : 2.0 + 2.0 * std::log2(1.0 + a); };
    const double dd = (double)d;
    for (int64_t i = 0; i < n; ++i) {
        double x = (double)w[i] / dd;
        if (!(std::fabs(x) <= 1.0e300)) x = 0.0;                    // NaN / Inf -> level 0, like the uniform quantiser
        for (int s = 0; s < 8; ++s) next_cost[s] = inf;
        Back* bk = &back[(size_t)i * 8];
        for (int s = 0; s < 8; ++s) {
            if (cost[s] == inf) continue;
            const int qsel = s & 1;
            int64_t k0 = (int64_t)std::floor(x / 2.0);
            while ((double)grid_of(k0 + 1, qsel) <= x) ++k0;
            while ((double)grid_of(k0, qsel) > x) --k0;
            for (int c = 0; c < 2; ++c) {
                const int64_t k = k0 + c;
                const double err = x - (double)grid_of(k, qsel);
                const double cnew = cost[s] + err * err + (lambda_scale != 0.0f ? (double)lambda_scale * bits_of(k) : 0.0);
                const int ns = kNextState[s][k & 1];
                if (cnew < next_cost[ns]) {
                    next_cost[ns] = cnew;
                    bk[ns].prev = (uint8_t)s;
                    bk[ns].k = (int32_t)k;
                }
            }
        }
        std::memcpy(cost, next_cost, sizeof(cost));
    }
    int best = 0;
    for (int s = 1; s < 8; ++s) if (cost[s] < cost[best]) best = s;
    int s = best;
    for (int64_t i = n - 1; i >= 0; --i) {
        const Back& b = back[(size_t)i * 8 + s];
        lvl[i] = (int32_t)grid_of(b.k, b.prev & 1);
        s = b.prev;
    }
    return 0;
}

int nncabac_encoder_encode_layer(nncabac_encoder* e, const int32_t* lvl, int64_t n, int dq_flag, int scan_order) {
    if (!e || n < 0 || (n > 0 && !lvl)) return -1;
    if (scan_order != 0) return -2;
    if (e->finished) { e->ae.reset(); e->finished = false; }
    e->bin.start_layer();
    // transmitted indices (dependent quantisation maps the union-grid integers back through the state machine)
    std::vector<int32_t> idx((size_t)n);
    std::vector<uint8_t> st((size_t)n);
    int state = 0;
    for (int64_t i = 0; i < n; ++i) {
        int64_t k = lvl[i];
        st[(size_t)i] = (uint8_t)state;
        if (dq_flag) {
            if (!index_of(lvl[i], state, &k)) return -1;            // not a path of the trellis
            state = kNextState[state][k & 1];
        }
        idx[(size_t)i] = (int32_t)k;
    }
    // per-context adaptation rates: with param_opt the bins of this layer are traced once and the cheapest of the
    // kNumRateSets rate pairs is chosen per context (contexts are independent given the bin sequence) and signalled
    uint8_t rate_set[kNumCtx] = {0};
    unsigned any = 0;
    if (e->param_opt && n > 0) {
        std::vector<uint8_t> per_ctx[kNumCtx];
        TraceSink ts{per_ctx};
        Binariser b = e->bin;
        for (int64_t i = 0; i < n; ++i) {
            binarise_level(ts, b, idx[(size_t)i], dq_flag ? st[(size_t)i] : 0);
            b.prev = idx[(size_t)i];
        }
        for (int c = 0; c < kNumCtx; ++c) {
            if (per_ctx[c].empty()) continue;
            double best = 0.0;
            for (int r = 0; r < kNumRateSets; ++r) {
                Ctx m;
                m.reset(r);
                double bits = 0.0;
                for (uint8_t bn : per_ctx[c]) { bits += g_cost.bits[m.state()][bn]; m.update(bn); }
                if (r == 0 || bits < best - 2.0) { best = bits; rate_set[c] = (uint8_t)r; }      // a switch must pay for its signalling
            }
            any |= rate_set[c];
        }
    }
    e->ae.encode_bypass(any ? 1 : 0);
    if (any)
        for (int c = 0; c < kNumCtx; ++c) e->ae.encode_bypass_bits(rate_set[c], 2);
    for (int c = 0; c < kNumCtx; ++c) e->ctx[c].reset(rate_set[c]);
    EncSink es{e->ae, e->ctx};
    for (int64_t i = 0; i < n; ++i) {
        binarise_level(es, e->bin, idx[(size_t)i], dq_flag ? st[(size_t)i] : 0);
        e->bin.prev = idx[(size_t)i];
    }
    return 0;
}

int nncabac_encoder_finish(nncabac_encoder* e, const uint8_t** data, int64_t* size) {
    if (!e || !data || !size) return -1;
    if (!e->finished) {
        e->ae.encode_terminate(1);
        e->ae.finish();
        e->finished = true;
    }
    *data = e->ae.out.data();
    *size = (int64_t)e->ae.out.size();
    return 0;
}

nncabac_decoder* nncabac_decoder_new(void) { return new (std::nothrow) nncabac_decoder(); }
void nncabac_decoder_free(nncabac_decoder* d) { delete d; }

int nncabac_decoder_set_stream(nncabac_decoder* d, const uint8_t* data, int64_t size) {
    if (!d || size < 0 || (size > 0 && !data)) return -1;
    d->ad.buf.assign(data, data + size);
    d->ad.start();
    d->started = true;
    return 0;
}

int nncabac_decoder_init_ctx(nncabac_decoder* d, int cabac_unary_length_minus1) {
    if (!d || cabac_unary_length_minus1 < 0 || cabac_unary_length_minus1 >= kMaxUnary) return -1;
    for (auto& c : d->ctx) c.reset();
    d->bin.unary = cabac_unary_length_minus1 + 1;
    d->bin.start_layer();
    return 0;
}

int nncabac_decoder_iae_v(nncabac_decoder* d, int n_bits, int* value) {
    if (!d || !d->started || !value || n_bits < 1 || n_bits > 32) return -1;
    const uint32_t v = d->ad.decode_bypass_bits(n_bits);
    if (n_bits < 32 && (v >> (n_bits - 1))) *value = (int)((int64_t)v - ((int64_t)1 << n_bits));        // sign extension
    else *value = (int)v;
    return d->ad.overrun ? -3 : 0;
}

int nncabac_decoder_decode_layer(nncabac_decoder* d, int32_t* lvl, int64_t n, int dq_flag, int scan_order) {
    if (!d || !d->started || n < 0 || (n > 0 && !lvl)) return -1;
    if (scan_order != 0) return -2;
    d->bin.start_layer();
    const unsigned any = d->ad.decode_bypass();
    for (int c = 0; c < kNumCtx; ++c) d->ctx[c].reset(any ? (int)d->ad.decode_bypass_bits(2) : 0);
    int state = 0;
    for (int64_t i = 0; i < n; ++i) {
        const int32_t k = decode_level(d->ad, d->ctx, d->bin, dq_flag ? state : 0);
        d->bin.prev = k;
        if (dq_flag) {
            lvl[i] = (int32_t)grid_of(k, state & 1);
            state = kNextState[state][k & 1];
        } else {
            lvl[i] = k;
        }
        if (d->ad.overrun) return -3;
    }
    return 0;
}

int nncabac_decoder_finish(nncabac_decoder* d, int64_t* bytes_read) {
    if (!d || !d->started || !bytes_read) return -1;
    if (!d->ad.decode_terminate() || d->ad.overrun) return -3;
    // the byte read last holds the stop bit (the encoder's finish writes it right behind the codeword)
    if (d->ad.pos == 0 || d->ad.pos > d->ad.buf.size()) return -3;
    const uint32_t last = d->ad.buf[d->ad.pos - 1];
    if (((last << (8 + d->ad.bits_needed)) & 0xff) != 0x80) return -3;
    *bytes_read = (int64_t)d->ad.pos;
    return 0;
}

int nncabac_dequant_layer(float* out, const int32_t* lvl, int64_t n, int qp_density, int qp) {
    if (n < 0 || (n > 0 && (!out || !lvl)) || qp_density < 0 || qp_density > 8) return -1;
    const float d = stepsize(qp, qp_density);
    for (int64_t i = 0; i < n; ++i) out[i] = (float)lvl[i] * d;
    return 0;
}

}  // extern "C"
