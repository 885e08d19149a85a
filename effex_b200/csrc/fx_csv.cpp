// fx_csv.cpp -- host-side text formatter for effex's .csv rows (effex/effex.py:693:
// `np.savetxt(fh, [row], delimiter=',')` of complex128).  numpy writes every element as
// " (%.18e%+.18ej)" joined by ',' and ends the row with '\n'.  float32 results are widened to
// float64 first, so the text round-trips through np.loadtxt(dtype=complex128) exactly.
//
// The reference's writer thread is bound by np.savetxt (~1.5 us per element; BASELINE config 5 writes
// 6000 x 1024 of them) and a snprintf("%.18e") per value still costs ~0.3 us.  Here the 19 significant
// digits come from exact integer arithmetic on the value itself: a finite double is m * 2^e with an integer
// m < 2^53, so it equals the integer D = m * 2^e (e >= 0) or D / 10^-e with D = m * 5^-e (e < 0); D is
// built in base-10^9 limbs (a float32-born value of the size a cross-spectrum has needs 3..6 limbs), its
// leading 19 digits are rounded half-to-even on the exact remainder -- what glibc's printf does -- and the
// decimal exponent is the digit count.  Byte-identical to snprintf/np.savetxt (tests/test_csv_cpu.py
// compares 10^6 values incl. ties, subnormals and the float32 extremes); rows are formatted in parallel.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/effex_fx.h"

namespace {
// worst case per element: " (" + 25 + 25 + "j)" + ',' = 55 bytes; + '\n' per row
constexpr size_t kElemMax = 56;
constexpr uint32_t kBase = 1000000000u;
constexpr int kMaxLimbs = 96;            // 10^864 > 2^1024 * ... covers every finite double

struct Big {
    uint32_t l[kMaxLimbs];               // little-endian base 10^9
    int n;
};
inline void big_mul_small(Big &b, uint32_t f) {
    uint64_t carry = 0;
    for (int i = 0; i < b.n; ++i) {
        const uint64_t t = (uint64_t)b.l[i] * f + carry;
        b.l[i] = (uint32_t)(t % kBase);
        carry = t / kBase;
    }
    while (carry) {
        b.l[b.n++] = (uint32_t)(carry % kBase);
        carry /= kBase;
    }
}

// 5^k, k = 0 .. kPow5Max, as base-10^9 limbs: a float32-born value has a mantissa below 2^32, so its D is ONE
// pass of big_mul_small over a table entry (float32 exponents reach 2^-149 -> 5^172 after the 23 mantissa bits)
constexpr int kPow5Max = 180, kPow5Limbs = 15;
struct Pow5Table {
    uint32_t l[kPow5Max + 1][kPow5Limbs];
    int n[kPow5Max + 1];
    Pow5Table() {
        Big b;
        b.n = 1; b.l[0] = 1;
        for (int k = 0; k <= kPow5Max; ++k) {
            n[k] = b.n;
            for (int i = 0; i < kPow5Limbs; ++i) l[k][i] = i < b.n ? b.l[i] : 0;
            big_mul_small(b, 5);
        }
    }
};
const Pow5Table &pow5_table() {
    static const Pow5Table t;
    return t;
}

// "%.18e" (plus == false) or "%+.18e" (plus == true) of v into p; returns the end
char *fmt_e18(double v, bool plus, char *p) {
    if (!std::isfinite(v)) return p + snprintf(p, 32, plus ? "%+.18e" : "%.18e", v);
    if (std::signbit(v)) { *p++ = '-'; v = -v; }
    else if (plus) *p++ = '+';
    char dig[24];
    int E = 0;
    if (v == 0.0) {
        memset(dig, '0', 19);
    } else {
        uint64_t bits;
        memcpy(&bits, &v, sizeof(bits));
        const int ex = (int)((bits >> 52) & 0x7ff);
        uint64_t m = bits & ((1ull << 52) - 1);
        int e2 = -1074;                                      // subnormal: v = frac * 2^-1074
        if (ex) { m |= 1ull << 52; e2 = ex - 1075; }
        const int tz = __builtin_ctzll(m);
        m >>= tz;
        e2 += tz;                                            // v = m * 2^e2, m odd
        Big b;
        b.n = 0;
        int dec_shift = 0;                                   // v = D * 10^dec_shift
        if (e2 < 0 && -e2 <= kPow5Max && m < (1ull << 32)) {
            const Pow5Table &t5 = pow5_table();
            b.n = t5.n[-e2];
            memcpy(b.l, t5.l[-e2], sizeof(uint32_t) * (size_t)b.n);
            big_mul_small(b, (uint32_t)m);
            dec_shift = e2;
        } else if (e2 >= 0) {
            for (uint64_t t = m; t; t /= kBase) b.l[b.n++] = (uint32_t)(t % kBase);
            int k = e2;
            for (; k >= 29; k -= 29) big_mul_small(b, 1u << 29);
            if (k) big_mul_small(b, 1u << k);
        } else {
            for (uint64_t t = m; t; t /= kBase) b.l[b.n++] = (uint32_t)(t % kBase);
            int k = -e2;
            dec_shift = e2;
            for (; k >= 13; k -= 13) big_mul_small(b, 1220703125u);       // 5^13
            static const uint32_t p5[13] = {1, 5, 25, 125, 625, 3125, 15625, 78125, 390625, 1953125, 9765625, 48828125, 244140625};
            if (k) big_mul_small(b, p5[k]);
        }
        // leading digits: the top limb without leading zeros, then whole 9-digit limbs (constant divisors)
        char buf[48];
        int nt = 0;
        {
            char top[12];
            for (uint32_t t = b.l[b.n - 1]; t; t /= 10) top[nt++] = (char)('0' + t % 10);
            for (int i = 0; i < nt; ++i) buf[i] = top[nt - 1 - i];
        }
        E = nt + 9 * (b.n - 1) - 1 + dec_shift;
        int got = nt, limb = b.n - 2;
        for (; got < 21 && limb >= 0; --limb, got += 9) {
            uint32_t x = b.l[limb];
            for (int i = 8; i >= 0; --i) { buf[got + i] = (char)('0' + x % 10); x /= 10; }
        }
        if (got <= 19) {
            memcpy(dig, buf, (size_t)got);
            for (; got < 19; ++got) dig[got] = '0';
        } else {
            memcpy(dig, buf, 19);
            // digit 20 and everything after it decide the rounding (half to even)
            const int d20 = buf[19] - '0';
            bool rest = false;
            for (int i = 20; i < got && !rest; ++i) rest = buf[i] != '0';
            for (int i = limb; i >= 0 && !rest; --i) rest = b.l[i] != 0;
            const bool up = d20 > 5 || (d20 == 5 && (rest || ((dig[18] - '0') & 1)));
            if (up) {
                int i = 18;
                for (; i >= 0 && dig[i] == '9'; --i) dig[i] = '0';
                if (i >= 0) ++dig[i];
                else { dig[0] = '1'; ++E; }                  // 9.99..9 -> 1.00..0e+1
            }
        }
    }
    *p++ = dig[0];
    *p++ = '.';
    memcpy(p, dig + 1, 18);
    p += 18;
    *p++ = 'e';
    if (E < 0) { *p++ = '-'; E = -E; } else *p++ = '+';
    if (E >= 100) { *p++ = (char)('0' + E / 100); E %= 100; }
    *p++ = (char)('0' + E / 10);
    *p++ = (char)('0' + E % 10);
    return p;
}

size_t format_row(const float *row, long long nbins, char *out) {
    char *p = out;
    for (long long c = 0; c < nbins; ++c) {
        if (c) *p++ = ',';
        *p++ = ' ';
        *p++ = '(';
        p = fmt_e18((double)row[2 * c], false, p);
        p = fmt_e18((double)row[2 * c + 1], true, p);
        *p++ = 'j';
        *p++ = ')';
    }
    *p++ = '\n';
    return (size_t)(p - out);
}
}  // namespace

extern "C" {

size_t fx_csv_rows_bound(int64_t n_rows, int64_t nbins) {
    return (size_t)n_rows * ((size_t)nbins * kElemMax + 2) + 1;
}

int fx_csv_format_rows(const float *h_rows, int64_t n_rows, int64_t nbins, int n_threads, char *h_out,
                       size_t out_cap, size_t *out_len) {
    if (!h_rows || !h_out || !out_len || n_rows < 0 || nbins < 1) return FX_ERR_INVALID;
    if (out_cap < fx_csv_rows_bound(n_rows, nbins)) return FX_ERR_INVALID;
    if (n_threads < 1) n_threads = (int)std::thread::hardware_concurrency();
    if (n_threads < 1) n_threads = 1;
    if ((int64_t)n_threads > n_rows) n_threads = (int)(n_rows > 0 ? n_rows : 1);
    auto run = [&](auto &&work) {
        std::vector<std::thread> pool;
        for (int i = 1; i < n_threads; ++i) pool.emplace_back(work, i);
        work(0);
        for (auto &t : pool) t.join();
    };
    // A finite float32 prints with a two-digit exponent, so an element is 53 characters plus one per minus
    // sign of the real part: the row lengths are known before a digit is produced and every row is formatted
    // straight into its final place.  (Rows holding inf/nan take the slot path below.)
    std::vector<size_t> len((size_t)n_rows, 0);
    std::vector<char> odd((size_t)n_threads, 0);
    run([&](int tid) {
        for (int64_t r = tid; r < n_rows; r += n_threads) {
            const float *row = h_rows + 2 * (size_t)nbins * (size_t)r;
            size_t neg = 0;
            bool finite = true;
            for (int64_t c = 0; c < nbins; ++c) {
                neg += std::signbit(row[2 * c]) ? 1 : 0;
                finite = finite && std::isfinite(row[2 * c]) && std::isfinite(row[2 * c + 1]);
            }
            if (!finite) odd[(size_t)tid] = 1;
            len[(size_t)r] = (size_t)nbins * 53 + neg + (size_t)(nbins - 1) + 1;
        }
    });
    bool any_odd = false;
    for (char o : odd) any_odd = any_odd || o;
    if (!any_odd) {
        std::vector<size_t> off((size_t)n_rows + 1, 0);
        for (int64_t r = 0; r < n_rows; ++r) off[(size_t)r + 1] = off[(size_t)r] + len[(size_t)r];
        run([&](int tid) {
            for (int64_t r = tid; r < n_rows; r += n_threads)
                format_row(h_rows + 2 * (size_t)nbins * (size_t)r, nbins, h_out + off[(size_t)r]);
        });
        *out_len = off[(size_t)n_rows];
        return FX_OK;
    }
    // general path: every row into its own worst-case slot in parallel, then compacted in row order
    const size_t stride = (size_t)nbins * kElemMax + 2;
    run([&](int tid) {
        for (int64_t r = tid; r < n_rows; r += n_threads)
            len[(size_t)r] = format_row(h_rows + 2 * (size_t)nbins * (size_t)r, nbins, h_out + stride * (size_t)r);
    });
    size_t off = 0;
    for (int64_t r = 0; r < n_rows; ++r) {
        memmove(h_out + off, h_out + stride * (size_t)r, len[(size_t)r]);
        off += len[(size_t)r];
    }
    *out_len = off;
    return FX_OK;
}

/* "%.18e" of one double, exactly as printf writes it (the formatter above, exposed for the tests and for
 * metadata lines); returns the number of characters, no terminating NUL (h_out must hold 32 bytes). */
int fx_csv_format_double(double v, int plus_sign, char *h_out) {
    if (!h_out) return FX_ERR_INVALID;
    return (int)(fmt_e18(v, plus_sign != 0, h_out) - h_out);
}

}  // extern "C"
