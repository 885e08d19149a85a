// fx_csv.cpp -- host-side text formatter for effex's .csv rows (effex/effex.py:693:
// `np.savetxt(fh, [row], delimiter=',')` of complex128).  numpy writes every element as
// " (%.18e%+.18ej)" joined by ',' and ends the row with '\n'.  float32 results are widened to
// float64 first, so the text round-trips through np.loadtxt(dtype=complex128) exactly.
// Rows are formatted in parallel (the reference's writer thread is bound by np.savetxt at
// ~1.5 us per element; BASELINE config 5 writes 6000 x 1024 of them).
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/effex_fx.h"

namespace {
// worst case per element: " (" + 25 + 25 + "j)" + ',' = 55 bytes; + '\n' per row
constexpr size_t kElemMax = 56;

size_t format_row(const float *row, long long nbins, char *out) {
    char *p = out;
    for (long long c = 0; c < nbins; ++c) {
        if (c) *p++ = ',';
        const double re = (double)row[2 * c], im = (double)row[2 * c + 1];
        p += snprintf(p, kElemMax, " (%.18e%+.18ej)", re, im);
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
    const size_t stride = (size_t)nbins * kElemMax + 2;
    if (n_threads < 1) n_threads = (int)std::thread::hardware_concurrency();
    if (n_threads < 1) n_threads = 1;
    if ((int64_t)n_threads > n_rows) n_threads = (int)(n_rows > 0 ? n_rows : 1);
    std::vector<size_t> len((size_t)n_rows, 0);
    // pass 1: every row into its own worst-case slot, in parallel
    auto work = [&](int tid) {
        for (int64_t r = tid; r < n_rows; r += n_threads)
            len[(size_t)r] = format_row(h_rows + 2 * (size_t)nbins * (size_t)r, nbins, h_out + stride * (size_t)r);
    };
    std::vector<std::thread> pool;
    for (int i = 1; i < n_threads; ++i) pool.emplace_back(work, i);
    work(0);
    for (auto &t : pool) t.join();
    // pass 2: compact in row order (memmove: slots never overlap their packed destination from behind)
    size_t off = 0;
    for (int64_t r = 0; r < n_rows; ++r) {
        memmove(h_out + off, h_out + stride * (size_t)r, len[(size_t)r]);
        off += len[(size_t)r];
    }
    *out_len = off;
    return FX_OK;
}

}  // extern "C"
