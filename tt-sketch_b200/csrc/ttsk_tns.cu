// FROSTT ".tns" text -> COO arrays (the on-disk format in front of the sketching path).
// Replaces the per-line Python loop of scripts/frostt.py:51-66 of the reference: one nonzero per line, d 1-based integer
// coordinates and a value separated by blanks; empty lines and lines starting with '#' are skipped.  Host code: the
// buffer is cut at line ends into one range per thread, lines are counted, then parsed in place into the caller's
// (d x nnz) int64 index rows (0-based) and fp64 values -- the layout ttsk_sparse_sketch_host streams to the GPU.
// Values go through strtod (correctly rounded, like Python's float()).
#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include "ttsk_common.cuh"

namespace {

inline bool blank(char c) { return c == ' ' || c == '\t' || c == '\r'; }

// is [p, e) (one line without its '\n') a data line?
inline bool data_line(const char* p, const char* e) {
    while (p < e && blank(*p)) p++;
    return p < e && *p != '#';
}

struct Range {
    const char* lo;
    const char* hi;
    int64_t lines = 0, first = 0;
    int status = 0;
};

void cut(const char* buf, int64_t len, int threads, std::vector<Range>& out) {
    const char* end = buf + len;
    const char* p = buf;
    for (int t = 0; t < threads && p < end; t++) {
        const char* q = (t == threads - 1) ? end : buf + len * (t + 1) / threads;
        if (q < p) q = p;
        while (q < end && *q != '\n') q++;
        if (q < end) q++;  // include the newline
        out.push_back({p, q});
        p = q;
    }
}

void count_lines(Range& r) {
    const char* p = r.lo;
    while (p < r.hi) {
        const char* e = (const char*)memchr(p, '\n', (size_t)(r.hi - p));
        if (!e) e = r.hi;
        if (data_line(p, e)) r.lines++;
        p = e + 1;
    }
}

// columns of the first data line
int first_line_fields(const char* buf, int64_t len) {
    const char* p = buf;
    const char* end = buf + len;
    while (p < end) {
        const char* e = (const char*)memchr(p, '\n', (size_t)(end - p));
        if (!e) e = end;
        if (data_line(p, e)) {
            int f = 0;
            while (p < e) {
                while (p < e && blank(*p)) p++;
                if (p >= e) break;
                f++;
                while (p < e && !blank(*p)) p++;
            }
            return f;
        }
        p = e + 1;
    }
    return 0;
}

void parse_range(Range& r, int d, int64_t nnz, int64_t* idx, double* val, int64_t* max_idx) {
    const char* p = r.lo;
    int64_t row = r.first;
    std::vector<int64_t> mx((size_t)d, -1);
    while (p < r.hi) {
        const char* e = (const char*)memchr(p, '\n', (size_t)(r.hi - p));
        if (!e) e = r.hi;
        if (data_line(p, e)) {
            if (row >= nnz) { r.status = 1; return; }
            const char* q = p;
            for (int m = 0; m < d; m++) {
                while (q < e && blank(*q)) q++;
                if (q >= e || *q < '0' || *q > '9') { r.status = 2; return; }
                int64_t v = 0;
                while (q < e && *q >= '0' && *q <= '9') v = v * 10 + (*q++ - '0');
                if (q < e && !blank(*q)) { r.status = 2; return; }
                if (v < 1) { r.status = 3; return; }  // coordinates are 1-based
                idx[(int64_t)m * nnz + row] = v - 1;
                mx[(size_t)m] = std::max(mx[(size_t)m], v - 1);
            }
            while (q < e && blank(*q)) q++;
            if (q >= e) { r.status = 2; return; }
            char tmp[64];
            const size_t n = std::min<size_t>((size_t)(e - q), sizeof(tmp) - 1);
            memcpy(tmp, q, n);
            tmp[n] = 0;
            char* stop = nullptr;
            val[row] = strtod(tmp, &stop);
            if (stop == tmp) { r.status = 2; return; }
            while (*stop && blank(*stop)) stop++;
            if (*stop) { r.status = 2; return; }
            row++;
        }
        p = e + 1;
    }
    for (int m = 0; m < d; m++) max_idx[m] = mx[(size_t)m];
}

}  // namespace

extern "C" int64_t ttsk_tns_count(const char* buf, int64_t len, int* d_out) {
    if (!buf || len < 0 || !d_out) { ttsk::set_error("tns_count: bad arguments"); return -1; }
    const int fields = first_line_fields(buf, len);
    *d_out = fields > 0 ? fields - 1 : 0;
    int threads = (int)std::min<int64_t>(std::max<unsigned>(1u, std::thread::hardware_concurrency()), std::max<int64_t>(1, len >> 20));
    std::vector<Range> rs;
    cut(buf, len, threads, rs);
    std::vector<std::thread> th;
    for (auto& r : rs) th.emplace_back(count_lines, std::ref(r));
    for (auto& t : th) t.join();
    int64_t n = 0;
    for (auto& r : rs) n += r.lines;
    return n;
}

extern "C" int ttsk_tns_parse(const char* buf, int64_t len, int d, int64_t nnz, int64_t* idx, double* val, int64_t* max_idx) {
    if (!buf || len < 0 || d < 1 || d > TTSK_MAX_ORDER || nnz < 0 || (nnz > 0 && (!idx || !val)) || !max_idx) {
        ttsk::set_error("tns_parse: bad arguments");
        return TTSK_E_ARG;
    }
    int threads = (int)std::min<int64_t>(std::max<unsigned>(1u, std::thread::hardware_concurrency()), std::max<int64_t>(1, len >> 20));
    std::vector<Range> rs;
    cut(buf, len, threads, rs);
    {
        std::vector<std::thread> th;
        for (auto& r : rs) th.emplace_back(count_lines, std::ref(r));
        for (auto& t : th) t.join();
    }
    int64_t total = 0;
    for (auto& r : rs) { r.first = total; total += r.lines; }
    if (total != nnz) {
        ttsk::set_error("tns_parse: the buffer holds %lld data lines, the caller expects %lld", (long long)total, (long long)nnz);
        return TTSK_E_ARG;
    }
    std::vector<std::vector<int64_t>> mx(rs.size(), std::vector<int64_t>((size_t)d, -1));
    {
        std::vector<std::thread> th;
        for (size_t t = 0; t < rs.size(); t++)
            th.emplace_back(parse_range, std::ref(rs[t]), d, nnz, idx, val, mx[t].data());
        for (auto& t : th) t.join();
    }
    for (auto& r : rs)
        if (r.status) {
            ttsk::set_error(r.status == 3 ? "tns_parse: coordinates must be >= 1 (the format is 1-based)"
                                          : "tns_parse: malformed line (expected %d integer coordinates and a value)", d);
            return TTSK_E_ARG;
        }
    for (int m = 0; m < d; m++) {
        max_idx[m] = -1;
        for (auto& v : mx) max_idx[m] = std::max(max_idx[m], v[(size_t)m]);
    }
    return TTSK_OK;
}
