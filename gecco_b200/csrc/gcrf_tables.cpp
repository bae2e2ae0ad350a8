// gcrf_tables.cpp — the table side of `gecco predict`, native and without per-row Python objects
// (SURVEY.md §8(f) rows N2 and N3).  Host code only; it feeds the CSR arrays of gcrf_marginals_windowed and
// writes the result tables straight from the probability array.
//
// What it restates, in the reference's order of operations (gecco/cli/commands/predict.py:62-100):
//   * GeneTable.load / FeatureTable.load: tab-separated, a header line naming the columns, any column order,
//     extra columns ignored, "\n" or "\r\n" line ends (gecco/_base.py:119-131, gecco/model.py:621-637, 773-789);
//   * annotate_genes: domain rows are attached to genes by protein_id, in row order, across all feature files;
//     duplicate gene names, unknown protein ids and rows that disagree with their gene are errors
//     (gecco/cli/commands/_common.py:211-262);
//   * genes sorted by (sequence_id, start, end), each gene's domains by (domain_start, domain_end), both stable
//     (predict.py:81-83); ClusterCRF's own sort by (sequence_id, start) and by domain start (crf/__init__.py:199-201)
//     leaves that order untouched;
//   * filter_domains: keep i_evalue < e_filter, then pvalue < p_filter; NaN compares false and is dropped
//     (_common.py:419-448);
//   * features: the SET of domain names of a gene, first occurrence first, names unknown to the model dropped
//     (crf/features.py:13-35); domain mode: one row per domain, one empty row per domain-less gene (:38-48);
//   * GeneTable.from_genes(...).dump / FeatureTable.from_genes(...).dump: columns in schema order, NaN written as
//     an empty field, a probability column that holds nothing but NaN is left out, "\n" line ends
//     (gecco/model.py:644-670, 791-813, gecco/_base.py:133-151).  Floats are written in their shortest
//     round-trip form with Python's repr() layout (1e-05, not 1e-5): that is what the reference's committed
//     result tables (tests/test_cli/data/BGC0001866.*.tsv, diffed by the Galaxy tool test) contain.
#include "../../include/gecco_crf_b200.h"

#include <fcntl.h>
#include <sys/mman.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <cerrno>
#include <charconv>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <new>
#include <numeric>
#include <string>
#include <string_view>
#include <thread>
#include <unordered_map>
#include <vector>

namespace {

using sv = std::string_view;

// Large buffers (row arrays, whole files) are first touched by many threads at once: with 4 KB pages that is tens of
// thousands of page faults on one address space; 2 MB-aligned blocks with MADV_HUGEPAGE (a hint: honoured where
// transparent huge pages are "madvise" or "always") cut them by 500x.  free() releases the block.
inline void *big_alloc(size_t bytes) {
    constexpr size_t kHuge = 2u << 20;
    if (bytes >= 4 * kHuge) {
        void *p = nullptr;
        if (posix_memalign(&p, kHuge, (bytes + kHuge - 1) & ~(kHuge - 1)) == 0) {
#ifdef MADV_HUGEPAGE
            madvise(p, bytes, MADV_HUGEPAGE);
#endif
            return p;
        }
    }
    return malloc(bytes ? bytes : 1);
}

// Row storage that is not filled before the parsing threads write it: resize() on a plain std::vector would zero
// (and fault in, on ONE thread) a few hundred MB that are overwritten right away.  Only for the trivially copyable
// row structs below; every element is assigned before it is read.
template <typename T>
struct NoInitAllocator : std::allocator<T> {
    template <typename U>
    struct rebind {
        using other = NoInitAllocator<U>;
    };
    NoInitAllocator() = default;
    template <typename U>
    NoInitAllocator(const NoInitAllocator<U> &) noexcept {}
    T *allocate(size_t n) {
        void *p = big_alloc(n * sizeof(T));
        if (!p) throw std::bad_alloc();
        return static_cast<T *>(p);
    }
    void deallocate(T *p, size_t) noexcept { free(p); }
    template <typename U>
    void construct(U *) noexcept {}  // resize(): leave the storage as it is
    template <typename U, typename... Args>
    void construct(U *p, Args &&...args) {
        ::new (static_cast<void *>(p)) U(std::forward<Args>(args)...);
    }
};
template <typename T>
using RowVector = std::vector<T, NoInitAllocator<T>>;

struct GeneRow {
    sv seq, prot, strand;
    int64_t start = 0, end = 0;
};

// What is kept of a feature row once it has been checked against its gene (48 bytes + two views).
struct DomainRow {
    sv name, hmm;
    int64_t dstart, dend;
    double i_evalue, pvalue;
    int32_t gene;  // index into the sorted genes; -1 = dropped by a filter
};

// protein_id -> row number of the genes table: open addressing, one 64-bit word per slot — the upper half of the
// name's hash and the row number — so that a slot is claimed with ONE compare-and-swap and several threads can fill
// the table at once (no node allocations; the std::unordered_map it replaces took a third of a second per million
// genes, the serial fill of this table 0.11 s).
struct NameIndex {
    static constexpr uint64_t kEmpty = ~0ull;
    std::vector<uint64_t> slot;
    uint64_t mask = 0;
    static uint64_t hash_of(sv s) {
        uint64_t h = 1469598103934665603ull;  // FNV-1a, folded
        for (unsigned char c : s) h = (h ^ c) * 1099511628211ull;
        return h ^ (h >> 29);
    }
    static uint64_t word_of(uint64_t h, int32_t id) { return (h & 0xffffffff00000000ull) | (uint32_t)id; }
    void reserve(size_t n) {
        size_t cap = 16;
        while (cap < 2 * n + 2) cap <<= 1;
        slot.assign(cap, kEmpty);
        mask = cap - 1;
    }
    size_t capacity() const { return slot.size(); }
    // returns false if the name is already present; safe to call from several threads at once (ids are >= 0 and
    // below 2^31, so no word equals kEmpty)
    template <typename KeyOf>
    bool insert(sv name, int32_t id, KeyOf &&key_of) {
        const uint64_t h = hash_of(name), mine = word_of(h, id);
        for (uint64_t i = h & mask;; i = (i + 1) & mask) {
            uint64_t cur = __atomic_load_n(&slot[i], __ATOMIC_ACQUIRE);
            if (cur == kEmpty) {
                if (__atomic_compare_exchange_n(&slot[i], &cur, mine, false, __ATOMIC_ACQ_REL, __ATOMIC_ACQUIRE)) return true;
                // lost the race: `cur` now holds the winner's word
            }
            if ((cur >> 32) == (h >> 32) && key_of((int32_t)(uint32_t)cur) == name) return false;
        }
    }
    template <typename KeyOf>
    int32_t find(sv name, KeyOf &&key_of) const {
        const uint64_t h = hash_of(name);
        for (uint64_t i = h & mask;; i = (i + 1) & mask) {
            const uint64_t cur = slot[i];
            if (cur == kEmpty) return -1;
            if ((cur >> 32) == (h >> 32) && key_of((int32_t)(uint32_t)cur) == name) return (int32_t)(uint32_t)cur;
        }
    }
};

struct ParseError {
    std::string message;
};

// ---- field-level helpers ---------------------------------------------------------------------------------
sv unquote(sv f) {
    if (f.size() >= 2 && f.front() == '"' && f.back() == '"') return f.substr(1, f.size() - 2);
    return f;
}

bool parse_int(sv f, int64_t *out) {
    f = unquote(f);
    if (f.empty()) return false;
    const char *b = f.data(), *e = f.data() + f.size();
    if (*b == '+') ++b;
    auto r = std::from_chars(b, e, *out);
    return r.ec == std::errc() && r.ptr == e;
}

// polars reads an empty float field as null -> NaN (gecco/_base.py:127-130)
bool parse_float(sv f, double *out) {
    f = unquote(f);
    if (f.empty()) {
        *out = std::numeric_limits<double>::quiet_NaN();
        return true;
    }
    const char *b = f.data(), *e = f.data() + f.size();
    if (*b == '+') ++b;
    auto r = std::from_chars(b, e, *out);
    if (r.ec == std::errc::result_out_of_range) {  // 1e-400 and friends: what strtod would give
        *out = std::strtod(std::string(f).c_str(), nullptr);
        return true;
    }
    return r.ec == std::errc() && r.ptr == e;
}

struct Header {
    std::vector<sv> names;
    int find(const char *name) const {
        for (size_t i = 0; i < names.size(); ++i)
            if (names[i] == name) return (int)i;
        return -1;
    }
};

sv strip_cr(sv line) {
    if (!line.empty() && line.back() == '\r') line.remove_suffix(1);
    return line;
}

void split_tabs(sv line, std::vector<sv> &out) {
    out.clear();
    size_t b = 0;
    for (;;) {
        const size_t t = line.find('\t', b);
        if (t == sv::npos) {
            out.push_back(line.substr(b));
            return;
        }
        out.push_back(line.substr(b, t - b));
        b = t + 1;
    }
}

// Header line -> names; returns the offset of the first data line.
size_t read_header(sv buf, Header *h) {
    size_t nl = buf.find('\n');
    sv line = strip_cr(nl == sv::npos ? buf : buf.substr(0, nl));
    std::vector<sv> f;
    split_tabs(line, f);
    for (sv x : f) h->names.push_back(unquote(x));
    return nl == sv::npos ? buf.size() : nl + 1;
}

// What a row callback may remember from the previous line of its chunk (the rows of a gene usually follow each other).
struct LineMemo {
    sv name;
    int32_t row = -1;
};

// Calls make_row(fields, offset, memo) for every non-empty data line of buf[begin, end) on `threads` threads over
// newline-aligned chunks.  Two passes: the lines of every chunk are counted first, so that each thread writes its
// rows straight into its slice of the result (file order, no reallocation, no concatenation).
template <typename Row, typename Fn>
void parse_lines(sv buf, size_t begin, int threads, RowVector<Row> *rows, Fn &&make_row) {
    const size_t n = buf.size();
    if (begin >= n) return;
    if (threads < 1) threads = 1;
    const size_t min_chunk = 1u << 20;
    size_t want = (n - begin + min_chunk - 1) / min_chunk;
    if ((size_t)threads > want) threads = (int)want;
    std::vector<size_t> cut(threads + 1, n);
    cut[0] = begin;
    for (int t = 1; t < threads; ++t) {
        size_t p = begin + (n - begin) / threads * t;
        const size_t nl = buf.find('\n', p);
        cut[t] = nl == sv::npos ? n : nl + 1;
        if (cut[t] < cut[t - 1]) cut[t] = cut[t - 1];
    }
    auto next_line = [&](size_t p, size_t stop, sv *line) -> size_t {  // returns the offset after the line
        const void *hit = p < stop ? memchr(buf.data() + p, '\n', stop - p) : nullptr;
        const size_t nl = hit ? (size_t)(static_cast<const char *>(hit) - buf.data()) : stop;
        *line = strip_cr(buf.substr(p, nl - p));
        return nl + 1;
    };
    std::vector<size_t> count(threads + 1, 0);
    std::vector<std::string> errors(threads);
    auto run = [&](auto &&work) {
        if (threads == 1) {
            work(0);
        } else {
            std::vector<std::thread> pool;
            for (int t = 0; t < threads; ++t) pool.emplace_back(work, t);
            for (auto &th : pool) th.join();
        }
    };
    run([&](int t) {
        size_t c = 0;
        sv line;
        for (size_t p = cut[t]; p < cut[t + 1];) {
            p = next_line(p, cut[t + 1], &line);
            c += !line.empty();
        }
        count[t + 1] = c;
    });
    const size_t base = rows->size();
    for (int t = 0; t < threads; ++t) count[t + 1] += count[t];
    rows->resize(base + count[threads]);
    run([&](int t) {
        std::vector<sv> f;
        sv line;
        LineMemo memo;  // lives as long as this thread's pass over its chunk
        Row *out = rows->data() + base + count[t];
        try {
            for (size_t p = cut[t]; p < cut[t + 1];) {
                const size_t at = p;
                p = next_line(p, cut[t + 1], &line);
                if (line.empty()) continue;
                split_tabs(line, f);
                *out++ = make_row(f, at, memo);
            }
        } catch (const ParseError &e) {
            errors[t] = e.message;
        }
    });
    for (const auto &e : errors)
        if (!e.empty()) throw ParseError{e};
}

size_t line_number(sv buf, size_t offset) { return 1 + (size_t)std::count(buf.begin(), buf.begin() + offset, '\n'); }

// ---- Python repr() of a double -----------------------------------------------------------------------------
// Shortest digits that round-trip (std::to_chars), laid out as float.__repr__ does: fixed notation for
// 1e-4 <= |x| < 1e16 (always with a fractional part), otherwise d[.ddd]e+XX with at least two exponent digits.
void append_repr(std::string &out, double x) {
    if (std::isnan(x)) {
        out += "nan";
        return;
    }
    if (std::isinf(x)) {
        out += x < 0 ? "-inf" : "inf";
        return;
    }
    if (x == 0.0) {
        out += std::signbit(x) ? "-0.0" : "0.0";
        return;
    }
    // shortest round-trip digits in scientific form: [-]d[.ddd]e[+-]XX
    char sci[40];
    const char *end = std::to_chars(sci, sci + sizeof(sci), x, std::chars_format::scientific).ptr;
    const char *p = sci;
    char res[48];  // the laid-out number: at most sign + "0.000" + 17 digits, or 17 digits + '.' + "e-308"
    char *o = res;
    if (*p == '-') *o++ = *p++;
    char digits[20];
    int nd = 0;
    for (; *p != 'e'; ++p)
        if (*p != '.') digits[nd++] = *p;
    ++p;  // 'e'
    const bool eneg = *p == '-';
    ++p;  // sign (to_chars always writes one)
    int a = 0;
    for (; p < end; ++p) a = a * 10 + (*p - '0');
    const int exp10 = eneg ? -a : a;
    if (exp10 >= -4 && exp10 < 16) {
        if (exp10 < 0) {
            *o++ = '0';
            *o++ = '.';
            for (int k = 0; k < -exp10 - 1; ++k) *o++ = '0';
            memcpy(o, digits, (size_t)nd);
            o += nd;
        } else {
            const int ip = exp10 + 1;  // digits in front of the point
            if (nd <= ip) {
                memcpy(o, digits, (size_t)nd);
                o += nd;
                for (int k = nd; k < ip; ++k) *o++ = '0';
                *o++ = '.';
                *o++ = '0';
            } else {
                memcpy(o, digits, (size_t)ip);
                o += ip;
                *o++ = '.';
                memcpy(o, digits + ip, (size_t)(nd - ip));
                o += nd - ip;
            }
        }
    } else {
        *o++ = digits[0];
        if (nd > 1) {
            *o++ = '.';
            memcpy(o, digits + 1, (size_t)(nd - 1));
            o += nd - 1;
        }
        *o++ = 'e';
        *o++ = eneg ? '-' : '+';
        if (a < 10) *o++ = '0';
        o = std::to_chars(o, res + sizeof(res), a).ptr;
    }
    out.append(res, (size_t)(o - res));
}

void append_int(std::string &out, int64_t v) {
    char buf[24];
    auto r = std::to_chars(buf, buf + sizeof(buf), v);
    out.append(buf, (size_t)(r.ptr - buf));
}

}  // namespace

// A file's bytes; not zero-filled before the read (a std::string would be), never copied.
struct FileBuffer {
    char *data = nullptr;
    size_t size = 0;
    FileBuffer() = default;
    FileBuffer(const FileBuffer &) = delete;
    FileBuffer &operator=(const FileBuffer &) = delete;
    FileBuffer(FileBuffer &&o) noexcept : data(o.data), size(o.size) { o.data = nullptr; o.size = 0; }
    FileBuffer &operator=(FileBuffer &&o) noexcept {
        if (this != &o) {
            free(data);
            data = o.data; size = o.size;
            o.data = nullptr; o.size = 0;
        }
        return *this;
    }
    ~FileBuffer() { free(data); }
    bool allocate(size_t n) {
        free(data);
        data = static_cast<char *>(big_alloc(n));
        size = data ? n : 0;
        return data != nullptr;
    }
    sv view() const { return sv(data, size); }
};

struct gcrf_table {
    std::vector<FileBuffer> buffers;  // the files; every string_view below points into one of them
    // genes, in the reference's order (sequence_id, start, end)
    RowVector<GeneRow> genes;
    std::vector<int32_t> contig_ptr;           // [C+1] into genes
    // contig c is called genes[contig_ptr[c]].seq; NUL-terminated copies are made when gcrf_table_contig_id first asks
    // (a metagenome table has a million of them)
    mutable std::vector<std::string> contig_ids;
    size_t contigs() const { return genes.empty() ? 0 : contig_ptr.size() - 1; }
    sv contig_name(size_t c) const { return genes[(size_t)contig_ptr[c]].seq; }
    std::vector<std::string> gene_ids;         // [G] NUL-terminated copies, built lazily for the accessor
    // domain rows kept by the filters, grouped by gene, ordered by (domain_start, domain_end)
    RowVector<DomainRow> domains;
    std::vector<int64_t> dom_ptr;              // [G+1] into domains
    std::vector<uint8_t> annotated;            // [G] gene has >= 1 domain left
    // gcrf_table_pack results
    int32_t packed_mode = -1;
    std::vector<int32_t> row_contig_ptr, row_ptr, attr_idx, row_gene;
};

namespace {

thread_local char t_error[512] = "";

int tfail(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(t_error, sizeof(t_error), fmt, ap);
    va_end(ap);
    return code;
}

int default_threads();

// The whole file into one malloc'd buffer; files of more than 16 MB are read by several threads (pread into disjoint
// slices: the copies out of the page cache and the first touch of the buffer's pages run in parallel).
int read_file(const char *path, FileBuffer *out) {
    const int fd = open(path, O_RDONLY);
    if (fd < 0) return tfail(GCRF_EINVAL, "cannot open %s", path);
    const off_t n = lseek(fd, 0, SEEK_END);
    if (n < 0 || !out->allocate((size_t)(n > 0 ? n : 0))) {
        close(fd);
        return n < 0 ? tfail(GCRF_EINVAL, "cannot read %s", path) : tfail(GCRF_ENOMEM, "out of host memory reading %s", path);
    }
    const size_t size = out->size;
    auto slice = [&](size_t b, size_t e) {
        while (b < e) {
            const ssize_t got = pread(fd, out->data + b, e - b, (off_t)b);
            if (got < 0 && errno == EINTR) continue;
            if (got <= 0) return false;
            b += (size_t)got;
        }
        return true;
    };
    const int nt = (int)std::min<size_t>((size_t)default_threads(), size >> 24);
    std::atomic<bool> ok{true};
    if (nt <= 1) {
        ok = slice(0, size);
    } else {
        std::vector<std::thread> pool;
        for (int k = 0; k < nt; ++k)
            pool.emplace_back([&, k] {
                if (!slice(size * k / nt, size * (k + 1) / nt)) ok.store(false);
            });
        for (auto &th : pool) th.join();
    }
    close(fd);
    if (!ok.load()) return tfail(GCRF_EINVAL, "short read on %s", path);
    return GCRF_OK;
}

int column(const Header &h, const char *name, const char *what) {
    const int i = h.find(name);
    if (i < 0) throw ParseError{std::string(what) + " table has no column '" + name + "'"};
    return i;
}

struct PhaseTimer {
    bool on = getenv("GCRF_TABLE_TIMING") != nullptr;
    std::chrono::steady_clock::time_point last = std::chrono::steady_clock::now();
    void mark(const char *what) {
        if (!on) return;
        const auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "[gcrf tables] %-28s %8.1f ms\n", what, std::chrono::duration<double, std::milli>(now - last).count());
        last = now;
    }
};

void build(gcrf_table *t, sv genes_buf, const std::vector<sv> &feature_bufs, double e_filter, double p_filter, int threads) {
    PhaseTimer timer;
    // ---- genes table
    RowVector<GeneRow> rows;
    {
        Header h;
        const size_t body = read_header(genes_buf, &h);
        const int c_seq = column(h, "sequence_id", "genes"), c_prot = column(h, "protein_id", "genes"),
                  c_start = column(h, "start", "genes"), c_end = column(h, "end", "genes"),
                  c_strand = column(h, "strand", "genes");
        const size_t need = (size_t)std::max({c_seq, c_prot, c_start, c_end, c_strand}) + 1;
        parse_lines(genes_buf, body, threads, &rows, [&](const std::vector<sv> &f, size_t off, LineMemo &) {
            if (f.size() < need) throw ParseError{"genes table: line " + std::to_string(line_number(genes_buf, off)) + " has too few fields"};
            GeneRow g;
            g.seq = unquote(f[c_seq]);
            g.prot = unquote(f[c_prot]);
            // GeneTable.to_genes: Strand.Coding if strand == "+" else Strand.Reverse (gecco/model.py:826)
            g.strand = unquote(f[c_strand]) == "+" ? sv("+") : sv("-");
            if (!parse_int(f[c_start], &g.start) || !parse_int(f[c_end], &g.end))
                throw ParseError{"genes table: bad coordinate on line " + std::to_string(line_number(genes_buf, off))};
            return g;
        });
    }
    timer.mark("parse genes");
    if (rows.size() > 0x7fffff00u) throw ParseError{"too many genes for int32 row pointers; shard the table"};
    // annotate_genes: gene names are unique (_common.py:217-219)
    auto prot_of = [&](int32_t i) { return rows[(size_t)i].prot; };
    NameIndex by_name;
    by_name.reserve(rows.size());
    {
        const size_t n = rows.size();
        const int nt = (int)std::min<size_t>((size_t)threads, std::max<size_t>(1, n / 50000));
        std::atomic<bool> duplicate{false};
        auto fill = [&](size_t i0, size_t i1) {
            for (size_t i = i0; i < i1; ++i)
                if (!by_name.insert(rows[i].prot, (int32_t)i, prot_of)) duplicate.store(true);
        };
        if (nt <= 1) {
            fill(0, n);
        } else {
            std::vector<std::thread> pool;
            for (int k = 0; k < nt; ++k) pool.emplace_back(fill, n * k / nt, n * (k + 1) / nt);
            for (auto &th : pool) th.join();
        }
        if (duplicate.load()) throw ParseError{"Duplicate gene names in input genes"};
    }
    timer.mark("index genes by name");
    // sort by (sequence_id, start, end), stable (predict.py:81).  The ids are compared once per CONTIG, not once per
    // comparison: distinct ids are ranked, genes are bucketed by rank (stable), and every bucket is sorted by numbers.
    const size_t G = rows.size();
    std::vector<int32_t> contig_of(G);
    std::vector<sv> ids;  // distinct sequence ids in order of first appearance
    {
        // Runs of equal ids (tables are usually grouped by contig) are found by all threads; the first row of every run
        // is a candidate, candidates are de-duplicated through a concurrently filled name index, and the smallest
        // candidate of a name stands for it — the ids come out in order of first appearance whatever the thread timing.
        const int nt = (int)std::min<size_t>((size_t)threads, std::max<size_t>(1, G / 50000));
        auto in_threads = [&](auto &&fn) {
            if (nt <= 1) {
                fn(0);
                return;
            }
            std::vector<std::thread> pool;
            for (int k = 0; k < nt; ++k) pool.emplace_back(fn, k);
            for (auto &th : pool) th.join();
        };
        std::vector<std::vector<int32_t>> starts((size_t)nt);
        in_threads([&](int k) {
            const size_t i0 = G * (size_t)k / (size_t)nt, i1 = G * (size_t)(k + 1) / (size_t)nt;
            for (size_t i = i0; i < i1; ++i)
                if (i == i0 || rows[i].seq != rows[i - 1].seq) starts[(size_t)k].push_back((int32_t)i);
        });
        std::vector<size_t> first_cand((size_t)nt + 1, 0);
        for (int k = 0; k < nt; ++k) first_cand[(size_t)k + 1] = first_cand[(size_t)k] + starts[(size_t)k].size();
        const size_t ncand = first_cand[(size_t)nt];
        std::vector<int32_t> cand(ncand), canon(ncand);
        for (int k = 0; k < nt; ++k) std::copy(starts[(size_t)k].begin(), starts[(size_t)k].end(), cand.begin() + (std::ptrdiff_t)first_cand[(size_t)k]);
        auto name_of = [&](int32_t c) { return rows[(size_t)cand[(size_t)c]].seq; };
        NameIndex seen;
        seen.reserve(ncand);
        std::vector<char> fresh(ncand, 0);
        in_threads([&](int k) {
            for (size_t c = first_cand[(size_t)k]; c < first_cand[(size_t)k + 1]; ++c)
                fresh[c] = seen.insert(name_of((int32_t)c), (int32_t)c, name_of) ? 1 : 0;
        });
        // every candidate learns which one holds its name in the index; the smallest of a name becomes its canonical one
        std::vector<int32_t> smallest(ncand);
        std::iota(smallest.begin(), smallest.end(), 0);
        in_threads([&](int k) {
            for (size_t c = first_cand[(size_t)k]; c < first_cand[(size_t)k + 1]; ++c) {
                const int32_t holder = fresh[c] ? (int32_t)c : seen.find(name_of((int32_t)c), name_of);
                canon[c] = holder;
                int32_t cur = __atomic_load_n(&smallest[(size_t)holder], __ATOMIC_RELAXED);
                while ((int32_t)c < cur &&
                       !__atomic_compare_exchange_n(&smallest[(size_t)holder], &cur, (int32_t)c, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {
                }
            }
        });
        std::vector<int32_t> id_of_cand(ncand, -1);
        for (size_t c = 0; c < ncand; ++c)
            if (smallest[(size_t)canon[c]] == (int32_t)c) {
                id_of_cand[c] = (int32_t)ids.size();
                ids.push_back(name_of((int32_t)c));
            }
        in_threads([&](int k) {
            const size_t i1 = G * (size_t)(k + 1) / (size_t)nt;
            for (size_t c = first_cand[(size_t)k]; c < first_cand[(size_t)k + 1]; ++c) {
                const int32_t id = id_of_cand[(size_t)smallest[(size_t)canon[c]]];
                const size_t e = c + 1 < first_cand[(size_t)k + 1] ? (size_t)cand[c + 1] : i1;
                for (size_t i = (size_t)cand[c]; i < e; ++i) contig_of[i] = id;
            }
        });
    }
    timer.mark("  distinct contig ids");
    const size_t C = ids.size();
    std::vector<int32_t> by_id(C), rank_of(C);
    std::iota(by_id.begin(), by_id.end(), 0);
    {
        // a metagenome table has a million distinct ids: nothing to do when they already come in order, else sorted
        // slices on all threads, merged pairwise
        auto less = [&](int32_t a, int32_t b) { return ids[(size_t)a] < ids[(size_t)b]; };
        bool in_order = true;
        for (size_t r = 1; r < C && in_order; ++r) in_order = !(ids[r] < ids[r - 1]);
        int nt = (int)std::min<size_t>((size_t)threads, C / 20000);
        if (in_order) {
            // ids are distinct: the order of first appearance is the sorted order
        } else if (nt <= 1) {
            std::sort(by_id.begin(), by_id.end(), less);
        } else {
            int pow2 = 1;
            while (pow2 * 2 <= nt) pow2 *= 2;
            nt = pow2;
            auto cut = [&](int k) { return by_id.begin() + (std::ptrdiff_t)(C * (size_t)k / (size_t)nt); };
            {
                std::vector<std::thread> pool;
                for (int k = 0; k < nt; ++k) pool.emplace_back([&, k] { std::sort(cut(k), cut(k + 1), less); });
                for (auto &th : pool) th.join();
            }
            for (int width = 1; width < nt; width *= 2) {
                std::vector<std::thread> pool;
                for (int k = 0; k + width < nt; k += 2 * width)
                    pool.emplace_back([&, k, width] { std::inplace_merge(cut(k), cut(k + width), cut(std::min(nt, k + 2 * width)), less); });
                for (auto &th : pool) th.join();
            }
        }
    }
    for (size_t r = 0; r < C; ++r) rank_of[(size_t)by_id[r]] = (int32_t)r;
    timer.mark("  rank contig ids");
    t->contig_ptr.assign(C + 1, 0);
    for (size_t i = 0; i < G; ++i) ++t->contig_ptr[(size_t)rank_of[(size_t)contig_of[i]] + 1];
    for (size_t r = 0; r < C; ++r) t->contig_ptr[r + 1] += t->contig_ptr[r];
    std::vector<int32_t> order(G);
    {
        std::vector<int32_t> cursor(t->contig_ptr.begin(), t->contig_ptr.end() - 1);
        for (size_t i = 0; i < G; ++i) order[(size_t)cursor[(size_t)rank_of[(size_t)contig_of[i]]]++] = (int32_t)i;
    }
    {
        auto sort_contigs = [&](size_t r0, size_t r1) {
            for (size_t r = r0; r < r1; ++r)
                std::stable_sort(order.begin() + t->contig_ptr[r], order.begin() + t->contig_ptr[r + 1], [&](int32_t a, int32_t b) {
                    const GeneRow &x = rows[(size_t)a], &y = rows[(size_t)b];
                    if (x.start != y.start) return x.start < y.start;
                    return x.end < y.end;
                });
        };
        const int nt = (int)std::min<size_t>((size_t)threads, std::max<size_t>(1, G / 50000));
        if (nt <= 1) {
            sort_contigs(0, C);
        } else {  // contiguous contig ranges of about equal gene counts
            std::vector<std::thread> pool;
            size_t r0 = 0;
            for (int k = 0; k < nt; ++k) {
                const int64_t want = (int64_t)G * (k + 1) / nt;
                size_t r1 = r0;
                while (r1 < C && t->contig_ptr[r1 + 1] <= want) ++r1;
                if (k == nt - 1) r1 = C;
                pool.emplace_back(sort_contigs, r0, r1);
                r0 = r1;
            }
            for (auto &th : pool) th.join();
        }
    }
    timer.mark("  bucket + sort inside contigs");
    std::vector<int32_t> rank(G);
    t->genes.resize(G);
    {
        auto gather = [&](size_t k0, size_t k1) {
            for (size_t k = k0; k < k1; ++k) {
                t->genes[k] = rows[(size_t)order[k]];
                rank[(size_t)order[k]] = (int32_t)k;
            }
        };
        const int nt = (int)std::min<size_t>((size_t)threads, std::max<size_t>(1, G / 50000));
        if (nt <= 1) {
            gather(0, G);
        } else {
            std::vector<std::thread> pool;
            for (int k = 0; k < nt; ++k) pool.emplace_back(gather, G * k / nt, G * (k + 1) / nt);
            for (auto &th : pool) th.join();
        }
    }
    if (G == 0) t->contig_ptr.assign(1, 0);
    timer.mark("sort genes, contigs");

    // ---- feature tables, concatenated in the order given (load_features, _common.py:193-208).  Every row is
    //      attached to its gene (annotate_genes, _common.py:226-249) and filtered (filter_domains, :419-448: NaN < x is
    //      false, so NaN rows go as well) by the thread that parses it; what is stored is the compact DomainRow.
    RowVector<DomainRow> drows;
    for (sv fb : feature_bufs) {
        Header h;
        const size_t body = read_header(fb, &h);
        if (h.names.empty() || (h.names.size() == 1 && h.names[0].empty())) continue;  // an empty file
        const int c_seq = column(h, "sequence_id", "features"), c_prot = column(h, "protein_id", "features"),
                  c_start = column(h, "start", "features"), c_end = column(h, "end", "features"),
                  c_strand = column(h, "strand", "features"), c_dom = column(h, "domain", "features"),
                  c_hmm = column(h, "hmm", "features"), c_ev = column(h, "i_evalue", "features"),
                  c_pv = column(h, "pvalue", "features"), c_ds = column(h, "domain_start", "features"),
                  c_de = column(h, "domain_end", "features");
        const size_t need = (size_t)std::max({c_seq, c_prot, c_start, c_end, c_strand, c_dom, c_hmm, c_ev, c_pv, c_ds, c_de}) + 1;
        parse_lines(fb, body, threads, &drows, [&](const std::vector<sv> &f, size_t off, LineMemo &memo) {
            auto where = [&] { return " on line " + std::to_string(line_number(fb, off)) + " of a features table"; };
            if (f.size() < need) throw ParseError{"too few fields" + where()};
            DomainRow d;
            d.name = unquote(f[c_dom]);
            d.hmm = unquote(f[c_hmm]);
            int64_t start, end;
            if (!parse_int(f[c_start], &start) || !parse_int(f[c_end], &end) || !parse_int(f[c_ds], &d.dstart) ||
                !parse_int(f[c_de], &d.dend))
                throw ParseError{"bad integer" + where()};
            if (!parse_float(f[c_ev], &d.i_evalue) || !parse_float(f[c_pv], &d.pvalue)) throw ParseError{"bad number" + where()};
            const sv prot = unquote(f[c_prot]), seq = unquote(f[c_seq]), strand = unquote(f[c_strand]);
            // the rows of a gene usually follow each other: remember the last answer (per parsing thread)
            int32_t row;
            if (memo.row >= 0 && prot == memo.name) {
                row = memo.row;
            } else {
                row = by_name.find(prot, prot_of);
                memo.name = prot;
                memo.row = row;
            }
            if (row < 0) throw ParseError{"Unknown protein " + std::string(prot) + " in features table"};
            const GeneRow &g = rows[(size_t)row];
            if (g.seq != seq || g.end - g.start != end - start || g.start != start || g.end != end || g.strand != strand) {
                const std::string id(prot);
                if (g.seq != seq) throw ParseError{"Mismatched source sequence for '" + id + "': '" + std::string(g.seq) + "' != '" + std::string(seq) + "'"};
                if (g.end - g.start != end - start)
                    throw ParseError{"Mismatched gene length for '" + id + "': " + std::to_string(g.end - g.start) + " != " + std::to_string(end - start)};
                if (g.start != start) throw ParseError{"Mismatched gene start for '" + id + "': " + std::to_string(g.start) + " != " + std::to_string(start)};
                if (g.end != end) throw ParseError{"Mismatched gene end for '" + id + "': " + std::to_string(g.end) + " != " + std::to_string(end)};
                throw ParseError{"Mismatched gene strand for '" + id + "': '" + std::string(g.strand) + "' != '" + std::string(strand) + "'"};
            }
            const bool keep = (std::isnan(e_filter) || d.i_evalue < e_filter) && (std::isnan(p_filter) || d.pvalue < p_filter);
            d.gene = keep ? rank[(size_t)row] : -1;
            return d;
        });
    }
    timer.mark("parse + annotate features");
    // group by gene keeping the row order (counting sort), then order every gene's rows by (start, end), stable
    t->dom_ptr.assign(G + 1, 0);
    {
        // Rows that already arrive in gene order (sorted tables: the usual case) only need compaction, which every
        // thread can do for its own slice; otherwise a sequential, order-preserving scatter.
        const size_t R = drows.size();
        const int nt = (int)std::min<size_t>((size_t)threads, std::max<size_t>(1, R / 200000));
        std::vector<size_t> kept(nt + 1, 0);
        std::vector<char> sorted(nt, 1);
        std::vector<int32_t> first(nt, -1), last(nt, -1);
        auto scan = [&](int k) {
            size_t c = 0;
            int32_t prev = -1;
            for (size_t i = R * k / nt; i < R * (k + 1) / nt; ++i) {
                const int32_t g = drows[i].gene;
                if (g < 0) continue;
                if (first[k] < 0) first[k] = g;
                if (g < prev) sorted[k] = 0;
                prev = g;
                ++c;
            }
            last[k] = prev;
            kept[k + 1] = c;
        };
        auto in_threads = [&](auto &&fn) {
            if (nt <= 1) {
                fn(0);
                return;
            }
            std::vector<std::thread> pool;
            for (int k = 0; k < nt; ++k) pool.emplace_back(fn, k);
            for (auto &th : pool) th.join();
        };
        in_threads(scan);
        bool in_order = true;
        int32_t prev = -1;
        for (int k = 0; k < nt; ++k) {
            in_order = in_order && sorted[k] && (first[k] < 0 || first[k] >= prev);
            if (last[k] >= 0) prev = last[k];
            kept[k + 1] += kept[k];
        }
        if (in_order) {
            // compaction per slice; the row pointers follow from where the gene changes in the compacted rows:
            // dom_ptr[g] = first kept row of a gene >= g, filled slice by slice (a slice owns the genes from the one
            // after the previous kept row's up to its own last row's)
            t->domains.resize(kept[nt]);
            in_threads([&](int k) {
                DomainRow *out = t->domains.data() + kept[k];
                int32_t prev_gene = -1;  // gene of the last kept row in front of this slice
                for (int j = k - 1; j >= 0 && prev_gene < 0; --j) prev_gene = last[j];
                size_t pos = kept[k];
                for (size_t i = R * k / nt; i < R * (k + 1) / nt; ++i) {
                    const int32_t g = drows[i].gene;
                    if (g < 0) continue;
                    *out++ = drows[i];
                    for (int32_t q = prev_gene + 1; q <= g; ++q) t->dom_ptr[(size_t)q] = (int64_t)pos;
                    prev_gene = g;
                    ++pos;
                }
            });
            int32_t last_gene = -1;
            for (int j = nt - 1; j >= 0 && last_gene < 0; --j) last_gene = last[j];
            for (size_t q = (size_t)(last_gene + 1); q <= G; ++q) t->dom_ptr[q] = (int64_t)kept[nt];
        } else {
            for (const DomainRow &d : drows)
                if (d.gene >= 0) ++t->dom_ptr[(size_t)d.gene + 1];
            for (size_t g = 0; g < G; ++g) t->dom_ptr[g + 1] += t->dom_ptr[g];
            t->domains.resize((size_t)t->dom_ptr[G]);
            std::vector<int64_t> cursor(t->dom_ptr.begin(), t->dom_ptr.end() - 1);
            for (const DomainRow &d : drows)
                if (d.gene >= 0) t->domains[(size_t)cursor[(size_t)d.gene]++] = d;
        }
    }
    drows.clear();
    drows.shrink_to_fit();
    t->annotated.assign(G, 0);
    {
        auto finish = [&](size_t g0, size_t g1) {
            for (size_t g = g0; g < g1; ++g) {
                auto b = t->domains.begin() + t->dom_ptr[g], e = t->domains.begin() + t->dom_ptr[g + 1];
                t->annotated[g] = b != e;
                if (e - b > 1)
                    std::stable_sort(b, e, [](const DomainRow &x, const DomainRow &y) {
                        if (x.dstart != y.dstart) return x.dstart < y.dstart;
                        return x.dend < y.dend;
                    });
            }
        };
        const int nt = (int)std::min<size_t>((size_t)threads, std::max<size_t>(1, G / 50000));
        if (nt <= 1) {
            finish(0, G);
        } else {
            std::vector<std::thread> pool;
            for (int k = 0; k < nt; ++k) pool.emplace_back(finish, G * k / nt, G * (k + 1) / nt);
            for (auto &th : pool) th.join();
        }
    }
    timer.mark("group + sort domains");
}

int default_threads() {
    if (const char *env = getenv("GCRF_TABLE_THREADS")) {
        const int v = atoi(env);
        if (v >= 1) return v > 64 ? 64 : v;
    }
    unsigned n = std::thread::hardware_concurrency();
    if (n == 0) n = 1;
    if (n > 32) n = 32;
    return (int)n;
}

int load_buffers(std::vector<FileBuffer> &&bufs, double e_filter, double p_filter, gcrf_table **out) {
    gcrf_table *t = new (std::nothrow) gcrf_table();
    if (!t) return tfail(GCRF_ENOMEM, "out of host memory");
    t->buffers = std::move(bufs);
    std::vector<sv> feats;
    for (size_t i = 1; i < t->buffers.size(); ++i) feats.emplace_back(t->buffers[i].view());
    try {
        build(t, t->buffers[0].view(), feats, e_filter, p_filter, default_threads());
    } catch (const ParseError &e) {
        delete t;
        return tfail(GCRF_EINVAL, "%s", e.message.c_str());
    } catch (const std::bad_alloc &) {
        delete t;
        return tfail(GCRF_ENOMEM, "out of host memory");
    }
    *out = t;
    return GCRF_OK;
}

// per-gene (average_p, max_p) from per-row probabilities; NaN = none (Gene.average_probability /
// maximum_probability, gecco/model.py:274-290)
void gene_probabilities(const gcrf_table *t, const double *row_prob, size_t g, double *avg, double *mx, int64_t *row_cursor) {
    const double nan = std::numeric_limits<double>::quiet_NaN();
    if (!row_prob) {
        *avg = *mx = nan;
        return;
    }
    const int64_t nd = t->dom_ptr[g + 1] - t->dom_ptr[g];
    if (t->packed_mode == 0 || nd == 0) {
        const double p = row_prob[(*row_cursor)++];  // with_probability: the gene's own value
        *avg = *mx = p;
        return;
    }
    long double sum = 0;
    double best = nan;
    int64_t n = 0;
    for (int64_t k = 0; k < nd; ++k) {
        const double p = row_prob[(*row_cursor)++];
        if (std::isnan(p)) continue;
        sum += p;
        best = n == 0 ? p : std::max(best, p);
        ++n;
    }
    *avg = n ? (double)(sum / n) : nan;
    *mx = best;
}

int write_all(const char *path, const std::string &text) {
    FILE *f = fopen(path, "wb");
    if (!f) return tfail(GCRF_EINVAL, "cannot open %s for writing", path);
    const size_t put = fwrite(text.data(), 1, text.size(), f);
    const int rc = fclose(f);
    if (put != text.size() || rc != 0) return tfail(GCRF_EINVAL, "short write on %s", path);
    return GCRF_OK;
}

// header + the rows of genes [0, G): `emit(out, g)` appends gene g's lines; gene ranges are formatted by separate
// threads into their own strings and written in order.
template <typename Emit>
int write_by_genes(const char *path, const std::string &header, size_t G, size_t bytes_per_gene, Emit &&emit) {
    // Genes are cut into chunks handed out in order; a thread formats its chunk into a buffer it reuses (no fresh pages
    // per chunk), learns its file offset from the chunk in front (offset + size, published as soon as that chunk is
    // formatted) and writes with pwrite, so formatting and the copies into the page cache both run in parallel.
    PhaseTimer timer;
    const int fd = open(path, O_WRONLY | O_CREAT | O_TRUNC, 0644);
    if (fd < 0) return tfail(GCRF_EINVAL, "cannot open %s for writing", path);
    auto put = [fd](const char *data, size_t n, int64_t at) {
        while (n) {
            const ssize_t w = pwrite(fd, data, n, (off_t)at);
            if (w < 0) {
                if (errno == EINTR) continue;
                return false;
            }
            data += w;
            n -= (size_t)w;
            at += w;
        }
        return true;
    };
    std::atomic<bool> ok{put(header.data(), header.size(), 0)};
    const size_t chunk_genes = std::max<size_t>(256, std::min<size_t>(8192, (1u << 20) / std::max<size_t>(1, bytes_per_gene)));
    const size_t nchunks = (G + chunk_genes - 1) / chunk_genes;
    const int nt = (int)std::max<size_t>(1, std::min<size_t>((size_t)default_threads(), nchunks / 2));
    std::vector<std::atomic<int64_t>> chunk_end(nchunks);
    for (auto &e : chunk_end) e.store(-1, std::memory_order_relaxed);
    std::atomic<size_t> next_chunk{0};
    auto work = [&]() {
        std::string out;
        out.reserve(chunk_genes * bytes_per_gene + 4096);
        for (;;) {
            const size_t i = next_chunk.fetch_add(1, std::memory_order_relaxed);
            if (i >= nchunks) break;
            out.clear();
            const size_t g1 = std::min(G, (i + 1) * chunk_genes);
            for (size_t g = i * chunk_genes; g < g1; ++g) emit(out, g);
            int64_t at = (int64_t)header.size();
            if (i > 0)
                while ((at = chunk_end[i - 1].load(std::memory_order_acquire)) < 0) std::this_thread::yield();
            chunk_end[i].store(at + (int64_t)out.size(), std::memory_order_release);
            if (!put(out.data(), out.size(), at)) ok.store(false);
        }
    };
    if (nt <= 1) {
        work();
    } else {
        std::vector<std::thread> pool;
        for (int k = 0; k < nt; ++k) pool.emplace_back(work);
        for (auto &th : pool) th.join();
    }
    const bool closed = close(fd) == 0;
    timer.mark("format + write rows");
    if (!ok.load() || !closed) return tfail(GCRF_EINVAL, "short write on %s", path);
    return GCRF_OK;
}

}  // namespace

extern "C" {

const char *gcrf_table_last_error(void) { return t_error; }

int gcrf_table_load(const char *genes_tsv, const char *const *features_tsv, int32_t n_features, double e_filter,
                    double p_filter, gcrf_table **out) {
    if (!out) return tfail(GCRF_EINVAL, "out is NULL");
    *out = nullptr;
    if (!genes_tsv || n_features < 0 || (n_features > 0 && !features_tsv)) return tfail(GCRF_EINVAL, "bad arguments");
    std::vector<FileBuffer> bufs((size_t)n_features + 1);
    int rc = read_file(genes_tsv, &bufs[0]);
    if (rc != GCRF_OK) return rc;
    for (int32_t i = 0; i < n_features; ++i) {
        if (!features_tsv[i]) return tfail(GCRF_EINVAL, "features path %d is NULL", i);
        rc = read_file(features_tsv[i], &bufs[(size_t)i + 1]);
        if (rc != GCRF_OK) return rc;
    }
    return load_buffers(std::move(bufs), e_filter, p_filter, out);
}

int gcrf_table_parse(const char *genes, uint64_t genes_len, const char *const *features, const uint64_t *features_len,
                     int32_t n_features, double e_filter, double p_filter, gcrf_table **out) {
    if (!out) return tfail(GCRF_EINVAL, "out is NULL");
    *out = nullptr;
    if ((!genes && genes_len) || n_features < 0 || (n_features > 0 && (!features || !features_len)))
        return tfail(GCRF_EINVAL, "bad arguments");
    std::vector<FileBuffer> bufs((size_t)n_features + 1);
    auto copy_in = [](FileBuffer &b, const char *p, uint64_t n) {
        if (!b.allocate((size_t)n)) return false;
        if (n) memcpy(b.data, p, (size_t)n);
        return true;
    };
    if (!copy_in(bufs[0], genes, genes_len)) return tfail(GCRF_ENOMEM, "out of host memory");
    for (int32_t i = 0; i < n_features; ++i) {
        if (features_len[i] && !features[i]) return tfail(GCRF_EINVAL, "features buffer %d is NULL", i);
        if (!copy_in(bufs[(size_t)i + 1], features[i], features_len[i])) return tfail(GCRF_ENOMEM, "out of host memory");
    }
    return load_buffers(std::move(bufs), e_filter, p_filter, out);
}

void gcrf_table_destroy(gcrf_table *t) { delete t; }

int64_t gcrf_table_contigs(const gcrf_table *t) { return t ? (int64_t)t->contigs() : 0; }
int64_t gcrf_table_genes(const gcrf_table *t) { return t ? (int64_t)t->genes.size() : 0; }
int64_t gcrf_table_domains(const gcrf_table *t) { return t ? (int64_t)t->domains.size() : 0; }

const char *gcrf_table_contig_id(const gcrf_table *t, int64_t c) {
    if (!t || c < 0 || c >= (int64_t)t->contigs()) return nullptr;
    if (t->contig_ids.empty()) {
        t->contig_ids.reserve(t->contigs());
        for (size_t k = 0; k < t->contigs(); ++k) t->contig_ids.emplace_back(t->contig_name(k));
    }
    return t->contig_ids[(size_t)c].c_str();
}

const char *gcrf_table_gene_id(gcrf_table *t, int64_t g) {
    if (!t || g < 0 || g >= (int64_t)t->genes.size()) return nullptr;
    if (t->gene_ids.empty()) {
        t->gene_ids.reserve(t->genes.size());
        for (const GeneRow &r : t->genes) t->gene_ids.emplace_back(r.prot);
    }
    return t->gene_ids[(size_t)g].c_str();
}

const int32_t *gcrf_table_contig_ptr(const gcrf_table *t) { return t && !t->contig_ptr.empty() ? t->contig_ptr.data() : nullptr; }
const uint8_t *gcrf_table_annotated(const gcrf_table *t) { return t && !t->annotated.empty() ? t->annotated.data() : nullptr; }

int gcrf_table_gene_coordinates(const gcrf_table *t, int64_t *start, int64_t *end) {
    if (!t) return tfail(GCRF_EINVAL, "table is NULL");
    for (size_t g = 0; g < t->genes.size(); ++g) {
        if (start) start[g] = t->genes[g].start;
        if (end) end[g] = t->genes[g].end;
    }
    return GCRF_OK;
}

int gcrf_table_pack(gcrf_table *t, const char *const *attr_names, int32_t A, int32_t feature_type,
                    const int32_t **contig_ptr, const int32_t **row_ptr, const int32_t **attr_idx, int64_t *rows,
                    int64_t *nnz) {
    if (!t) return tfail(GCRF_EINVAL, "table is NULL");
    if (A < 0 || (A > 0 && !attr_names)) return tfail(GCRF_EINVAL, "bad vocabulary");
    if (feature_type != 0 && feature_type != 1) return tfail(GCRF_EINVAL, "invalid feature type: %d", feature_type);
    try {
        std::unordered_map<sv, int32_t> vocab;
        vocab.reserve((size_t)A * 2);
        for (int32_t a = 0; a < A; ++a) {
            if (!attr_names[a]) return tfail(GCRF_EINVAL, "attribute name %d is NULL", a);
            if (!vocab.emplace(sv(attr_names[a]), a).second) return tfail(GCRF_EINVAL, "attribute %s appears twice", attr_names[a]);
        }
        const size_t G = t->genes.size(), C = t->contigs();
        t->row_ptr.assign(1, 0);
        t->attr_idx.clear();
        t->row_gene.clear();
        t->row_contig_ptr.assign(1, 0);
        std::vector<int32_t> seen((size_t)A, -1);  // last gene that used the attribute: the per-gene set
        size_t c = 0;
        for (size_t g = 0; g < G; ++g) {
            while (c + 1 < C && (int64_t)g >= t->contig_ptr[c + 1]) {
                t->row_contig_ptr.push_back((int32_t)(t->row_ptr.size() - 1));
                ++c;
            }
            const int64_t b = t->dom_ptr[g], e = t->dom_ptr[g + 1];
            if (feature_type == 0) {
                for (int64_t k = b; k < e; ++k) {
                    auto it = vocab.find(t->domains[(size_t)k].name);
                    if (it == vocab.end() || seen[(size_t)it->second] == (int32_t)g) continue;
                    seen[(size_t)it->second] = (int32_t)g;
                    t->attr_idx.push_back(it->second);
                }
                t->row_ptr.push_back((int32_t)t->attr_idx.size());
                t->row_gene.push_back((int32_t)g);
            } else {
                for (int64_t k = b; k < e; ++k) {
                    auto it = vocab.find(t->domains[(size_t)k].name);
                    if (it != vocab.end()) t->attr_idx.push_back(it->second);
                    t->row_ptr.push_back((int32_t)t->attr_idx.size());
                    t->row_gene.push_back((int32_t)g);
                }
                if (b == e) {  // a gene without domains is one empty position (features.py:44-46)
                    t->row_ptr.push_back((int32_t)t->attr_idx.size());
                    t->row_gene.push_back((int32_t)g);
                }
            }
            if (t->attr_idx.size() > 0x7fffff00u || t->row_ptr.size() > 0x7fffff00u)
                return tfail(GCRF_EINVAL, "table too large for int32 row pointers; shard it");
        }
        if (G) t->row_contig_ptr.push_back((int32_t)(t->row_ptr.size() - 1));
        t->packed_mode = feature_type;
    } catch (const std::bad_alloc &) {
        return tfail(GCRF_ENOMEM, "out of host memory");
    }
    if (contig_ptr) *contig_ptr = t->row_contig_ptr.data();
    if (row_ptr) *row_ptr = t->row_ptr.data();
    if (attr_idx) *attr_idx = t->attr_idx.data();
    if (rows) *rows = (int64_t)t->row_ptr.size() - 1;
    if (nnz) *nnz = (int64_t)t->attr_idx.size();
    return GCRF_OK;
}

int gcrf_table_pack_accessions(gcrf_table *t, int32_t feature_type, int32_t digits, const int32_t **contig_ptr,
                               const int32_t **row_ptr, const int32_t **accession, int64_t *rows, int64_t *nnz) {
    if (!t) return tfail(GCRF_EINVAL, "table is NULL");
    if (digits < 0 || digits > 9) return tfail(GCRF_EINVAL, "digits must be 0..9");
    if (feature_type != 0 && feature_type != 1) return tfail(GCRF_EINVAL, "invalid feature type: %d", feature_type);
    try {
        const size_t G = t->genes.size(), C = t->contigs(), D = t->domains.size();
        if (D > 0x7fffff00u || G + D > 0x7fffff00u) return tfail(GCRF_EINVAL, "table too large for int32 row pointers; shard it");
        // "PF" + digits -> the number; anything else is not in a Pfam-only vocabulary
        t->attr_idx.resize(D);
        auto numbers = [&](size_t d0, size_t d1) {
            for (size_t d = d0; d < d1; ++d) {
                const sv name = t->domains[d].name;
                int32_t v = -1;
                if (name.size() > 2 && name.size() <= 11 && name[0] == 'P' && name[1] == 'F' &&
                    (digits == 0 || name.size() == (size_t)digits + 2)) {
                    int64_t x = 0;
                    bool digits = true;
                    for (size_t i = 2; i < name.size(); ++i) {
                        if (name[i] < '0' || name[i] > '9') {
                            digits = false;
                            break;
                        }
                        x = x * 10 + (name[i] - '0');
                    }
                    if (digits && x <= 0x7fffffff) v = (int32_t)x;
                }
                t->attr_idx[d] = v;
            }
        };
        const int nt = (int)std::min<size_t>((size_t)default_threads(), std::max<size_t>(1, D / 100000));
        if (nt <= 1) {
            numbers(0, D);
        } else {
            std::vector<std::thread> pool;
            for (int k = 0; k < nt; ++k) pool.emplace_back(numbers, D * k / nt, D * (k + 1) / nt);
            for (auto &th : pool) th.join();
        }
        t->row_contig_ptr.resize(C + 1);
        if (feature_type == 0) {  // one row per gene: the domain pointers are the row pointers
            t->row_ptr.resize(G + 1);
            for (size_t g = 0; g <= G; ++g) t->row_ptr[g] = (int32_t)t->dom_ptr[g];
            t->row_gene.resize(G);
            std::iota(t->row_gene.begin(), t->row_gene.end(), 0);
            for (size_t c = 0; c <= C; ++c) t->row_contig_ptr[c] = t->contig_ptr[c];
        } else {  // one row per domain, one empty row per domain-less gene (features.py:44-46)
            t->row_ptr.assign(1, 0);
            t->row_gene.clear();
            t->row_ptr.reserve(G + D + 1);
            t->row_gene.reserve(G + D);
            size_t c = 0;
            for (size_t g = 0; g < G; ++g) {
                while (c < C && (int64_t)g == t->contig_ptr[c]) t->row_contig_ptr[c++] = (int32_t)(t->row_ptr.size() - 1);
                const int64_t b = t->dom_ptr[g], e = t->dom_ptr[g + 1];
                for (int64_t k = b; k < e; ++k) {
                    t->row_ptr.push_back((int32_t)(k + 1));
                    t->row_gene.push_back((int32_t)g);
                }
                if (b == e) {
                    t->row_ptr.push_back((int32_t)b);
                    t->row_gene.push_back((int32_t)g);
                }
            }
            t->row_contig_ptr[C] = (int32_t)(t->row_ptr.size() - 1);
        }
        if (G == 0) t->row_contig_ptr.assign(1, 0);
        t->packed_mode = feature_type;
    } catch (const std::bad_alloc &) {
        return tfail(GCRF_ENOMEM, "out of host memory");
    }
    if (contig_ptr) *contig_ptr = t->row_contig_ptr.data();
    if (row_ptr) *row_ptr = t->row_ptr.data();
    if (accession) *accession = t->attr_idx.data();
    if (rows) *rows = (int64_t)t->row_ptr.size() - 1;
    if (nnz) *nnz = (int64_t)t->attr_idx.size();
    return GCRF_OK;
}

const int32_t *gcrf_table_row_gene(const gcrf_table *t) { return t && !t->row_gene.empty() ? t->row_gene.data() : nullptr; }

int gcrf_table_gene_probabilities(const gcrf_table *t, const double *row_prob, double *average_p, double *max_p) {
    if (!t) return tfail(GCRF_EINVAL, "table is NULL");
    if (t->packed_mode < 0) return tfail(GCRF_EINVAL, "gcrf_table_pack has not been called");
    int64_t cursor = 0;
    for (size_t g = 0; g < t->genes.size(); ++g) {
        double a, m;
        gene_probabilities(t, row_prob, g, &a, &m, &cursor);
        if (average_p) average_p[g] = a;
        if (max_p) max_p[g] = m;
    }
    return GCRF_OK;
}

int gcrf_table_write_genes(const gcrf_table *t, const double *row_prob, const char *path) {
    if (!t || !path) return tfail(GCRF_EINVAL, "bad arguments");
    if (row_prob && t->packed_mode < 0) return tfail(GCRF_EINVAL, "gcrf_table_pack has not been called");
    const size_t G = t->genes.size();
    std::vector<double> avg(G), mx(G);
    bool any_avg = false, any_max = false;
    int64_t cursor = 0;
    for (size_t g = 0; g < G; ++g) {
        gene_probabilities(t, row_prob, g, &avg[g], &mx[g], &cursor);
        any_avg |= !std::isnan(avg[g]);
        any_max |= !std::isnan(mx[g]);
    }
    std::string header = "sequence_id\tprotein_id\tstart\tend\tstrand";
    if (any_avg) header += "\taverage_p";  // a column of nothing but its default is left out (gecco/_base.py:136-144)
    if (any_max) header += "\tmax_p";
    header += '\n';
    return write_by_genes(path, header, G, 96, [&](std::string &out, size_t g) {
        const GeneRow &r = t->genes[g];
        out.append(r.seq);
        out += '\t';
        out.append(r.prot);
        out += '\t';
        append_int(out, r.start);
        out += '\t';
        append_int(out, r.end);
        out += '\t';
        out.append(r.strand);
        if (any_avg) {
            out += '\t';
            if (!std::isnan(avg[g])) append_repr(out, avg[g]);
        }
        if (any_max) {
            out += '\t';
            if (!std::isnan(mx[g])) append_repr(out, mx[g]);
        }
        out += '\n';
    });
}

int gcrf_table_write_clusters(const gcrf_table *t, const double *row_prob, const int32_t *seg_contig, const int32_t *seg_begin,
                              const int32_t *seg_end, const int32_t *seg_ordinal, int64_t n_segments, const char *path) {
    if (!t || !path || n_segments < 0) return tfail(GCRF_EINVAL, "bad arguments");
    if (n_segments > 0 && (!seg_contig || !seg_begin || !seg_end || !seg_ordinal)) return tfail(GCRF_EINVAL, "NULL segment array");
    if (row_prob && t->packed_mode < 0) return tfail(GCRF_EINVAL, "gcrf_table_pack has not been called");
    const size_t G = t->genes.size();
    std::vector<double> avg(G), mx(G);
    int64_t cursor = 0;
    for (size_t g = 0; g < G; ++g) gene_probabilities(t, row_prob, g, &avg[g], &mx[g], &cursor);
    // ClusterTable.from_clusters on clusters without a predicted type (gecco/model.py:735-762): the data columns in
    // insertion order, then the schema columns that are missing with their default (`type` = "Unknown").
    std::string out = "sequence_id\tcluster_id\tstart\tend\taverage_p\tmax_p\tproteins\tdomains\ttype\n";
    std::vector<sv> names;
    for (int64_t k = 0; k < n_segments; ++k) {
        const int64_t c = seg_contig[k], b = seg_begin[k], e = seg_end[k];
        if (c < 0 || c >= (int64_t)t->contigs() || b < 0 || e > (int64_t)G || b >= e)
            return tfail(GCRF_EINVAL, "segment %lld is out of range", (long long)k);
        const sv seq = t->contig_name((size_t)c);
        int64_t start = t->genes[(size_t)b].start, end = t->genes[(size_t)b].end;
        long double sum = 0;
        double best = std::numeric_limits<double>::quiet_NaN();
        int64_t n = 0;
        for (int64_t g = b; g < e; ++g) {
            start = std::min(start, t->genes[(size_t)g].start);
            end = std::max(end, t->genes[(size_t)g].end);
            if (!std::isnan(avg[(size_t)g])) {  // Cluster.average_probability / maximum_probability (gecco/model.py:442-455)
                sum += avg[(size_t)g];
                ++n;
            }
            if (!std::isnan(mx[(size_t)g])) best = std::isnan(best) ? mx[(size_t)g] : std::max(best, mx[(size_t)g]);
        }
        out += seq;
        out += '\t';
        out += seq;
        out += "_cluster_";
        append_int(out, seg_ordinal[k]);
        out += '\t';
        append_int(out, start);
        out += '\t';
        append_int(out, end);
        out += '\t';
        if (n) append_repr(out, (double)(sum / n));
        out += '\t';
        if (!std::isnan(best)) append_repr(out, best);
        out += '\t';
        names.clear();
        for (int64_t g = b; g < e; ++g) names.push_back(t->genes[(size_t)g].prot);
        std::sort(names.begin(), names.end());  // ";".join(sorted(gene.protein.id ...))
        for (size_t i = 0; i < names.size(); ++i) {
            if (i) out += ';';
            out.append(names[i]);
        }
        out += '\t';
        names.clear();
        for (int64_t g = b; g < e; ++g)
            for (int64_t d = t->dom_ptr[(size_t)g]; d < t->dom_ptr[(size_t)g + 1]; ++d) names.push_back(t->domains[(size_t)d].name);
        std::sort(names.begin(), names.end());
        for (size_t i = 0; i < names.size(); ++i) {
            if (i) out += ';';
            out.append(names[i]);
        }
        out += "\tUnknown\n";
    }
    return write_all(path, out);
}

int gcrf_table_write_features(const gcrf_table *t, const double *row_prob, const char *path) {
    if (!t || !path) return tfail(GCRF_EINVAL, "bad arguments");
    if (row_prob && t->packed_mode < 0) return tfail(GCRF_EINVAL, "gcrf_table_pack has not been called");
    const size_t G = t->genes.size();
    PhaseTimer timer;
    // per-domain probability: the gene's value in protein mode (row g of row_prob: read in place), the row's own in
    // domain mode (a per-domain array, since domain-less genes own a row too)
    const bool protein_rows = row_prob && t->packed_mode == 0;
    std::vector<double> dp;
    bool any = false;
    if (protein_rows) {
        for (size_t g = 0; g < G && !any; ++g) any = t->dom_ptr[g] != t->dom_ptr[g + 1] && !std::isnan(row_prob[g]);
    } else if (row_prob) {
        dp.assign(t->domains.size(), std::numeric_limits<double>::quiet_NaN());
        int64_t cursor = 0;
        for (size_t g = 0; g < G; ++g) {
            const int64_t b = t->dom_ptr[g], e = t->dom_ptr[g + 1];
            if (t->packed_mode == 0 || b == e) {
                const double p = row_prob[cursor++];
                for (int64_t k = b; k < e; ++k) dp[(size_t)k] = p;
            } else {
                for (int64_t k = b; k < e; ++k) dp[(size_t)k] = row_prob[cursor++];
            }
        }
        for (double p : dp) any |= !std::isnan(p);
    }
    timer.mark("per-domain probabilities");
    std::string header = "sequence_id\tprotein_id\tstart\tend\tstrand\tdomain\thmm\ti_evalue\tpvalue\tdomain_start\tdomain_end";
    if (any) header += "\tcluster_probability";
    header += '\n';
    const size_t per_gene = G ? 160 * (t->domains.size() / G + 1) : 160;
    const bool per_gene_prob = protein_rows;  // protein mode: every domain row of a gene carries the gene's value
    return write_by_genes(path, header, G, per_gene, [&](std::string &out, size_t g) {
        const GeneRow &r = t->genes[g];
        const int64_t b = t->dom_ptr[g], e = t->dom_ptr[g + 1];
        if (b == e) return;
        // the five gene columns are the same text on every row of the gene: laid out once
        const size_t p0 = out.size();
        out.append(r.seq);
        out += '\t';
        out.append(r.prot);
        out += '\t';
        append_int(out, r.start);
        out += '\t';
        append_int(out, r.end);
        out += '\t';
        out.append(r.strand);
        out += '\t';
        const size_t plen = out.size() - p0;
        char prob[40];
        size_t prob_len = 0;
        std::string tmp;
        if (any && per_gene_prob && !std::isnan(row_prob[g])) {
            append_repr(tmp, row_prob[g]);
            prob_len = std::min(tmp.size(), sizeof(prob));
            memcpy(prob, tmp.data(), prob_len);
        }
        for (int64_t k = b; k < e; ++k) {
            const DomainRow &d = t->domains[(size_t)k];
            if (k > b) {
                out.reserve(out.size() + plen);  // no reallocation while the string copies from itself
                out.append(out.data() + p0, plen);
            }
            out.append(d.name);
            out += '\t';
            out.append(d.hmm);
            out += '\t';
            if (!std::isnan(d.i_evalue)) append_repr(out, d.i_evalue);
            out += '\t';
            if (!std::isnan(d.pvalue)) append_repr(out, d.pvalue);
            out += '\t';
            append_int(out, d.dstart);
            out += '\t';
            append_int(out, d.dend);
            if (any) {
                out += '\t';
                if (per_gene_prob)
                    out.append(prob, prob_len);
                else if (!std::isnan(dp[(size_t)k]))
                    append_repr(out, dp[(size_t)k]);
            }
            out += '\n';
        }
    });
}

}  // extern "C"
