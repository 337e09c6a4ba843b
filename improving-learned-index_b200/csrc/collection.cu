// collection.cu — host-side parser / writer of the doc-major text collection
// ("term: score, term: score", one document per line, docid = line number).
//
// Pure host code (no CUDA): it replaces the Python string loops that otherwise dominate the wall
// clock of quantize_file (quantize.py:17-24, 39-47) and of InvertedIndexCreator
// (deep_impact_collection.py:11-25, create.py:19-35) once the arithmetic runs on the GPU. The exact
// tokenisation of the reference is kept:
//   line.strip()                       the same whitespace set as str.strip(), matched on the UTF-8 bytes
//   .split(', ') then .split(': ')     exact two-byte separators; a pair must split into exactly 2 parts
//   float(score)                       glibc strtod (correctly rounded, like CPython's float())
//   universal newlines                 "\n", "\r\n" and a lone "\r" all end a line
// mode DI_PARSE_DICT     (InvertedIndexCreator): blank line = empty document; a term repeated in a
//                        line keeps its LAST score at its FIRST position (dict semantics).
// mode DI_PARSE_SEQUENCE (quantize_file): every pair is kept in order (pairs are .strip()ped first,
//                        quantize.py:41-42); a blank line is an error ("not enough values to unpack").
// Numeric literals CPython accepts but this parser does not model (digit group separators '_', non-ASCII
// digits) return DI_ERR_UNSUPPORTED and the Python caller falls back to the reference-shaped pure-Python
// parser; so does any input that is not valid UTF-8 (checked by the caller, which decodes nothing otherwise).
#include <algorithm>
#include <cerrno>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iterator>
#include <new>
#include <string>
#include <string_view>
#include <thread>
#include <unordered_map>
#include <vector>

#include "common.cuh"  // error buffer shared with the rest of the library (di_last_error)

namespace {

int fail(int code, const char *fmt, unsigned long long line, const char *what)
{
    return di::set_error(code, fmt, line, what);
}

// str.strip() / float() whitespace: ASCII \t-\r, \x1c-\x1f, space, and the Unicode White_Space code
// points, matched here in their UTF-8 encodings so that stripping is byte-exact with CPython's.
size_t space_prefix(std::string_view s)
{
    if (s.empty()) return 0;
    const unsigned char c = (unsigned char)s[0];
    if (c == ' ' || (c >= '\t' && c <= '\r') || (c >= 0x1c && c <= 0x1f)) return 1;
    if (c == 0xC2 && s.size() >= 2 && ((unsigned char)s[1] == 0x85 || (unsigned char)s[1] == 0xA0)) return 2;
    if (s.size() >= 3) {
        const unsigned char d = (unsigned char)s[1], e = (unsigned char)s[2];
        if (c == 0xE1 && d == 0x9A && e == 0x80) return 3;                                   // U+1680
        if (c == 0xE2 && d == 0x80 && ((e >= 0x80 && e <= 0x8A) || e == 0xA8 || e == 0xA9 || e == 0xAF)) return 3;
        if (c == 0xE2 && d == 0x81 && e == 0x9F) return 3;                                   // U+205F
        if (c == 0xE3 && d == 0x80 && e == 0x80) return 3;                                   // U+3000
    }
    return 0;
}

size_t space_suffix(std::string_view s)
{
    for (size_t len = 1; len <= 3 && len <= s.size(); ++len) {
        const std::string_view tail = s.substr(s.size() - len);
        if (((unsigned char)tail[0] & 0xC0) == 0x80) continue;  // continuation byte: look one further back
        return space_prefix(tail) == len ? len : 0;
    }
    return 0;
}

std::string_view strip(std::string_view s)
{
    for (size_t n; (n = space_prefix(s)) != 0;) s.remove_prefix(n);
    for (size_t n; (n = space_suffix(s)) != 0;) s.remove_suffix(n);
    return s;
}

// float(text) for the literals this parser models; 0 = ok, 1 = ValueError, 2 = unsupported
int parse_float(std::string_view text, double *out)
{
    std::string_view t = strip(text);
    if (t.empty() || t.size() > 400) return t.empty() ? 1 : 2;
    {
        // Exact fast path (Clinger): [-+]digits[.digits] with <= 15 significant digits and <= 22
        // fraction digits is mantissa / 10^frac with both operands exact in binary64, so ONE correctly
        // rounded division gives the correctly rounded result — the value strtod / float() return.
        static const double kPow10[23] = {1e0,  1e1,  1e2,  1e3,  1e4,  1e5,  1e6,  1e7,  1e8,  1e9,  1e10, 1e11,
                                          1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};
        size_t i = 0;
        const bool neg = t[0] == '-';
        if (t[0] == '-' || t[0] == '+') ++i;
        uint64_t mant = 0;
        int digits = 0, frac = 0, n_int = 0;
        for (; i < t.size() && t[i] >= '0' && t[i] <= '9'; ++i, ++n_int) {
            mant = mant * 10 + (uint64_t)(t[i] - '0');
            digits += (mant != 0);
        }
        if (i < t.size() && t[i] == '.') {
            for (++i; i < t.size() && t[i] >= '0' && t[i] <= '9'; ++i, ++frac) {
                mant = mant * 10 + (uint64_t)(t[i] - '0');
                digits += (mant != 0);
            }
        }
        if (i == t.size() && n_int + frac > 0 && digits <= 15 && frac <= 22 && n_int + frac <= 18) {
            const double v = (double)mant / kPow10[frac];
            *out = neg ? -v : v;
            return 0;
        }
    }
    bool plain = true;
    for (char c : t) {
        if (!((c >= '0' && c <= '9') || c == '+' || c == '-' || c == '.' || c == 'e' || c == 'E')) plain = false;
        if ((unsigned char)c >= 0x80) return 2;  // float() also accepts Unicode digits / spaces
    }
    if (!plain) {
        std::string low(t);
        for (char &c : low) c = (char)tolower((unsigned char)c);
        std::string_view body(low);
        if (!body.empty() && (body[0] == '+' || body[0] == '-')) body.remove_prefix(1);
        if (body != "inf" && body != "infinity" && body != "nan") return low.find('_') != std::string::npos ? 2 : 1;
    }
    char buf[408];
    memcpy(buf, t.data(), t.size());
    buf[t.size()] = 0;
    char *end = nullptr;
    errno = 0;
    const double v = strtod(buf, &end);
    if (end != buf + t.size()) return 1;
    *out = v;  // overflow -> inf, underflow -> 0/denormal: same as CPython
    return 0;
}

}  // namespace

struct di_collection {
    int mode = 0;
    std::vector<uint64_t> doc_offsets{0};
    std::vector<uint32_t> term_ids;  // final ids = rank of the term in bytewise (= code point) order
    std::vector<double> scores;
    std::string vocab_blob;          // sorted terms, concatenated
    std::vector<uint64_t> vocab_offsets{0};
};

namespace {

// Strict UTF-8, as CPython's decoder accepts it (no overlong forms, no surrogates, nothing above U+10FFFF)
bool valid_utf8(std::string_view s)
{
    const unsigned char *p = reinterpret_cast<const unsigned char *>(s.data()), *end = p + s.size();
    while (p < end) {
        if (end - p >= 8) {  // ASCII run, 8 bytes at a time
            uint64_t w;
            memcpy(&w, p, 8);
            if (!(w & 0x8080808080808080ull)) { p += 8; continue; }
        }
        const unsigned char c = *p;
        if (c < 0x80) { ++p; continue; }
        auto cont = [&](int i, unsigned char lo, unsigned char hi) { return p + i < end && p[i] >= lo && p[i] <= hi; };
        if (c >= 0xC2 && c <= 0xDF) { if (!cont(1, 0x80, 0xBF)) return false; p += 2; }
        else if (c == 0xE0) { if (!cont(1, 0xA0, 0xBF) || !cont(2, 0x80, 0xBF)) return false; p += 3; }
        else if ((c >= 0xE1 && c <= 0xEC) || c == 0xEE || c == 0xEF) { if (!cont(1, 0x80, 0xBF) || !cont(2, 0x80, 0xBF)) return false; p += 3; }
        else if (c == 0xED) { if (!cont(1, 0x80, 0x9F) || !cont(2, 0x80, 0xBF)) return false; p += 3; }
        else if (c == 0xF0) { if (!cont(1, 0x90, 0xBF) || !cont(2, 0x80, 0xBF) || !cont(3, 0x80, 0xBF)) return false; p += 4; }
        else if (c >= 0xF1 && c <= 0xF3) { if (!cont(1, 0x80, 0xBF) || !cont(2, 0x80, 0xBF) || !cont(3, 0x80, 0xBF)) return false; p += 4; }
        else if (c == 0xF4) { if (!cont(1, 0x80, 0x8F) || !cont(2, 0x80, 0xBF) || !cont(3, 0x80, 0xBF)) return false; p += 4; }
        else return false;
    }
    return true;
}

// One contiguous piece of the file (cut after a '\n'), parsed by one thread: documents, postings with
// piece-local term ids, and the piece's own term table. Pieces share nothing until the merge.
struct Piece {
    std::string_view text;
    std::vector<uint64_t> doc_ends;          // postings of the piece before the end of each document
    std::vector<uint32_t> term_ids;          // piece-local ids, remapped to final ranks by the merge
    std::vector<double> scores;
    std::vector<std::string_view> terms;     // piece-local id -> term (views into the file text)
    std::vector<std::string_view> sorted;    // the same terms in bytewise order (sorted by the piece's own thread)
    uint64_t lines = 0;                      // lines consumed (the failing line included)
    int rc = DI_OK;
    const char *what = "";
    bool bad_utf8 = false;
};

void parse_piece(Piece &pc, int mode)
{
    pc.bad_utf8 = !valid_utf8(pc.text);
    if (pc.bad_utf8) return;
    // term interning: open-addressing table of local ids, keys are views into the text
    std::vector<std::string_view> &terms = pc.terms;
    std::vector<uint32_t> table(1u << 16, 0xFFFFFFFFu);
    std::vector<uint64_t> hashes;
    auto hash_of = [](std::string_view t) {
        uint64_t h = 0xcbf29ce484222325ull;  // FNV-1a, then a finaliser for the low bits
        for (unsigned char ch : t) h = (h ^ ch) * 0x100000001b3ull;
        h ^= h >> 32;
        return h * 0x9E3779B97F4A7C15ull;
    };
    auto intern_term = [&](std::string_view t) {
        const uint64_t h = hash_of(t);
        size_t mask = table.size() - 1;
        for (size_t i = (h >> 20) & mask;; i = (i + 1) & mask) {
            const uint32_t id = table[i];
            if (id == 0xFFFFFFFFu) {
                const uint32_t nid = (uint32_t)terms.size();
                terms.push_back(t);
                hashes.push_back(h);
                table[i] = nid;
                if (terms.size() * 2 > table.size()) {  // grow and re-insert
                    std::vector<uint32_t> bigger(table.size() * 4, 0xFFFFFFFFu);
                    mask = bigger.size() - 1;
                    for (uint32_t k = 0; k < terms.size(); ++k) {
                        size_t j = (hashes[k] >> 20) & mask;
                        while (bigger[j] != 0xFFFFFFFFu) j = (j + 1) & mask;
                        bigger[j] = k;
                    }
                    table.swap(bigger);
                }
                return nid;
            }
            if (hashes[id] == h && terms[id] == t) return id;
        }
    };
    auto stop = [&](int code, const char *what) { pc.rc = code; pc.what = what; };
    // "term: 1.234, " is rarely shorter than 12 bytes: one allocation instead of ~30 doublings
    pc.term_ids.reserve(pc.text.size() / 12 + 16);
    pc.scores.reserve(pc.text.size() / 12 + 16);
    std::vector<uint64_t> seen_in_doc;     // DICT mode: 1 + index of the term's posting if seen in the current doc
    size_t pos = 0;
    const std::string_view all = pc.text;
    while (pos < all.size() && pc.rc == DI_OK) {
        size_t eol = pos;
        while (eol < all.size() && all[eol] != '\n' && all[eol] != '\r') ++eol;
        const std::string_view raw = all.substr(pos, eol - pos);
        pos = eol < all.size() ? eol + ((all[eol] == '\r' && eol + 1 < all.size() && all[eol + 1] == '\n') ? 2 : 1) : eol;
        ++pc.lines;
        const std::string_view line = strip(raw);
        const size_t doc_begin = pc.term_ids.size();
        if (line.empty()) {
            if (mode == DI_PARSE_SEQUENCE) { stop(DI_ERR_FORMAT, "not enough values to unpack (expected 2, got 1)"); break; }
            pc.doc_ends.push_back(doc_begin);
            continue;
        }
        size_t p = 0;
        while (pc.rc == DI_OK) {
            size_t sep = line.find(", ", p);
            std::string_view pair = line.substr(p, sep == std::string_view::npos ? std::string_view::npos : sep - p);
            if (mode == DI_PARSE_SEQUENCE) pair = strip(pair);
            const size_t colon = pair.find(": ");
            if (colon == std::string_view::npos) { stop(DI_ERR_FORMAT, "not enough values to unpack (expected 2, got 1)"); break; }
            if (pair.find(": ", colon + 2) != std::string_view::npos) { stop(DI_ERR_FORMAT, "too many values to unpack (expected 2)"); break; }
            const std::string_view term = pair.substr(0, colon);
            double v = 0;
            const int fr = parse_float(pair.substr(colon + 2), &v);
            if (fr == 1) { stop(DI_ERR_FORMAT, "could not convert string to float"); break; }
            if (fr == 2) { stop(DI_ERR_UNSUPPORTED, "numeric literal outside the fast parser's grammar"); break; }
            const uint32_t id = intern_term(term);
            bool replaced = false;
            if (mode == DI_PARSE_DICT) {  // dict: last value wins, first position kept
                if (seen_in_doc.size() <= id) seen_in_doc.resize((size_t)id * 2 + 64, 0);
                const uint64_t at = seen_in_doc[id];      // postings before doc_begin belong to earlier docs
                if (at > doc_begin) { pc.scores[at - 1] = v; replaced = true; }
                else seen_in_doc[id] = pc.term_ids.size() + 1;
            }
            if (!replaced) { pc.term_ids.push_back(id); pc.scores.push_back(v); }
            if (sep == std::string_view::npos) break;
            p = sep + 2;
        }
        if (pc.rc == DI_OK) pc.doc_ends.push_back(pc.term_ids.size());
    }
    if (pc.rc == DI_OK) {
        pc.sorted = pc.terms;
        std::sort(pc.sorted.begin(), pc.sorted.end());
    }
}

unsigned parse_threads(uint64_t n_bytes)
{
    // DI_B200_PARSE_THREADS pins the number of pieces (tests cut tiny inputs into many pieces with it)
    if (const char *e = getenv("DI_B200_PARSE_THREADS")) return (unsigned)std::min(256, std::max(1, atoi(e)));
    const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
    const uint64_t by_size = std::max<uint64_t>(1, n_bytes >> 20);  // at least 1 MB of text per thread
    return (unsigned)std::min<uint64_t>(std::min<uint64_t>(hw, 256), by_size);
}

template <typename F> void run_parallel(unsigned n, F &&body)
{
    std::vector<std::thread> pool;
    for (unsigned t = 1; t < n; ++t) pool.emplace_back([&body, t] { body(t); });
    body(0);
    for (std::thread &t : pool) t.join();
}

}  // namespace

// The file is cut into one piece per host thread (after a '\n'; a "\r\n" pair is never split, and a lone '\r'
// still ends a line inside its piece), the pieces are parsed independently, and the merge gives every term
// its rank in the sorted global vocabulary. The outcome — arrays, vocabulary, or the FIRST error of the file
// with its line number — does not depend on the number of threads.
extern "C" int di_collection_parse(const char *text, uint64_t n_bytes, int mode, di_collection_t **out)
{
    if (!out || (!text && n_bytes)) return fail(DI_ERR_ARG, "line %llu: %s", 0, "NULL argument");
    *out = nullptr;
    if (mode != DI_PARSE_DICT && mode != DI_PARSE_SEQUENCE) return fail(DI_ERR_ARG, "line %llu: %s", 0, "bad mode");
    di_collection *c = new (std::nothrow) di_collection();
    if (!c) return fail(DI_ERR_NOMEM, "line %llu: %s", 0, "out of memory");
    c->mode = mode;
    const std::string_view all(text ? text : "", (size_t)n_bytes);
    const unsigned n_pieces = parse_threads(n_bytes);
    std::vector<Piece> pieces(n_pieces);
    size_t begin = 0;
    for (unsigned i = 0; i < n_pieces; ++i) {
        size_t end = all.size();
        if (i + 1 < n_pieces) {
            end = std::max(begin, (size_t)((uint64_t)all.size() * (i + 1) / n_pieces));
            const size_t nl = all.find('\n', end);
            end = nl == std::string_view::npos ? all.size() : nl + 1;
        }
        pieces[i].text = all.substr(begin, end - begin);
        begin = end;
    }
    try {
        run_parallel(n_pieces, [&](unsigned i) { parse_piece(pieces[i], mode); });
        // invalid UTF-8 anywhere: the caller's decode raises, whatever else the file holds
        for (const Piece &pc : pieces)
            if (pc.bad_utf8) { delete c; return fail(DI_ERR_UNSUPPORTED, "line %llu: %s", 0, "input is not valid UTF-8"); }
        uint64_t lines_before = 0;
        for (const Piece &pc : pieces) {  // the first error in file order, with its global line number
            if (pc.rc != DI_OK) { delete c; return fail(pc.rc, "line %llu: %s", lines_before + pc.lines, pc.what); }
            lines_before += pc.lines;
        }
        // global vocabulary: sorted (bytewise == code point order for UTF-8) union of the pieces' terms,
        // as sorted(set(terms)) gives; final id = rank
        std::vector<std::string_view> vocab, merged;
        for (const Piece &pc : pieces) {  // union of the pieces' sorted term lists: linear per piece
            merged.clear();
            merged.reserve(vocab.size() + pc.sorted.size());
            std::set_union(vocab.begin(), vocab.end(), pc.sorted.begin(), pc.sorted.end(), std::back_inserter(merged));
            vocab.swap(merged);
        }
        for (const std::string_view t : vocab) {
            c->vocab_blob.append(t);
            c->vocab_offsets.push_back(c->vocab_blob.size());
        }
        std::vector<uint64_t> post_base(n_pieces + 1, 0), doc_base(n_pieces + 1, 0);
        for (unsigned i = 0; i < n_pieces; ++i) {
            post_base[i + 1] = post_base[i] + pieces[i].term_ids.size();
            doc_base[i + 1] = doc_base[i] + pieces[i].doc_ends.size();
        }
        c->term_ids.resize(post_base[n_pieces]);
        c->scores.resize(post_base[n_pieces]);
        c->doc_offsets.resize(doc_base[n_pieces] + 1);
        c->doc_offsets[0] = 0;
        run_parallel(n_pieces, [&](unsigned i) {
            Piece &pc = pieces[i];
            std::vector<uint32_t> rank(pc.terms.size());
            for (size_t k = 0; k < pc.terms.size(); ++k)
                rank[k] = (uint32_t)(std::lower_bound(vocab.begin(), vocab.end(), pc.terms[k]) - vocab.begin());
            uint32_t *ids = c->term_ids.data() + post_base[i];
            for (size_t k = 0; k < pc.term_ids.size(); ++k) ids[k] = rank[pc.term_ids[k]];
            if (!pc.scores.empty()) memcpy(c->scores.data() + post_base[i], pc.scores.data(), pc.scores.size() * sizeof(double));
            uint64_t *offs = c->doc_offsets.data() + doc_base[i] + 1;
            for (size_t d = 0; d < pc.doc_ends.size(); ++d) offs[d] = post_base[i] + pc.doc_ends[d];
            std::vector<uint32_t>().swap(pc.term_ids);
            std::vector<double>().swap(pc.scores);
        });
    } catch (const std::bad_alloc &) {
        delete c;
        return fail(DI_ERR_NOMEM, "line %llu: %s", 0, "out of memory");
    }
    *out = c;
    return DI_OK;
}

extern "C" void di_collection_free(di_collection_t *c) { delete c; }

extern "C" int di_collection_info(const di_collection_t *c, uint64_t *n_docs, uint64_t *n_postings, uint32_t *n_terms)
{
    if (!c) return fail(DI_ERR_ARG, "line %llu: %s", 0, "NULL collection");
    if (n_docs) *n_docs = c->doc_offsets.size() - 1;
    if (n_postings) *n_postings = c->term_ids.size();
    if (n_terms) *n_terms = (uint32_t)(c->vocab_offsets.size() - 1);
    return DI_OK;
}

extern "C" int di_collection_arrays(const di_collection_t *c, const uint64_t **doc_offsets, const uint32_t **term_ids,
                                    const double **scores, const char **vocab_blob, const uint64_t **vocab_offsets)
{
    if (!c) return fail(DI_ERR_ARG, "line %llu: %s", 0, "NULL collection");
    if (doc_offsets) *doc_offsets = c->doc_offsets.data();
    if (term_ids) *term_ids = c->term_ids.data();
    if (scores) *scores = c->scores.data();
    if (vocab_blob) *vocab_blob = c->vocab_blob.data();
    if (vocab_offsets) *vocab_offsets = c->vocab_offsets.data();
    return DI_OK;
}

// quantize.py:40-47 — one output line per document: "term: value" for every value > 0, joined by ", ".
// Formatting is done by all host threads on blocks of ~2 M postings; the blocks are written in order.
extern "C" int di_collection_write_quantized(const di_collection_t *c, const int32_t *values, const char *path)
{
    if (!c || (!values && !c->term_ids.empty()) || !path) return fail(DI_ERR_ARG, "line %llu: %s", 0, "NULL argument");
    FILE *f = fopen(path, "wb");
    if (!f) return fail(DI_ERR_ARG, "line %llu: cannot open %s for writing", 0, path);
    const size_t n_docs = c->doc_offsets.size() - 1;
    constexpr uint64_t kBlockPostings = 2u << 20;
    std::vector<size_t> block_begin{0};  // first document of every block
    while (block_begin.back() < n_docs) {
        const uint64_t target = c->doc_offsets[block_begin.back()] + kBlockPostings;
        size_t next = (size_t)(std::upper_bound(c->doc_offsets.begin(), c->doc_offsets.end(), target) - c->doc_offsets.begin());
        next = std::min(n_docs, std::max(next, block_begin.back() + 1));
        // blocks of empty documents: at most 1 M lines each
        block_begin.push_back(std::min(next, block_begin.back() + (1u << 20)));
    }
    const size_t n_blocks = block_begin.size() - 1;
    auto format_block = [&](size_t b, std::string &buf) {
        buf.clear();
        char num[16];
        for (size_t d = block_begin[b]; d < block_begin[b + 1]; ++d) {
            bool first = true;
            for (uint64_t i = c->doc_offsets[d]; i < c->doc_offsets[d + 1]; ++i) {
                if (values[i] <= 0) continue;
                if (!first) buf += ", ";
                first = false;
                const uint32_t t = c->term_ids[i];
                buf.append(c->vocab_blob, c->vocab_offsets[t], c->vocab_offsets[t + 1] - c->vocab_offsets[t]);
                buf += ": ";
                int len = 0;  // values[i] > 0 here
                for (uint32_t x = (uint32_t)values[i]; x; x /= 10) num[sizeof num - 1 - len++] = (char)('0' + x % 10);
                buf.append(num + sizeof num - len, (size_t)len);
            }
            buf += '\n';
        }
    };
    const unsigned n_threads = (unsigned)std::max<size_t>(1, std::min<size_t>(parse_threads(c->term_ids.size() * 16ull), n_blocks));
    bool ok = true;
    try {
        std::vector<std::string> bufs(n_threads);
        for (size_t b0 = 0; b0 < n_blocks && ok; b0 += n_threads) {
            const unsigned n = (unsigned)std::min<size_t>(n_threads, n_blocks - b0);
            run_parallel(n, [&](unsigned t) { format_block(b0 + t, bufs[t]); });
            for (unsigned t = 0; t < n && ok; ++t) ok = fwrite(bufs[t].data(), 1, bufs[t].size(), f) == bufs[t].size();
        }
    } catch (const std::bad_alloc &) {
        fclose(f);
        return fail(DI_ERR_NOMEM, "line %llu: %s", 0, "out of memory");
    }
    if (fclose(f) != 0 || !ok) return fail(DI_ERR_ARG, "line %llu: %s", 0, "short write");
    return DI_OK;
}
