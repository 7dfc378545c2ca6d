// Native host-side text processing (pybind11, no torch dependency).
//
//  count_words   : word -> count over an iterable of token sequences.  The reference does this
//                  with a Spark shuffle (flatMap/map/reduceByKey, MLLIB:259-262); here it is an
//                  open-addressing hash map keyed by the UTF-8 bytes, sharded over threads once the
//                  Python strings have been flattened (the GIL-bound part is only the flattening).
//  encode_corpus : words -> vocabulary indices with OOV dropped and sentences chunked at
//                  maxSentenceLength (MLLIB:335-343).
//  count_words_file / encode_file : the same two passes straight from a text file (one sentence per line),
//                  mmap'ed and split over threads at line boundaries - the data-loader path for corpora
//                  that do not fit a Python list of lists (the reference reads them as Spark RDD partitions).
//  vose_alias    : Vose alias construction for the unigram^0.75 noise distribution (replaces the
//                  Glint servers' 1e8-entry unigram table, ML:204-206).
#include <pybind11/pybind11.h>
#include <pybind11/numpy.h>
#include <pybind11/stl.h>

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <string_view>
#include <thread>
#include <unordered_map>
#include <vector>

namespace py = pybind11;

namespace {

struct StrHash {
    size_t operator()(std::string_view s) const noexcept {
        uint64_t h = 1469598103934665603ull;                 // FNV-1a
        for (unsigned char c : s) { h ^= c; h *= 1099511628211ull; }
        return (size_t)h;
    }
};

// Open-addressing string map (linear probing, power-of-two capacity, load <= 0.5).  Keys are views into storage
// that outlives the map (the mmap'ed file or an arena); std::unordered_map cost ~2x more per token here.
struct FlatStrMap {
    struct Slot { uint64_t hash; const char* p; uint32_t len; int64_t val; };
    std::vector<Slot> slots;
    size_t mask = 0, n = 0;
    static inline const char kEmptyKey = 0;                  // non-null address for the empty string key

    explicit FlatStrMap(size_t cap_pow2 = 1 << 16) : slots(cap_pow2, Slot{0, nullptr, 0, 0}), mask(cap_pow2 - 1) {}

    static uint64_t hash(std::string_view s) noexcept { return StrHash{}(s) * 0x9E3779B97F4A7C15ull; }

    void grow() {
        std::vector<Slot> old;
        old.swap(slots);
        slots.assign(old.size() * 2, Slot{0, nullptr, 0, 0});
        mask = slots.size() - 1;
        for (const Slot& o : old) {
            if (!o.p) continue;
            size_t i = (size_t)(o.hash >> 7) & mask;
            while (slots[i].p) i = (i + 1) & mask;
            slots[i] = o;
        }
    }
    Slot& find_or_insert(std::string_view s, uint64_t h) {
        if ((n + 1) * 2 > slots.size()) grow();
        size_t i = (size_t)(h >> 7) & mask;
        while (true) {
            Slot& sl = slots[i];
            if (!sl.p) {
                sl = Slot{h, s.empty() ? &kEmptyKey : s.data(), (uint32_t)s.size(), 0};
                ++n;
                return sl;
            }
            if (sl.hash == h && sl.len == s.size() && std::memcmp(sl.p, s.data(), s.size()) == 0) return sl;
            i = (i + 1) & mask;
        }
    }
    const Slot* find(std::string_view s, uint64_t h) const noexcept {
        size_t i = (size_t)(h >> 7) & mask;
        while (true) {
            const Slot& sl = slots[i];
            if (!sl.p) return nullptr;
            if (sl.hash == h && sl.len == s.size() && std::memcmp(sl.p, s.data(), s.size()) == 0) return &sl;
            i = (i + 1) & mask;
        }
    }
};

// flatten an iterable of token sequences into one byte arena + offsets (holds the GIL)
struct Flat {
    std::string arena;
    std::vector<uint64_t> tok_off;     // token i = arena[tok_off[i], tok_off[i+1])
    std::vector<uint64_t> sent_off;    // sentence s = tokens [sent_off[s], sent_off[s+1])
};

Flat flatten(const py::iterable& sentences) {
    Flat f;
    f.tok_off.push_back(0);
    f.sent_off.push_back(0);
    for (py::handle sent : sentences) {
        if (!sent.is_none()) {
            for (py::handle tok : py::reinterpret_borrow<py::iterable>(sent)) {
                Py_ssize_t len = 0;
                const char* data = PyUnicode_AsUTF8AndSize(tok.ptr(), &len);
                if (!data) throw py::error_already_set();
                f.arena.append(data, (size_t)len);
                f.tok_off.push_back(f.arena.size());
            }
        }
        f.sent_off.push_back(f.tok_off.size() - 1);
    }
    return f;
}

py::dict count_words(const py::iterable& sentences) {
    Flat f = flatten(sentences);
    const size_t ntok = f.tok_off.size() - 1;
    unsigned nthreads = std::max(1u, std::min(8u, std::thread::hardware_concurrency()));
    if (ntok < 200000) nthreads = 1;
    std::vector<std::unordered_map<std::string_view, int64_t, StrHash>> maps(nthreads);
    {
        py::gil_scoped_release rel;
        auto work = [&](unsigned t) {
            auto& m = maps[t];
            m.reserve(1 << 16);
            size_t lo = ntok * t / nthreads, hi = ntok * (t + 1) / nthreads;
            for (size_t i = lo; i < hi; ++i) {
                std::string_view w(f.arena.data() + f.tok_off[i], f.tok_off[i + 1] - f.tok_off[i]);
                ++m[w];
            }
        };
        std::vector<std::thread> th;
        for (unsigned t = 1; t < nthreads; ++t) th.emplace_back(work, t);
        work(0);
        for (auto& x : th) x.join();
        for (unsigned t = 1; t < nthreads; ++t)
            for (auto& kv : maps[t]) maps[0][kv.first] += kv.second;
    }
    py::dict out;
    for (auto& kv : maps[0])
        out[py::reinterpret_steal<py::str>(PyUnicode_DecodeUTF8(kv.first.data(), (Py_ssize_t)kv.first.size(), "strict"))] =
            kv.second;
    return out;
}

py::tuple encode_corpus(const py::iterable& sentences, const py::dict& index, int64_t max_len) {
    // build a native word -> id map once
    std::unordered_map<std::string, int32_t> idx;
    idx.reserve(index.size() * 2);
    for (auto kv : index) {
        Py_ssize_t len = 0;
        const char* data = PyUnicode_AsUTF8AndSize(kv.first.ptr(), &len);
        if (!data) throw py::error_already_set();
        idx.emplace(std::string(data, (size_t)len), kv.second.cast<int32_t>());
    }
    Flat f = flatten(sentences);
    std::vector<int32_t> toks;
    std::vector<int64_t> offs{0};
    {
        py::gil_scoped_release rel;
        toks.reserve(f.tok_off.size());
        std::string key;
        for (size_t s = 0; s + 1 < f.sent_off.size(); ++s) {
            int64_t in_chunk = 0;
            bool any = false;
            for (uint64_t i = f.sent_off[s]; i < f.sent_off[s + 1]; ++i) {
                key.assign(f.arena.data() + f.tok_off[i], f.tok_off[i + 1] - f.tok_off[i]);
                auto it = idx.find(key);
                if (it == idx.end()) continue;
                if (in_chunk == max_len) { offs.push_back((int64_t)toks.size()); in_chunk = 0; }
                toks.push_back(it->second);
                ++in_chunk;
                any = true;
            }
            if (any) offs.push_back((int64_t)toks.size());
        }
    }
    py::array_t<int32_t> a((py::ssize_t)toks.size());
    if (!toks.empty()) std::memcpy(a.mutable_data(), toks.data(), toks.size() * sizeof(int32_t));
    py::array_t<int64_t> o((py::ssize_t)offs.size());
    std::memcpy(o.mutable_data(), offs.data(), offs.size() * sizeof(int64_t));
    return py::make_tuple(a, o);
}

// ------------------------------------------------------------------------------------------ text files
// Read-only mmap of a whole file.
struct MappedFile {
    const char* data = nullptr;
    size_t size = 0;
    int fd = -1;
    explicit MappedFile(const std::string& path) {
        fd = ::open(path.c_str(), O_RDONLY);
        if (fd < 0) throw std::runtime_error("cannot open " + path);
        struct stat st;
        if (fstat(fd, &st) != 0) { ::close(fd); throw std::runtime_error("cannot stat " + path); }
        size = (size_t)st.st_size;
        if (size > 0) {
            void* p = mmap(nullptr, size, PROT_READ, MAP_PRIVATE, fd, 0);
            if (p == MAP_FAILED) { ::close(fd); throw std::runtime_error("cannot mmap " + path); }
            data = static_cast<const char*>(p);
            madvise(p, size, MADV_SEQUENTIAL);
        }
    }
    ~MappedFile() {
        if (data) munmap(const_cast<char*>(data), size);
        if (fd >= 0) ::close(fd);
    }
    MappedFile(const MappedFile&) = delete;
    MappedFile& operator=(const MappedFile&) = delete;
};

// [begin, end) byte ranges that start and end on line boundaries, one per thread
std::vector<std::pair<size_t, size_t>> line_chunks(const MappedFile& f, unsigned n) {
    std::vector<std::pair<size_t, size_t>> out;
    size_t begin = 0;
    for (unsigned t = 0; t < n && begin < f.size; ++t) {
        size_t end = (t + 1 == n) ? f.size : std::max(begin, f.size * (t + 1) / n);
        while (end < f.size && f.data[end - 1] != '\n') ++end;        // extend to the end of the line
        if (end > begin) out.emplace_back(begin, end);
        begin = end;
    }
    return out;
}

// Calls fn(token) for every token of the line [p, e) (no newline inside).
//   java_mode: String.split(" ") of the JVM - split on every single space, keep interior empty tokens,
//              drop trailing empty tokens; a line without any token left yields one empty token iff the
//              line itself is empty
//   otherwise: split on runs of spaces / tabs / CR, never yields empty tokens
template <typename F>
inline void for_tokens(const char* p, const char* e, bool java_mode, F&& fn) {
    if (e > p && e[-1] == '\r') --e;
    if (!java_mode) {
        while (p < e) {
            while (p < e && (*p == ' ' || *p == '\t' || *p == '\r')) ++p;
            const char* b = p;
            while (p < e && !(*p == ' ' || *p == '\t' || *p == '\r')) ++p;
            if (p > b) fn(std::string_view(b, (size_t)(p - b)));
        }
        return;
    }
    if (p == e) { fn(std::string_view()); return; }                    // "" -> [""]
    const char* last = e;
    while (last > p && last[-1] == ' ') --last;                        // trailing empties are dropped
    if (last == p) return;                                             // only spaces -> no tokens
    const char* b = p;
    for (const char* q = p; q <= last; ++q) {
        if (q == last || *q == ' ') { fn(std::string_view(b, (size_t)(q - b))); b = q + 1; }
    }
}

template <typename F>
inline void for_lines(const char* p, const char* e, F&& fn) {
    while (p < e) {
        const char* nl = static_cast<const char*>(memchr(p, '\n', (size_t)(e - p)));
        const char* le = nl ? nl : e;
        fn(p, le);
        p = nl ? nl + 1 : e;
    }
}

unsigned pick_threads(int64_t requested, size_t bytes) {
    unsigned n = requested > 0 ? (unsigned)requested : std::max(1u, std::min(32u, std::thread::hardware_concurrency()));
    if (bytes < (1u << 20)) n = 1;
    return n;
}

// word -> count over a text file; returns (list of words, int64 counts) in unspecified order
py::tuple count_words_file(const std::string& path, bool java_mode, int64_t num_threads) {
    std::vector<std::pair<std::string, int64_t>> merged;
    {
        py::gil_scoped_release rel;
        MappedFile f(path);
        auto chunks = line_chunks(f, pick_threads(num_threads, f.size));
        std::vector<FlatStrMap> maps(chunks.size());
        auto work = [&](size_t t) {
            auto& m = maps[t];
            for_lines(f.data + chunks[t].first, f.data + chunks[t].second, [&](const char* b, const char* e) {
                for_tokens(b, e, java_mode, [&](std::string_view w) { ++m.find_or_insert(w, FlatStrMap::hash(w)).val; });
            });
        };
        std::vector<std::thread> th;
        for (size_t t = 1; t < chunks.size(); ++t) th.emplace_back(work, t);
        if (!chunks.empty()) work(0);
        for (auto& x : th) x.join();
        for (size_t t = 1; t < maps.size(); ++t)
            for (const auto& sl : maps[t].slots)
                if (sl.p) maps[0].find_or_insert(std::string_view(sl.p, sl.len), sl.hash).val += sl.val;
        if (!maps.empty()) {
            merged.reserve(maps[0].n);
            for (const auto& sl : maps[0].slots)
                if (sl.p) merged.emplace_back(std::string(sl.p, sl.len), sl.val);          // copy before munmap
        }
    }
    py::list words;
    py::array_t<int64_t> counts((py::ssize_t)merged.size());
    int64_t* c = counts.mutable_data();
    for (size_t i = 0; i < merged.size(); ++i) {
        words.append(py::reinterpret_steal<py::str>(
            PyUnicode_DecodeUTF8(merged[i].first.data(), (Py_ssize_t)merged[i].first.size(), "replace")));
        c[i] = merged[i].second;
    }
    return py::make_tuple(words, counts);
}

// text file -> (tokens int32, sentence offsets int64): OOV words dropped, sentences chunked at max_len,
// sentences without any in-vocabulary word skipped (MLLIB:335-343); `words[i]` has index i
py::tuple encode_file(const std::string& path, const py::list& words, int64_t max_len, bool java_mode,
                      int64_t num_threads) {
    // vocabulary: one arena holding the UTF-8 bytes + a flat map word -> index (read-only while encoding)
    std::string arena;
    std::vector<std::pair<size_t, size_t>> spans;
    spans.reserve((size_t)py::len(words));
    for (py::handle w : words) {
        Py_ssize_t len = 0;
        const char* data = PyUnicode_AsUTF8AndSize(w.ptr(), &len);
        if (!data) throw py::error_already_set();
        spans.emplace_back(arena.size(), (size_t)len);
        arena.append(data, (size_t)len);
    }
    size_t cap = 1 << 10;
    while (cap < spans.size() * 2 + 2) cap <<= 1;
    FlatStrMap idx(cap);
    for (size_t i = 0; i < spans.size(); ++i) {
        std::string_view w(arena.data() + spans[i].first, spans[i].second);
        auto& sl = idx.find_or_insert(w, FlatStrMap::hash(w));
        sl.val = (int64_t)i;                              // duplicates keep the last index (callers pass unique words)
    }
    std::vector<std::vector<int32_t>> toks;
    std::vector<std::vector<int64_t>> lens;          // sentence (chunk) lengths per thread
    {
        py::gil_scoped_release rel;
        MappedFile f(path);
        auto chunks = line_chunks(f, pick_threads(num_threads, f.size));
        toks.resize(chunks.size());
        lens.resize(chunks.size());
        auto work = [&](size_t t) {
            auto& tk = toks[t];
            auto& ln = lens[t];
            for_lines(f.data + chunks[t].first, f.data + chunks[t].second, [&](const char* b, const char* e) {
                int64_t in_chunk = 0;
                for_tokens(b, e, java_mode, [&](std::string_view w) {
                    const FlatStrMap::Slot* sl = idx.find(w, FlatStrMap::hash(w));
                    if (!sl) return;
                    if (in_chunk == max_len) { ln.push_back(in_chunk); in_chunk = 0; }
                    tk.push_back((int32_t)sl->val);
                    ++in_chunk;
                });
                if (in_chunk > 0) ln.push_back(in_chunk);
            });
        };
        std::vector<std::thread> th;
        for (size_t t = 1; t < chunks.size(); ++t) th.emplace_back(work, t);
        if (!chunks.empty()) work(0);
        for (auto& x : th) x.join();
    }
    size_t ntok = 0, nsent = 0;
    for (auto& v : toks) ntok += v.size();
    for (auto& v : lens) nsent += v.size();
    py::array_t<int32_t> a((py::ssize_t)ntok);
    py::array_t<int64_t> o((py::ssize_t)nsent + 1);
    int32_t* ap = a.mutable_data();
    int64_t* op = o.mutable_data();
    size_t at = 0, os = 0;
    int64_t run = 0;
    op[os++] = 0;
    for (size_t t = 0; t < toks.size(); ++t) {
        if (!toks[t].empty()) std::memcpy(ap + at, toks[t].data(), toks[t].size() * sizeof(int32_t));
        at += toks[t].size();
        for (int64_t l : lens[t]) { run += l; op[os++] = run; }
    }
    return py::make_tuple(a, o);
}

// Streaming variant of encode_file for corpora that must not be materialised in memory (the reference never holds
// the corpus on one node either: RDD partitions, MLLIB:335-345).  The text file is processed in blocks of
// `block_bytes` (cut at line boundaries); every block is encoded on all cores and appended to
//   <out_prefix>.tokens.i32   (int32 token ids)        <out_prefix>.offsets.i64  (int64 sentence offsets, starts with 0)
// Peak memory is one block of tokens, whatever the corpus size.  Returns (tokens, sentences).
py::tuple encode_file_to(const std::string& path, const py::list& words, int64_t max_len, bool java_mode,
                         int64_t num_threads, const std::string& out_prefix, int64_t block_bytes) {
    std::string arena;
    std::vector<std::pair<size_t, size_t>> spans;
    spans.reserve((size_t)py::len(words));
    for (py::handle w : words) {
        Py_ssize_t len = 0;
        const char* data = PyUnicode_AsUTF8AndSize(w.ptr(), &len);
        if (!data) throw py::error_already_set();
        spans.emplace_back(arena.size(), (size_t)len);
        arena.append(data, (size_t)len);
    }
    size_t cap = 1 << 10;
    while (cap < spans.size() * 2 + 2) cap <<= 1;
    FlatStrMap idx(cap);
    for (size_t i = 0; i < spans.size(); ++i) {
        std::string_view w(arena.data() + spans[i].first, spans[i].second);
        idx.find_or_insert(w, FlatStrMap::hash(w)).val = (int64_t)i;
    }
    int64_t ntok = 0, nsent = 0;
    {
        py::gil_scoped_release rel;
        MappedFile f(path);
        FILE* ft = fopen((out_prefix + ".tokens.i32").c_str(), "wb");
        FILE* fo = fopen((out_prefix + ".offsets.i64").c_str(), "wb");
        if (!ft || !fo) { if (ft) fclose(ft); if (fo) fclose(fo); throw std::runtime_error("cannot create " + out_prefix + ".*"); }
        const int64_t zero = 0;
        fwrite(&zero, sizeof(int64_t), 1, fo);
        const size_t blk = (size_t)std::max<int64_t>(block_bytes, 1 << 20);
        const unsigned nth = pick_threads(num_threads, f.size);
        size_t begin = 0;
        std::vector<int64_t> offs;
        while (begin < f.size) {
            size_t end = std::min(f.size, begin + blk);
            while (end < f.size && f.data[end - 1] != '\n') ++end;
            // thread chunks of this block, cut at line boundaries
            std::vector<std::pair<size_t, size_t>> chunks;
            size_t b = begin;
            for (unsigned t = 0; t < nth && b < end; ++t) {
                size_t e = (t + 1 == nth) ? end : std::max(b, begin + (end - begin) * (t + 1) / nth);
                while (e < end && f.data[e - 1] != '\n') ++e;
                if (e > b) chunks.emplace_back(b, e);
                b = e;
            }
            std::vector<std::vector<int32_t>> toks(chunks.size());
            std::vector<std::vector<int64_t>> lens(chunks.size());
            auto work = [&](size_t t) {
                auto& tk = toks[t];
                auto& ln = lens[t];
                for_lines(f.data + chunks[t].first, f.data + chunks[t].second, [&](const char* lb, const char* le) {
                    int64_t in_chunk = 0;
                    for_tokens(lb, le, java_mode, [&](std::string_view w) {
                        const FlatStrMap::Slot* sl = idx.find(w, FlatStrMap::hash(w));
                        if (!sl) return;
                        if (in_chunk == max_len) { ln.push_back(in_chunk); in_chunk = 0; }
                        tk.push_back((int32_t)sl->val);
                        ++in_chunk;
                    });
                    if (in_chunk > 0) ln.push_back(in_chunk);
                });
            };
            std::vector<std::thread> th;
            for (size_t t = 1; t < chunks.size(); ++t) th.emplace_back(work, t);
            if (!chunks.empty()) work(0);
            for (auto& x : th) x.join();
            for (size_t t = 0; t < chunks.size(); ++t) {
                if (!toks[t].empty()) fwrite(toks[t].data(), sizeof(int32_t), toks[t].size(), ft);
                offs.clear();
                for (int64_t l : lens[t]) { ntok += l; offs.push_back(ntok); }
                if (!offs.empty()) fwrite(offs.data(), sizeof(int64_t), offs.size(), fo);
                nsent += (int64_t)lens[t].size();
            }
            madvise(const_cast<char*>(f.data + begin), end - begin, MADV_DONTNEED);      // drop the pages of the block
            begin = end;
        }
        const bool ok = fclose(ft) == 0;
        const bool ok2 = fclose(fo) == 0;
        if (!ok || !ok2) throw std::runtime_error("write error on " + out_prefix + ".*");
    }
    return py::make_tuple(ntok, nsent);
}

py::tuple vose_alias(py::array_t<double, py::array::c_style | py::array::forcecast> p) {
    const int64_t v = (int64_t)p.size();
    const double* pp = p.data();
    py::array_t<double> prob(v);
    py::array_t<int64_t> alias(v);
    double* pr = prob.mutable_data();
    int64_t* al = alias.mutable_data();
    {
        py::gil_scoped_release rel;
        std::vector<double> scaled(v);
        std::vector<int64_t> small, large;
        small.reserve(v); large.reserve(v);
        for (int64_t i = 0; i < v; ++i) {
            scaled[i] = pp[i] * (double)v;
            pr[i] = 1.0; al[i] = i;
            (scaled[i] < 1.0 ? small : large).push_back(i);
        }
        while (!small.empty() && !large.empty()) {
            int64_t s = small.back(); small.pop_back();
            int64_t l = large.back(); large.pop_back();
            pr[s] = scaled[s];
            al[s] = l;
            scaled[l] = (scaled[l] + scaled[s]) - 1.0;
            (scaled[l] < 1.0 ? small : large).push_back(l);
        }
    }
    return py::make_tuple(prob, alias);
}

}  // namespace

PYBIND11_MODULE(_host, m) {
    m.doc() = "glint_word2vec_b200 native host library";
    m.def("count_words", &count_words);
    m.def("encode_corpus", &encode_corpus);
    m.def("count_words_file", &count_words_file, py::arg("path"), py::arg("java_mode") = true, py::arg("num_threads") = 0);
    m.def("encode_file", &encode_file, py::arg("path"), py::arg("words"), py::arg("max_len"),
          py::arg("java_mode") = true, py::arg("num_threads") = 0);
    m.def("encode_file_to", &encode_file_to, py::arg("path"), py::arg("words"), py::arg("max_len"), py::arg("java_mode") = true,
          py::arg("num_threads") = 0, py::arg("out_prefix"), py::arg("block_bytes") = (int64_t)256 << 20);
    m.def("vose_alias", &vose_alias);
}
