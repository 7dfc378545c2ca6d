// Native host-side text processing (pybind11, no torch dependency).
//
//  count_words   : word -> count over an iterable of token sequences.  The reference does this
//                  with a Spark shuffle (flatMap/map/reduceByKey, MLLIB:259-262); here it is an
//                  open-addressing hash map keyed by the UTF-8 bytes, sharded over threads once the
//                  Python strings have been flattened (the GIL-bound part is only the flattening).
//  encode_corpus : words -> vocabulary indices with OOV dropped and sentences chunked at
//                  maxSentenceLength (MLLIB:335-343).
//  vose_alias    : Vose alias construction for the unigram^0.75 noise distribution (replaces the
//                  Glint servers' 1e8-entry unigram table, ML:204-206).
#include <pybind11/pybind11.h>
#include <pybind11/numpy.h>
#include <pybind11/stl.h>

#include <algorithm>
#include <cstdint>
#include <cstring>
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
    m.def("vose_alias", &vose_alias);
}
