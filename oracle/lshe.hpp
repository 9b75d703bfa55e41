// ORACLE — TEST INFRASTRUCTURE ONLY (see nthash.hpp header).
//
// CPU restatement of src/lshe/lshe.go (ContainmentIndex: AddWindow, LoadFromBytes bootstrap,
// Query + full containment check) and of what it calls in github.com/ekzhu/lshensemble v1.1.0
// (go.mod:10; call sites lshe.go:134,145,157,165): BootstrapLshEnsembleEquiDepth, LshForest
// (32-bit hash keys), OptimalKL with its numeric integration, Query and Containment.
//
// PARITY STATUS: lshensemble is not vendored -> restated from its published algorithm (Zhu et al.,
// "LSH Ensemble", VLDB 2016, and the package's probability.go / lshforest.go). Candidate-level
// behaviour is "parity unpinned" by the reference repo; the end-to-end behaviour is pinned by
// src/pipeline/3_sketch_test.go:49-58 and testing/run_travis_tests.sh:43-56 (tests/test_oracle_kat.py).
// pow() here is libm's, Go uses its own math.Pow: a last-ulp difference could only matter on an
// exact tie of the (K,L) optimiser's error sums.
#pragma once
#include <algorithm>
#include <array>
#include <cmath>
#include <cstring>
#include <map>
#include <string>
#include <tuple>
#include <vector>

#include "graph.hpp"

namespace oracle {

// ---- lshensemble/probability.go ----
constexpr double kIntegrationPrecision = 0.01;

template <class F>
inline double integral(F f, double a, double b, double precision) {
    double area = 0;
    for (double x = a; x < b; x += precision) area += f(x + 0.5 * precision) * precision;
    return area;
}
inline double prob_false_negative(int x, int q, int l, int k, double t, double precision) {
    auto fn = [=](double s) { return 1.0 - (1.0 - std::pow(1.0 - std::pow(s / (1.0 + double(x) / double(q) - s), double(k)), double(l))); };
    double xq = double(x) / double(q);
    if (xq >= 1.0) return integral(fn, t, 1.0, precision);
    if (xq >= t) return integral(fn, t, xq, precision);
    return 0.0;
}
inline double prob_false_positive(int x, int q, int l, int k, double t, double precision) {
    auto fp = [=](double s) { return 1.0 - std::pow(1.0 - std::pow(s / (1.0 + double(x) / double(q) - s), double(k)), double(l)); };
    double xq = double(x) / double(q);
    if (xq >= 1.0) return integral(fp, 0.0, t, precision);
    if (xq >= t) return integral(fp, 0.0, t, precision);
    return integral(fp, 0.0, xq, precision);
}
// LshForest.OptimalKL: argmin over l=1..L (outer), k=1..K (inner), strict '<'
inline void optimal_kl(int maxK, int maxL, int x, int q, double t, int* optK, int* optL) {
    double minError = 1.7976931348623157e308;
    *optK = 0; *optL = 0;
    for (int l = 1; l <= maxL; l++)
        for (int k = 1; k <= maxK; k++) {
            double fp = prob_false_positive(x, q, l, k, t, kIntegrationPrecision);
            double fn = prob_false_negative(x, q, l, k, t, kIntegrationPrecision);
            double err = fn + fp;
            if (minError > err) { minError = err; *optK = k; *optL = l; }
        }
}

// lshensemble.Containment (called at lshe.go:165)
inline double containment(const uint64_t* q, const uint64_t* x, int sigLen, int qSize, int xSize) {
    if (qSize == 0 || xSize == 0) return 0.0;
    int eq = 0;
    for (int i = 0; i < sigLen; i++) if (x[i] == q[i]) eq++;
    if (eq == 0) return 0.0;
    double jaccard = double(eq) / double(sigLen);
    return (double(xSize) / double(qSize) + 1.0) * jaccard / (1.0 + jaccard);
}
// smallest number of equal slots for which containment(...) > threshold; S+1 when unsatisfiable
inline int eq_min_for(int S, int qSize, int xSize, double threshold) {
    for (int eq = 1; eq <= S; eq++) {
        double jaccard = double(eq) / double(S);
        double c = (double(xSize) / double(qSize) + 1.0) * jaccard / (1.0 + jaccard);
        if (qSize != 0 && xSize != 0 && c > threshold) return eq;
    }
    return S + 1;
}

// ---- LshForest with 32-bit hash values (the package default) ----
using BandKey = std::array<uint32_t, 8>;  // low 32 bits of up to maxK(<=8) consecutive hashes
struct LshForest {
    int k = 0, l = 0;
    // per band: (key, window index) sorted by key bytes == sorted lexicographically by the
    // little-endian 4-byte groups, as the Go string compare does
    std::vector<std::vector<std::pair<BandKey, uint32_t>>> tables;
    static bool key_less(const BandKey& a, const BandKey& b, int n) {
        for (int i = 0; i < n; i++) {
            if (a[i] == b[i]) continue;
            uint32_t x = __builtin_bswap32(a[i]), y = __builtin_bswap32(b[i]);  // byte-wise LE string order
            return x < y;
        }
        return false;
    }
    void init(int k_, int l_) { k = k_; l = l_; tables.assign(l, {}); }
    void add(uint32_t win, const uint64_t* sig) {
        for (int i = 0; i < l; i++) {
            BandKey bk{};
            for (int j = 0; j < k; j++) bk[j] = static_cast<uint32_t>(sig[i * k + j]);
            tables[i].push_back({bk, win});
        }
    }
    void index() {
        for (auto& t : tables)
            std::stable_sort(t.begin(), t.end(), [this](const auto& a, const auto& b) { return key_less(a.first, b.first, k); });
    }
    // LshForest.Query(sig, K, L): prefix probe of the first K hashes of bands 0..L-1, de-duplicated
    void query(const uint64_t* sig, int K, int L, std::vector<uint32_t>* out) const {
        size_t first = out->size();
        for (int i = 0; i < L; i++) {
            BandKey probe{};
            for (int j = 0; j < K; j++) probe[j] = static_cast<uint32_t>(sig[i * k + j]);
            auto& ht = tables[i];
            auto lo = std::lower_bound(ht.begin(), ht.end(), probe, [K](const auto& e, const BandKey& p) { return key_less(e.first, p, K); });
            for (auto it = lo; it != ht.end(); ++it) {
                bool same = true;
                for (int j = 0; j < K; j++) if (it->first[j] != probe[j]) { same = false; break; }
                if (!same) break;
                bool seen = false;
                for (size_t s = first; s < out->size(); s++) if ((*out)[s] == it->second) { seen = true; break; }
                if (!seen) out->push_back(it->second);
            }
        }
    }
};

// ---- src/lshe/lshe.go ContainmentIndex ----
struct ContainmentIndex {
    int numPart = 0, maxK = 0, numWindowKmers = 0, sketchSize = 0;
    std::vector<std::string> lookupNames;  // "g%dn%do%d-%d" (pipeline/index.go:199)
    std::vector<Key> windows;              // WindowLookup values, canonical order (sorted by graph, node, offset, dup idx)
    // bootstrap result
    std::vector<LshForest> parts;
    std::vector<int> partUpper;
    std::map<std::tuple<int, int, double>, std::pair<int, int>> paramCache;

    void init(int np, int mk, int nwk, int ss) { numPart = np; maxK = mk; numWindowKmers = nwk; sketchSize = ss; }
    void addWindow(const std::string& name, const Key& w) { lookupNames.push_back(name); windows.push_back(w); }

    // lshe.go:108-146 -> lshensemble.BootstrapLshEnsembleEquiDepth(numPart, numHash, maxK, total, recs):
    // depth = total/numPart records per partition in arrival order (Go: map order; here: canonical
    // window order), Upper = size of the last record of the partition (all sizes are equal here).
    void bootstrap() {
        int total = static_cast<int>(windows.size());
        if (total == 0) throw std::runtime_error("loaded an empty index file");
        int np = std::max(1, numPart);
        parts.assign(np, LshForest());
        partUpper.assign(np, numWindowKmers);
        for (auto& p : parts) p.init(maxK, sketchSize / maxK);
        int depth = total / np;
        int cur = 0, count = 0;
        for (int w = 0; w < total; w++) {
            parts[cur].add(static_cast<uint32_t>(w), windows[w].sketch.data());
            count++;
            if (depth > 0 && count % depth == 0 && cur < np - 1) cur++;
        }
        for (auto& p : parts) p.index();
    }

    std::pair<int, int> params(int x, int q, double t) {
        auto key = std::make_tuple(x, q, t);
        auto it = paramCache.find(key);
        if (it != paramCache.end()) return it->second;
        int K, L;
        optimal_kl(maxK, sketchSize / maxK, x, q, t, &K, &L);
        paramCache[key] = {K, L};
        return {K, L};
    }

    // lshe.go:153-175: window indices (ascending) that pass the LSH probe AND the containment check
    void query(const uint64_t* sig, int querySize, double threshold, std::vector<uint32_t>* hits) {
        hits->clear();
        std::vector<uint32_t> cand;
        for (size_t p = 0; p < parts.size(); p++) {
            auto kl = params(partUpper[p], querySize, threshold);
            parts[p].query(sig, kl.first, kl.second, &cand);
        }
        for (uint32_t w : cand)
            if (containment(sig, windows[w].sketch.data(), sketchSize, querySize, numWindowKmers) > threshold) hits->push_back(w);
        std::sort(hits->begin(), hits->end());
    }
};

}  // namespace oracle
