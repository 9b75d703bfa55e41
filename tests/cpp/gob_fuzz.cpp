// Fuzz of the gob reader (groot_b200/csrc/host/gob_reader.cpp) and the index validation behind it: takes a valid
// groot.gg + groot.lshe pair, damages one of them (byte flips, truncation, inserted / duplicated / zeroed stretches,
// inflated varints) and loads the pair. Every outcome but a clean load or a std::exception is a bug — built with ASan / UBSan.
// The same for the library's own flat index file: the pair is loaded, saved as groot.grootb200 (save_index) and that
// file is damaged and loaded (load_index + validate_index).
//   gob_fuzz <groot.gg> <groot.lshe> <tmp dir> <seed> <iterations> [flat]     prints "ok <loaded> <refused>"
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <random>
#include <string>
#include <vector>

#include "../../groot_b200/csrc/flat_index.h"

static std::vector<char> slurp(const char* p) { std::ifstream f(p, std::ios::binary); return std::vector<char>((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>()); }
static void spit(const std::string& p, const std::vector<char>& v) { std::ofstream f(p, std::ios::binary); f.write(v.data(), static_cast<std::streamsize>(v.size())); }

int main(int argc, char** argv) {
    if (argc < 6) return 1;
    const std::string dir = argv[3];
    const bool flat = argc > 6 && std::string(argv[6]) == "flat";
    if (flat) { groot::FlatIndex ix; groot::load_index_gob(ix, argv[1], argv[2]); groot::save_index(ix, dir + "/ok.grootb200"); }
    const std::vector<char> gg = flat ? slurp((dir + "/ok.grootb200").c_str()) : slurp(argv[1]), lshe = flat ? gg : slurp(argv[2]);
    std::mt19937 rng(static_cast<unsigned>(atoi(argv[4])));
    const int iters = atoi(argv[5]);
    auto rnd = [&](size_t n) { return static_cast<size_t>(rng() % n); };
    int loaded = 0, refused = 0;
    for (int it = 0; it < iters; it++) {
        std::vector<char> a = gg, b = lshe;
        std::vector<char>& v = rnd(2) ? a : b;
        for (int k = 1 + static_cast<int>(rnd(3)); k > 0 && !v.empty(); k--) {
            const size_t at = rnd(4) ? rnd(std::min<size_t>(v.size(), 600)) : rnd(v.size());    // mostly in the type definitions and headers
            switch (rnd(7)) {
                case 0: v[at] = static_cast<char>(rnd(256)); break;
                case 1: v[at] = static_cast<char>(v[at] ^ (1 << rnd(8))); break;
                case 2: v.resize(at); break;
                case 3: v.insert(v.begin() + static_cast<long>(at), static_cast<size_t>(1 + rnd(9)), static_cast<char>(rnd(256))); break;
                case 4: { const size_t n = std::min(v.size() - at, 1 + rnd(64)); std::vector<char> piece(v.begin() + static_cast<long>(at), v.begin() + static_cast<long>(at + n)); v.insert(v.begin() + static_cast<long>(at), piece.begin(), piece.end()); break; }
                case 5: { const size_t n = std::min(v.size() - at, 1 + rnd(16)); std::fill(v.begin() + static_cast<long>(at), v.begin() + static_cast<long>(at + n), 0); break; }
                default: { const char big[9] = {static_cast<char>(0xF8), 0x7f, -1, -1, -1, -1, -1, -1, -1}; v.insert(v.begin() + static_cast<long>(at), big, big + 9); break; }   // a varint of 2^63-ish
            }
        }
        if (flat) spit(dir + "/f.grootb200", v); else { spit(dir + "/f.gg", a); spit(dir + "/f.lshe", b); }
        try {
            groot::FlatIndex ix;
            if (flat) groot::load_index(ix, dir + "/f.grootb200"); else groot::load_index_gob(ix, dir + "/f.gg", dir + "/f.lshe");
            groot::validate_index(ix);
            loaded++;
        } catch (std::exception&) { refused++; }
    }
    printf("ok %d %d\n", loaded, refused);
    return 0;
}
