// CPU-only check of poa_b200::dedup_sequences (reference src/smooth.cpp:217-241) and poa_b200_encode_bases
// (reference src/smooth.cpp:304-313): reads "name strand seq" lines, prints the groups.
#include <cstdio>
#include <iostream>
#include <sstream>
#include "poa_b200_smooth.hpp"
int main(int argc, char **argv) {
    if (argc > 1 && std::string(argv[1]) == "identity") {  // stdin: "block seq" lines -> one threshold per block, or the error text
        std::vector<std::vector<std::string>> ranges;
        std::string line;
        while (std::getline(std::cin, line)) {
            std::istringstream is(line); size_t b; std::string q;
            if (!(is >> b >> q)) continue;
            if (ranges.size() <= b) ranges.resize(b + 1);
            ranges[b].push_back(q);
        }
        try {
            for (float t : poa_b200::estimate_block_identity(0, ranges)) printf("%.9g\n", t);
        } catch (const std::exception &e) { printf("error: %s\n", e.what()); return 3; }
        return 0;
    }
    if (argc > 1) {  // preset mode: thresholds on the command line -> "m n g e q c" or "-" per line (src/smooth.cpp:2026-2062)
        for (int a = 1; a < argc; ++a) {
            int m = 0, n = 0, g = 0, e = 0, q = 0, c = 0;
            if (poa_b200::adaptive_poa_preset(std::stof(argv[a]), m, n, g, e, q, c)) printf("%d %d %d %d %d %d\n", m, n, g, e, q, c);
            else printf("-\n");
        }
        return 0;
    }
    std::vector<std::string> seqs, names; std::vector<bool> revs;
    std::string line;
    while (std::getline(std::cin, line)) {
        std::istringstream is(line); std::string n, s, q;
        if (!(is >> n >> s)) continue;
        is >> q; names.push_back(n); revs.push_back(s == "-"); seqs.push_back(q);
    }
    const poa_b200::block_sequences b = poa_b200::dedup_sequences(seqs, names, revs);
    for (size_t i = 0; i < b.seqs.size(); ++i) {
        std::vector<uint8_t> codes(b.seqs[i].size());
        poa_b200_encode_bases(b.seqs[i].data(), (int64_t)codes.size(), codes.data());
        printf("%d", b.weights[i]);
        for (size_t z = 0; z < b.dup_seq_names[i].size(); ++z) printf(" %s%c", b.dup_seq_names[i][z].c_str(), b.dup_is_revs[i][z] ? '-' : '+');
        printf(" ");
        for (uint8_t c : codes) printf("%d", (int)c);
        printf("\n");
    }
    return 0;
}
