// poa_b200::extract_range_sequence (include/poa_b200_smooth.hpp; reference src/smooth.cpp:75-126, :177-214) on a mock path graph:
// one path of oriented nodes; prints the padded, oriented string of a step range.  (Against the real thing -- xg::XG inside the
// patched smoothxg -- it is compared range by range in verify mode, integration/smoothxg_poa_b200.patch.)
//   usage: extract_test <padding> <begin step> <end step> node[+|-] ...     e.g. extract_test 3 1 3 ACGT+ GG- TTA+ C+
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>
#include "poa_b200_smooth.hpp"

struct MockGraph {
    struct Node { std::string fwd; bool rev; };
    std::vector<Node> steps;  // the single path
    static std::string rc(std::string s) {
        std::reverse(s.begin(), s.end());
        for (auto &c : s) c = c == 'A' ? 'T' : c == 'C' ? 'G' : c == 'G' ? 'C' : c == 'T' ? 'A' : 'N';
        return s;
    }
    int get_path_handle_of_step(int) const { return 0; }
    int path_begin(int) const { return 0; }
    int path_end(int) const { return (int)steps.size(); }
    int get_handle_of_step(int s) const { return s; }
    uint64_t get_length(int h) const { return steps[(size_t)h].fwd.size(); }
    std::string get_sequence(int h) const { return steps[(size_t)h].rev ? rc(steps[(size_t)h].fwd) : steps[(size_t)h].fwd; }
    bool get_is_reverse(int h) const { return steps[(size_t)h].rev; }
    int get_previous_step(int s) const { return s - 1; }
    int get_next_step(int s) const { return s + 1; }
};

int main(int argc, char **argv) {
    if (argc < 5) { fprintf(stderr, "usage\n"); return 2; }
    MockGraph g;
    for (int i = 4; i < argc; ++i) { std::string a = argv[i]; g.steps.push_back({a.substr(0, a.size() - 1), a.back() == '-'}); }
    bool rev = false;
    const std::string s = poa_b200::extract_range_sequence(g, atoi(argv[2]), atoi(argv[3]), atoi(argv[1]), [](std::string &x) { x = MockGraph::rc(x); }, &rev);
    printf("%s %d\n", s.c_str(), rev ? 1 : 0);
    return 0;
}
