// bow_oracle.cpp — CPU restatement of Frame::ComputeBoW (src/Frame.cc:721-735), i.e. DBoW2's
// TemplatedVocabulary<FORB::TDescriptor, FORB>::transform(features, BowVector&, FeatureVector&, levelsup)
// (Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1127-1194, :1218-1259), FORB::distance (FORB.cpp:95-110),
// BowVector::addWeight / addIfNotExist / normalize (BowVector.cpp:34-84), FeatureVector::addFeature (FeatureVector.cpp:29-43)
// and the text loader loadFromTextFile (TemplatedVocabulary.h:1338-1422).  SURVEY.md §8(f) rank 1.
// TEST INFRASTRUCTURE ONLY: the checker of the CUDA path (tests/, smoke(), bench.py's CPU legs), never the product.
// DBoW2 is vendored in the reference but includes OpenCV headers (cv::Mat descriptors, cv::FileStorage), which this image
// lacks, so it cannot be compiled here: parity of this row is "unpinned" in the strict sense (cross-checked by a naive
// Python restatement in tests/test_oracle_bow.py).
#include <cmath>
#include <cstdint>
#include <cstring>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <string>
#include <vector>

#include "oracle.h"

namespace {
struct Node {
    std::vector<unsigned> children;
    unsigned parent = 0, word_id = 0;
    double weight = 0;
    uint8_t desc[32] = {0};
};
// FORB::distance (FORB.cpp:85-110): bit-parallel popcount over eight 32-bit words
int forb_distance(const uint8_t* a, const uint8_t* b) {
    int dist = 0;
    for (int i = 0; i < 8; i++) {
        uint32_t x, y;
        memcpy(&x, a + 4 * i, 4);
        memcpy(&y, b + 4 * i, 4);
        unsigned int v = x ^ y;
        v = v - ((v >> 1) & 0x55555555);
        v = (v & 0x33333333) + ((v >> 2) & 0x33333333);
        dist += (((v + (v >> 4)) & 0xF0F0F0F) * 0x1010101) >> 24;
    }
    return dist;
}
}  // namespace

struct orc_voc {
    int k = 0, L = 0, scoring = 0, weighting = 0, n_words = 0;
    std::vector<Node> nodes;
};

extern "C" orc_voc* orc_voc_create(int k, int L, int scoring, int weighting, int n_nodes, const int* parent, const uint8_t* is_leaf,
                                   const uint8_t* desc, const double* weight) {
    orc_voc* v = new orc_voc;
    v->k = k; v->L = L; v->scoring = scoring; v->weighting = weighting;
    v->nodes.resize((size_t)n_nodes + 1);
    for (int i = 0; i < n_nodes; i++) {  // file order: node id = line number (TemplatedVocabulary.h:1385-1417)
        const int nid = i + 1;
        Node& nd = v->nodes[nid];
        if (parent[i] < 0 || parent[i] >= nid) { delete v; return nullptr; }
        nd.parent = (unsigned)parent[i];
        v->nodes[parent[i]].children.push_back((unsigned)nid);
        memcpy(nd.desc, desc + 32 * (size_t)i, 32);
        nd.weight = weight[i];
        if (is_leaf[i]) nd.word_id = (unsigned)v->n_words++;
    }
    return v;
}

// C stdio on purpose: this library carries a static libstdc++ and must not touch iostream globals inside a foreign process
extern "C" orc_voc* orc_voc_load_text(const char* filename) {
    FILE* f = fopen(filename, "r");
    if (!f) return nullptr;
    std::vector<char> line(1 << 16);
    int k = -1, L = -1, n1 = -1, n2 = -1;
    if (!fgets(line.data(), (int)line.size(), f) || sscanf(line.data(), "%d %d %d %d", &k, &L, &n1, &n2) != 4 || k < 0 || k > 20 || L < 1 || L > 10 ||
        n1 < 0 || n1 > 5 || n2 < 0 || n2 > 3) {
        fclose(f);
        return nullptr;
    }
    std::vector<int> parent;
    std::vector<uint8_t> leaf, desc;
    std::vector<double> weight;
    while (fgets(line.data(), (int)line.size(), f)) {
        char* p = line.data();
        if (p[strspn(p, " \t\r\n")] == 0) continue;  // the reference would turn a trailing blank line into a bogus node
        char* e;
        const long pid = strtol(p, &e, 10); p = e;
        const long nIsLeaf = strtol(p, &e, 10); p = e;
        parent.push_back((int)pid);
        leaf.push_back(nIsLeaf > 0);
        for (int i = 0; i < 32; i++) { desc.push_back((uint8_t)strtol(p, &e, 10)); p = e; }
        weight.push_back(strtod(p, &e));
    }
    fclose(f);
    return orc_voc_create(k, L, n1, n2, (int)parent.size(), parent.data(), leaf.data(), desc.data(), weight.data());
}

extern "C" void orc_voc_destroy(orc_voc* v) { delete v; }
extern "C" int orc_voc_info(const orc_voc* v, int* k, int* L, int* n_nodes, int* n_words) {
    *k = v->k; *L = v->L; *n_nodes = (int)v->nodes.size(); *n_words = v->n_words;
    return 0;
}

extern "C" int orc_voc_transform(const orc_voc* v, const uint8_t* desc, int n, int levelsup, int* n_words, unsigned* word_id, double* word_value,
                                 int* n_fv_nodes, unsigned* node_id, int* node_off, unsigned* feat_idx) {
    std::map<unsigned, double> bow;
    std::map<unsigned, std::vector<unsigned>> fv;
    *n_words = 0;
    *n_fv_nodes = 0;
    node_off[0] = 0;
    if (v->nodes.size() <= 1) return 0;  // empty()
    const bool must = v->scoring != 5;                 // DotProductScoring does not normalise (ScoringObject.h:73-95)
    const bool l2 = v->scoring == 1;
    const bool tf = v->weighting == 0 || v->weighting == 1;  // TF_IDF, TF
    for (int i = 0; i < n; i++) {
        const uint8_t* feature = desc + 32 * (size_t)i;
        const int nid_level = v->L - levelsup;
        unsigned nid = 0;  // the reference leaves it uninitialised when the leaf is shallower than nid_level; 0 here
        unsigned final_id = 0;
        int current_level = 0;
        do {
            ++current_level;
            const std::vector<unsigned>& nodes = v->nodes[final_id].children;
            final_id = nodes[0];
            double best_d = forb_distance(feature, v->nodes[final_id].desc);
            for (size_t c = 1; c < nodes.size(); c++) {
                const double d = forb_distance(feature, v->nodes[nodes[c]].desc);
                if (d < best_d) { best_d = d; final_id = nodes[c]; }
            }
            if (current_level == nid_level) nid = final_id;
        } while (!v->nodes[final_id].children.empty());
        const unsigned id = v->nodes[final_id].word_id;
        const double w = v->nodes[final_id].weight;
        if (w > 0) {
            if (tf) bow[id] += w;                       // addWeight
            else if (!bow.count(id)) bow[id] = w;       // addIfNotExist
            fv[nid].push_back((unsigned)i);
        }
    }
    if (tf && !bow.empty() && !must) {
        const double nd = (double)bow.size();
        for (auto& kv : bow) kv.second /= nd;
    }
    if (must) {
        double norm = 0.0;
        if (!l2) { for (auto& kv : bow) norm += std::fabs(kv.second); }
        else { for (auto& kv : bow) norm += kv.second * kv.second; norm = std::sqrt(norm); }
        if (norm > 0.0) for (auto& kv : bow) kv.second /= norm;
    }
    int nw = 0;
    for (auto& kv : bow) { word_id[nw] = kv.first; word_value[nw] = kv.second; nw++; }
    int nn = 0, nf = 0;
    for (auto& kv : fv) {
        node_id[nn] = kv.first;
        for (unsigned x : kv.second) feat_idx[nf++] = x;
        node_off[++nn] = nf;
    }
    *n_words = nw;
    *n_fv_nodes = nn;
    return 0;
}
