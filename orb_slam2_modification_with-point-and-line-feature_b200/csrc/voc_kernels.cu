// voc_kernels.cu — Frame::ComputeBoW on sm_100a: DBoW2's vocabulary-tree transform, the CUDA path
// behind pl_voc_* (include/plslam_c.h).
//
// Reference functions replaced:
//   Frame::ComputeBoW                                        src/Frame.cc:721-735
//   TemplatedVocabulary::transform (features -> BowVector, FeatureVector)   Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1127-1194
//   TemplatedVocabulary::transform (one feature down the tree)              :1218-1259
//   FORB::distance                                           Thirdparty/DBoW2/DBoW2/FORB.cpp:85-110
//   BowVector::addWeight / addIfNotExist / normalize         Thirdparty/DBoW2/DBoW2/BowVector.cpp:34-84
//   FeatureVector::addFeature                                Thirdparty/DBoW2/DBoW2/FeatureVector.cpp:29-43
//   TemplatedVocabulary::loadFromTextFile                    Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1338-1422
//
// Layout in HBM: the children of a node are stored next to each other (32-byte descriptors, child order = file order), so one
// warp reads the <= 32 children of a node as one contiguous block; ORBvoc (k = 10, L = 6: 1.1M nodes) is 36 MB and stays in L2.
// k_voc_descend: one warp per feature, lane = child, popc distance + shuffle arg-min (first minimum, as `d < best_d`).
// k_voc_assemble: one CTA per frame turns the per-feature (word, weight, node) triples into the two std::map's of the
// reference — sorted runs out of a shared-memory bitonic sort; a word's value is formed by repeated addition in feature order
// and the norm is summed in ascending word order, so every double is the reference's.
#include <stdlib.h>

#include "match_common.cuh"

namespace pl {

struct VocDev {
    const int* child_begin;   // per node: first slot of its children in the child-ordered arrays
    const int* child_count;   // per node (0 = leaf)
    const unsigned* slot_node;  // child slot -> node id
    const uint4* slot_desc;     // child slot -> descriptor (2 x uint4)
    const unsigned* word_id;    // per node
    const double* weight;       // per node
    int L;
};

__global__ void __launch_bounds__(256) k_voc_descend(VocDev V, const uint4* __restrict__ desc, int total, int nid_level, unsigned* __restrict__ o_word,
                                                     double* __restrict__ o_weight, unsigned* __restrict__ o_nid) {
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (i >= total) return;
    const uint4 f0 = desc[2 * (size_t)i], f1 = desc[2 * (size_t)i + 1];
    unsigned final_id = 0, nid = 0;
    int level = 0;
    int cnt = V.child_count[0];
    while (cnt > 0) {
        ++level;
        const int beg = V.child_begin[final_id];
        unsigned key = 0xFFFFFFFFu;
        if (lane < cnt) key = ((unsigned)hamming256(f0, f1, V.slot_desc[2 * (size_t)(beg + lane)], V.slot_desc[2 * (size_t)(beg + lane) + 1]) << 8) | (unsigned)lane;
#pragma unroll
        for (int sft = 16; sft > 0; sft >>= 1) key = min(key, __shfl_xor_sync(0xffffffffu, key, sft));
        final_id = V.slot_node[beg + (int)(key & 0xFFu)];
        if (level == nid_level) nid = final_id;
        cnt = V.child_count[final_id];
    }
    if (lane == 0) {
        o_word[i] = V.word_id[final_id];
        o_weight[i] = V.weight[final_id];
        o_nid[i] = nid;
    }
}

constexpr int kVocThreads = 1024;
constexpr int kVocMaxFeat = 8192;

__device__ void block_bitonic_u64(unsigned long long* k, int n2) {
    for (int a = 2; a <= n2; a <<= 1)
        for (int j = a >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < n2; i += blockDim.x) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const bool up = (i & a) == 0;
                    const unsigned long long x = k[i], y = k[ixj];
                    if (up ? (x > y) : (x < y)) { k[i] = y; k[ixj] = x; }
                }
            }
            __syncthreads();
        }
}

// exclusive scan of s_v[0..n) in place (n <= kVocMaxFeat); returns the total.  blockDim.x == kVocThreads
__device__ int block_excl_scan(int* s_v, int n, int* s_warp, int* s_carry) {
    const int tid = threadIdx.x;
    if (tid == 0) *s_carry = 0;
    __syncthreads();
    for (int base = 0; base < n; base += kVocThreads) {
        const int i = base + tid;
        const int v = i < n ? s_v[i] : 0;
        int incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if ((tid & 31) >= o) incl += t;
        }
        if ((tid & 31) == 31) s_warp[tid >> 5] = incl;
        __syncthreads();
        if (tid < 32) {
            const int w0 = s_warp[tid];
            int w = w0;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, w, o);
                if (tid >= o) w += t;
            }
            s_warp[tid] = w - w0;
            if (tid == 31) s_warp[32] = w;
        }
        __syncthreads();
        const int carry = *s_carry;
        if (i < n) s_v[i] = carry + s_warp[tid >> 5] + incl - v;
        __syncthreads();
        if (tid == 0) *s_carry = carry + s_warp[32];
        __syncthreads();
    }
    return *s_carry;
}

__global__ void __launch_bounds__(kVocThreads) k_voc_assemble(const int* __restrict__ off, const unsigned* __restrict__ f_word,
                                                              const double* __restrict__ f_weight, const unsigned* __restrict__ f_nid, int scoring,
                                                              int weighting, int* __restrict__ n_words, unsigned* __restrict__ word_id,
                                                              double* __restrict__ word_value, int* __restrict__ n_fv_nodes,
                                                              unsigned* __restrict__ node_id, int* __restrict__ node_off,
                                                              unsigned* __restrict__ feat_idx) {
    extern __shared__ __align__(16) unsigned char s_raw[];
    __shared__ int s_warp[33];
    __shared__ int s_carry;
    __shared__ double s_norm;
    const int f = blockIdx.x, tid = threadIdx.x;
    const int o = off[f], n = off[f + 1] - o;
    int n2 = 1;
    while (n2 < n) n2 <<= 1;
    unsigned long long* keys = (unsigned long long*)s_raw;
    int* pos = (int*)(keys + n2);
    const bool tf = weighting == 0 || weighting == 1;
    const bool must = scoring != 5, l2 = scoring == 1;
    // ---- BowVector: runs of equal word id, in ascending word order ----
    for (int i = tid; i < n2; i += kVocThreads) {
        unsigned long long k = ~0ull;
        if (i < n && f_weight[o + i] > 0) k = ((unsigned long long)f_word[o + i] << 32) | (unsigned)i;
        keys[i] = k;
    }
    __syncthreads();
    block_bitonic_u64(keys, n2);
    for (int i = tid; i < n; i += kVocThreads) {
        const unsigned long long k = keys[i];
        pos[i] = (k != ~0ull && (i == 0 || (unsigned)(keys[i - 1] >> 32) != (unsigned)(k >> 32))) ? 1 : 0;
    }
    __syncthreads();
    // pos[] becomes the exclusive scan; a head is an element whose flag was 1: recomputed from the keys below
    const int nw = block_excl_scan(pos, n, s_warp, &s_carry);
    for (int i = tid; i < n; i += kVocThreads) {
        const unsigned long long k = keys[i];
        if (k == ~0ull) continue;
        const unsigned w = (unsigned)(k >> 32);
        if (i != 0 && (unsigned)(keys[i - 1] >> 32) == w) continue;
        int cnt = 1;
        while (i + cnt < n && keys[i + cnt] != ~0ull && (unsigned)(keys[i + cnt] >> 32) == w) cnt++;
        const double wt = f_weight[o + (int)(k & 0xFFFFFFFFu)];
        double v = wt;                                      // addWeight: v += w once per feature, in feature order (BowVector.cpp:34-46)
        if (tf) for (int c = 1; c < cnt; c++) v = __dadd_rn(v, wt);
        word_id[o + pos[i]] = w;
        word_value[o + pos[i]] = v;
    }
    __threadfence_block();
    __syncthreads();
    if (tid == 0) {
        double norm = 0.0;
        if (must) {                                         // BowVector::normalize (BowVector.cpp:62-84), ascending word id
            if (!l2) for (int k = 0; k < nw; k++) norm = __dadd_rn(norm, fabs(word_value[o + k]));
            else { for (int k = 0; k < nw; k++) norm = __dadd_rn(norm, __dmul_rn(word_value[o + k], word_value[o + k])); norm = sqrt(norm); }
        } else if (tf) {
            norm = (double)nw;                              // `vit->second /= nd` (TemplatedVocabulary.h:1163-1169)
        }
        s_norm = norm;
        n_words[f] = nw;
    }
    __syncthreads();
    if (s_norm > 0.0)
        for (int k = tid; k < nw; k += kVocThreads) word_value[o + k] = word_value[o + k] / s_norm;
    __syncthreads();
    // ---- FeatureVector: (node, feature) ascending ----
    for (int i = tid; i < n2; i += kVocThreads) {
        unsigned long long k = ~0ull;
        if (i < n && f_weight[o + i] > 0) k = ((unsigned long long)f_nid[o + i] << 32) | (unsigned)i;
        keys[i] = k;
    }
    __syncthreads();
    block_bitonic_u64(keys, n2);
    for (int i = tid; i < n; i += kVocThreads) {
        const unsigned long long k = keys[i];
        pos[i] = (k != ~0ull && (i == 0 || (unsigned)(keys[i - 1] >> 32) != (unsigned)(k >> 32))) ? 1 : 0;
        if (k != ~0ull) feat_idx[o + i] = (unsigned)(k & 0xFFFFFFFFu);
    }
    __syncthreads();
    const int nn = block_excl_scan(pos, n, s_warp, &s_carry);
    int* noff = node_off + o + f;
    for (int i = tid; i < n; i += kVocThreads) {
        const unsigned long long k = keys[i];
        if (k == ~0ull) continue;
        const unsigned nd = (unsigned)(k >> 32);
        if (i != 0 && (unsigned)(keys[i - 1] >> 32) == nd) continue;
        node_id[o + pos[i]] = nd;
        noff[pos[i]] = i;
    }
    if (tid == 0) {
        int kept = 0;  // valid keys sort to the front: their number is the position of the first invalid one
        int lo = 0, hi = n;
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (keys[mid] != ~0ull) lo = mid + 1; else hi = mid; }
        kept = lo;
        noff[nn] = kept;
        n_fv_nodes[f] = nn;
    }
}

}  // namespace pl

using namespace pl;

struct pl_voc {
    int device = 0, k = 0, L = 0, scoring = 0, weighting = 0, n_nodes = 0, n_words = 0;
    cudaStream_t stream = nullptr;
    VocDev dev{};
    void* d_mem[6] = {nullptr};
    PlStage in;
    uint8_t* d_scratch = nullptr;
    size_t scratch_cap = 0;
};

namespace {
inline size_t padb(size_t b) { return PlStage::pad(b); }
}

extern "C" {

PL_API int pl_voc_create(pl_voc** out, int device, int k, int L, int scoring, int weighting, int n_nodes, const int* parent, const uint8_t* is_leaf,
                         const uint8_t* desc, const double* weight) {
    PL_CHECK_ARG(out && k >= 0 && L >= 1 && scoring >= 0 && scoring <= 5 && weighting >= 0 && weighting <= 3 && n_nodes >= 0);
    PL_CHECK_ARG(n_nodes == 0 || (parent && is_leaf && desc && weight));
    *out = nullptr;
    const int nn = n_nodes + 1;
    std::vector<std::vector<unsigned>> children(nn);
    std::vector<unsigned> word(nn, 0);
    std::vector<double> wt(nn, 0.0);
    int n_words = 0;
    for (int i = 0; i < n_nodes; i++) {
        PL_CHECK_ARG(parent[i] >= 0 && parent[i] <= i);  // a parent precedes its children in the file
        children[parent[i]].push_back((unsigned)(i + 1));
        wt[i + 1] = weight[i];
        if (is_leaf[i]) word[i + 1] = (unsigned)n_words++;
    }
    for (int i = 0; i < n_nodes; i++) PL_CHECK_ARG((is_leaf[i] != 0) == children[i + 1].empty());  // isLeaf() == children.empty() (:328)
    std::vector<int> cbeg(nn, 0), ccnt(nn, 0);
    std::vector<unsigned> slot_node;
    std::vector<uint8_t> slot_desc;
    slot_node.reserve(n_nodes);
    slot_desc.reserve((size_t)n_nodes * 32);
    for (int nd = 0; nd < nn; nd++) {
        PL_CHECK_ARG(children[nd].size() <= 32);
        cbeg[nd] = (int)slot_node.size();
        ccnt[nd] = (int)children[nd].size();
        for (unsigned c : children[nd]) {
            slot_node.push_back(c);
            slot_desc.insert(slot_desc.end(), desc + 32 * (size_t)(c - 1), desc + 32 * (size_t)c);
        }
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) {
        set_error("pl_voc_create: no usable CUDA device %d (there is no CPU fallback)", device);
        return PL_ERR_CUDA;
    }
    PL_CUDA_TRY(cudaSetDevice(device));
    pl_voc* v = new pl_voc;
    v->device = device; v->k = k; v->L = L; v->scoring = scoring; v->weighting = weighting; v->n_nodes = nn; v->n_words = n_words;
    const void* src[6] = {cbeg.data(), ccnt.data(), slot_node.data(), slot_desc.data(), word.data(), wt.data()};
    const size_t bytes[6] = {(size_t)nn * 4, (size_t)nn * 4, slot_node.size() * 4, slot_desc.size(), (size_t)nn * 4, (size_t)nn * 8};
    cudaError_t e = cudaStreamCreateWithFlags(&v->stream, cudaStreamNonBlocking);
    for (int i = 0; i < 6 && e == cudaSuccess; i++) {
        e = cudaMalloc(&v->d_mem[i], std::max(bytes[i], (size_t)32));
        if (e == cudaSuccess && bytes[i]) e = cudaMemcpy(v->d_mem[i], src[i], bytes[i], cudaMemcpyHostToDevice);
    }
    if (e != cudaSuccess) {
        set_error("pl_voc_create: %s", cudaGetErrorString(e));
        pl_voc_destroy(v);
        return PL_ERR_CUDA;
    }
    v->dev = VocDev{(const int*)v->d_mem[0], (const int*)v->d_mem[1], (const unsigned*)v->d_mem[2], (const uint4*)v->d_mem[3],
                    (const unsigned*)v->d_mem[4], (const double*)v->d_mem[5], L};
    *out = v;
    return PL_OK;
}

PL_API int pl_voc_load_text(pl_voc** out, int device, const char* filename) {
    PL_CHECK_ARG(out && filename);
    FILE* f = fopen(filename, "r");
    if (!f) { set_error("pl_voc_load_text: cannot open %s", filename); return PL_ERR_ARG; }
    std::vector<char> line(1 << 16);
    int k = -1, L = -1, n1 = -1, n2 = -1;
    if (!fgets(line.data(), (int)line.size(), f) || sscanf(line.data(), "%d %d %d %d", &k, &L, &n1, &n2) != 4 || k < 0 || k > 20 || L < 1 || L > 10 ||
        n1 < 0 || n1 > 5 || n2 < 0 || n2 > 3) {  // :1359-1363
        fclose(f);
        set_error("Vocabulary loading failure: This is not a correct text file!");
        return PL_ERR_ARG;
    }
    std::vector<int> parent;
    std::vector<uint8_t> leaf, desc;
    std::vector<double> weight;
    while (fgets(line.data(), (int)line.size(), f)) {
        char* p = line.data();
        if (p[strspn(p, " \t\r\n")] == 0) continue;
        char* e;
        const long pid = strtol(p, &e, 10);
        bool ok = e != p;
        p = e;
        const long nIsLeaf = strtol(p, &e, 10);
        ok = ok && e != p;
        p = e;
        parent.push_back((int)pid);
        leaf.push_back(nIsLeaf > 0);
        for (int i = 0; i < 32; i++) {
            desc.push_back((uint8_t)strtol(p, &e, 10));
            ok = ok && e != p;
            p = e;
        }
        weight.push_back(strtod(p, &e));
        ok = ok && e != p;
        if (!ok) {
            fclose(f);
            set_error("pl_voc_load_text: malformed node line %zu", parent.size() + 1);
            return PL_ERR_ARG;
        }
    }
    fclose(f);
    return pl_voc_create(out, device, k, L, n1, n2, (int)parent.size(), parent.data(), leaf.data(), desc.data(), weight.data());
}

PL_API void pl_voc_destroy(pl_voc* v) {
    if (!v) return;
    cudaSetDevice(v->device);
    for (void* p : v->d_mem) if (p) cudaFree(p);
    if (v->d_scratch) cudaFree(v->d_scratch);
    v->in.release();
    if (v->stream) cudaStreamDestroy(v->stream);
    delete v;
}

PL_API int pl_voc_info(const pl_voc* v, int* k, int* L, int* n_nodes, int* n_words) {
    PL_CHECK_ARG(v);
    if (k) *k = v->k;
    if (L) *L = v->L;
    if (n_nodes) *n_nodes = v->n_nodes;
    if (n_words) *n_words = v->n_words;
    return PL_OK;
}

PL_API int pl_voc_transform_batch(pl_voc* v, int n_frames, const int* off, const uint8_t* desc, int levelsup, int* n_words, unsigned int* word_id,
                                  double* word_value, int* n_fv_nodes, unsigned int* node_id, int* node_off, unsigned int* feat_idx) {
    PL_CHECK_ARG(v && n_frames >= 0);
    if (n_frames == 0) return PL_OK;
    PL_CHECK_ARG(off && off[0] == 0 && n_words && n_fv_nodes && node_off);
    int max_n = 0;
    for (int f = 0; f < n_frames; f++) {
        PL_CHECK_ARG(off[f + 1] >= off[f] && off[f + 1] - off[f] <= kVocMaxFeat);
        max_n = std::max(max_n, off[f + 1] - off[f]);
    }
    const size_t total = (size_t)off[n_frames];
    PL_CHECK_ARG(total == 0 || (desc && word_id && word_value && node_id && feat_idx));
    PL_CUDA_TRY(cudaSetDevice(v->device));
    cudaStream_t st = v->stream;
    if (v->n_nodes <= 1) {  // empty() vocabulary: both vectors stay empty (TemplatedVocabulary.h:1134-1137)
        for (int f = 0; f < n_frames; f++) { n_words[f] = 0; n_fv_nodes[f] = 0; node_off[off[f] + f] = 0; }
        return PL_OK;
    }
    const size_t tp = std::max(total, (size_t)1);
    // packed buffer: inputs (off, desc) then outputs; per-feature scratch lives in d_scratch
    int rc = v->in.reserve(padb((size_t)(n_frames + 1) * 4) + padb(tp * 32) + padb((size_t)n_frames * 4) * 2 + padb(tp * 4) * 3 + padb(tp * 8) +
                           padb((tp + n_frames) * 4));
    if (rc != PL_OK) return rc;
    const size_t need = padb(tp * 4) * 2 + padb(tp * 8);
    if (v->scratch_cap < need) {
        if (v->d_scratch) cudaFree(v->d_scratch);
        v->d_scratch = nullptr;
        v->scratch_cap = 0;
        PL_CUDA_TRY(cudaMalloc((void**)&v->d_scratch, need + need / 4));
        v->scratch_cap = need + need / 4;
    }
    unsigned* f_word = (unsigned*)v->d_scratch;
    unsigned* f_nid = (unsigned*)(v->d_scratch + padb(tp * 4));
    double* f_weight = (double*)(v->d_scratch + padb(tp * 4) * 2);
    const int* d_off = v->in.put(off, (size_t)n_frames + 1);
    const uint4* d_desc = (const uint4*)v->in.put(desc, total * 32);
    int *h_nw, *h_nn, *h_noff;
    unsigned *h_wid, *h_nid, *h_fi;
    double* h_wv;
    int* d_nw = v->in.out<int>((size_t)n_frames, &h_nw);
    int* d_nn = v->in.out<int>((size_t)n_frames, &h_nn);
    unsigned* d_wid = v->in.out<unsigned>(tp, &h_wid);
    unsigned* d_nid = v->in.out<unsigned>(tp, &h_nid);
    unsigned* d_fi = v->in.out<unsigned>(tp, &h_fi);
    double* d_wv = v->in.out<double>(tp, &h_wv);
    int* d_noff = v->in.out<int>(tp + n_frames, &h_noff);
    const size_t out_begin = (size_t)((uint8_t*)d_nw - v->in.d);
    if ((rc = v->in.upload(st)) != PL_OK) return rc;
    if (total) k_voc_descend<<<(unsigned)((total * 32 + 255) / 256), 256, 0, st>>>(v->dev, d_desc, (int)total, v->L - levelsup, f_word, f_weight, f_nid);
    int n2 = 1;
    while (n2 < max_n) n2 <<= 1;
    const size_t sm = (size_t)n2 * 12;
    if (sm > 48 * 1024) PL_CUDA_TRY(cudaFuncSetAttribute(k_voc_assemble, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    k_voc_assemble<<<n_frames, kVocThreads, sm, st>>>(d_off, f_word, f_weight, f_nid, v->scoring, v->weighting, d_nw, d_wid, d_wv, d_nn, d_nid, d_noff, d_fi);
    PL_CUDA_TRY(cudaGetLastError());
    PL_CUDA_TRY(cudaMemcpyAsync(v->in.h + out_begin, v->in.d + out_begin, v->in.cur - out_begin, cudaMemcpyDeviceToHost, st));
    PL_CUDA_TRY(pl::stream_sync(st));
    memcpy(n_words, h_nw, (size_t)n_frames * 4);
    memcpy(n_fv_nodes, h_nn, (size_t)n_frames * 4);
    for (int f = 0; f < n_frames; f++) {
        const size_t o = (size_t)off[f];
        memcpy(word_id + o, h_wid + o, (size_t)h_nw[f] * 4);
        memcpy(word_value + o, h_wv + o, (size_t)h_nw[f] * 8);
        memcpy(node_id + o, h_nid + o, (size_t)h_nn[f] * 4);
        memcpy(node_off + o + f, h_noff + o + f, (size_t)(h_nn[f] + 1) * 4);
        const int kept = h_noff[o + f + h_nn[f]];
        memcpy(feat_idx + o, h_fi + o, (size_t)kept * 4);
    }
    return PL_OK;
}

}  // extern "C"
