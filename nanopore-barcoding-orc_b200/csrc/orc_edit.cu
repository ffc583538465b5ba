// orc_edit.cu -- C ABI of the pairwise edit-distance kernel (orc_edit.cuh): the drop-in for
// amplicon_sorter's distance() / distance_finetune() (amplicon_sorter.py:225-235, :838-849).
#include "../../include/orcdemux.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "orc_edit.cuh"

namespace {

struct DevBuf {
    void *p = nullptr;
    ~DevBuf() { if (p) cudaFree(p); }
    cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes ? bytes : 1); }
};

int fail(char *err, size_t err_len, int code, const std::string &what)
{
    if (err && err_len) snprintf(err, err_len, "%s", what.c_str());
    return code;
}

}  // namespace

#define ECK(call)                                                                                   \
    do {                                                                                            \
        cudaError_t e_ = (call);                                                                    \
        if (e_ != cudaSuccess)                                                                      \
            return fail(err, err_len, ORC_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
    } while (0)

extern "C" int orc_edit_distances(int device, const uint8_t *seqs, const uint64_t *offsets, const uint32_t *lengths,
                                  uint32_t n_seqs, const uint32_t *pair_a, const uint32_t *pair_b, uint64_t n_pairs,
                                  int mode, uint32_t *dist, float *kernel_ms, char *err, size_t err_len)
{
    using namespace orc;
    if ((n_seqs && (!seqs || !offsets || !lengths)) || (n_pairs && (!pair_a || !pair_b || !dist)) ||
        (mode != ORC_EDIT_NW && mode != ORC_EDIT_HW) || n_pairs > 0xFFFFFFFFull)
        return fail(err, err_len, ORC_EINVAL, "orc_edit_distances: bad argument");
    if (kernel_ms) *kernel_ms = 0.0f;
    if (n_pairs == 0) return ORC_OK;
    // the batch's alphabet: every distinct byte is its own symbol (edlib compares bytes)
    uint64_t n_bytes = 0;
    for (uint32_t i = 0; i < n_seqs; i++)
        if (offsets[i] + lengths[i] > n_bytes) n_bytes = offsets[i] + lengths[i];
    int map[256];
    for (int i = 0; i < 256; i++) map[i] = -1;
    int n_sym = 0;
    std::vector<uint8_t> sym(n_bytes);
    for (uint32_t i = 0; i < n_seqs; i++)
        for (uint64_t p = offsets[i]; p < offsets[i] + lengths[i]; p++) {
            const uint8_t c = seqs[p];
            if (map[c] < 0) {
                if (n_sym == EDIT_SYMS)
                    return fail(err, err_len, ORC_EINVAL, "unsupported: more than 8 distinct characters in the batch");
                map[c] = n_sym++;
            }
            sym[p] = (uint8_t)map[c];
        }
    // pairs by the lanes a query needs (8, 16 or 32, one 64-row block per lane) and, beyond 2048
    // rows, by the blocks a lane has to hold (2 or 4); longest target first inside a class
    // classes by the 64-row blocks of the query: G lanes per pair with one block each up to 32 blocks,
    // then 2 and 4 blocks per lane
    constexpr int N_CLS = 9;
    static const int cls_g[N_CLS] = {5, 6, 7, 8, 10, 16, 32, 32, 32};
    static const int cls_wb[N_CLS] = {1, 1, 1, 1, 1, 1, 1, 2, 4};
    constexpr uint32_t KEY_CAP = 1u << 16;          // targets beyond 65535 share the first bucket
    auto cls_of = [](uint32_t m) {
        const uint32_t nb = (m + 63u) / 64u;
        return nb <= 5u ? 0 : nb == 6u ? 1 : nb == 7u ? 2 : nb == 8u ? 3 : nb <= 10u ? 4 : nb <= 16u ? 5 : nb <= 32u ? 6
                                                                                             : nb <= 64u ? 7 : 8;
    };
    // counting sort by (class, target length descending), stable
    std::vector<uint64_t> count((size_t)N_CLS * (KEY_CAP + 1u) + 1u, 0);
    for (uint64_t k = 0; k < n_pairs; k++) {
        if (pair_a[k] >= n_seqs || pair_b[k] >= n_seqs)
            return fail(err, err_len, ORC_EINVAL, "orc_edit_distances: pair index out of range");
        const uint32_t la = lengths[pair_a[k]], lb = lengths[pair_b[k]];
        const uint32_t m = la < lb ? la : lb, n = la < lb ? lb : la;
        if (m > 32u * 64u * EDIT_MAX_WB)
            return fail(err, err_len, ORC_EINVAL, "unsupported: sequences longer than 8192 on both sides of a pair");
        count[(size_t)cls_of(m) * (KEY_CAP + 1u) + (KEY_CAP - (n < KEY_CAP ? n : KEY_CAP)) + 1u]++;
    }
    for (size_t i = 1; i < count.size(); i++) count[i] += count[i - 1];
    std::vector<uint32_t> order(n_pairs);
    for (uint64_t k = 0; k < n_pairs; k++) {
        const uint32_t la = lengths[pair_a[k]], lb = lengths[pair_b[k]];
        const uint32_t m = la < lb ? la : lb, n = la < lb ? lb : la;
        order[count[(size_t)cls_of(m) * (KEY_CAP + 1u) + (KEY_CAP - (n < KEY_CAP ? n : KEY_CAP))]++] = (uint32_t)k;
    }
    uint64_t cls_begin[N_CLS + 1];
    cls_begin[0] = 0;
    for (int c = 0; c < N_CLS; c++) cls_begin[c + 1] = count[(size_t)c * (KEY_CAP + 1u) + KEY_CAP];
    if (cudaSetDevice(device) != cudaSuccess)
        return fail(err, err_len, ORC_ECUDA, "no usable CUDA device (there is no CPU fallback)");
    cudaDeviceProp prop;
    ECK(cudaGetDeviceProperties(&prop, device));
    DevBuf d_sym, d_off, d_len, d_pa, d_pb, d_out, d_todo;
    ECK(d_sym.alloc(n_bytes + 64));
    ECK(d_off.alloc(8ull * n_seqs));
    ECK(d_len.alloc(4ull * n_seqs));
    ECK(d_pa.alloc(4ull * n_pairs));
    ECK(d_pb.alloc(4ull * n_pairs));
    ECK(d_out.alloc(4ull * n_pairs));
    ECK(d_todo.alloc(4ull * n_pairs));
    ECK(cudaMemcpy(d_sym.p, sym.data(), n_bytes, cudaMemcpyHostToDevice));
    ECK(cudaMemcpy(d_off.p, offsets, 8ull * n_seqs, cudaMemcpyHostToDevice));
    ECK(cudaMemcpy(d_len.p, lengths, 4ull * n_seqs, cudaMemcpyHostToDevice));
    ECK(cudaMemcpy(d_pa.p, pair_a, 4ull * n_pairs, cudaMemcpyHostToDevice));
    ECK(cudaMemcpy(d_pb.p, pair_b, 4ull * n_pairs, cudaMemcpyHostToDevice));
    ECK(cudaMemcpy(d_todo.p, order.data(), 4ull * n_pairs, cudaMemcpyHostToDevice));
    struct Ev {                      // destroyed on every return path, like DevBuf
        cudaEvent_t e = nullptr;
        ~Ev() { if (e) cudaEventDestroy(e); }
    } ev0, ev1;
    ECK(cudaEventCreate(&ev0.e));
    ECK(cudaEventCreate(&ev1.e));
    cudaEvent_t e0 = ev0.e, e1 = ev1.e;
    ECK(cudaEventRecord(e0, 0));
    for (int cls = 0; cls < N_CLS; cls++) {
        const uint32_t nt = (uint32_t)(cls_begin[cls + 1] - cls_begin[cls]);
        if (!nt) continue;
        uint32_t *d_list = (uint32_t *)d_todo.p + cls_begin[cls];
        const int wb = cls_wb[cls];
        const uint32_t per_warp = 32u / (uint32_t)cls_g[cls];
        const size_t smem = 4u * EDIT_SYMS * wb * 32u * sizeof(uint64_t);
        const uint32_t want = (nt + 4u * per_warp - 1u) / (4u * per_warp), cap = (uint32_t)prop.multiProcessorCount * 8u;
        const uint32_t blocks = want < cap ? want : cap;
        const uint8_t *ds = (const uint8_t *)d_sym.p;
        const uint64_t *dof = (const uint64_t *)d_off.p;
        const uint32_t *dl = (const uint32_t *)d_len.p, *da = (const uint32_t *)d_pa.p, *db = (const uint32_t *)d_pb.p;
        uint32_t *dout = (uint32_t *)d_out.p;
#define ORC_EDIT_LAUNCH(WB, G) edit_kernel<WB, G><<<blocks, 128, smem, 0>>>(ds, dof, dl, da, db, d_list, nt, mode, dout)
        switch (cls) {
        case 0: ORC_EDIT_LAUNCH(1, 5); break;
        case 1: ORC_EDIT_LAUNCH(1, 6); break;
        case 2: ORC_EDIT_LAUNCH(1, 7); break;
        case 3: ORC_EDIT_LAUNCH(1, 8); break;
        case 4: ORC_EDIT_LAUNCH(1, 10); break;
        case 5: ORC_EDIT_LAUNCH(1, 16); break;
        case 6: ORC_EDIT_LAUNCH(1, 32); break;
        case 7: ORC_EDIT_LAUNCH(2, 32); break;
        default: ORC_EDIT_LAUNCH(4, 32); break;
        }
#undef ORC_EDIT_LAUNCH
        ECK(cudaGetLastError());
    }
    ECK(cudaEventRecord(e1, 0));
    ECK(cudaEventSynchronize(e1));
    float ms = 0.0f;
    ECK(cudaEventElapsedTime(&ms, e0, e1));
    if (kernel_ms) *kernel_ms = ms;
    ECK(cudaMemcpy(dist, d_out.p, 4ull * n_pairs, cudaMemcpyDeviceToHost));
    return ORC_OK;
}
