// k_group.cu -- groups the signatures of a verification batch by public key, on the device (no host round trip).
//
// hash 57 key bytes -> radix-sort (hash, index) pairs -> mark group heads by comparing the BYTES of neighbours
// (a hash collision can only split a group, never merge two keys) -> scans -> work lists (verify_plan.cuh).
// The sort and the scans are CUB primitives: plumbing around the path, not the path.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include "verify_plan.cuh"

namespace {
constexpr int GB = 256;
constexpr uint32_t MIN_SHARE = 2; /* a key table pays for itself from the second signature on */
inline size_t al(size_t b) { return (b + 255) & ~(size_t)255; }
inline unsigned blocks(size_t n) { return (unsigned)((n + GB - 1) / GB); }

__device__ __forceinline__ uint64_t mix64(uint64_t x) { x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33; return x; }

// `csize`: signatures per chunk when the chunk id is part of the grouping key (rlc.cuh: one equation per chunk); 0 = no chunks
__global__ void k_hash(uint64_t *h, uint32_t *idx, const uint8_t *pk, uint32_t n, uint32_t csize = 0) {
    const uint32_t i = blockIdx.x * GB + threadIdx.x;
    if (i >= n) return;
    const uint8_t *p = pk + 57 * (size_t)i;
    uint64_t acc = 0x9e3779b97f4a7c15ull;
    for (int w = 0; w < 8; w++) {
        uint64_t x = 0;
        for (int b = 0; b < 8; b++) { const int k = 8 * w + b; if (k < 57) x |= (uint64_t)p[k] << (8 * b); }
        acc = mix64(acc ^ x) + 0x9e3779b97f4a7c15ull * (uint64_t)(w + 1);
    }
    if (csize) acc = mix64(acc ^ (uint64_t)(i / csize));
    h[i] = acc;
    idx[i] = i;
}
__global__ void k_heads(uint32_t *head, const uint32_t *is, const uint8_t *pk, uint32_t n, uint32_t csize = 0) {
    const uint32_t j = blockIdx.x * GB + threadIdx.x;
    if (j >= n) return;
    uint32_t differ = 1;
    if (j) {
        const uint8_t *a = pk + 57 * (size_t)is[j], *b = pk + 57 * (size_t)is[j - 1];
        uint32_t d = 0;
        for (int k = 0; k < 57; k++) d |= (uint32_t)(a[k] ^ b[k]);
        if (csize) d |= (is[j] / csize) ^ (is[j - 1] / csize);
        differ = d ? 1u : 0u;
    }
    head[j] = differ;
}
__global__ void k_gstart(uint32_t *gstart, uint32_t *ngroups, const uint32_t *head, const uint32_t *gid, uint32_t n) {
    const uint32_t j = blockIdx.x * GB + threadIdx.x;
    if (j >= n) return;
    if (head[j]) gstart[gid[j] - 1] = j;
    if (j == n - 1) { gstart[gid[j]] = n; *ngroups = gid[j]; }
}
__global__ void k_gflag(uint32_t *sflag, const uint32_t *gstart, const uint32_t *ngroups, uint32_t n) {
    const uint32_t g = blockIdx.x * GB + threadIdx.x;
    if (g >= n) return;
    sflag[g] = (g < *ngroups && gstart[g + 1] - gstart[g] >= MIN_SHARE) ? 1u : 0u;
}
__global__ void k_admit(uint32_t *adm, uint32_t *tab_rep, const uint32_t *head, const uint32_t *gid, const uint32_t *sflag, const uint32_t *tslot,
                        const uint32_t *is, uint32_t cap, uint32_t n) {
    const uint32_t j = blockIdx.x * GB + threadIdx.x;
    if (j >= n) return;
    const uint32_t g = gid[j] - 1;
    const uint32_t a = (sflag[g] && tslot[g] < cap) ? 1u : 0u;
    adm[j] = a;
    if (a && head[j]) tab_rep[tslot[g]] = is[j];
}
__global__ void k_lists(uint32_t *shared_sig, uint32_t *shared_tab, uint32_t *unique_sig, uint32_t *counts, const uint32_t *adm, const uint32_t *pos,
                        const uint32_t *gid, const uint32_t *sflag, const uint32_t *tslot, const uint32_t *is, uint32_t cap, uint32_t n) {
    const uint32_t j = blockIdx.x * GB + threadIdx.x;
    if (j >= n) return;
    const uint32_t g = gid[j] - 1;
    if (adm[j]) { shared_sig[pos[j]] = is[j]; shared_tab[pos[j]] = tslot[g]; }
    else unique_sig[j - pos[j]] = is[j];
    if (j == n - 1) {
        const uint32_t ns = pos[j] + adm[j], nt = tslot[g] + sflag[g];
        counts[0] = ns; counts[1] = n - ns; counts[2] = nt < cap ? nt : cap; counts[3] = 0; counts[4] = 0; counts[5] = 0;
    }
}
struct Layout {
    uint64_t *h, *hs; uint32_t *idx, *is, *head, *gid, *gstart, *sflag, *tslot, *adm, *pos, *shared_sig, *shared_tab, *unique_sig, *tab_rep, *counts, *ngroups;
    void *cub_tmp; size_t cub_bytes, total;
};
Layout lay(void *scratch, size_t n, size_t cap) {
    Layout L;
    char *p = (char *)scratch;
    auto take = [&](size_t bytes) { char *r = p; p += al(bytes); return (void *)r; };
    L.h = (uint64_t *)take(8 * n); L.hs = (uint64_t *)take(8 * n);
    L.idx = (uint32_t *)take(4 * n); L.is = (uint32_t *)take(4 * n); L.head = (uint32_t *)take(4 * n); L.gid = (uint32_t *)take(4 * n);
    L.gstart = (uint32_t *)take(4 * (n + 1)); L.sflag = (uint32_t *)take(4 * n); L.tslot = (uint32_t *)take(4 * n);
    L.adm = (uint32_t *)take(4 * n); L.pos = (uint32_t *)take(4 * n);
    L.shared_sig = (uint32_t *)take(4 * n); L.shared_tab = (uint32_t *)take(4 * n); L.unique_sig = (uint32_t *)take(4 * n);
    L.tab_rep = (uint32_t *)take(4 * (cap + 1)); L.counts = (uint32_t *)take(32); L.ngroups = (uint32_t *)take(4);
    size_t b1 = 0, b2 = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, b1, (const uint64_t *)nullptr, (uint64_t *)nullptr, (const uint32_t *)nullptr, (uint32_t *)nullptr, (int)n);
    cub::DeviceScan::InclusiveSum(nullptr, b2, (const uint32_t *)nullptr, (uint32_t *)nullptr, (int)n);
    L.cub_bytes = (b1 > b2 ? b1 : b2) + 256;
    L.cub_tmp = take(L.cub_bytes);
    L.total = (size_t)(p - (char *)scratch);
    return L;
}
}  // namespace

size_t group_scratch_bytes(size_t n, size_t cap) { return n ? lay(nullptr, n, cap).total : 256; }

cudaError_t group_keys(const uint8_t *pk, size_t n, uint32_t cap, void *scratch, verify_plan *plan, cudaStream_t s, uint64_t *launches) {
    const Layout L = lay(scratch, n, cap);
    const uint32_t m = (uint32_t)n;
    size_t tmp = L.cub_bytes;
    cudaError_t e;
    k_hash<<<blocks(n), GB, 0, s>>>(L.h, L.idx, pk, m);
    if ((e = cub::DeviceRadixSort::SortPairs(L.cub_tmp, tmp, L.h, L.hs, L.idx, L.is, (int)n, 0, 64, s)) != cudaSuccess) return e;
    k_heads<<<blocks(n), GB, 0, s>>>(L.head, L.is, pk, m);
    tmp = L.cub_bytes;
    if ((e = cub::DeviceScan::InclusiveSum(L.cub_tmp, tmp, L.head, L.gid, (int)n, s)) != cudaSuccess) return e;
    k_gstart<<<blocks(n), GB, 0, s>>>(L.gstart, L.ngroups, L.head, L.gid, m);
    k_gflag<<<blocks(n), GB, 0, s>>>(L.sflag, L.gstart, L.ngroups, m);
    tmp = L.cub_bytes;
    if ((e = cub::DeviceScan::ExclusiveSum(L.cub_tmp, tmp, L.sflag, L.tslot, (int)n, s)) != cudaSuccess) return e;
    k_admit<<<blocks(n), GB, 0, s>>>(L.adm, L.tab_rep, L.head, L.gid, L.sflag, L.tslot, L.is, cap, m);
    tmp = L.cub_bytes;
    if ((e = cub::DeviceScan::ExclusiveSum(L.cub_tmp, tmp, L.adm, L.pos, (int)n, s)) != cudaSuccess) return e;
    k_lists<<<blocks(n), GB, 0, s>>>(L.shared_sig, L.shared_tab, L.unique_sig, L.counts, L.adm, L.pos, L.gid, L.sflag, L.tslot, L.is, cap, m);
    if (launches) *launches += 6; /* our kernels; the CUB passes are library launches */
    plan->shared_sig = L.shared_sig; plan->shared_tab = L.shared_tab; plan->unique_sig = L.unique_sig; plan->tab_rep = L.tab_rep; plan->counts = L.counts;
    return cudaGetLastError();
}

// ---- every group of the batch, no table budget: the key view of the random-linear-combination path (rlc.cuh) --------
// order[j] = signature at sorted position j, gid[j] = 1-based group of that position, gstart[g] = first position of
// group g (gstart[ngroups] = n), *ngroups = number of distinct keys.  Hash collisions can only split a group.
size_t group_all_scratch_bytes(size_t n) {
    if (!n) return 256;
    size_t b1 = 0, b2 = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, b1, (const uint64_t *)nullptr, (uint64_t *)nullptr, (const uint32_t *)nullptr, (uint32_t *)nullptr, (int)n);
    cub::DeviceScan::InclusiveSum(nullptr, b2, (const uint32_t *)nullptr, (uint32_t *)nullptr, (int)n);
    return 2 * al(8 * n) + 4 * al(4 * n) + al(4 * (n + 1)) + al(4) + al((b1 > b2 ? b1 : b2) + 256);
}
__global__ void k_kchunk(uint32_t *kchunk, const uint32_t *is, const uint32_t *gstart, uint32_t m, uint32_t nch, uint32_t csize) {
    const uint32_t l = blockIdx.x * GB + threadIdx.x;
    if (l < m) kchunk[l] = csize ? is[gstart[l]] / csize : 0u;   /* chunk of key group l */
    else if (l < m + nch) kchunk[l] = l - m;                      /* the B of chunk l - m */
}
cudaError_t key_chunks(uint32_t *kchunk, const key_groups *g, uint32_t m, uint32_t nch, uint32_t csize, cudaStream_t s) {
    k_kchunk<<<blocks((size_t)m + nch), GB, 0, s>>>(kchunk, g->order, g->gstart, m, nch, csize);
    return cudaGetLastError();
}
cudaError_t group_keys_all(const uint8_t *pk, size_t n, void *scratch, key_groups *out, cudaStream_t s, uint64_t *launches, uint32_t csize) {
    char *p = (char *)scratch;
    auto take = [&](size_t bytes) { char *r = p; p += al(bytes); return (void *)r; };
    uint64_t *h = (uint64_t *)take(8 * n), *hs = (uint64_t *)take(8 * n);
    uint32_t *idx = (uint32_t *)take(4 * n), *is = (uint32_t *)take(4 * n), *head = (uint32_t *)take(4 * n), *gid = (uint32_t *)take(4 * n);
    uint32_t *gstart = (uint32_t *)take(4 * (n + 1)), *ngroups = (uint32_t *)take(4);
    size_t b1 = 0, b2 = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, b1, (const uint64_t *)nullptr, (uint64_t *)nullptr, (const uint32_t *)nullptr, (uint32_t *)nullptr, (int)n);
    cub::DeviceScan::InclusiveSum(nullptr, b2, (const uint32_t *)nullptr, (uint32_t *)nullptr, (int)n);
    const size_t cub_bytes = (b1 > b2 ? b1 : b2) + 256;
    void *cub_tmp = take(cub_bytes);
    const uint32_t m = (uint32_t)n;
    size_t tmp = cub_bytes;
    cudaError_t e;
    k_hash<<<blocks(n), GB, 0, s>>>(h, idx, pk, m, csize);
    if ((e = cub::DeviceRadixSort::SortPairs(cub_tmp, tmp, h, hs, idx, is, (int)n, 0, 64, s)) != cudaSuccess) return e;
    k_heads<<<blocks(n), GB, 0, s>>>(head, is, pk, m, csize);
    tmp = cub_bytes;
    if ((e = cub::DeviceScan::InclusiveSum(cub_tmp, tmp, head, gid, (int)n, s)) != cudaSuccess) return e;
    k_gstart<<<blocks(n), GB, 0, s>>>(gstart, ngroups, head, gid, m);
    if (launches) *launches += 3;
    out->order = is; out->gid = gid; out->gstart = gstart; out->ngroups = ngroups;
    return cudaGetLastError();
}

// Radix sort of the (window | digit) -> point pairs of the bucket method (rlc.cuh); a CUB primitive, plumbing.
size_t pair_sort_scratch_bytes(size_t npairs) {
    size_t b = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, b, (const uint32_t *)nullptr, (uint32_t *)nullptr, (const uint32_t *)nullptr, (uint32_t *)nullptr, (int)npairs);
    return b + 256;
}
cudaError_t pair_sort(void *tmp, size_t tmp_bytes, const uint32_t *keys_in, uint32_t *keys_out, const uint32_t *vals_in, uint32_t *vals_out, size_t npairs,
                      int key_bits, cudaStream_t s) {
    return cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, keys_in, keys_out, vals_in, vals_out, (int)npairs, 0, key_bits, s);
}
