// sort.cuh -- stable LSD radix sort of (u64 key, u32 value) pairs, 8 bits per pass.
// Used only by the reference-literal traversal order (PF_ORDER_SORTED), which restates
// sort!(putative, alg=QuickSort, by=weight) (src/fearnhead.jl:148) with the declared
// tie-break "putative index" (a stable sort of index-ordered input).  Non-negative doubles
// order like their bit patterns, so the keys are the weights themselves.
#pragma once
#include "common.cuh"

#define RS_TILE 4096
#define RS_THREADS 256
#define RS_WARPS (RS_THREADS / 32)
#define RS_PER_WARP (RS_TILE / RS_WARPS)

__global__ void k_rs_init(const u64 *__restrict__ in, u64 *__restrict__ keys, unsigned int *__restrict__ vals, long long M)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < M) { keys[i] = in[i]; vals[i] = (unsigned int)i; }
}

// per-warp digit counts of one tile into s_cnt[w][256]
__device__ __forceinline__ void rs_count_tile(const u64 *__restrict__ keys, long long M, int shift, unsigned int (*s_cnt)[256])
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int t = threadIdx.x; t < RS_WARPS * 256; t += RS_THREADS) (&s_cnt[0][0])[t] = 0;
    __syncthreads();
    const long long base = (long long)blockIdx.x * RS_TILE + (long long)wid * RS_PER_WARP;
    for (int c = 0; c < RS_PER_WARP; c += 32) {
        long long i = base + c + lane;
        bool have = i < M;
        unsigned int dgt = have ? (unsigned int)((keys[i] >> shift) & 0xFF) : 0x100u + lane;   // distinct dummies
        unsigned int peers = __match_any_sync(0xffffffffu, dgt);
        if (have && (__ffs(peers) - 1) == lane) s_cnt[wid][dgt] += __popc(peers);
        __syncwarp();
    }
    __syncthreads();
}

__global__ void __launch_bounds__(RS_THREADS) k_rs_hist(const u64 *__restrict__ keys, long long M, int shift, int nblk,
                                                        unsigned int *__restrict__ hist)
{
    __shared__ unsigned int s_cnt[RS_WARPS][256];
    rs_count_tile(keys, M, shift, s_cnt);
    for (int dgt = threadIdx.x; dgt < 256; dgt += RS_THREADS) {
        unsigned int s = 0;
        for (int w = 0; w < RS_WARPS; w++) s += s_cnt[w][dgt];
        hist[(long long)dgt * nblk + blockIdx.x] = s;
    }
}

// exclusive scan of hist[256 * nblk] in place (digit-major): block sums, one block over the block sums, apply
#define RS_SCAN_TILE 4096          // entries per block (1024 threads x 4)
__device__ __forceinline__ unsigned int rs_block_excl_scan(unsigned int v, unsigned int *s_w /* >= 33 */, unsigned int *total)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    unsigned int inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { unsigned int o = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += o; }
    if (lane == 31) s_w[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        unsigned int w = s_w[lane], wi = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { unsigned int o = __shfl_up_sync(0xffffffffu, wi, d); if (lane >= d) wi += o; }
        s_w[lane] = wi - w;
        if (lane == 31) s_w[32] = wi;
    }
    __syncthreads();
    const unsigned int r = s_w[wid] + inc - v;
    *total = s_w[32];
    __syncthreads();
    return r;
}
__global__ void __launch_bounds__(1024) k_rs_scan_sums(const unsigned int *__restrict__ hist, long long n, unsigned int *__restrict__ part)
{
    __shared__ unsigned int s_w[33];
    const long long b = (long long)blockIdx.x * RS_SCAN_TILE;
    unsigned int loc = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) { const long long i = b + k * 1024 + threadIdx.x; if (i < n) loc += hist[i]; }
    unsigned int tot;
    rs_block_excl_scan(loc, s_w, &tot);
    if (threadIdx.x == 0) part[blockIdx.x] = tot;
}
__global__ void __launch_bounds__(1024) k_rs_scan_top(unsigned int *part, long long nb)
{
    __shared__ unsigned int s_w[33];
    const long long per = (nb + 1023) / 1024;
    const long long b = (long long)threadIdx.x * per, e = b + per < nb ? b + per : nb;
    unsigned int loc = 0;
    for (long long i = b; i < e; i++) loc += part[i];
    unsigned int tot;
    unsigned int run = rs_block_excl_scan(loc, s_w, &tot);
    for (long long i = b; i < e; i++) { const unsigned int x = part[i]; part[i] = run; run += x; }
}
__global__ void __launch_bounds__(1024) k_rs_scan_apply(unsigned int *__restrict__ hist, long long n, const unsigned int *__restrict__ part)
{
    __shared__ unsigned int s_w[33];
    const long long b = (long long)blockIdx.x * RS_SCAN_TILE + (long long)threadIdx.x * 4;
    unsigned int v[4], loc = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) { v[k] = b + k < n ? hist[b + k] : 0u; loc += v[k]; }
    unsigned int tot;
    unsigned int run = part[blockIdx.x] + rs_block_excl_scan(loc, s_w, &tot);
#pragma unroll
    for (int k = 0; k < 4; k++) { if (b + k < n) hist[b + k] = run; run += v[k]; }
}

__global__ void __launch_bounds__(RS_THREADS) k_rs_scatter(const u64 *__restrict__ keys, const unsigned int *__restrict__ vals,
                                                           u64 *__restrict__ keys_out, unsigned int *__restrict__ vals_out,
                                                           long long M, int shift, int nblk, const unsigned int *__restrict__ hist)
{
    __shared__ unsigned int s_cnt[RS_WARPS][256];
    rs_count_tile(keys, M, shift, s_cnt);
    // s_cnt[w][d] -> global offset of warp w's first element with digit d
    for (int dgt = threadIdx.x; dgt < 256; dgt += RS_THREADS) {
        unsigned int run = hist[(long long)dgt * nblk + blockIdx.x];
        for (int w = 0; w < RS_WARPS; w++) { unsigned int x = s_cnt[w][dgt]; s_cnt[w][dgt] = run; run += x; }
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const long long base = (long long)blockIdx.x * RS_TILE + (long long)wid * RS_PER_WARP;
    for (int c = 0; c < RS_PER_WARP; c += 32) {
        long long i = base + c + lane;
        bool have = i < M;
        u64 key = have ? keys[i] : 0;
        unsigned int dgt = have ? (unsigned int)((key >> shift) & 0xFF) : 0x100u + lane;
        unsigned int peers = __match_any_sync(0xffffffffu, dgt);
        unsigned int rank = __popc(peers & ((1u << lane) - 1));
        if (have) {
            unsigned int pos = s_cnt[wid][dgt] + rank;
            keys_out[pos] = key;
            vals_out[pos] = vals[i];
        }
        __syncwarp();
        if (have && (__ffs(peers) - 1) == lane) s_cnt[wid][dgt] += __popc(peers);
        __syncwarp();
    }
}

// sorts M pairs; result in (keys, vals).  hist needs 256 * nblk + nblk / 16 + 1024 entries, nblk = ceil(M / RS_TILE).
static int radix_sort_pairs(cudaStream_t stream, const u64 *in, u64 *keys, u64 *keys2, unsigned int *vals, unsigned int *vals2,
                            unsigned int *hist, long long M, long long *launches)
{
    const int nblk = (int)((M + RS_TILE - 1) / RS_TILE);
    k_rs_init<<<(unsigned)((M + 255) / 256), 256, 0, stream>>>(in, keys, vals, M);
    u64 *ka = keys, *kb = keys2;
    unsigned int *va = vals, *vb = vals2;
    for (int pass = 0; pass < 8; pass++) {
        const int shift = 8 * pass;
        k_rs_hist<<<nblk, RS_THREADS, 0, stream>>>(ka, M, shift, nblk, hist);
        {
            const long long nh = 256ll * nblk, nb = (nh + RS_SCAN_TILE - 1) / RS_SCAN_TILE;
            unsigned int *part = hist + nh;                    // (the caller allocates 256 * nblk + nblk / 16 + 1024 entries)
            k_rs_scan_sums<<<(unsigned)nb, 1024, 0, stream>>>(hist, nh, part);
            k_rs_scan_top<<<1, 1024, 0, stream>>>(part, nb);
            k_rs_scan_apply<<<(unsigned)nb, 1024, 0, stream>>>(hist, nh, part);
        }
        k_rs_scatter<<<nblk, RS_THREADS, 0, stream>>>(ka, va, kb, vb, M, shift, nblk, hist);
        u64 *tk = ka; ka = kb; kb = tk;
        unsigned int *tv = va; va = vb; vb = tv;
    }
    if (launches) *launches += 1 + 5 * 8;
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}
