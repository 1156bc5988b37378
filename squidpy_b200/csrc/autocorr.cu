// autocorr.cu — Moran's I / Geary's C for every feature over a sparse weight matrix W (sm_100a).
//
// Replaces the scanpy.metrics.morans_i / gearys_c call of the reference (src/squidpy/gr/_ppatterns.py:216 and,
// per permutation, :267-272).  Arithmetic follows scanpy's published kernels, float64 throughout:
//     I = N/S0 * sum_i z_i * sum_j w_ij z_j / sum_i z_i^2,   z = x - mean(x)
//     C = (N-1) * sum_ij w_ij (x_i - x_j)^2 / (2 * S0 * sum_i (x_i - mean)^2)
//
// SPARSE X (AnnData's native CSR, the path BASELINE configs[2] is quoted on) — work proportional to the non-zeros:
//   The feature matrix is kept as CSR *by feature* with observation indices ascending inside a feature and every
//   segment padded to a multiple of 4 entries (16-byte vector loads).  One CTA walks one feature at a time:
//     1. the feature's stored observations S are marked in a shared-memory bitmap; word w of the bitmap also holds the
//        rank of its first stored observation (entries are sorted, so the rank is the entry's position): a lookup
//        "is j stored, and what is x_j" is ONE 8-byte shared-memory read, + popc + one cached read of the value;
//     2. for every stored observation i the row of W is fetched (packed 64/128-byte rows, 8/16 lanes per row, one
//        slot each) and  acc_i = sum_j w_ij z_j  is formed with z_j = x_j - m for stored j and -m otherwise;
//     3. the rows of W that belong to unstored observations are never touched; their contribution follows from the
//        column sums c_j = sum_i w_ij:   sum_{i not in S} (Wz)_i = c'z - sum_{i in S} (Wz)_i,
//           z'Wz   = A - m * (D - m*(S0 - E) - B)        A = sum_S z_i acc_i, B = sum_S acc_i,
//                                                         D = sum_S c_j z_j,   E = sum_S c_j
//           sum z^2 = sum_S z_i^2 + (N - |S|) m^2
//        (every term is formed from centred values, so nothing of order m^2*N is subtracted from something of the same
//        size except S0 - E, which is exact when nothing is stored and ~0 when everything is);  Geary likewise:
//           sum_ij w_ij (x_i-x_j)^2 = sum_{i in S} sum_j w_ij [(x_i-x_j)^2 - [j in S] x_j^2] + sum_{j in S} c_j x_j^2.
//   A row permutation of W (g[idx, :], the permutation variant) changes neither c nor S0; D, E, sum z^2, m and the
//   Geary column term are kept per feature from the unpermuted pass.
//   Reductions are fixed-order (entry -> lane assignment, shuffle tree, sequential combine over warps) and one CTA
//   owns a feature whatever the grid size: results are bit-reproducible run to run and GPU to GPU.
// DENSE X keeps the tile formulation of round 1: features in tiles of 32 (lane = feature), a dense [N][32] slab per
//   tile, W streamed once per tile.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include <stdlib.h>

#include <algorithm>
#include <thread>

#include "common.cuh"

#define TILE 32

// =====================================================================================================
// dense path (features x obs or obs x features, row-major)
// =====================================================================================================
// ---- dense features x obs (row-major) -> dense tile (transpose through shared memory) -------------------
template <typename XT>
__global__ void ac_transpose_kernel(const XT* __restrict__ x, int64_t g0, int64_t n_feat, int64_t n, XT* __restrict__ D) {
    __shared__ XT tile[TILE][33];
    const int64_t tl = blockIdx.y;
    const int64_t c0 = (int64_t)blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
    for (int r = ty; r < TILE; r += 8) {
        const int64_t g = g0 + tl * TILE + r;
        const int64_t c = c0 + tx;
        tile[r][tx] = (g < n_feat && c < n) ? x[g * n + c] : (XT)0;
    }
    __syncthreads();
    XT* __restrict__ Dt = D + tl * n * TILE;
    for (int r = ty; r < 32; r += 8) {
        const int64_t c = c0 + r;
        if (c < n) Dt[c * TILE + tx] = tile[tx][r];
    }
}

// ---- column sums of a dense tile (fixed order): partial[tile][warp][lane] --------------------------------
template <typename XT>
__global__ void ac_colsum_kernel(const XT* __restrict__ D, int64_t pitch, int64_t tile_stride, int64_t n,
                                 int64_t obs_per_warp, double* __restrict__ partial, int64_t g0, int64_t n_feat) {
    const int lane = threadIdx.x & 31;
    const int64_t w = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int64_t tl = blockIdx.y;
    const bool valid = g0 + tl * TILE + lane < n_feat;
    const XT* __restrict__ Dt = D + tl * tile_stride;
    int64_t r0 = w * obs_per_warp, r1 = r0 + obs_per_warp;
    if (r1 > n) r1 = n;
    double s = 0.0;
    if (valid)
        for (int64_t r = r0; r < r1; ++r) s += (double)Dt[r * pitch + lane];
    partial[(tl * nw + w) * TILE + lane] = s;
}

__global__ void ac_colsum_final_kernel(const double* __restrict__ partial, int64_t nw, double* __restrict__ sums) {
    const int lane = threadIdx.x;
    const int64_t tl = blockIdx.x;
    double s = 0.0;
    for (int64_t w = 0; w < nw; ++w) s += partial[(tl * nw + w) * TILE + lane];
    sums[tl * TILE + lane] = s;
}

// ---- dense SpMM-style kernel -----------------------------------------------------------------------------
// MODE 0: Moran (num = sum_r z_r * sum_e w_e z_{j_e}), MODE 1: Geary (num = sum_r sum_e w_e (x_r - x_{j_e})^2).
// den = sum_r z_r^2 in both modes.  lane = feature; warp = contiguous observation range.
template <typename XT, int MODE>
__global__ void __launch_bounds__(256) ac_main_kernel(const int32_t* __restrict__ wp, const int32_t* __restrict__ wi,
                                                      const double* __restrict__ wd, const XT* __restrict__ D,
                                                      int64_t pitch, int64_t tile_stride, int64_t n, int64_t obs_per_warp,
                                                      const double* __restrict__ sums, const int64_t* __restrict__ row_perm,
                                                      double* __restrict__ pnum, double* __restrict__ pden, int64_t g0,
                                                      int64_t n_feat) {
    const int lane = threadIdx.x & 31;
    const int64_t w = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t tl = blockIdx.y;
    const bool valid = g0 + tl * TILE + lane < n_feat;
    const XT* __restrict__ Dt = D + tl * tile_stride + lane;
    const double mean = sums[tl * TILE + lane] / (double)n;
    int64_t r0 = w * obs_per_warp, r1 = r0 + obs_per_warp;
    if (r1 > n) r1 = n;
    double num = 0.0, den = 0.0;
    if (valid) {
        for (int64_t r = r0; r < r1; ++r) {
            const int64_t src = row_perm ? row_perm[r] : r;
            const int32_t beg = wp[src], end = wp[src + 1];
            const double xr = (double)Dt[r * pitch];
            const double zr = xr - mean;
            double acc = 0.0;
#pragma unroll 4
            for (int32_t e = beg; e < end; ++e) {
                const double xj = (double)Dt[(int64_t)wi[e] * pitch];
                if (MODE == 0) {
                    acc = fma(wd[e], xj - mean, acc);
                } else {
                    const double d = xr - xj;
                    acc = fma(wd[e], d * d, acc);
                }
            }
            if (MODE == 0)
                num = fma(acc, zr, num);
            else
                num += acc;
            den = fma(zr, zr, den);
        }
    }
    // fixed-order combine of the 8 warps of this CTA -> one partial per (tile, CTA)
    __shared__ double s_num[8][TILE], s_den[8][TILE];
    s_num[threadIdx.x >> 5][lane] = num;
    s_den[threadIdx.x >> 5][lane] = den;
    __syncthreads();
    if (threadIdx.x < TILE) {
        double a = 0.0, b = 0.0;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            a += s_num[k][lane];
            b += s_den[k][lane];
        }
        pnum[(tl * gridDim.x + blockIdx.x) * TILE + lane] = a;
        pden[(tl * gridDim.x + blockIdx.x) * TILE + lane] = b;
    }
}

template <int MODE>
__global__ void __launch_bounds__(256) ac_final_kernel(const double* __restrict__ pnum, const double* __restrict__ pden,
                                                       int64_t nparts, int64_t n, double s0, int64_t g0, int64_t n_feat,
                                                       double* __restrict__ out) {
    // 8 warps sum contiguous ranges of the per-CTA partials (lane = feature), then warp 0 combines them in order
    __shared__ double s_num[8][TILE], s_den[8][TILE];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t tl = blockIdx.x;
    const int64_t chunk = (nparts + 7) / 8;
    int64_t p0 = warp * chunk, p1 = p0 + chunk;
    if (p1 > nparts) p1 = nparts;
    double num = 0.0, den = 0.0;
    for (int64_t w = p0; w < p1; ++w) {
        num += pnum[(tl * nparts + w) * TILE + lane];
        den += pden[(tl * nparts + w) * TILE + lane];
    }
    s_num[warp][lane] = num;
    s_den[warp][lane] = den;
    __syncthreads();
    const int64_t g = g0 + tl * TILE + lane;
    if (warp != 0 || g >= n_feat) return;
    num = 0.0;
    den = 0.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        num += s_num[k][lane];
        den += s_den[k][lane];
    }
    double r;
    if (den == 0.0) {
        r = __longlong_as_double(0x7ff8000000000000LL);  // constant feature -> NaN (scanpy)
    } else if (MODE == 0) {
        r = (double)n / s0 * num / den;
    } else {
        r = ((double)(n - 1) * num) / (2.0 * s0 * den);
    }
    out[g] = r;
}

// =====================================================================================================
// sparse path
// =====================================================================================================
#define AC_T 256          // threads per CTA of the per-feature kernels
#define AC_NW (AC_T / 32)

enum { AC_ERR_INDEX = 1, AC_ERR_DUP = 2, AC_ERR_INDPTR = 4, AC_ERR_PERM = 8 };

// ---- input validation / transposition (load time) ----------------------------------------------------
// indptr must start at 0, be non-decreasing and end at nnz
__global__ void ac_check_indptr_kernel(const int64_t* __restrict__ ptr, int64_t rows, int64_t nnz, int* __restrict__ err) {
    for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < rows; r += (int64_t)gridDim.x * blockDim.x) {
        const int64_t a = ptr[r], b = ptr[r + 1];
        if (a > b || a < 0 || b > nnz || (r == 0 && a != 0)) atomicOr(err, AC_ERR_INDPTR);
    }
}

// per-column entry counts of a CSR-by-observation matrix; out-of-range column -> error flag.
// Shared-memory histogram per CTA when n_feat fits (hist_smem), else global atomics.
__global__ void ac_colcount_kernel(const int32_t* __restrict__ xi, int64_t nnz, int64_t n_feat, int32_t col_lo, int hist_smem,
                                   unsigned int* __restrict__ cnt, int* __restrict__ err) {
    extern __shared__ unsigned int s_hist[];
    if (hist_smem)
        for (int64_t c = threadIdx.x; c < n_feat; c += blockDim.x) s_hist[c] = 0;
    __syncthreads();
    const int64_t per = (nnz + gridDim.x - 1) / gridDim.x;
    const int64_t e0 = blockIdx.x * per, e1 = e0 + per < nnz ? e0 + per : nnz;
    for (int64_t e = e0 + threadIdx.x; e < e1; e += blockDim.x) {
        const int32_t c = xi[e] - col_lo;
        if (c < 0 || c >= n_feat) {
            atomicOr(err, AC_ERR_INDEX);
            continue;
        }
        if (hist_smem)
            atomicAdd(&s_hist[c], 1u);
        else
            atomicAdd(&cnt[c], 1u);
    }
    if (!hist_smem) return;
    __syncthreads();
    for (int64_t c = threadIdx.x; c < n_feat; c += blockDim.x) {
        const unsigned int v = s_hist[c];
        if (v) atomicAdd(&cnt[c], v);
    }
}

// ptr[g] = exclusive prefix of cnt, pad[g] = exclusive prefix of cnt rounded up to 4, cursor[g] = ptr[g]
// (single CTA; n_feat + 1 outputs)
__global__ void __launch_bounds__(1024) ac_scan_counts_kernel(const unsigned int* __restrict__ cnt, int64_t n_feat,
                                                              int64_t* __restrict__ ptr, int64_t* __restrict__ pad,
                                                              unsigned long long* __restrict__ cursor) {
    __shared__ int64_t s_a[32], s_b[32];
    __shared__ int64_t s_run[2];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_run[0] = s_run[1] = 0;
    __syncthreads();
    for (int64_t base = 0; base <= n_feat; base += 1024) {
        const int64_t g = base + threadIdx.x;
        const int64_t c = g < n_feat ? (int64_t)cnt[g] : 0;
        int64_t a = c, b = (c + 3) & ~(int64_t)3;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int64_t ta = __shfl_up_sync(0xffffffffu, a, d), tb = __shfl_up_sync(0xffffffffu, b, d);
            if (lane >= d) {
                a += ta;
                b += tb;
            }
        }
        if (lane == 31) {
            s_a[warp] = a;
            s_b[warp] = b;
        }
        __syncthreads();
        if (warp == 0) {
            int64_t va = s_a[lane], vb = s_b[lane];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int64_t ta = __shfl_up_sync(0xffffffffu, va, d), tb = __shfl_up_sync(0xffffffffu, vb, d);
                if (lane >= d) {
                    va += ta;
                    vb += tb;
                }
            }
            s_a[lane] = va - s_a[lane];  // exclusive over warps
            s_b[lane] = vb - s_b[lane];
        }
        __syncthreads();
        const int64_t ea = s_run[0] + s_a[warp] + a - c, eb = s_run[1] + s_b[warp] + b - ((c + 3) & ~(int64_t)3);
        if (g <= n_feat) {
            if (ptr) ptr[g] = ea;
            pad[g] = eb;
            if (cursor) cursor[g] = (unsigned long long)ea;
        }
        __syncthreads();
        if (threadIdx.x == 1023) {
            s_run[0] = ea + c;
            s_run[1] = eb + ((c + 3) & ~(int64_t)3);
        }
        __syncthreads();
    }
}

// padded starts for a caller-supplied CSR by feature: cnt[g] = ptr[g+1] - ptr[g]
__global__ void ac_seglen_kernel(const int64_t* __restrict__ ptr, int64_t n_feat, int64_t n, unsigned int* __restrict__ cnt,
                                 int* __restrict__ err) {
    for (int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; g < n_feat; g += (int64_t)gridDim.x * blockDim.x) {
        int64_t d = ptr[g + 1] - ptr[g];
        if (d < 0) d = 0;  // flagged by ac_check_indptr_kernel
        if (d > n) {
            atomicOr(err, AC_ERR_DUP);  // more entries than observations: some index repeats
            d = n;
        }
        cnt[g] = (unsigned int)d;
    }
}

// CSR by observation -> unsorted segments by feature (atomic cursor; the rank sort below orders them)
template <typename XT>
__global__ void ac_coltranspose_kernel(const int64_t* __restrict__ xp, const int32_t* __restrict__ xi,
                                       const XT* __restrict__ xv, int64_t n_obs, int64_t n_feat, int32_t col_lo,
                                       unsigned long long* __restrict__ cursor, int32_t* __restrict__ oi, XT* __restrict__ ov) {
    const int lane = threadIdx.x & 31;
    const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5; r < n_obs; r += nw) {
        for (int64_t e = xp[r] + lane; e < xp[r + 1]; e += 32) {
            const int32_t c = xi[e] - col_lo;
            if (c < 0 || c >= n_feat) continue;  // flagged by the count kernel
            const unsigned long long pos = atomicAdd(&cursor[c], 1ULL);
            oi[pos] = (int32_t)r;
            ov[pos] = xv[e];
        }
    }
}

// ---- bitmap + rank structure ------------------------------------------------------------------------------
// S[w] = { bits of the observations 32w .. 32w+31 stored for the feature, rank of the first of them }
__device__ __forceinline__ void ac_zero_words(uint2* S, int nwords) {
    for (int w = threadIdx.x; w < nwords; w += AC_T) S[w] = make_uint2(0u, 0u);
}

// exclusive prefix of popc(S[w].x) into S[w].y; returns the total through s_tmp[32] (all threads must call)
__device__ __forceinline__ unsigned int ac_scan_words(uint2* S, int nwords, unsigned int* s_tmp) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int chunk = (nwords + AC_T - 1) / AC_T;
    int w0 = threadIdx.x * chunk, w1 = w0 + chunk;
    if (w0 > nwords) w0 = nwords;
    if (w1 > nwords) w1 = nwords;
    unsigned int cnt = 0;
    for (int w = w0; w < w1; ++w) cnt += __popc(S[w].x);
    unsigned int inc = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned int t = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += t;
    }
    if (lane == 31) s_tmp[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        const unsigned int v = lane < AC_NW ? s_tmp[lane] : 0u;
        unsigned int s = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned int t = __shfl_up_sync(0xffffffffu, s, d);
            if (lane >= d) s += t;
        }
        if (lane < AC_NW) s_tmp[lane] = s - v;
        if (lane == AC_NW - 1) s_tmp[AC_NW] = s;
    }
    __syncthreads();
    unsigned int run = inc - cnt + s_tmp[warp];
    for (int w = w0; w < w1; ++w) {
        const unsigned int b = S[w].x;
        S[w].y = run;
        run += __popc(b);
    }
    const unsigned int total = s_tmp[AC_NW];
    __syncthreads();
    return total;
}

// One CTA per feature: order the (observation, value) pairs of a segment by observation (rank = number of stored
// observations below), write them to the padded segment, pad with index -1 / value 0.  Out-of-range or repeated
// observation indices raise the error flag.
template <typename XT>
__global__ void __launch_bounds__(AC_T) ac_rank_sort_kernel(const int64_t* __restrict__ in_ptr, const int32_t* __restrict__ in_idx,
                                                             const XT* __restrict__ in_val, const int64_t* __restrict__ out_start,
                                                             int32_t* __restrict__ out_idx, XT* __restrict__ out_val,
                                                             int32_t* __restrict__ seg_len, int64_t n, int64_t n_feat,
                                                             int nwords, uint2* __restrict__ s_global, int* __restrict__ err) {
    extern __shared__ __align__(16) unsigned char ac_smem[];
    __shared__ unsigned int s_tmp[AC_NW + 1];
    uint2* S = s_global ? s_global + (size_t)blockIdx.x * nwords : reinterpret_cast<uint2*>(ac_smem);
    ac_zero_words(S, nwords);
    __syncthreads();
    for (int64_t g = blockIdx.x; g < n_feat; g += gridDim.x) {
        const int64_t rb = in_ptr[g];
        int64_t len64 = in_ptr[g + 1] - rb;
        if (len64 < 0) len64 = 0;
        if (len64 > n) len64 = n;  // flagged by ac_seglen_kernel
        const int len = (int)len64;
        const int64_t os = out_start[g];
        for (int e = threadIdx.x; e < len; e += AC_T) {
            const int32_t i = in_idx[rb + e];
            if (i < 0 || i >= n) {
                atomicOr(err, AC_ERR_INDEX);
                continue;
            }
            const unsigned int bit = 1u << (i & 31);
            if (atomicOr(&S[i >> 5].x, bit) & bit) atomicOr(err, AC_ERR_DUP);
        }
        __syncthreads();
        ac_scan_words(S, nwords, s_tmp);
        for (int e = threadIdx.x; e < len; e += AC_T) {
            const int32_t i = in_idx[rb + e];
            if (i < 0 || i >= n) continue;
            const uint2 s = S[i >> 5];
            const unsigned int bit = 1u << (i & 31);
            const int64_t pos = os + s.y + __popc(s.x & (bit - 1u));
            out_idx[pos] = i;
            out_val[pos] = in_val[rb + e];
        }
        const int lenp = (len + 3) & ~3;
        for (int e = len + threadIdx.x; e < lenp; e += AC_T) {
            out_idx[os + e] = -1;
            out_val[os + e] = (XT)0;
        }
        if (threadIdx.x == 0) seg_len[g] = len;
        __syncthreads();
        ac_zero_words(S, nwords);
        __syncthreads();
    }
}

// ---- W preprocessing ---------------------------------------------------------------------------------------
// packed rows: LPR slots of 8 bytes {column (int32, -1 = empty), weight (float32)}
template <int LPR>
__global__ void ac_pack_rows_kernel(const int32_t* __restrict__ wp, const int32_t* __restrict__ wi, const double* __restrict__ wd,
                                    int64_t n, uint2* __restrict__ rows) {
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n * LPR; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = t / LPR;
        const int32_t e = wp[r] + (int)(t % LPR);
        rows[t] = e < wp[r + 1] ? make_uint2((unsigned int)wi[e], __float_as_uint((float)wd[e])) : make_uint2(0xffffffffu, 0u);
    }
}

// compact rows (FMT 3): 8 slots of 4 bytes = 7 columns (-1 = empty) + the float32 weight ALL entries of the row share (the
// row-normalised binary graph of the standard pipeline): one 32-byte sector per row instead of a 64-byte line
__global__ void ac_pack_rows32_kernel(const int32_t* __restrict__ wp, const int32_t* __restrict__ wi, const double* __restrict__ wd,
                                      int64_t n, uint32_t* __restrict__ rows) {
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n * 8; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = t >> 3;
        const int s = (int)(t & 7);
        const int32_t b = wp[r], e = wp[r + 1];
        if (s < 7) rows[t] = b + s < e ? (uint32_t)wi[b + s] : 0xffffffffu;
        else rows[t] = e > b ? __float_as_uint((float)wd[b]) : 0u;
    }
}

// row permutation: int64 host layout -> int32, validated on the device (bitmap of seen sources)
__global__ void ac_perm_prepare_kernel(const int64_t* __restrict__ in, int64_t n, int32_t* __restrict__ out,
                                       unsigned int* __restrict__ seen, int* __restrict__ err) {
    for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x) {
        const int64_t s = in[r];
        if (s < 0 || s >= n) {
            atomicOr(err, AC_ERR_PERM);
            out[r] = 0;
            continue;
        }
        const unsigned int bit = 1u << (s & 31);
        if (atomicOr(&seen[s >> 5], bit) & bit) atomicOr(err, AC_ERR_PERM);
        out[r] = (int32_t)s;
    }
}

// ---- the per-feature kernel ---------------------------------------------------------------------------------
struct AcSparseParams {
    int64_t n, n_feat;
    int nwords;
    uint2* s_global;  // bitmap scratch in global memory (nwords per CTA) when it does not fit shared memory, else null
    const int64_t* seg_start;
    const int32_t* seg_len;
    const int32_t* order;  // features by decreasing length (longest first), or null
    const int32_t* xi;
    const void* xv;
    // W
    const uint2* rows;  // packed rows (FMT 0 / 1)
    const int32_t* wp;  // CSR (FMT 2)
    const int32_t* wi;
    const double* wd;
    const double* csum;
    double s0;
    const int32_t* perm;  // PERM: n_perm row permutations [n_perm][n] (one feature pass serves all of them), else null
    int n_perm;           // PERM: permutations per launch; outputs go to out[b * n_feat + g]
    double* aux;          // [5][n_feat]: mean, sum z^2 over stored, D, E, Geary column term (written by unpermuted runs)
    double* out;
    int* counter;  // dynamic feature queue
};

// fixed-order block reduction of NQ doubles; result valid in thread 0
template <int NQ>
__device__ __forceinline__ void ac_block_reduce(double (&q)[NQ], double (*s_red)[AC_NW]) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < NQ; ++k) {
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) q[k] += __shfl_xor_sync(0xffffffffu, q[k], d);
        if (lane == 0) s_red[k][warp] = q[k];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < NQ; ++k) {
            double t = 0.0;
#pragma unroll
            for (int w = 0; w < AC_NW; ++w) t += s_red[k][w];
            q[k] = t;
        }
    }
}

// FMT 0: packed rows, 8 lanes per row (<= 8 entries);  FMT 1: packed rows, 16 lanes per row (<= 16 entries);
// FMT 2: CSR rows, 8 lanes per row (any length; float64 weights);  FMT 3: compact rows, 8 lanes per row (<= 7 entries of one
// common float32 weight: 32 bytes per row).
template <typename XT, int MODE, int FMT, bool PERM, int GRP>
__global__ void __launch_bounds__(AC_T, GRP == 8 ? 3 : 4) ac_sparse_kernel(const __grid_constant__ AcSparseParams p) {
    constexpr int LPR = FMT == 1 ? 16 : 8;
    constexpr int RPW = 32 / LPR;        // rows per warp step
    constexpr int NSUB = 32 / RPW;       // warp steps per block of 32 entries
    // GRP = warp steps handled together: GRP independent row loads / look-ups / value loads in flight per warp
    extern __shared__ __align__(16) unsigned char ac_smem[];
    __shared__ double s_red[6][AC_NW];
    __shared__ double s_mean;
    __shared__ int s_next;
    uint2* S = p.s_global ? p.s_global + (size_t)blockIdx.x * p.nwords : reinterpret_cast<uint2*>(ac_smem);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int slot = lane % LPR, sub = lane / LPR;
    const double dn = (double)p.n;
    ac_zero_words(S, p.nwords);
    __syncthreads();
    for (;;) {
        if (threadIdx.x == 0) s_next = atomicAdd(p.counter, 1);
        __syncthreads();
        const int k = s_next;
        if (k >= p.n_feat) break;
        const int g = p.order ? p.order[k] : k;
        const int64_t b = p.seg_start[g];
        const int len = p.seg_len[g];
        const int32_t* __restrict__ xi = p.xi + b;
        const XT* __restrict__ xv = reinterpret_cast<const XT*>(p.xv) + b;
        const int len4 = (len + 3) >> 2;
        // ---- phase 1: mark the stored observations, rank of the first one of every word, sum of the values ----
        double q[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
#pragma unroll 2
        for (int v = threadIdx.x; v < len4; v += AC_T) {
            const int4 iv = __ldg(reinterpret_cast<const int4*>(xi) + v);
            const int prev = v > 0 ? __ldg(xi + 4 * v - 1) : -1;
            const int ii[4] = {iv.x, iv.y, iv.z, iv.w};
            if (!PERM) {
                double xs[4];
                if (sizeof(XT) == 4) {
                    const float4 f = __ldg(reinterpret_cast<const float4*>(xv) + v);
                    xs[0] = f.x, xs[1] = f.y, xs[2] = f.z, xs[3] = f.w;
                } else {
                    const double2 d0 = __ldg(reinterpret_cast<const double2*>(xv) + 2 * v);
                    const double2 d1 = __ldg(reinterpret_cast<const double2*>(xv) + 2 * v + 1);
                    xs[0] = d0.x, xs[1] = d0.y, xs[2] = d1.x, xs[3] = d1.y;
                }
                q[0] += ((xs[0] + xs[1]) + xs[2]) + xs[3];  // pads hold 0
            }
            int cur_w = prev >= 0 ? prev >> 5 : -1;  // word of the previous entry
            int open_w = -1;
            unsigned int mask = 0;
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const int i = ii[t];
                if (i < 0) break;
                const int w = i >> 5;
                if (w != open_w) {
                    if (open_w >= 0) atomicOr(&S[open_w].x, mask);
                    open_w = w;
                    mask = 0;
                }
                mask |= 1u << (i & 31);
                if (w != cur_w) {
                    S[w].y = (unsigned int)(4 * v + t);  // first stored observation of this word
                    cur_w = w;
                }
            }
            if (open_w >= 0) atomicOr(&S[open_w].x, mask);
        }
        double m;
        if (!PERM) {
            double q1[1] = {q[0]};
            ac_block_reduce<1>(q1, s_red);  // contains a __syncthreads after the shuffles
            if (threadIdx.x == 0) s_mean = q1[0] / dn;
            __syncthreads();
            m = s_mean;
            q[0] = 0.0;
            // ---- phase 1b (unpermuted pass only): the column-sum terms D = sum c z, E = sum c, sum z^2, Geary's sum c x^2 ----
#pragma unroll 2
            for (int v = threadIdx.x; v < len4; v += AC_T) {
                const int4 iv = __ldg(reinterpret_cast<const int4*>(xi) + v);
                const int ii[4] = {iv.x, iv.y, iv.z, iv.w};
                double xs[4], cs[4];
                if (sizeof(XT) == 4) {
                    const float4 f = __ldg(reinterpret_cast<const float4*>(xv) + v);
                    xs[0] = f.x, xs[1] = f.y, xs[2] = f.z, xs[3] = f.w;
                } else {
                    const double2 d0 = __ldg(reinterpret_cast<const double2*>(xv) + 2 * v);
                    const double2 d1 = __ldg(reinterpret_cast<const double2*>(xv) + 2 * v + 1);
                    xs[0] = d0.x, xs[1] = d0.y, xs[2] = d1.x, xs[3] = d1.y;
                }
#pragma unroll
                for (int t = 0; t < 4; ++t) cs[t] = ii[t] >= 0 ? __ldg(p.csum + ii[t]) : 0.0;
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    if (ii[t] < 0) continue;
                    const double zi = xs[t] - m;
                    q[2] = fma(cs[t], zi, q[2]);
                    q[3] += cs[t];
                    q[4] = fma(zi, zi, q[4]);
                    q[5] = fma(cs[t], xs[t] * xs[t], q[5]);
                }
            }
        } else {
            m = p.aux[g];
            __syncthreads();
        }
        // ---- phase 2: walk the rows of W of the stored observations; q[0], q[1] = slot accumulators ----
        // Moran: q[0] += z_i * w z_j, q[1] += w z_j;   Geary: q[0] += w ((x_i - x_j)^2 - [j stored] x_j^2)
        // PERM: the bitmap of the feature is built once and walked once per row permutation of the batch
        for (int pb = 0; pb < (PERM ? p.n_perm : 1); ++pb) {
        const int32_t* __restrict__ perm = PERM ? p.perm + (int64_t)pb * p.n : nullptr;
        if (PERM) q[0] = q[1] = 0.0;
        int nxt_i = -1;
        XT nxt_x = (XT)0;
        {
            const int e = warp * 32 + lane;
            if (e < len) {
                nxt_i = __ldg(xi + e);
                nxt_x = __ldg(xv + e);
            }
        }
        for (int base = warp * 32; base < len; base += AC_NW * 32) {
            const int my_i = nxt_i;
            const XT my_x = nxt_x;
            {  // prefetch this warp's next block of 32 entries
                const int e = base + AC_NW * 32 + lane;
                nxt_i = -1;
                nxt_x = (XT)0;
                if (e < len) {
                    nxt_i = __ldg(xi + e);
                    nxt_x = __ldg(xv + e);
                }
            }
#pragma unroll
            for (int s0 = 0; s0 < NSUB; s0 += GRP) {
                int ri[GRP];
                double rx[GRP];
                if (FMT != 2) {
                    // stage 1: the row slots; stage 2: bitmap lookups; stage 3: values of the hits; stage 4: accumulate.
                    // No branches: an empty slot / a missing row has weight 0, column 0 and a suppressed value load.
                    uint2 rs[GRP];
#pragma unroll
                    for (int s = 0; s < GRP; ++s) {
                        const int src = (s0 + s) * RPW + sub;
                        ri[s] = __shfl_sync(0xffffffffu, my_i, src);
                        rx[s] = (double)__shfl_sync(0xffffffffu, my_x, src);
                        rs[s] = make_uint2(0xffffffffu, 0u);
                        if (FMT == 3) {  // 7 columns + the row's weight: 4 bytes per lane, the weight comes from lane 7 of the row
                            unsigned int cw = 0xffffffffu;
                            if (ri[s] >= 0) {
                                const int64_t r = PERM ? (int64_t)__ldg(perm + ri[s]) : (int64_t)ri[s];
                                cw = __ldg(reinterpret_cast<const unsigned int*>(p.rows) + r * 8 + slot);
                            }
                            const unsigned int wb = __shfl_sync(0xffffffffu, cw, (lane & ~7) | 7);
                            rs[s] = make_uint2(slot == 7 ? 0xffffffffu : cw, wb);
                        } else if (ri[s] >= 0) {
                            const int64_t r = PERM ? (int64_t)__ldg(perm + ri[s]) : (int64_t)ri[s];
                            rs[s] = __ldg(p.rows + r * LPR + slot);
                        }
                    }
                    bool hit[GRP];
                    unsigned int rank[GRP];
#pragma unroll
                    for (int s = 0; s < GRP; ++s) {
                        const int j = (int)rs[s].x;
                        const int jj = j >= 0 ? j : 0;
                        const uint2 w = S[jj >> 5];
                        const unsigned int bit = 1u << (jj & 31);
                        hit[s] = j >= 0 && (w.x & bit) != 0u;
                        rank[s] = w.y + __popc(w.x & (bit - 1u));
                    }
                    XT xj[GRP];
#pragma unroll
                    for (int s = 0; s < GRP; ++s) xj[s] = hit[s] ? __ldg(xv + rank[s]) : (XT)0;
#pragma unroll
                    for (int s = 0; s < GRP; ++s) {
                        const double w = (int)rs[s].x >= 0 ? (double)__uint_as_float(rs[s].y) : 0.0;
                        const double xjd = (double)xj[s];
                        if (MODE == 0) {
                            const double t = w * (hit[s] ? xjd - m : -m);
                            q[0] = fma(rx[s] - m, t, q[0]);
                            q[1] += t;
                        } else {
                            const double d = rx[s] - xjd;
                            q[0] = fma(w, d * d - xjd * xjd, q[0]);  // xj == 0 when not stored
                        }
                    }
                } else {
#pragma unroll
                    for (int s = 0; s < GRP; ++s) {
                        const int src = (s0 + s) * RPW + sub;
                        ri[s] = __shfl_sync(0xffffffffu, my_i, src);
                        rx[s] = (double)__shfl_sync(0xffffffffu, my_x, src);
                    }
#pragma unroll
                    for (int s = 0; s < GRP; ++s) {
                        if (ri[s] < 0) continue;
                        const int64_t r = PERM ? (int64_t)__ldg(perm + ri[s]) : (int64_t)ri[s];
                        const int rbeg = __ldg(p.wp + r), rend = __ldg(p.wp + r + 1);
                        const double zi = rx[s] - m;
                        for (int ew = rbeg + slot; ew < rend; ew += LPR) {
                            const int j = __ldg(p.wi + ew);
                            const double w = __ldg(p.wd + ew);
                            const uint2 sw = S[j >> 5];
                            const unsigned int bit = 1u << (j & 31);
                            const bool h = (sw.x & bit) != 0u;
                            const double xjd = h ? (double)__ldg(xv + (sw.y + __popc(sw.x & (bit - 1u)))) : 0.0;
                            if (MODE == 0) {
                                const double t = w * (h ? xjd - m : -m);
                                q[0] = fma(zi, t, q[0]);
                                q[1] += t;
                            } else {
                                const double d = rx[s] - xjd;
                                q[0] = fma(w, d * d - xjd * xjd, q[0]);
                            }
                        }
                    }
                }
            }
        }
        // ---- phase 3: combine, finish the feature, clear the touched words ----
        if (!PERM) {
            ac_block_reduce<6>(q, s_red);
        } else {
            double q2[2] = {q[0], q[1]};
            ac_block_reduce<2>(q2, s_red);
            q[0] = q2[0];
            q[1] = q2[1];
        }
        if (threadIdx.x == 0) {
            double D, E, Z2, CX;
            if (!PERM) {
                D = q[2], E = q[3], Z2 = q[4], CX = q[5];
                p.aux[g] = m;
                p.aux[p.n_feat + g] = Z2;
                p.aux[2 * p.n_feat + g] = D;
                p.aux[3 * p.n_feat + g] = E;
                p.aux[4 * p.n_feat + g] = CX;
            } else {
                Z2 = p.aux[p.n_feat + g];
                D = p.aux[2 * p.n_feat + g];
                E = p.aux[3 * p.n_feat + g];
                CX = p.aux[4 * p.n_feat + g];
            }
            const double den = Z2 + (dn - (double)len) * m * m;
            double r;
            if (den == 0.0) {
                r = __longlong_as_double(0x7ff8000000000000LL);  // constant feature -> NaN (scanpy)
            } else if (MODE == 0) {
                const double num = q[0] - m * (D - m * (p.s0 - E) - q[1]);
                r = dn / p.s0 * num / den;
            } else {
                const double num = q[0] + CX;
                r = ((dn - 1.0) * num) / (2.0 * p.s0 * den);
            }
            p.out[(int64_t)pb * p.n_feat + g] = r;
        }
        if (PERM) __syncthreads();  // s_red is reused by the next permutation of the batch
        }  // permutations of the batch
#pragma unroll 4
        for (int v = threadIdx.x; v < len4; v += AC_T) {
            const int4 iv = __ldg(reinterpret_cast<const int4*>(xi) + v);
            if (iv.x >= 0) S[iv.x >> 5] = make_uint2(0u, 0u);
            if (iv.y >= 0) S[iv.y >> 5] = make_uint2(0u, 0u);
            if (iv.z >= 0) S[iv.z >> 5] = make_uint2(0u, 0u);
            if (iv.w >= 0) S[iv.w >> 5] = make_uint2(0u, 0u);
        }
        __syncthreads();
    }
}

// ================================================================================================
struct sqb_autocorr {
    sqb_ctx* ctx = nullptr;
    int64_t n = 0, nnz = 0;
    double s0 = 0.0;
    int w_fmt = 2;  // 0 packed 8 lanes, 1 packed 16 lanes, 2 CSR, 3 compact (7 columns + one common weight in 32 bytes)
    DevBuf<int32_t> d_wp, d_wi;
    DevBuf<double> d_wd, d_csum;
    DevBuf<uint2> d_rows;
    // loaded X
    int kind = 0;     // 0 none, 1 dense feat x obs, 2 dense obs x feat (zero copy tiles), 3 sparse (padded CSR by feature)
    int x_dtype = 0;  // 0 f32, 1 f64
    int64_t n_feat = 0, x_nnz = 0;
    DevBuf<uint8_t> d_x;  // dense matrix or padded sparse values
    DevBuf<int64_t> d_xp;  // padded segment starts (n_feat + 1)
    DevBuf<int32_t> d_xi, d_len, d_order;
    DevBuf<uint8_t> d_tile;
    DevBuf<double> d_sums, d_partial, d_pnum, d_pden, d_out, d_aux, d_out_perms;
    DevBuf<int64_t> d_perm;
    DevBuf<int32_t> d_perm32;
    DevBuf<unsigned int> d_seen;
    DevBuf<uint2> d_bitmap;  // global-memory bitmap scratch (large n)
    DevBuf<int> d_flags;     // [0] error flags, [1] feature queue counter
    bool aux_valid = false;
    int tiles_per_launch = 8;  // dense path: 256 features per launch
    bool ran = false;
};

static int xsize(int dt) { return dt == 0 ? 4 : 8; }

static const char* ac_err_text(int f) {
    if (f & AC_ERR_INDPTR) return "indptr is not a non-decreasing sequence from 0 to nnz";
    if (f & AC_ERR_INDEX) return "index out of range";
    if (f & AC_ERR_DUP) return "an observation is stored twice for one feature (sum duplicates first)";
    if (f & AC_ERR_PERM) return "row_perm is not a permutation";
    return "invalid input";
}

// bitmap placement: shared memory when 8 bytes per 32 observations fit, else one global scratch slice per CTA
struct AcGeom {
    int nwords = 0;
    size_t smem = 0;
    int ctas = 0;
    bool global = false;
};

static int ac_geometry(sqb_autocorr* h, const void* kernel, AcGeom* g) {
    sqb_ctx* c = h->ctx;
    g->nwords = (int)ceil_div64(h->n, 32);
    const size_t need = (size_t)g->nwords * sizeof(uint2);
    const size_t avail = c->smem_optin > 4096 ? c->smem_optin - 4096 : 0;  // static shared memory of the kernels
    g->global = need > avail;
    g->smem = g->global ? 0 : need;
    if (!g->global && g->smem > 48 * 1024) SQB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g->smem));
    int per_sm = 0;
    SQB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, AC_T, g->smem));
    if (per_sm < 1) per_sm = 1;
    g->ctas = per_sm * c->sm_count;
    if (g->global) SQB_TRY(h->d_bitmap.alloc((size_t)g->ctas * g->nwords));
    return SQB_OK;
}

template <typename XT, int MODE, int FMT, bool PERM, int GRP>
static int ac_sparse_launch_g(sqb_autocorr* h, const AcSparseParams& base) {
    sqb_ctx* c = h->ctx;
    const void* kernel = (const void*)ac_sparse_kernel<XT, MODE, FMT, PERM, GRP>;
    AcGeom geo;
    SQB_TRY(ac_geometry(h, kernel, &geo));
    AcSparseParams p = base;
    p.nwords = geo.nwords;
    p.s_global = geo.global ? h->d_bitmap.p : nullptr;
    int64_t ctas = geo.ctas;
    if (ctas > h->n_feat) ctas = h->n_feat;
    SQB_CUDA(cudaMemsetAsync(h->d_flags.p + 1, 0, sizeof(int), c->stream));
    SqbLaunchScope scope(c, SQB_K_AUTOCORR_MAIN);
    ac_sparse_kernel<XT, MODE, FMT, PERM, GRP><<<(unsigned)ctas, AC_T, geo.smem, c->stream>>>(p);
    SQB_POST_LAUNCH();
    return SQB_OK;
}

template <typename XT, int MODE, int FMT, bool PERM>
static int ac_sparse_launch(sqb_autocorr* h, const AcSparseParams& base) {
    // 8 row groups in flight per warp (measured at configs[2]: 7.15 ms; 4 groups: 8.7 ms)
    return ac_sparse_launch_g<XT, MODE, FMT, PERM, 8>(h, base);
}

template <typename XT, int MODE, bool PERM>
static int ac_sparse_fmt(sqb_autocorr* h, const AcSparseParams& p) {
    switch (h->w_fmt) {
        case 0: return ac_sparse_launch<XT, MODE, 0, PERM>(h, p);
        case 1: return ac_sparse_launch<XT, MODE, 1, PERM>(h, p);
        case 3: return ac_sparse_launch<XT, MODE, 3, PERM>(h, p);
        default: return ac_sparse_launch<XT, MODE, 2, PERM>(h, p);
    }
}

// one pass over all features; d_perm32 == nullptr: unpermuted (also refreshes the per-feature invariants)
static int ac_run_sparse(sqb_autocorr* h, int mode, const int32_t* d_perm32, double* d_out, int n_perm = 1) {
    AcSparseParams p;
    memset(&p, 0, sizeof(p));
    p.n = h->n;
    p.n_feat = h->n_feat;
    p.seg_start = h->d_xp.p;
    p.seg_len = h->d_len.p;
    p.order = h->d_order.p;
    p.xi = h->d_xi.p;
    p.xv = h->d_x.p;
    p.rows = h->d_rows.p;
    p.wp = h->d_wp.p;
    p.wi = h->d_wi.p;
    p.wd = h->d_wd.p;
    p.csum = h->d_csum.p;
    p.s0 = h->s0;
    p.perm = d_perm32;
    p.n_perm = n_perm;
    p.aux = h->d_aux.p;
    p.out = d_out;
    p.counter = h->d_flags.p + 1;
    const bool f32 = h->x_dtype == 0;
    if (!d_perm32) {
        int rc;
        if (mode == 0)
            rc = f32 ? ac_sparse_fmt<float, 0, false>(h, p) : ac_sparse_fmt<double, 0, false>(h, p);
        else
            rc = f32 ? ac_sparse_fmt<float, 1, false>(h, p) : ac_sparse_fmt<double, 1, false>(h, p);
        if (rc == SQB_OK) h->aux_valid = true;
        return rc;
    }
    if (mode == 0) return f32 ? ac_sparse_fmt<float, 0, true>(h, p) : ac_sparse_fmt<double, 0, true>(h, p);
    return f32 ? ac_sparse_fmt<float, 1, true>(h, p) : ac_sparse_fmt<double, 1, true>(h, p);
}

template <typename XT>
static int ac_run_dense(sqb_autocorr* h, int mode, const int64_t* d_perm) {
    sqb_ctx* c = h->ctx;
    const int64_t n = h->n, G = h->n_feat;
    const int64_t ntiles = ceil_div64(G, TILE);
    const int TPL = (int)(ntiles < h->tiles_per_launch ? ntiles : h->tiles_per_launch);
    // warps: contiguous observation ranges, >= 32 observations per warp
    int64_t ctas = (int64_t)c->sm_count * 4;
    int64_t nw = ctas * 8;
    int64_t obs_per_warp = ceil_div64(n, nw);
    if (obs_per_warp < 32) obs_per_warp = 32;
    nw = ceil_div64(n, obs_per_warp);
    ctas = ceil_div64(nw, 8);
    nw = ctas * 8;
    SQB_TRY(h->d_sums.alloc((size_t)TPL * TILE));
    SQB_TRY(h->d_partial.alloc((size_t)TPL * nw * TILE));
    SQB_TRY(h->d_pnum.alloc((size_t)TPL * nw * TILE));
    SQB_TRY(h->d_pden.alloc((size_t)TPL * nw * TILE));
    const bool zero_copy = (h->kind == 2);
    if (!zero_copy) SQB_TRY(h->d_tile.alloc((size_t)TPL * n * TILE * sizeof(XT)));
    XT* Dbuf = reinterpret_cast<XT*>(h->d_tile.p);
    const XT* X = reinterpret_cast<const XT*>(h->d_x.p);
    for (int64_t t0 = 0; t0 < ntiles; t0 += TPL) {
        const int nt = (int)(ntiles - t0 < TPL ? ntiles - t0 : TPL);
        const int64_t g0 = t0 * TILE;
        const XT* D;
        int64_t pitch, tile_stride;
        if (zero_copy) {
            D = X + g0;
            pitch = G;
            tile_stride = TILE;
        } else {
            D = Dbuf;
            pitch = TILE;
            tile_stride = n * TILE;
            SqbLaunchScope scope(c, SQB_K_AUTOCORR_PREP);
            dim3 grid((unsigned)ceil_div64(n, 32), (unsigned)nt);
            ac_transpose_kernel<XT><<<grid, 256, 0, c->stream>>>(X, g0, G, n, Dbuf);
            SQB_POST_LAUNCH();
        }
        {
            SqbLaunchScope scope(c, SQB_K_AUTOCORR_PREP);
            dim3 grid((unsigned)ctas, (unsigned)nt);
            ac_colsum_kernel<XT><<<grid, 256, 0, c->stream>>>(D, pitch, tile_stride, n, obs_per_warp, h->d_partial.p, g0, G);
            SQB_POST_LAUNCH();
        }
        {
            SqbLaunchScope scope(c, SQB_K_AUTOCORR_PREP);
            ac_colsum_final_kernel<<<nt, 32, 0, c->stream>>>(h->d_partial.p, nw, h->d_sums.p);
            SQB_POST_LAUNCH();
        }
        {
            SqbLaunchScope scope(c, SQB_K_AUTOCORR_MAIN);
            dim3 grid((unsigned)ctas, (unsigned)nt);
            if (mode == 0)
                ac_main_kernel<XT, 0><<<grid, 256, 0, c->stream>>>(h->d_wp.p, h->d_wi.p, h->d_wd.p, D, pitch, tile_stride, n,
                                                                   obs_per_warp, h->d_sums.p, d_perm, h->d_pnum.p,
                                                                   h->d_pden.p, g0, G);
            else
                ac_main_kernel<XT, 1><<<grid, 256, 0, c->stream>>>(h->d_wp.p, h->d_wi.p, h->d_wd.p, D, pitch, tile_stride, n,
                                                                   obs_per_warp, h->d_sums.p, d_perm, h->d_pnum.p,
                                                                   h->d_pden.p, g0, G);
            SQB_POST_LAUNCH();
        }
        {
            SqbLaunchScope scope(c, SQB_K_AUTOCORR_FINAL);
            if (mode == 0)
                ac_final_kernel<0><<<nt, 256, 0, c->stream>>>(h->d_pnum.p, h->d_pden.p, ctas, n, h->s0, g0, G, h->d_out.p);
            else
                ac_final_kernel<1><<<nt, 256, 0, c->stream>>>(h->d_pnum.p, h->d_pden.p, ctas, n, h->s0, g0, G, h->d_out.p);
            SQB_POST_LAUNCH();
        }
    }
    return SQB_OK;
}

// read the error flag word (synchronises the stream)
static int ac_check_flags(sqb_autocorr* h, const char* where) {
    sqb_ctx* c = h->ctx;
    int f = 0;
    SQB_CUDA(cudaMemcpyAsync(&f, h->d_flags.p, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    SQB_CUDA(cudaStreamSynchronize(c->stream));
    if (f != 0) {
        SQB_CUDA(cudaMemsetAsync(h->d_flags.p, 0, sizeof(int), c->stream));
        sqb_set_error("%s: %s", where, ac_err_text(f));
        return SQB_ERR_INVALID;
    }
    return SQB_OK;
}

// raw segments by feature (ptr / idx / val on the device, any order inside a segment) -> padded, ordered layout
template <typename XT>
static int ac_build_segments(sqb_autocorr* h, const int64_t* d_ptr, const unsigned int* d_cnt, int64_t* d_pad,
                             const int32_t* d_idx, const XT* d_val, int64_t nnz, int64_t n_feat) {
    sqb_ctx* c = h->ctx;
    const int64_t cap = nnz + 4 * n_feat + 4;
    SQB_TRY(h->d_xi.alloc((size_t)cap));
    SQB_TRY(h->d_x.alloc((size_t)cap * sizeof(XT)));
    SQB_TRY(h->d_len.alloc((size_t)n_feat));
    SQB_TRY(h->d_order.alloc((size_t)n_feat));
    SQB_TRY(h->d_aux.alloc((size_t)5 * n_feat));
    SQB_TRY(h->d_out.alloc((size_t)n_feat));
    const void* kernel = (const void*)ac_rank_sort_kernel<XT>;
    AcGeom geo;
    SQB_TRY(ac_geometry(h, kernel, &geo));
    int64_t ctas = geo.ctas < n_feat ? geo.ctas : n_feat;
    {
        SqbLaunchScope scope(c, SQB_K_AUTOCORR_PREP);
        ac_rank_sort_kernel<XT><<<(unsigned)ctas, AC_T, geo.smem, c->stream>>>(d_ptr, d_idx, d_val, d_pad, h->d_xi.p,
                                                                             reinterpret_cast<XT*>(h->d_x.p), h->d_len.p, h->n,
                                                                             n_feat, geo.nwords, geo.global ? h->d_bitmap.p : nullptr,
                                                                             h->d_flags.p);
        SQB_POST_LAUNCH();
    }
    // features by decreasing length: the queue hands out the long ones first (LPT), the tail of the grid stays short
    {
        DevBuf<unsigned int> keys_out;
        DevBuf<int32_t> iota;
        DevBuf<uint8_t> tmp;
        keys_out.bind(c->stream);
        iota.bind(c->stream);
        tmp.bind(c->stream);
        int rc;
        if ((rc = keys_out.alloc(n_feat)) || (rc = iota.alloc(n_feat))) {
            keys_out.release();
            iota.release();
            return rc;
        }
        std::vector<int32_t> hi((size_t)n_feat);
        for (int64_t g = 0; g < n_feat; ++g) hi[g] = (int32_t)g;
        cudaError_t e = cudaMemcpyAsync(iota.p, hi.data(), n_feat * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream);
        size_t tmp_bytes = 0;
        if (e == cudaSuccess)
            e = cub::DeviceRadixSort::SortPairsDescending(nullptr, tmp_bytes, d_cnt, keys_out.p, iota.p, h->d_order.p, (int)n_feat, 0, 32,
                                                          c->stream);
        if (e == cudaSuccess && tmp.alloc(tmp_bytes > 0 ? tmp_bytes : 1) != SQB_OK) e = cudaErrorMemoryAllocation;
        if (e == cudaSuccess)
            e = cub::DeviceRadixSort::SortPairsDescending(tmp.p, tmp_bytes, d_cnt, keys_out.p, iota.p, h->d_order.p, (int)n_feat, 0, 32,
                                                          c->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);  // hi must outlive the copy
        keys_out.release();
        iota.release();
        tmp.release();
        if (e != cudaSuccess) {
            sqb_set_error("autocorr load: %s", cudaGetErrorString(e));
            return SQB_ERR_CUDA;
        }
        c->launches += 2;
    }
    // the padded starts now belong to the handle
    SQB_TRY(h->d_xp.alloc((size_t)n_feat + 1));
    SQB_CUDA(cudaMemcpyAsync(h->d_xp.p, d_pad, (n_feat + 1) * sizeof(int64_t), cudaMemcpyDeviceToDevice, c->stream));
    return SQB_OK;
}

template <typename XT>
static int ac_load_csr_typed(sqb_autocorr* h, const int64_t* h_xp, const int32_t* h_xi, const void* h_xv, int layout,
                             int64_t n_feat, int64_t nnz, int32_t col_lo = 0, const int64_t* sub_start = nullptr,
                             const int64_t* sub_cnt = nullptr) {
    // sub_start / sub_cnt (layout 1 only): upload only the piece [sub_start[r], +sub_cnt[r]) of every observation's row (the
    // columns col_lo .. col_lo + n_feat of rows with ascending indices); h_xp then is the indptr of the pieces, nnz their total
    sqb_ctx* c = h->ctx;
    const int64_t n = h->n;
    const int64_t rows = layout == 0 ? n_feat : n;
    const int64_t nz1 = nnz > 0 ? nnz : 1;
    DevBuf<int64_t> r_ptr, t_ptr, pad;
    DevBuf<int32_t> r_idx, t_idx;
    DevBuf<uint8_t> r_val, t_val;
    DevBuf<unsigned int> cnt;
    DevBuf<unsigned long long> cursor;
    r_ptr.bind(c->stream), t_ptr.bind(c->stream), pad.bind(c->stream), r_idx.bind(c->stream), t_idx.bind(c->stream);
    r_val.bind(c->stream), t_val.bind(c->stream), cnt.bind(c->stream), cursor.bind(c->stream);
    auto cleanup = [&]() {
        r_ptr.release(), t_ptr.release(), pad.release(), r_idx.release(), t_idx.release(), r_val.release(), t_val.release();
        cnt.release(), cursor.release();
    };
    int rc;
    if ((rc = r_ptr.alloc(rows + 1)) || (rc = r_idx.alloc(nz1)) || (rc = r_val.alloc((size_t)nz1 * sizeof(XT))) ||
        (rc = cnt.alloc(n_feat + 1)) || (rc = pad.alloc(n_feat + 1)) || (rc = h->d_flags.alloc(2))) {
        cleanup();
        return rc;
    }
    auto fail = [&](cudaError_t e) {
        cleanup();
        sqb_set_error("autocorr load: %s", cudaGetErrorString(e));
        return SQB_ERR_CUDA;
    };
    cudaError_t e = cudaMemsetAsync(h->d_flags.p, 0, 2 * sizeof(int), c->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(r_ptr.p, h_xp, (rows + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, c->stream);
    if (e != cudaSuccess) return fail(e);
    if (nnz > 0) {
        if (sub_start)
            rc = sqb_h2d_gather(c, r_idx.p, h_xi, sizeof(int32_t), sub_start, sub_cnt, h_xp, rows) ||
                 sqb_h2d_gather(c, r_val.p, h_xv, sizeof(XT), sub_start, sub_cnt, h_xp, rows);
        else
            rc = sqb_h2d(c, r_idx.p, h_xi, nnz * sizeof(int32_t)) || sqb_h2d(c, r_val.p, h_xv, (size_t)nnz * sizeof(XT));
        if (rc) {
            cleanup();
            return SQB_ERR_CUDA;
        }
    }
    {
        SqbLaunchScope scope(c, SQB_K_AUTOCORR_PREP);
        ac_check_indptr_kernel<<<c->sm_count * 4, 256, 0, c->stream>>>(r_ptr.p, rows, nnz, h->d_flags.p);
    }
    const int64_t* seg_ptr;
    const int32_t* seg_idx;
    const XT* seg_val;
    if (layout == 0) {
        SqbLaunchScope scope(c, SQB_K_AUTOCORR_PREP);
        ac_seglen_kernel<<<c->sm_count * 2, 256, 0, c->stream>>>(r_ptr.p, n_feat, n, cnt.p, h->d_flags.p);
        seg_ptr = r_ptr.p;
        seg_idx = r_idx.p;
        seg_val = reinterpret_cast<const XT*>(r_val.p);
    } else {
        if ((rc = t_ptr.alloc(n_feat + 1)) || (rc = t_idx.alloc(nz1)) || (rc = t_val.alloc((size_t)nz1 * sizeof(XT))) ||
            (rc = cursor.alloc(n_feat + 1))) {
            cleanup();
            return rc;
        }
        e = cudaMemsetAsync(cnt.p, 0, (n_feat + 1) * sizeof(unsigned int), c->stream);
        if (e != cudaSuccess) return fail(e);
        {
            SqbLaunchScope scope(c, SQB_K_AUTOCORR_PREP);
            const size_t hist_bytes = (size_t)n_feat * sizeof(unsigned int);
            const int hist_smem = hist_bytes <= 160 * 1024 ? 1 : 0;
            if (hist_smem && hist_bytes > 48 * 1024)
                cudaFuncSetAttribute(ac_colcount_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hist_bytes);
            ac_colcount_kernel<<<c->sm_count * (hist_smem ? 1 : 8), 1024, hist_smem ? hist_bytes : 0, c->stream>>>(r_idx.p, nnz, n_feat, col_lo,
                                                                                                              hist_smem, cnt.p, h->d_flags.p);
        }
    }
    {
        SqbLaunchScope scope(c, SQB_K_AUTOCORR_PREP);
        // layout 0: the caller's indptr already is the plain prefix, only the padded starts are needed
        ac_scan_counts_kernel<<<1, 1024, 0, c->stream>>>(cnt.p, n_feat, layout == 0 ? nullptr : t_ptr.p, pad.p,
                                                         layout == 0 ? nullptr : cursor.p);
    }
    if (layout != 0) {
        SqbLaunchScope scope(c, SQB_K_AUTOCORR_PREP);
        ac_coltranspose_kernel<XT><<<c->sm_count * 8, 256, 0, c->stream>>>(r_ptr.p, r_idx.p, reinterpret_cast<const XT*>(r_val.p), n, n_feat,
                                                                           col_lo, cursor.p, t_idx.p, reinterpret_cast<XT*>(t_val.p));
        seg_ptr = t_ptr.p;
        seg_idx = t_idx.p;
        seg_val = reinterpret_cast<const XT*>(t_val.p);
    }
    e = cudaGetLastError();
    if (e != cudaSuccess) return fail(e);
    rc = ac_build_segments<XT>(h, seg_ptr, cnt.p, pad.p, seg_idx, seg_val, nnz, n_feat);
    if (rc == SQB_OK) rc = ac_check_flags(h, "sqb_autocorr_load_csr");  // synchronises: the temporaries may go
    else cudaStreamSynchronize(c->stream);
    cleanup();
    h->x_nnz = nnz;
    return rc;
}

extern "C" {

int sqb_autocorr_create(sqb_ctx* ctx, int64_t n, int64_t nnz, const int32_t* w_indptr, const int32_t* w_indices,
                        const void* w_data, int w_dtype, sqb_autocorr** out) {
    SQB_CHECK(ctx && out && w_indptr, SQB_ERR_INVALID, "sqb_autocorr_create: null argument");
    SQB_CHECK(n >= 2 && n < 2147483647LL, SQB_ERR_INVALID, "sqb_autocorr_create: n=%lld out of range", (long long)n);
    SQB_CHECK(nnz >= 0 && nnz < 2147483647LL, SQB_ERR_INVALID, "sqb_autocorr_create: nnz=%lld does not fit int32", (long long)nnz);
    SQB_CHECK(w_dtype == 0 || w_dtype == 1, SQB_ERR_INVALID, "sqb_autocorr_create: w_dtype must be 0 (f32) or 1 (f64)");
    SQB_CHECK(w_indptr[0] == 0 && (int64_t)w_indptr[n] == nnz, SQB_ERR_INVALID, "sqb_autocorr_create: inconsistent indptr");
    SQB_CHECK(nnz == 0 || (w_indices && w_data), SQB_ERR_INVALID, "sqb_autocorr_create: null W arrays");
    // one host pass over W (it is small next to X): validation, float64 weights (scanpy casts W.data to float64), S0 =
    // sum(W.data), column sums in CSR order (deterministic), longest row
    int64_t maxdeg = 0;
    for (int64_t r = 0; r < n; ++r) {
        const int64_t d = (int64_t)w_indptr[r + 1] - (int64_t)w_indptr[r];
        SQB_CHECK(d >= 0 && w_indptr[r + 1] <= nnz, SQB_ERR_INVALID, "sqb_autocorr_create: indptr decreases at row %lld", (long long)r);
        if (d > maxdeg) maxdeg = d;
    }
    std::vector<double> wd((size_t)(nnz > 0 ? nnz : 1)), csum((size_t)n, 0.0);
    double s0 = 0.0;
    for (int64_t e = 0; e < nnz; ++e) {
        const int32_t j = w_indices[e];
        SQB_CHECK(j >= 0 && j < n, SQB_ERR_INVALID, "sqb_autocorr_create: column index %d out of range at entry %lld", j, (long long)e);
        wd[e] = w_dtype == 0 ? (double)((const float*)w_data)[e] : ((const double*)w_data)[e];
        s0 += wd[e];
        csum[j] += wd[e];
    }
    // rows of at most 7 entries that all carry the same float32 weight (a row-normalised binary graph) take the compact format
    bool uniform7 = (w_dtype == 0 && maxdeg <= 7 && nnz > 0);
    if (uniform7) {
        const uint32_t* wb = reinterpret_cast<const uint32_t*>(w_data);
        for (int64_t r = 0; r < n && uniform7; ++r)
            for (int64_t e = (int64_t)w_indptr[r] + 1; e < (int64_t)w_indptr[r + 1]; ++e)
                if (wb[e] != wb[w_indptr[r]]) {
                    uniform7 = false;
                    break;
                }
    }
    SQB_CUDA(cudaSetDevice(ctx->device));
    sqb_autocorr* h = new sqb_autocorr();
    h->ctx = ctx;
    {  // stream-ordered allocation from the device pool: a handle is created and destroyed per API call, and GB-sized
       // cudaMalloc / cudaFree pairs would synchronise the device every time
        cudaStream_t st = ctx->stream;
        h->d_wp.bind(st), h->d_wi.bind(st), h->d_wd.bind(st), h->d_csum.bind(st), h->d_rows.bind(st), h->d_x.bind(st), h->d_xp.bind(st);
        h->d_xi.bind(st), h->d_len.bind(st), h->d_order.bind(st), h->d_tile.bind(st), h->d_sums.bind(st), h->d_partial.bind(st);
        h->d_pnum.bind(st), h->d_pden.bind(st), h->d_out.bind(st), h->d_aux.bind(st), h->d_out_perms.bind(st), h->d_perm.bind(st);
        h->d_perm32.bind(st), h->d_seen.bind(st), h->d_bitmap.bind(st), h->d_flags.bind(st);
    }
    h->n = n;
    h->nnz = nnz;
    h->s0 = s0;
    // packed rows hold float32 weights: only for float32 W (lossless); float64 W walks the CSR
    h->w_fmt = uniform7 ? 3 : (w_dtype == 0 && maxdeg <= 8) ? 0 : (w_dtype == 0 && maxdeg <= 16) ? 1 : 2;
    const int lpr = h->w_fmt == 3 ? 4 : h->w_fmt == 0 ? 8 : 16;  // uint2 units per row
    int rc;
    if ((rc = h->d_wp.alloc(n + 1)) || (rc = h->d_wi.alloc(nnz > 0 ? nnz : 1)) || (rc = h->d_wd.alloc(nnz > 0 ? nnz : 1)) ||
        (rc = h->d_csum.alloc(n)) || (rc = h->d_flags.alloc(2)) || (h->w_fmt != 2 && (rc = h->d_rows.alloc((size_t)n * lpr)))) {
        sqb_autocorr_destroy(h);
        return rc;
    }
    cudaError_t e = cudaMemcpyAsync(h->d_wp.p, w_indptr, (n + 1) * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess && nnz > 0) e = cudaMemcpyAsync(h->d_wi.p, w_indices, nnz * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess && nnz > 0) e = cudaMemcpyAsync(h->d_wd.p, wd.data(), nnz * sizeof(double), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(h->d_csum.p, csum.data(), n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(h->d_flags.p, 0, 2 * sizeof(int), ctx->stream);
    if (e == cudaSuccess && h->w_fmt != 2) {
        ctx->launches += 1;
        if (h->w_fmt == 3)
            ac_pack_rows32_kernel<<<ctx->sm_count * 4, 256, 0, ctx->stream>>>(h->d_wp.p, h->d_wi.p, h->d_wd.p, n,
                                                                             reinterpret_cast<uint32_t*>(h->d_rows.p));
        else if (h->w_fmt == 0)
            ac_pack_rows_kernel<8><<<ctx->sm_count * 4, 256, 0, ctx->stream>>>(h->d_wp.p, h->d_wi.p, h->d_wd.p, n, h->d_rows.p);
        else
            ac_pack_rows_kernel<16><<<ctx->sm_count * 4, 256, 0, ctx->stream>>>(h->d_wp.p, h->d_wi.p, h->d_wd.p, n, h->d_rows.p);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);  // wd / csum are host temporaries
    if (e != cudaSuccess) {
        sqb_autocorr_destroy(h);
        sqb_set_error("sqb_autocorr_create: %s", cudaGetErrorString(e));
        return SQB_ERR_CUDA;
    }
    *out = h;
    return SQB_OK;
}

int sqb_autocorr_destroy(sqb_autocorr* h) {
    if (!h) return SQB_OK;
    cudaSetDevice(h->ctx->device);
    h->d_wp.release();
    h->d_wi.release();
    h->d_wd.release();
    h->d_csum.release();
    h->d_rows.release();
    h->d_x.release();
    h->d_xp.release();
    h->d_xi.release();
    h->d_len.release();
    h->d_order.release();
    h->d_tile.release();
    h->d_sums.release();
    h->d_partial.release();
    h->d_pnum.release();
    h->d_pden.release();
    h->d_out.release();
    h->d_aux.release();
    h->d_out_perms.release();
    h->d_perm.release();
    h->d_perm32.release();
    h->d_seen.release();
    h->d_bitmap.release();
    h->d_flags.release();
    delete h;
    return SQB_OK;
}

int sqb_autocorr_load_dense(sqb_autocorr* h, const void* x, int x_dtype, int layout, int64_t n_features) {
    SQB_CHECK(h && x, SQB_ERR_INVALID, "sqb_autocorr_load_dense: null argument");
    SQB_CHECK(x_dtype == 0 || x_dtype == 1, SQB_ERR_INVALID, "sqb_autocorr_load_dense: x_dtype must be 0 or 1");
    SQB_CHECK(layout == 0 || layout == 1, SQB_ERR_INVALID, "sqb_autocorr_load_dense: layout must be 0 or 1");
    SQB_CHECK(n_features >= 1, SQB_ERR_INVALID, "sqb_autocorr_load_dense: n_features must be positive");
    sqb_ctx* c = h->ctx;
    SQB_CUDA(cudaSetDevice(c->device));
    const size_t bytes = (size_t)n_features * h->n * xsize(x_dtype);
    // zero-copy tiles read up to 31 elements past the last feature of a row: pad the allocation
    SQB_TRY(h->d_x.alloc(bytes + 64 * 8));
    SQB_TRY(h->d_out.alloc((size_t)n_features));
    SQB_TRY(sqb_h2d(c, h->d_x.p, x, bytes));
    SQB_CUDA(cudaStreamSynchronize(c->stream));
    h->kind = layout == 0 ? 1 : 2;
    h->x_dtype = x_dtype;
    h->n_feat = n_features;
    h->ran = false;
    h->aux_valid = false;
    return SQB_OK;
}

int sqb_autocorr_load_csr(sqb_autocorr* h, const int64_t* x_indptr, const int32_t* x_indices, const void* x_data, int x_dtype,
                          int layout, int64_t n_features) {
    SQB_CHECK(h && x_indptr, SQB_ERR_INVALID, "sqb_autocorr_load_csr: null argument");
    SQB_CHECK(x_dtype == 0 || x_dtype == 1, SQB_ERR_INVALID, "sqb_autocorr_load_csr: x_dtype must be 0 or 1");
    SQB_CHECK(layout == 0 || layout == 1, SQB_ERR_INVALID, "sqb_autocorr_load_csr: layout must be 0 or 1");
    SQB_CHECK(n_features >= 1 && n_features < 2147483647LL, SQB_ERR_INVALID, "sqb_autocorr_load_csr: n_features out of range");
    sqb_ctx* c = h->ctx;
    SQB_CUDA(cudaSetDevice(c->device));
    const int64_t rows = layout == 0 ? n_features : h->n;
    const int64_t nnz = x_indptr[rows];
    SQB_CHECK(x_indptr[0] == 0 && nnz >= 0, SQB_ERR_INVALID, "sqb_autocorr_load_csr: bad indptr");
    SQB_CHECK(nnz == 0 || (x_indices && x_data), SQB_ERR_INVALID, "sqb_autocorr_load_csr: null X arrays");
    h->kind = 0;
    h->aux_valid = false;
    h->ran = false;
    // index range, indptr monotonicity and duplicate observations are checked on the device while the matrix is
    // re-laid out (no host pass over the non-zeros)
    if (x_dtype == 0)
        SQB_TRY(ac_load_csr_typed<float>(h, x_indptr, x_indices, x_data, layout, n_features, nnz));
    else
        SQB_TRY(ac_load_csr_typed<double>(h, x_indptr, x_indices, x_data, layout, n_features, nnz));
    h->kind = 3;
    h->x_dtype = x_dtype;
    h->n_feat = n_features;
    return SQB_OK;
}

int sqb_autocorr_load_csr_cols(sqb_autocorr* h, const int64_t* x_indptr, const int32_t* x_indices, const void* x_data, int x_dtype,
                               int64_t n_features_total, int64_t col_lo, int64_t col_hi) {
    SQB_CHECK(h && x_indptr, SQB_ERR_INVALID, "sqb_autocorr_load_csr_cols: null argument");
    SQB_CHECK(x_dtype == 0 || x_dtype == 1, SQB_ERR_INVALID, "sqb_autocorr_load_csr_cols: x_dtype must be 0 or 1");
    SQB_CHECK(0 <= col_lo && col_lo < col_hi && col_hi <= n_features_total && n_features_total < 2147483647LL, SQB_ERR_INVALID,
              "sqb_autocorr_load_csr_cols: bad column range [%lld, %lld) of %lld", (long long)col_lo, (long long)col_hi,
              (long long)n_features_total);
    sqb_ctx* c = h->ctx;
    SQB_CUDA(cudaSetDevice(c->device));
    const int64_t n = h->n;
    const int64_t nnz_all = x_indptr[n];
    SQB_CHECK(x_indptr[0] == 0 && nnz_all >= 0, SQB_ERR_INVALID, "sqb_autocorr_load_csr_cols: bad indptr");
    SQB_CHECK(nnz_all == 0 || (x_indices && x_data), SQB_ERR_INVALID, "sqb_autocorr_load_csr_cols: null X arrays");
    // per observation: the run of entries with col_lo <= column < col_hi (rows hold ascending columns: two binary searches)
    std::vector<int64_t> start((size_t)n), cnt((size_t)n), ptr((size_t)n + 1);
    {
        int threads = (int)std::thread::hardware_concurrency() / 8;
        threads = threads < 2 ? 2 : (threads > 16 ? 16 : threads);
        std::vector<std::thread> pool;
        std::vector<int> bad((size_t)threads, 0);
        for (int t = 0; t < threads; ++t) {
            pool.emplace_back([&, t]() {
                for (int64_t r = n * t / threads; r < n * (t + 1) / threads; ++r) {
                    const int64_t a = x_indptr[r], b = x_indptr[r + 1];
                    if (a > b || a < 0 || b > nnz_all) {
                        bad[t] = 1;
                        start[r] = 0;
                        cnt[r] = 0;
                        continue;
                    }
                    const int32_t* lo = std::lower_bound(x_indices + a, x_indices + b, (int32_t)col_lo);
                    const int32_t* hi = std::lower_bound(lo, x_indices + b, (int32_t)col_hi);
                    start[r] = lo - x_indices;
                    cnt[r] = hi - lo;
                }
            });
        }
        for (auto& th : pool) th.join();
        for (int t = 0; t < threads; ++t) SQB_CHECK(!bad[t], SQB_ERR_INVALID, "sqb_autocorr_load_csr_cols: indptr is not non-decreasing");
    }
    ptr[0] = 0;
    for (int64_t r = 0; r < n; ++r) ptr[r + 1] = ptr[r] + cnt[r];
    h->kind = 0;
    h->aux_valid = false;
    h->ran = false;
    // the device checks that every uploaded column lies in [col_lo, col_hi) (an unsorted row shows up there or as a
    // duplicate) while it re-lays the slice out
    const int64_t nf = col_hi - col_lo;
    if (x_dtype == 0)
        SQB_TRY(ac_load_csr_typed<float>(h, ptr.data(), x_indices, x_data, 1, nf, ptr[n], (int32_t)col_lo, start.data(), cnt.data()));
    else
        SQB_TRY(ac_load_csr_typed<double>(h, ptr.data(), x_indices, x_data, 1, nf, ptr[n], (int32_t)col_lo, start.data(), cnt.data()));
    h->kind = 3;
    h->x_dtype = x_dtype;
    h->n_feat = nf;
    return SQB_OK;
}

// upload + validate one or more row permutations (int64 host layout); d_perm (int64) and d_perm32 hold them afterwards
static int ac_upload_perms(sqb_autocorr* h, const int64_t* row_perm, int64_t count) {
    sqb_ctx* c = h->ctx;
    const int64_t n = h->n;
    const int64_t words = ceil_div64(n, 32);
    SQB_TRY(h->d_perm.alloc((size_t)(n * count)));
    SQB_TRY(h->d_perm32.alloc((size_t)(n * count)));
    SQB_TRY(h->d_seen.alloc((size_t)(words * count)));
    SQB_TRY(sqb_h2d(c, h->d_perm.p, row_perm, (size_t)(n * count) * sizeof(int64_t)));
    SQB_CUDA(cudaMemsetAsync(h->d_seen.p, 0, (size_t)(words * count) * sizeof(unsigned int), c->stream));
    for (int64_t k = 0; k < count; ++k) {
        SqbLaunchScope scope(c, SQB_K_AUTOCORR_PREP);
        ac_perm_prepare_kernel<<<c->sm_count * 2, 256, 0, c->stream>>>(h->d_perm.p + k * n, n, h->d_perm32.p + k * n, h->d_seen.p + k * words,
                                                                       h->d_flags.p);
        SQB_POST_LAUNCH();
    }
    return SQB_OK;
}

static int ac_run_one(sqb_autocorr* h, int mode, int64_t perm_index /* -1: none */, double* d_out) {
    if (h->kind == 3) {
        if (perm_index >= 0 && !h->aux_valid) SQB_TRY(ac_run_sparse(h, mode, nullptr, h->d_out.p));
        return ac_run_sparse(h, mode, perm_index >= 0 ? h->d_perm32.p + perm_index * h->n : nullptr, d_out);
    }
    const int64_t* dp = perm_index >= 0 ? h->d_perm.p + perm_index * h->n : nullptr;
    SQB_TRY(h->x_dtype == 0 ? ac_run_dense<float>(h, mode, dp) : ac_run_dense<double>(h, mode, dp));
    if (d_out != h->d_out.p)
        SQB_CUDA(cudaMemcpyAsync(d_out, h->d_out.p, h->n_feat * sizeof(double), cudaMemcpyDeviceToDevice, h->ctx->stream));
    return SQB_OK;
}

int sqb_autocorr_run_async(sqb_autocorr* h, int mode, const int64_t* row_perm) {
    SQB_CHECK(h, SQB_ERR_INVALID, "sqb_autocorr_run_async: null handle");
    SQB_CHECK(h->kind != 0, SQB_ERR_STATE, "sqb_autocorr_run_async: load X first");
    SQB_CHECK(mode == 0 || mode == 1, SQB_ERR_INVALID, "sqb_autocorr_run_async: mode must be 0 (moran) or 1 (geary)");
    sqb_ctx* c = h->ctx;
    SQB_CUDA(cudaSetDevice(c->device));
    if (row_perm) {
        SQB_TRY(ac_upload_perms(h, row_perm, 1));
        SQB_TRY(ac_check_flags(h, "sqb_autocorr_run_async"));
    }
    int rc = ac_run_one(h, mode, row_perm ? 0 : -1, h->d_out.p);
    if (rc == SQB_OK) h->ran = true;
    return rc;
}

int sqb_autocorr_download(sqb_autocorr* h, double* out) {
    SQB_CHECK(h && out, SQB_ERR_INVALID, "sqb_autocorr_download: null argument");
    SQB_CHECK(h->ran, SQB_ERR_STATE, "sqb_autocorr_download: nothing has run");
    sqb_ctx* c = h->ctx;
    SQB_CUDA(cudaSetDevice(c->device));
    SQB_CUDA(cudaMemcpyAsync(out, h->d_out.p, h->n_feat * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    SQB_CUDA(cudaStreamSynchronize(c->stream));
    return SQB_OK;
}

int sqb_autocorr_run_perms(sqb_autocorr* h, int mode, const int64_t* row_perms, int64_t n_perms, double* out) {
    SQB_CHECK(h && row_perms && out, SQB_ERR_INVALID, "sqb_autocorr_run_perms: null argument");
    SQB_CHECK(h->kind != 0, SQB_ERR_STATE, "sqb_autocorr_run_perms: load X first");
    SQB_CHECK(mode == 0 || mode == 1, SQB_ERR_INVALID, "sqb_autocorr_run_perms: mode must be 0 (moran) or 1 (geary)");
    SQB_CHECK(n_perms >= 1, SQB_ERR_INVALID, "sqb_autocorr_run_perms: n_perms must be positive");
    sqb_ctx* c = h->ctx;
    SQB_CUDA(cudaSetDevice(c->device));
    const int64_t G = h->n_feat;
    // permutations are staged in batches of <= 32 (8 bytes + 4 bytes + 1 bit per entry on the device)
    const int64_t batch = n_perms < 32 ? n_perms : 32;
    SQB_TRY(h->d_out_perms.alloc((size_t)(batch * G)));
    for (int64_t p0 = 0; p0 < n_perms; p0 += batch) {
        const int64_t pb = n_perms - p0 < batch ? n_perms - p0 : batch;
        SQB_TRY(ac_upload_perms(h, row_perms + p0 * h->n, pb));
        if (h->kind == 3) {  // sparse X: one launch scores the whole batch (the feature's bitmap is built once per batch)
            if (!h->aux_valid) SQB_TRY(ac_run_sparse(h, mode, nullptr, h->d_out.p));
            SQB_TRY(ac_run_sparse(h, mode, h->d_perm32.p, h->d_out_perms.p, (int)pb));
        } else {
            for (int64_t k = 0; k < pb; ++k) SQB_TRY(ac_run_one(h, mode, k, h->d_out_perms.p + k * G));
        }
        SQB_CUDA(cudaMemcpyAsync(out + p0 * G, h->d_out_perms.p, (size_t)(pb * G) * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        SQB_TRY(ac_check_flags(h, "sqb_autocorr_run_perms"));  // synchronises
    }
    return SQB_OK;
}

int sqb_autocorr_dense(sqb_autocorr* h, int mode, const void* x, int x_dtype, int layout, int64_t n_features,
                       const int64_t* row_perm, double* out) {
    SQB_TRY(sqb_autocorr_load_dense(h, x, x_dtype, layout, n_features));
    SQB_TRY(sqb_autocorr_run_async(h, mode, row_perm));
    return sqb_autocorr_download(h, out);
}

int sqb_autocorr_csr(sqb_autocorr* h, int mode, const int64_t* x_indptr, const int32_t* x_indices, const void* x_data,
                     int x_dtype, int layout, int64_t n_features, const int64_t* row_perm, double* out) {
    SQB_TRY(sqb_autocorr_load_csr(h, x_indptr, x_indices, x_data, x_dtype, layout, n_features));
    SQB_TRY(sqb_autocorr_run_async(h, mode, row_perm));
    return sqb_autocorr_download(h, out);
}

}  // extern "C"
