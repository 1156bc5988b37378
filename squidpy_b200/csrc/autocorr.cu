// autocorr.cu — Moran's I / Geary's C for every feature over a sparse weight matrix W (sm_100a).
//
// Replaces the scanpy.metrics.morans_i / gearys_c call of the reference (src/squidpy/gr/_ppatterns.py:216 and,
// per permutation, :267-272).  Arithmetic follows scanpy's published kernels, float64 throughout:
//     I = N/S0 * sum_i z_i * sum_j w_ij z_j / sum_i z_i^2,   z = x - mean(x)
//     C = (N-1) * sum_ij w_ij (x_i - x_j)^2 / (2 * S0 * sum_i (x_i - mean)^2)
// Layout: features are processed in tiles of 32 (one warp lane per feature).  A tile is a dense [N][32]
// slab in HBM/L2 (feature-minor), so that the SpMM-style main kernel reads one coalesced 128-byte (f32) row
// per observation and gathers the neighbour rows through L1/L2.  W is streamed once per tile; X once per
// call.  Reductions are fixed-order (per-warp partials over contiguous observation ranges, then a
// sequential combine), so results are bit-reproducible run to run.
#include "common.cuh"

#define TILE 32

// ---- CSR by feature -> dense tile ------------------------------------------------------------------
// one CTA (256 threads) per feature of the super-tile: scatter the feature's non-zeros into D[tile][obs][lane] and
// reduce its sum (for the mean) in a fixed order (per-thread strided partial, warp shuffle tree, then warp 0).
template <typename XT>
__global__ void __launch_bounds__(256) ac_scatter_kernel(const int64_t* __restrict__ xp, const int32_t* __restrict__ xi,
                                                         const XT* __restrict__ xv, int64_t g0, int64_t n_feat, int64_t n,
                                                         XT* __restrict__ D, double* __restrict__ sums) {
    __shared__ double s_part[8];
    const int64_t wglobal = blockIdx.x;  // feature index inside the super-tile
    const int64_t g = g0 + wglobal;
    if (g >= n_feat) return;
    const int64_t tile = wglobal / TILE;
    const int t = (int)(wglobal % TILE);
    XT* __restrict__ Dt = D + tile * n * TILE;
    double s = 0.0;
    for (int64_t e = xp[g] + threadIdx.x; e < xp[g + 1]; e += 256) {
        const XT v = xv[e];
        Dt[(int64_t)xi[e] * TILE + t] = v;
        s += (double)v;
    }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double tot = 0.0;
        for (int w = 0; w < 8; ++w) tot += s_part[w];
        sums[wglobal] = tot;
    }
}

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

// ---- main SpMM-style kernel ------------------------------------------------------------------------------
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
    const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
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

// ---- CSR by observation -> CSR by feature (device transposition; counting sort by column) ----------------
__global__ void ac_colcount_kernel(const int32_t* __restrict__ xi, int64_t nnz, unsigned long long* __restrict__ cnt) {
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < nnz; e += (int64_t)gridDim.x * blockDim.x)
        atomicAdd(&cnt[xi[e]], 1ULL);
}

template <typename XT>
__global__ void ac_coltranspose_kernel(const int64_t* __restrict__ xp, const int32_t* __restrict__ xi,
                                       const XT* __restrict__ xv, int64_t n_obs, unsigned long long* __restrict__ cursor,
                                       int32_t* __restrict__ oi, XT* __restrict__ ov) {
    const int lane = threadIdx.x & 31;
    const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5; r < n_obs; r += nw) {
        for (int64_t e = xp[r] + lane; e < xp[r + 1]; e += 32) {
            const unsigned long long pos = atomicAdd(&cursor[xi[e]], 1ULL);
            oi[pos] = (int32_t)r;
            ov[pos] = xv[e];
        }
    }
}

// ================================================================================================
struct sqb_autocorr {
    sqb_ctx* ctx = nullptr;
    int64_t n = 0, nnz = 0;
    double s0 = 0.0;
    DevBuf<int32_t> d_wp, d_wi;
    DevBuf<double> d_wd;
    // loaded X
    int kind = 0;     // 0 none, 1 dense feat x obs, 2 dense obs x feat (zero copy tiles), 3 CSR by feature
    int x_dtype = 0;  // 0 f32, 1 f64
    int64_t n_feat = 0, x_nnz = 0;
    DevBuf<uint8_t> d_x;   // dense matrix or CSR values
    DevBuf<int64_t> d_xp;
    DevBuf<int32_t> d_xi;
    DevBuf<uint8_t> d_tile;
    DevBuf<double> d_sums, d_partial, d_pnum, d_pden, d_out;
    DevBuf<int64_t> d_perm;
    int tiles_per_launch = 8;  // 256 features per launch: enough CTAs to fill 148 SMs in every phase
    bool ran = false;
};

static int xsize(int dt) { return dt == 0 ? 4 : 8; }

template <typename XT>
static int ac_run_typed(sqb_autocorr* h, int mode, const int64_t* d_perm) {
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
    SQB_TRY(h->d_out.alloc((size_t)G));
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
        }
        if (h->kind == 3) {
            SQB_CUDA(cudaMemsetAsync(Dbuf, 0, (size_t)nt * n * TILE * sizeof(XT), c->stream));
            SqbLaunchScope scope(c, SQB_K_AUTOCORR_PREP);
            ac_scatter_kernel<XT><<<(unsigned)(nt * TILE), 256, 0, c->stream>>>(h->d_xp.p, h->d_xi.p, X, g0, G, n, Dbuf,
                                                                                  h->d_sums.p);
            SQB_POST_LAUNCH();
        } else {
            if (h->kind == 1) {
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

template <typename XT>
static int ac_transpose_csr(sqb_autocorr* h, const int64_t* h_xp, const int32_t* h_xi, const void* h_xv, int64_t n_feat) {
    // input: CSR by observation (n rows, n_feat columns) on the host -> CSR by feature on the device
    sqb_ctx* c = h->ctx;
    const int64_t n = h->n, nnz = h_xp[n];
    DevBuf<int64_t> t_xp;
    DevBuf<int32_t> t_xi;
    DevBuf<uint8_t> t_xv;
    DevBuf<unsigned long long> cnt;
    int rc;
    auto cleanup = [&]() {
        t_xp.release();
        t_xi.release();
        t_xv.release();
        cnt.release();
    };
    if ((rc = t_xp.alloc(n + 1)) || (rc = t_xi.alloc(nnz > 0 ? nnz : 1)) || (rc = t_xv.alloc((size_t)(nnz > 0 ? nnz : 1) * sizeof(XT))) ||
        (rc = cnt.alloc(n_feat + 1)) || (rc = h->d_xp.alloc(n_feat + 1)) || (rc = h->d_xi.alloc(nnz > 0 ? nnz : 1)) ||
        (rc = h->d_x.alloc((size_t)(nnz > 0 ? nnz : 1) * sizeof(XT)))) {
        cleanup();
        return rc;
    }
    cudaError_t e = cudaMemcpyAsync(t_xp.p, h_xp, (n + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess && nnz > 0 && sqb_h2d(c, t_xi.p, h_xi, nnz * sizeof(int32_t)) != SQB_OK) e = cudaErrorUnknown;
    if (e == cudaSuccess && nnz > 0 && sqb_h2d(c, t_xv.p, h_xv, nnz * sizeof(XT)) != SQB_OK) e = cudaErrorUnknown;
    if (e == cudaSuccess) e = cudaMemsetAsync(cnt.p, 0, (n_feat + 1) * sizeof(unsigned long long), c->stream);
    if (e != cudaSuccess) {
        cleanup();
        sqb_set_error("autocorr load: %s", cudaGetErrorString(e));
        return SQB_ERR_CUDA;
    }
    {
        SqbLaunchScope scope(c, SQB_K_AUTOCORR_PREP);
        ac_colcount_kernel<<<c->sm_count * 8, 256, 0, c->stream>>>(t_xi.p, nnz, cnt.p);
    }
    std::vector<unsigned long long> hc(n_feat + 1);
    e = cudaMemcpyAsync(hc.data(), cnt.p, (n_feat + 1) * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    if (e != cudaSuccess) {
        cleanup();
        sqb_set_error("autocorr load: %s", cudaGetErrorString(e));
        return SQB_ERR_CUDA;
    }
    std::vector<int64_t> ptr(n_feat + 1);
    std::vector<unsigned long long> cur(n_feat + 1);
    int64_t run = 0;
    for (int64_t g = 0; g < n_feat; ++g) {
        ptr[g] = run;
        cur[g] = (unsigned long long)run;
        run += (int64_t)hc[g];
    }
    ptr[n_feat] = run;
    cur[n_feat] = (unsigned long long)run;
    e = cudaMemcpyAsync(h->d_xp.p, ptr.data(), (n_feat + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(cnt.p, cur.data(), (n_feat + 1) * sizeof(unsigned long long), cudaMemcpyHostToDevice, c->stream);
    if (e != cudaSuccess) {
        cleanup();
        sqb_set_error("autocorr load: %s", cudaGetErrorString(e));
        return SQB_ERR_CUDA;
    }
    {
        SqbLaunchScope scope(c, SQB_K_AUTOCORR_PREP);
        ac_coltranspose_kernel<XT><<<c->sm_count * 8, 256, 0, c->stream>>>(t_xp.p, t_xi.p, reinterpret_cast<const XT*>(t_xv.p), n,
                                                                           cnt.p, h->d_xi.p, reinterpret_cast<XT*>(h->d_x.p));
    }
    e = cudaStreamSynchronize(c->stream);
    cleanup();
    if (e != cudaSuccess) {
        sqb_set_error("autocorr load: %s", cudaGetErrorString(e));
        return SQB_ERR_CUDA;
    }
    h->x_nnz = nnz;
    return SQB_OK;
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
    SQB_CUDA(cudaSetDevice(ctx->device));
    sqb_autocorr* h = new sqb_autocorr();
    h->ctx = ctx;
    h->n = n;
    h->nnz = nnz;
    std::vector<double> wd((size_t)(nnz > 0 ? nnz : 1));
    double s0 = 0.0;
    for (int64_t e = 0; e < nnz; ++e) {
        wd[e] = w_dtype == 0 ? (double)((const float*)w_data)[e] : ((const double*)w_data)[e];
        s0 += wd[e];  // S0 = sum(W.data) in float64 (scanpy casts W.data to float64 first)
    }
    h->s0 = s0;
    int rc;
    if ((rc = h->d_wp.alloc(n + 1)) || (rc = h->d_wi.alloc(nnz > 0 ? nnz : 1)) || (rc = h->d_wd.alloc(nnz > 0 ? nnz : 1))) {
        sqb_autocorr_destroy(h);
        return rc;
    }
    SQB_CUDA(cudaMemcpyAsync(h->d_wp.p, w_indptr, (n + 1) * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
    if (nnz > 0) {
        SQB_CUDA(cudaMemcpyAsync(h->d_wi.p, w_indices, nnz * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
        SQB_CUDA(cudaMemcpyAsync(h->d_wd.p, wd.data(), nnz * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    }
    SQB_CUDA(cudaStreamSynchronize(ctx->stream));
    *out = h;
    return SQB_OK;
}

int sqb_autocorr_destroy(sqb_autocorr* h) {
    if (!h) return SQB_OK;
    cudaSetDevice(h->ctx->device);
    h->d_wp.release();
    h->d_wi.release();
    h->d_wd.release();
    h->d_x.release();
    h->d_xp.release();
    h->d_xi.release();
    h->d_tile.release();
    h->d_sums.release();
    h->d_partial.release();
    h->d_pnum.release();
    h->d_pden.release();
    h->d_out.release();
    h->d_perm.release();
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
    SQB_TRY(sqb_h2d(c, h->d_x.p, x, bytes));
    SQB_CUDA(cudaStreamSynchronize(c->stream));
    h->kind = layout == 0 ? 1 : 2;
    h->x_dtype = x_dtype;
    h->n_feat = n_features;
    h->ran = false;
    return SQB_OK;
}

int sqb_autocorr_load_csr(sqb_autocorr* h, const int64_t* x_indptr, const int32_t* x_indices, const void* x_data, int x_dtype,
                          int layout, int64_t n_features) {
    SQB_CHECK(h && x_indptr, SQB_ERR_INVALID, "sqb_autocorr_load_csr: null argument");
    SQB_CHECK(x_dtype == 0 || x_dtype == 1, SQB_ERR_INVALID, "sqb_autocorr_load_csr: x_dtype must be 0 or 1");
    SQB_CHECK(layout == 0 || layout == 1, SQB_ERR_INVALID, "sqb_autocorr_load_csr: layout must be 0 or 1");
    SQB_CHECK(n_features >= 1, SQB_ERR_INVALID, "sqb_autocorr_load_csr: n_features must be positive");
    sqb_ctx* c = h->ctx;
    SQB_CUDA(cudaSetDevice(c->device));
    const int64_t rows = layout == 0 ? n_features : h->n;
    const int64_t cols = layout == 0 ? h->n : n_features;
    const int64_t nnz = x_indptr[rows];
    SQB_CHECK(x_indptr[0] == 0 && nnz >= 0, SQB_ERR_INVALID, "sqb_autocorr_load_csr: bad indptr");
    SQB_CHECK(nnz == 0 || (x_indices && x_data), SQB_ERR_INVALID, "sqb_autocorr_load_csr: null X arrays");
    for (int64_t e = 0; e < nnz; ++e)
        SQB_CHECK(x_indices[e] >= 0 && x_indices[e] < cols, SQB_ERR_INVALID, "sqb_autocorr_load_csr: index %d out of range at %lld",
                  x_indices[e], (long long)e);
    if (layout == 0) {
        SQB_TRY(h->d_xp.alloc(n_features + 1));
        SQB_TRY(h->d_xi.alloc(nnz > 0 ? nnz : 1));
        SQB_TRY(h->d_x.alloc((size_t)(nnz > 0 ? nnz : 1) * xsize(x_dtype)));
        SQB_CUDA(cudaMemcpyAsync(h->d_xp.p, x_indptr, (n_features + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, c->stream));
        if (nnz > 0) {
            SQB_TRY(sqb_h2d(c, h->d_xi.p, x_indices, nnz * sizeof(int32_t)));
            SQB_TRY(sqb_h2d(c, h->d_x.p, x_data, (size_t)nnz * xsize(x_dtype)));
        }
        SQB_CUDA(cudaStreamSynchronize(c->stream));
        h->x_nnz = nnz;
    } else {
        if (x_dtype == 0)
            SQB_TRY(ac_transpose_csr<float>(h, x_indptr, x_indices, x_data, n_features));
        else
            SQB_TRY(ac_transpose_csr<double>(h, x_indptr, x_indices, x_data, n_features));
    }
    h->kind = 3;
    h->x_dtype = x_dtype;
    h->n_feat = n_features;
    h->ran = false;
    return SQB_OK;
}

int sqb_autocorr_run_async(sqb_autocorr* h, int mode, const int64_t* row_perm) {
    SQB_CHECK(h, SQB_ERR_INVALID, "sqb_autocorr_run_async: null handle");
    SQB_CHECK(h->kind != 0, SQB_ERR_STATE, "sqb_autocorr_run_async: load X first");
    SQB_CHECK(mode == 0 || mode == 1, SQB_ERR_INVALID, "sqb_autocorr_run_async: mode must be 0 (moran) or 1 (geary)");
    sqb_ctx* c = h->ctx;
    SQB_CUDA(cudaSetDevice(c->device));
    const int64_t* d_perm = nullptr;
    if (row_perm) {
        std::vector<uint8_t> seen((size_t)h->n, 0);
        for (int64_t r = 0; r < h->n; ++r) {
            SQB_CHECK(row_perm[r] >= 0 && row_perm[r] < h->n && !seen[row_perm[r]], SQB_ERR_INVALID,
                      "sqb_autocorr_run_async: row_perm is not a permutation (entry %lld)", (long long)r);
            seen[row_perm[r]] = 1;
        }
        SQB_TRY(h->d_perm.alloc(h->n));
        SQB_CUDA(cudaMemcpyAsync(h->d_perm.p, row_perm, h->n * sizeof(int64_t), cudaMemcpyHostToDevice, c->stream));
        SQB_CUDA(cudaStreamSynchronize(c->stream));
        d_perm = h->d_perm.p;
    }
    int rc = h->x_dtype == 0 ? ac_run_typed<float>(h, mode, d_perm) : ac_run_typed<double>(h, mode, d_perm);
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
