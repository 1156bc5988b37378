// interaction.cu — interaction_matrix on B200 (sm_100a): cluster x cluster sums over the stored entries of a CSR graph.
//
// Replaces the reference's numba loop `_interaction_matrix` (src/squidpy/gr/_nhood.py:412-429) together with the NaN masking
// of `interaction_matrix` (:386-395): observations without a label (code < 0) are removed from rows AND columns, which for
// the sums is the same as skipping every stored entry with an unlabelled end point (no sub-matrix is built).
//   weights = false: out[a][b] = number of stored entries (i -> j) with code(i) = a, code(j) = b          (int64, exact)
//   weights = true : out[a][b] = sum of their values, accumulated in float64                             (order free:
//                    exact for integer-valued weights, |rel err| ~1e-16 * entries otherwise)
// HBM-bound: the CSR is streamed once (8-12 bytes per stored entry); the label gather hits L2.
#include "common.cuh"

template <typename VT, bool WEIGHTED>
__global__ void __launch_bounds__(256) interaction_kernel(const uint32_t* __restrict__ indptr, const uint32_t* __restrict__ indices,
                                                          const VT* __restrict__ data, const int32_t* __restrict__ codes, int64_t n,
                                                          int C, int use_smem, double* __restrict__ out_w,
                                                          unsigned long long* __restrict__ out_c) {
    extern __shared__ unsigned char sqb_inter_smem[];
    double* s_w = reinterpret_cast<double*>(sqb_inter_smem);
    unsigned long long* s_c = reinterpret_cast<unsigned long long*>(sqb_inter_smem);
    const int CC = C * C;
    if (use_smem) {
        for (int i = threadIdx.x; i < CC; i += blockDim.x) {
            if (WEIGHTED)
                s_w[i] = 0.0;
            else
                s_c[i] = 0ull;
        }
        __syncthreads();
    }
    // one warp per row: lanes stride over the row's entries (coalesced index / value loads)
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t i = warp; i < n; i += nwarps) {
        const int32_t a = codes[i];
        if (a < 0) continue;  // warp uniform
        const uint32_t beg = indptr[i], end = indptr[i + 1];
        for (uint32_t e = beg + lane; e < end; e += 32) {
            const int32_t b = codes[indices[e]];
            if (b < 0) continue;
            const int bin = a * C + b;
            if (WEIGHTED) {
                const double v = (double)data[e];
                if (use_smem)
                    atomicAdd(&s_w[bin], v);
                else
                    atomicAdd(&out_w[bin], v);
            } else {
                if (use_smem)
                    atomicAdd(&s_c[bin], 1ull);
                else
                    atomicAdd(&out_c[bin], 1ull);
            }
        }
    }
    if (use_smem) {
        __syncthreads();
        for (int i = threadIdx.x; i < CC; i += blockDim.x) {
            if (WEIGHTED) {
                const double v = s_w[i];
                if (v != 0.0) atomicAdd(&out_w[i], v);
            } else {
                const unsigned long long v = s_c[i];
                if (v != 0ull) atomicAdd(&out_c[i], v);
            }
        }
    }
}

extern "C" int sqb_interaction_matrix(sqb_ctx* c, int64_t n, int64_t nnz, const uint32_t* indptr, const uint32_t* indices,
                                      const void* data, int data_dtype, const int32_t* codes, int n_cls, double* out_weighted,
                                      int64_t* out_counts) {
    SQB_CHECK(c, SQB_ERR_INVALID, "sqb_interaction_matrix: null ctx");
    SQB_CHECK(n >= 1 && n < 0x7FFFFFFFLL, SQB_ERR_INVALID, "sqb_interaction_matrix: n=%lld out of range", (long long)n);
    SQB_CHECK(nnz >= 0 && nnz < 0xFFFFFFFFLL, SQB_ERR_INVALID, "sqb_interaction_matrix: nnz=%lld does not fit uint32", (long long)nnz);
    SQB_CHECK(indptr && (indices || nnz == 0) && codes, SQB_ERR_INVALID, "sqb_interaction_matrix: null argument");
    SQB_CHECK(n_cls >= 1 && n_cls <= 4096, SQB_ERR_INVALID, "sqb_interaction_matrix: n_cls=%d out of range [1, 4096]", n_cls);
    const bool weighted = data != nullptr;
    SQB_CHECK(!weighted || data_dtype == 0 || data_dtype == 1, SQB_ERR_INVALID, "sqb_interaction_matrix: data_dtype must be 0 (f32) or 1 (f64)");
    SQB_CHECK(weighted ? out_weighted != nullptr : out_counts != nullptr, SQB_ERR_INVALID, "sqb_interaction_matrix: null output");
    SQB_CHECK(indptr[0] == 0 && (int64_t)indptr[n] == nnz, SQB_ERR_INVALID, "sqb_interaction_matrix: indptr inconsistent with nnz");
    for (int64_t i = 0; i < n; ++i) {
        SQB_CHECK(indptr[i] <= indptr[i + 1], SQB_ERR_INVALID, "sqb_interaction_matrix: indptr not monotone at row %lld", (long long)i);
        SQB_CHECK(codes[i] < n_cls, SQB_ERR_INVALID, "sqb_interaction_matrix: codes[%lld]=%d >= n_cls=%d", (long long)i, codes[i], n_cls);
    }
    for (int64_t e = 0; e < nnz; ++e)
        SQB_CHECK((int64_t)indices[e] < n, SQB_ERR_INVALID, "sqb_interaction_matrix: column index %u at entry %lld out of range", indices[e],
                  (long long)e);
    SQB_CUDA(cudaSetDevice(c->device));
    const int CC = n_cls * n_cls;
    const size_t vsize = weighted ? (data_dtype == 0 ? 4 : 8) : 0;
    DevBuf<uint32_t> d_ptr, d_idx;
    DevBuf<uint8_t> d_val;
    DevBuf<int32_t> d_codes;
    DevBuf<unsigned long long> d_out;  // 8 bytes per bin: double or uint64
    d_ptr.bind(c->stream);
    d_idx.bind(c->stream);
    d_val.bind(c->stream);
    d_codes.bind(c->stream);
    d_out.bind(c->stream);
    auto cleanup = [&]() {
        d_ptr.release();
        d_idx.release();
        d_val.release();
        d_codes.release();
        d_out.release();
    };
    int rc;
    if ((rc = d_ptr.alloc(n + 1)) || (rc = d_idx.alloc(nnz > 0 ? nnz : 1)) || (rc = d_val.alloc(weighted && nnz > 0 ? nnz * vsize : 1)) ||
        (rc = d_codes.alloc(n)) || (rc = d_out.alloc(CC))) {
        cleanup();
        return rc;
    }
    cudaError_t e = cudaMemcpyAsync(d_ptr.p, indptr, (n + 1) * sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess && nnz > 0 && sqb_h2d(c, d_idx.p, indices, nnz * sizeof(uint32_t)) != SQB_OK) e = cudaErrorUnknown;
    if (e == cudaSuccess && weighted && nnz > 0 && sqb_h2d(c, d_val.p, data, nnz * vsize) != SQB_OK) e = cudaErrorUnknown;
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_codes.p, codes, n * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(d_out.p, 0, (size_t)CC * 8, c->stream);
    if (e == cudaSuccess) {
        const int use_smem = (size_t)CC * 8 <= 48 * 1024;
        const size_t smem = use_smem ? (size_t)CC * 8 : 0;
        int64_t blocks = ceil_div64(n * 32, 256);
        const int64_t max_blocks = (int64_t)c->sm_count * 8;
        if (blocks > max_blocks) blocks = max_blocks;
        if (blocks < 1) blocks = 1;
        SqbLaunchScope scope(c, SQB_K_MISC);
        double* ow = reinterpret_cast<double*>(d_out.p);
        if (!weighted)
            interaction_kernel<float, false><<<(unsigned)blocks, 256, smem, c->stream>>>(d_ptr.p, d_idx.p, nullptr, d_codes.p, n, n_cls, use_smem, ow, d_out.p);
        else if (data_dtype == 0)
            interaction_kernel<float, true><<<(unsigned)blocks, 256, smem, c->stream>>>(d_ptr.p, d_idx.p, reinterpret_cast<const float*>(d_val.p), d_codes.p, n, n_cls, use_smem, ow, d_out.p);
        else
            interaction_kernel<double, true><<<(unsigned)blocks, 256, smem, c->stream>>>(d_ptr.p, d_idx.p, reinterpret_cast<const double*>(d_val.p), d_codes.p, n, n_cls, use_smem, ow, d_out.p);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess)
        e = cudaMemcpyAsync(weighted ? (void*)out_weighted : (void*)out_counts, d_out.p, (size_t)CC * 8, cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    cleanup();
    if (e != cudaSuccess) {
        sqb_set_error("sqb_interaction_matrix: %s", cudaGetErrorString(e));
        return SQB_ERR_CUDA;
    }
    return SQB_OK;
}
