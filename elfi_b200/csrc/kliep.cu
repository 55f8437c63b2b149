// kliep.cu -- KLIEP density-ratio estimation for AdaptiveThresholdSMC:
//   elfi/methods/density_ratio_estimation.py:71-207 (fit, _compute_A, _compute_b, _KLIEP,
//   max_ratio), driven by elfi/methods/inference/samplers.py:757-813.
//
// r(x) = sum_j alpha_j exp(-|x - theta_j|^2 / (2 sigma^2)), theta = first n_basis numerator points.
// The reference builds A (N x n_basis) with Python double loops (infeasible at 1e6 particles) and
// runs <= max_iter projected-gradient steps  alpha += eps * A^T (w / (A alpha)); projection onto
// {alpha >= 0, b^T alpha = 1}.  Here A is materialised once in HBM (N x n_basis fp64, 800 MB at
// N = 1e6, n_basis = 100) and every step is ONE pass over it (row-wise A alpha and the rank-1
// accumulation of A^T (w / v) fused): 800 MB per step, HBM bound.
// Parity is tolerance-level (different summation order), pinned against the reference at small N.
#include "common.cuh"

namespace elfi {

constexpr int KL_MAXB = 128;   // max basis functions (4 per lane)

// A[i, j] = exp(-0.5 |x_i - theta_j|^2 / sigma^2), theta_j = x_j (j < nb)
__global__ void __launch_bounds__(256)
kliep_basis_kernel(const double* __restrict__ x, int64_t ldx, int64_t N, int p, int nb,
                   double neg_half_inv_s2, double* __restrict__ A) {
    extern __shared__ double th[];   // nb * p
    for (int t = threadIdx.x; t < nb * p; t += blockDim.x) th[t] = x[int64_t(t / p) * ldx + t % p];
    __syncthreads();
    const int64_t total = N * nb;
    const int64_t stride = int64_t(gridDim.x) * blockDim.x;
    for (int64_t e = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; e < total; e += stride) {
        const int64_t i = e / nb;
        const int j = int(e - i * nb);
        double r2 = 0.0;
        for (int a = 0; a < p; ++a) {
            const double d = x[i * ldx + a] - th[j * p + a];
            r2 = fma(d, d, r2);
        }
        A[e] = exp(r2 * neg_half_inv_s2);
    }
}

// b[j] = sum_k (w_y[k] / sum w_y) exp(-0.5 |theta_j - y_k|^2 / sigma^2); one block per basis
__global__ void __launch_bounds__(256)
kliep_b_kernel(const double* __restrict__ x, int64_t ldx, const double* __restrict__ y, int64_t ldy,
               int64_t Ny, int p, const double* __restrict__ wy, const double* __restrict__ wy_sum,
               double neg_half_inv_s2, double* __restrict__ b) {
    __shared__ double red[8];
    const int j = blockIdx.x;
    double acc = 0.0;
    for (int64_t k = threadIdx.x; k < Ny; k += blockDim.x) {
        double r2 = 0.0;
        for (int a = 0; a < p; ++a) {
            const double d = x[int64_t(j) * ldx + a] - y[k * ldy + a];
            r2 = fma(d, d, r2);
        }
        acc = fma(wy ? wy[k] : 1.0, exp(r2 * neg_half_inv_s2), acc);
    }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int k = 0; k < 8; ++k) s += red[k];
        b[j] = s / (wy ? wy_sum[0] : double(Ny));
    }
}

// One pass over A: v_i = A_i . alpha; rows with any A_ij > 1e-64 contribute w_i / v_i * A_i to the
// gradient.  Warp per row (lane owns columns lane, lane+32, ...); per-block partial gradients.
// mode 1 additionally writes t = A alpha, accumulates |t - t_prev|^2 and max(t) (no gradient).
__global__ void __launch_bounds__(256)
kliep_pass_kernel(const double* __restrict__ A, int64_t N, int nb, const double* __restrict__ alpha,
                  const double* __restrict__ wx, int mode, double* __restrict__ tprev,
                  double* __restrict__ partial) {
    __shared__ double sh[8][KL_MAXB + 2];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double al[4], g[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
    for (int q = 0; q < 4; ++q) al[q] = (lane + 32 * q < nb) ? alpha[lane + 32 * q] : 0.0;
    double diff2 = 0.0, tmax = -INFINITY;
    const int64_t wstride = int64_t(gridDim.x) * 8;
    for (int64_t i = int64_t(blockIdx.x) * 8 + warp; i < N; i += wstride) {
        double a[4];
        double v = 0.0, amax = 0.0;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            a[q] = (lane + 32 * q < nb) ? A[i * nb + lane + 32 * q] : 0.0;
            v = fma(a[q], al[q], v);
            amax = fmax(amax, a[q]);
        }
        for (int o = 16; o > 0; o >>= 1) {
            v += __shfl_xor_sync(0xffffffffu, v, o);
            amax = fmax(amax, __shfl_xor_sync(0xffffffffu, amax, o));
        }
        if (mode == 0) {
            if (amax > 1e-64) {
                const double r = (wx ? wx[i] : 1.0) / v;
#pragma unroll
                for (int q = 0; q < 4; ++q) g[q] = fma(a[q], r, g[q]);
            }
        } else if (lane == 0) {
            const double d = v - tprev[i];
            diff2 = fma(d, d, diff2);
            tmax = fmax(tmax, v);
            tprev[i] = v;
        }
    }
    if (mode == 0) {
#pragma unroll
        for (int q = 0; q < 4; ++q)
            if (lane + 32 * q < KL_MAXB) sh[warp][lane + 32 * q] = g[q];
    } else if (lane == 0) {
        sh[warp][0] = diff2;
        sh[warp][1] = tmax;
    }
    __syncthreads();
    if (mode == 0) {
        if (threadIdx.x < nb) {
            double s = 0.0;
            for (int k = 0; k < 8; ++k) s += sh[k][threadIdx.x];
            partial[int64_t(blockIdx.x) * KL_MAXB + threadIdx.x] = s;
        }
    } else if (threadIdx.x == 0) {
        double s = 0.0, m = -INFINITY;
        for (int k = 0; k < 8; ++k) { s += sh[k][0]; m = fmax(m, sh[k][1]); }
        partial[int64_t(blockIdx.x) * 2] = s;
        partial[int64_t(blockIdx.x) * 2 + 1] = m;
    }
}

// alpha update (density_ratio_estimation.py:191-194); single block of KL_MAXB threads
__global__ void __launch_bounds__(KL_MAXB)
kliep_update_kernel(const double* __restrict__ partial, int nblocks, int nb, const double* __restrict__ b,
                    double eps, double* __restrict__ alpha) {
    __shared__ double red[KL_MAXB];
    const int j = threadIdx.x;
    double g = 0.0;
    if (j < nb)
        for (int k = 0; k < nblocks; ++k) g += partial[int64_t(k) * KL_MAXB + j];
    const double bj = j < nb ? b[j] : 0.0;
    double a = j < nb ? alpha[j] + eps * g : 0.0;
    auto dot = [&](double v) -> double {
        red[j] = v;
        __syncthreads();
        for (int o = KL_MAXB / 2; o > 0; o >>= 1) {
            if (j < o) red[j] += red[j + o];
            __syncthreads();
        }
        const double r = red[0];
        __syncthreads();
        return r;
    };
    const double bb = dot(bj * bj);
    const double ba = dot(bj * a);
    a = fmax(0.0, a + (1.0 - ba) * (bj / bb));
    const double ba2 = dot(bj * a);
    if (j < nb) alpha[j] = a / ba2;
}

__global__ void kliep_reduce2_kernel(const double* __restrict__ partial, int nblocks,
                                     double* __restrict__ out) {
    if (threadIdx.x == 0) {
        double s = 0.0, m = -INFINITY;
        for (int k = 0; k < nblocks; ++k) { s += partial[2 * k]; m = fmax(m, partial[2 * k + 1]); }
        out[0] = sqrt(s);
        out[1] = m;
    }
}

__global__ void kliep_fill_kernel(double* __restrict__ p, int n, double v) {
    if (threadIdx.x < n) p[threadIdx.x] = v;
}

__global__ void kliep_sum_kernel(const double* __restrict__ v, int64_t n, double* __restrict__ out) {
    __shared__ double ws[32];
    double acc = 0.0;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) acc += v[i];
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        double t = threadIdx.x < (blockDim.x >> 5) ? ws[threadIdx.x] : 0.0;
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (threadIdx.x == 0) out[0] = t;
    }
}

}  // namespace elfi

extern "C" int elfi_b200_kliep_fit_f64(elfi_b200_ctx* ctx, const double* x, int64_t ldx, int64_t Nx,
                                       const double* y, int64_t ldy, int64_t Ny, int64_t p,
                                       const double* wx, const double* wy, double sigma,
                                       int64_t n_basis, double epsilon, int64_t max_iter,
                                       double abs_tol, int64_t conv_check_interval,
                                       double* alpha_out, double* result_host) {
    using namespace elfi;
    ELFI_REQUIRE(ctx && x && y && alpha_out && result_host, "kliep: NULL argument");
    ELFI_REQUIRE(p >= 1 && p <= 16 && ldx >= p && ldy >= p && Ny >= 1, "kliep: bad shape");
    ELFI_REQUIRE(n_basis >= 1 && n_basis <= KL_MAXB, "kliep: n_basis must be in [1, %d]", KL_MAXB);
    ELFI_REQUIRE(Nx >= n_basis, "Number of RBFs (%lld) can't be larger than number of samples (%lld).",
                 (long long)n_basis, (long long)Nx);
    ELFI_REQUIRE(sigma > 0 && max_iter >= 0 && conv_check_interval >= 1, "kliep: bad parameters");
    ELFI_CUDA_OK(cudaSetDevice(ctx->device));
    cudaStream_t stream = 0;
    const int nb = int(n_basis);
    const int blocks = ctx->sm_count * 8;
    auto align = [](size_t v) { return (v + 255) & ~size_t(255); };
    const size_t off_t = align(size_t(Nx) * nb * 8);
    const size_t off_part = off_t + align(size_t(Nx) * 8);
    const size_t off_b = off_part + align(size_t(blocks) * KL_MAXB * 8);
    const size_t off_misc = off_b + align(KL_MAXB * 8);
    uint8_t* base = static_cast<uint8_t*>(ctx_scratch(ctx, off_misc + 1024));
    if (!base) return ELFI_B200_ERR_NOMEM;
    double* A = reinterpret_cast<double*>(base);
    double* tprev = reinterpret_cast<double*>(base + off_t);
    double* partial = reinterpret_cast<double*>(base + off_part);
    double* b = reinterpret_cast<double*>(base + off_b);
    double* misc = reinterpret_cast<double*>(base + off_misc);   // [0..1] norm,max  [2] wy sum
    const double f = -0.5 / (sigma * sigma);
    kliep_basis_kernel<<<blocks, 256, size_t(nb) * p * 8, stream>>>(x, ldx, Nx, int(p), nb, f, A);
    if (wy) kliep_sum_kernel<<<1, 1024, 0, stream>>>(wy, Ny, misc + 2);
    kliep_b_kernel<<<nb, 256, 0, stream>>>(x, ldx, y, ldy, Ny, int(p), wy, misc + 2, f, b);
    kliep_fill_kernel<<<1, KL_MAXB, 0, stream>>>(alpha_out, nb, 1.0 / double(nb));
    ELFI_CUDA_OK(cudaMemsetAsync(tprev, 0, size_t(Nx) * 8, stream));
    // target_fun_prev = A alpha0
    kliep_pass_kernel<<<blocks, 256, 0, stream>>>(A, Nx, nb, alpha_out, wx, 1, tprev, partial);
    double host2[2] = {0.0, 0.0};
    int64_t iters = 0;
    for (int64_t i = 0; i < max_iter; ++i) {
        kliep_pass_kernel<<<blocks, 256, 0, stream>>>(A, Nx, nb, alpha_out, wx, 0, tprev, partial);
        kliep_update_kernel<<<1, KL_MAXB, 0, stream>>>(partial, blocks, nb, b, epsilon, alpha_out);
        ++iters;
        if (i % conv_check_interval == 0) {
            kliep_pass_kernel<<<blocks, 256, 0, stream>>>(A, Nx, nb, alpha_out, wx, 1, tprev, partial);
            kliep_reduce2_kernel<<<1, 32, 0, stream>>>(partial, blocks, misc);
            ELFI_CUDA_OK(cudaMemcpyAsync(host2, misc, 16, cudaMemcpyDeviceToHost, stream));
            ELFI_CUDA_OK(cudaStreamSynchronize(stream));
            if (host2[0] < abs_tol) break;
        }
    }
    // max_ratio = max_i (A alpha)_i over the numerator sample
    kliep_pass_kernel<<<blocks, 256, 0, stream>>>(A, Nx, nb, alpha_out, wx, 1, tprev, partial);
    kliep_reduce2_kernel<<<1, 32, 0, stream>>>(partial, blocks, misc);
    ELFI_CUDA_OK(cudaMemcpyAsync(host2, misc, 16, cudaMemcpyDeviceToHost, stream));
    ELFI_CUDA_OK(cudaStreamSynchronize(stream));
    ELFI_CUDA_OK(cudaGetLastError());
    result_host[0] = host2[1];          // max ratio
    result_host[1] = double(iters);
    return ELFI_B200_OK;
}
