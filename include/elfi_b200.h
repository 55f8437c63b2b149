/*
 * elfi_b200.h -- C ABI of libelfi_b200.so: the B200 (sm_100a) implementation of ELFI's
 * data-parallel sampler / BOLFI hot path.
 *
 * This is the drop-in boundary.  ELFI is pure Python: the binding a
 * maintainer adds on the reference side is a ctypes stub called from
 * the node operations that `elfi.executor.Executor._run` invokes (elfi/executor.py:143-159)
 * and from the sampler bookkeeping in elfi/methods/inference/samplers.py.
 *
 * Conventions
 *   - every function returns 0 on success, <0 on error (ELFI_B200_ERR_*); the message for
 *     the calling host thread is available from elfi_b200_last_error();
 *   - all array arguments are caller-allocated DEVICE pointers unless the name ends in
 *     `_host`; matrices are row-major with a leading dimension given in ELEMENTS;
 *   - sizes are int64_t; `stream` is a cudaStream_t passed as void* (NULL = legacy default
 *     stream); kernels are asynchronous on that stream;
 *   - the context owns only scratch memory; it never takes ownership of caller buffers;
 *   - one context per device; a context may be used by one host thread at a time.
 */
#ifndef ELFI_B200_H
#define ELFI_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ELFI_B200_VERSION 100

#define ELFI_B200_OK 0
#define ELFI_B200_ERR_ARG (-1)     /* invalid argument (shape, alignment, NULL) */
#define ELFI_B200_ERR_CUDA (-2)    /* a CUDA runtime / driver call failed */
#define ELFI_B200_ERR_NOMEM (-3)   /* scratch allocation failed */
#define ELFI_B200_ERR_UNSUPPORTED (-4)

#define ELFI_B200_MAX_NESTED 32    /* max nested distance columns K */

typedef struct elfi_b200_ctx elfi_b200_ctx;

/* ---- context ------------------------------------------------------------------------ */

int elfi_b200_version(void);
const char* elfi_b200_last_error(void);

/* Creates the context of CUDA device `device` (must be compute capability 10.x). */
int elfi_b200_ctx_create(int device, elfi_b200_ctx** out);
int elfi_b200_ctx_destroy(elfi_b200_ctx* ctx);
/* Number of SMs of the context's device (grid sizing is a multiple of this). */
int elfi_b200_ctx_sm_count(const elfi_b200_ctx* ctx);

/* ---- distance + acceptance ------------------------------------------------------------
 * Replaces, for the Euclidean family, the body of
 *   elfi/model/utils.py:37-52        distance_as_discrepancy  (column_stack + dist + flatten)
 *   elfi/model/elfi_model.py:1037    Distance -> scipy.spatial.distance.cdist(X, obs, 'euclidean')
 *   elfi/model/elfi_model.py:1135-1151  AdaptiveDistance.nested_distance (K weighted columns)
 * and the acceptance test of
 *   elfi/methods/inference/samplers.py:223-225   accepted = all_k(d[:, k] <= thr[k])
 *
 *   d[i, k] = sqrt( sum_j  W[k, j] * (S[i, j] - obs[j])^2 ),   j = 0 .. D-1 strictly in order,
 * every multiply and add rounded separately in fp64 (SciPy's order: bit-identical results).
 *
 *   S        (B, D) row-major, leading dimension ldS (elements)
 *   obs      (D)
 *   W        (K, D) row-major weights (cdist's `w`, i.e. 1/scale^2), or NULL = unweighted
 *            (K must then be 1; a row of ones is bit-identical to the unweighted form)
 *   K        number of nested distance columns, 1 <= K <= ELFI_B200_MAX_NESTED
 *   thr_host HOST pointer to K thresholds, or NULL = no acceptance test
 *   d_out    (B, K) row-major distances
 *   acc_idx  int32[B] ascending indices of accepted rows, or NULL (requires thr_host)
 *   n_acc    device int64[1] number of accepted rows (written when thr_host != NULL), or NULL
 */
int elfi_b200_dist_euclid_thr_f64(elfi_b200_ctx* ctx, const double* S, int64_t ldS, int64_t B,
                                  int64_t D, const double* obs, const double* W, int64_t K,
                                  const double* thr_host, double* d_out, int32_t* acc_idx,
                                  int64_t* n_acc, void* stream);

/* Same computation with HOST buffers (pageable or pinned): rows are streamed to the device
 * in chunks on two copy streams overlapped with the kernel; distances, accepted indices
 * and the count are copied back.  Blocks until the results are in host memory.
 * d_out_host may be NULL (then only accepted indices / count come back). */
int elfi_b200_dist_euclid_thr_f64_host(elfi_b200_ctx* ctx, const double* S_host, int64_t ldS,
                                       int64_t B, int64_t D, const double* obs_host,
                                       const double* W_host, int64_t K, const double* thr_host,
                                       double* d_out_host, int32_t* acc_idx_host,
                                       int64_t* n_acc_host);

#ifdef __cplusplus
}
#endif

#endif /* ELFI_B200_H */
