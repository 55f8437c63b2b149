// pairwise.cuh -- streaming evaluation of NumPy's pairwise summation (DOUBLE_pairwise_sum).
//
// NumPy reduces a contiguous run with blocks of <= 128 elements summed by 8 strided accumulators
// r[0..7] combined as ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)) plus a sequential tail; longer runs
// are split recursively at n/2 rounded down to a multiple of 8.  Every left part is a multiple
// of 8 long, so every leaf starts at an index that is a multiple of 8 and only the last leaf has
// a tail.  PairwiseStream consumes the terms strictly in order, 8 at a time, and reproduces that
// tree bit for bit.
#pragma once

#include "common.cuh"

namespace elfi {

constexpr int PW_MAX_DEPTH = 26;  // recursion depth bound: rows up to 128 * 2^26 elements

// Streaming evaluation of NumPy's pairwise sum over m terms fed one aligned group of 8 at a
// time (the last group may be partial).  All lanes of a warp run identical control flow
// because every row has the same length.
struct PairwiseStream {
    double r[8];
    double res;
    int64_t leaf_end;     // first term index after the current leaf
    int64_t tail_start;   // first term index of the sequential tail of the current leaf
    int64_t leaf_start;
    int depth;
    bool in_tail;
    int64_t pending_right[PW_MAX_DEPTH];
    double left_val[PW_MAX_DEPTH];
    bool has_left[PW_MAX_DEPTH];

    __device__ __forceinline__ void descend(int64_t start, int64_t n) {
        while (n > 128) {
            int64_t n2 = n / 2;
            n2 -= n2 % 8;
            pending_right[depth] = n - n2;
            has_left[depth] = false;
            ++depth;
            n = n2;
        }
        leaf_start = start;
        leaf_end = start + n;
        tail_start = n < 8 ? start : start + (n - n % 8);
        in_tail = n < 8;
        res = 0.0;
    }
    __device__ __forceinline__ void begin(int64_t m) {
        depth = 0;
        descend(0, m);
    }
    __device__ __forceinline__ double leaf_value() const {
        if (leaf_end - leaf_start < 8) return res;
        if (in_tail) return res;
        return __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])),
                         __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
    }
    // Called when term index j0 (a multiple of 8) is about to be fed and j0 == leaf_end.
    __device__ __forceinline__ void close_leaf_and_open_next() {
        double v = leaf_value();
        const int64_t next = leaf_end;
        while (depth > 0) {
            if (!has_left[depth - 1]) {
                left_val[depth - 1] = v;
                has_left[depth - 1] = true;
                const int64_t n = pending_right[depth - 1];
                descend(next, n);
                return;
            }
            v = __dadd_rn(left_val[depth - 1], v);
            --depth;
        }
        res = v;  // not reached while terms remain
    }
    // Feed up to 8 terms t[0..cnt) with global indices j0..j0+cnt-1, j0 % 8 == 0.
    __device__ __forceinline__ void feed8(int64_t j0, const double* t, int cnt) {
        if (j0 == leaf_end) close_leaf_and_open_next();
        if (!in_tail && j0 == tail_start && tail_start != leaf_start) {
            // leaf has a tail: fold the strided accumulators first, then go sequential
            res = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])),
                            __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
            in_tail = true;
        }
        if (in_tail) {
#pragma unroll
            for (int k = 0; k < 8; ++k)
                if (k < cnt) res = __dadd_rn(res, t[k]);
        } else if (j0 == leaf_start) {
#pragma unroll
            for (int k = 0; k < 8; ++k) r[k] = t[k];
        } else {
#pragma unroll
            for (int k = 0; k < 8; ++k) r[k] = __dadd_rn(r[k], t[k]);
        }
    }
    __device__ __forceinline__ double finish() {
        double v = leaf_value();
        while (depth > 0) {
            v = __dadd_rn(left_val[depth - 1], v);
            --depth;
        }
        return v;
    }
};

}  // namespace elfi
