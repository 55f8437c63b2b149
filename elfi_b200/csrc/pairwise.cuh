// pairwise.cuh -- streaming evaluation of NumPy's pairwise summation (DOUBLE_pairwise_sum).
//
// NumPy reduces a contiguous run with blocks of <= 128 elements summed by 8 strided accumulators
// r[0..7] combined as ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)) plus a sequential tail; longer runs
// are split recursively at n/2 rounded down to a multiple of 8.  Every left part is a multiple
// of 8 long, so every leaf starts at an index that is a multiple of 8 and only the last leaf has
// a tail.  PairwiseStream consumes the terms strictly in order, 8 at a time, and reproduces that
// tree bit for bit.
//
// The recursion stack (pending right-part lengths and finished left sums) has MAXD levels and
// is kept as a register SHIFT stack: the top is always element 0, push/pop move every level by
// one with compile-time indices.  (A runtime-indexed array - or a compare-and-select loop, which
// the compiler folds back into an indexed access - would demote the whole object to local
// memory, an order of magnitude slower on the bandwidth-bound summary kernels.)  Pushes and pops
// happen once per <=128-term leaf.  MAXD levels cover runs of up to 120 * 2^MAXD + 8 terms; the
// one-thread-per-row fallback uses a deeper stack.
#pragma once

#include "common.cuh"

namespace elfi {

template <int MAXD>
struct PairwiseStream {
    double r[8];
    double res;
    double left_val[MAXD];     // [0] = top of stack
    int pending_right[MAXD];   // [0] = top of stack
    uint32_t has_left;         // bit 0 = top frame already holds its left sum
    int leaf_start, leaf_end, tail_start;
    int depth;
    bool in_tail;

    // Longest run whose recursion never holds more than MAXD open splits.  A right part has up
    // to n/2 + 7 terms, so the depth needed is not log2(n / 128): runs of 120 * 2^MAXD + 8 terms
    // fit, and 7689 is the first length that needs a 7th level (found by the host harness of
    // treesum.cuh; enumerated in tests/test_leafsum_host.py).
    static __host__ __device__ constexpr int64_t max_terms() { return (int64_t(120) << MAXD) + 8; }

    __device__ __forceinline__ void push_frame(int right) {
#pragma unroll
        for (int d = MAXD - 1; d > 0; --d) {
            pending_right[d] = pending_right[d - 1];
            left_val[d] = left_val[d - 1];
        }
        pending_right[0] = right;
        left_val[0] = 0.0;
        has_left <<= 1;
        ++depth;
    }
    __device__ __forceinline__ void pop_frame() {
#pragma unroll
        for (int d = 0; d < MAXD - 1; ++d) {
            pending_right[d] = pending_right[d + 1];
            left_val[d] = left_val[d + 1];
        }
        has_left >>= 1;
        --depth;
    }

    __device__ __forceinline__ void descend(int start, int n) {
        while (n > 128) {
            int n2 = n / 2;
            n2 -= n2 % 8;
            push_frame(n - n2);
            n = n2;
        }
        leaf_start = start;
        leaf_end = start + n;
        tail_start = n < 8 ? start : start + (n - n % 8);
        in_tail = n < 8;
        res = 0.0;
    }
    __device__ __forceinline__ void begin(int m) {
        depth = 0;
        has_left = 0;
        descend(0, m);
    }
    __device__ __forceinline__ double fold() const {
        return __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])),
                         __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
    }
    __device__ __forceinline__ double leaf_value() const { return in_tail ? res : fold(); }

    // Called when term index j0 (a multiple of 8) is about to be fed and j0 == leaf_end.
    __device__ __forceinline__ void close_leaf_and_open_next() {
        double v = leaf_value();
        const int next = leaf_end;
        while (depth > 0) {
            if (!(has_left & 1u)) {
                left_val[0] = v;
                has_left |= 1u;
                descend(next, pending_right[0]);
                return;
            }
            v = __dadd_rn(left_val[0], v);
            pop_frame();
        }
        res = v;  // not reached while terms remain
    }
    // Feed up to 8 terms t[0..cnt) with global indices j0..j0+cnt-1, j0 % 8 == 0.
    __device__ __forceinline__ void feed8(int j0, const double (&t)[8], int cnt) {
        if (j0 == leaf_end) close_leaf_and_open_next();
        if (!in_tail && j0 == tail_start && tail_start != leaf_start) {
            res = fold();     // the leaf has a tail: fold the strided accumulators, go sequential
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
            v = __dadd_rn(left_val[0], v);
            pop_frame();
        }
        return v;
    }
};

}  // namespace elfi
