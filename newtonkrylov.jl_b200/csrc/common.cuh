// Device-side helpers: wide streaming loads/stores, deterministic grid reductions.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "ak_internal.h"

#define AK_DEV __device__ __forceinline__

namespace ak {

// ---- 256-bit / 128-bit streaming accesses (sm_100a has LDG.E.256 / STG.E.256) ------
struct alignas(32) d4 {
    double x, y, z, w;
};

// read-only data that is touched once per kernel: non-coherent path, do not allocate in L1
AK_DEV d4 ld4_stream(const double* p) {
    d4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];"
                 : "=d"(r.x), "=d"(r.y), "=d"(r.z), "=d"(r.w)
                 : "l"(p));
    return r;
}
// data the same kernel also writes (no .nc)
AK_DEV d4 ld4(const double* p) {
    d4 r;
    asm volatile("ld.global.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];"
                 : "=d"(r.x), "=d"(r.y), "=d"(r.z), "=d"(r.w)
                 : "l"(p));
    return r;
}
AK_DEV void st4(double* p, const d4& v) {
    asm volatile("st.global.L1::no_allocate.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(v.x), "d"(v.y), "d"(v.z),
                 "d"(v.w)
                 : "memory");
}
AK_DEV double2 ld2_stream(const double* p) {
    double2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
    return r;
}
AK_DEV double ld1_stream(const double* p) {
    double r;
    asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(r) : "l"(p));
    return r;
}

// ---- deterministic reductions ---------------------------------------------------------
// Butterfly: every lane ends with the same value, order fixed.
AK_DEV double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Sum over the block (any block shape); result valid in flat thread 0.
// `sh` must hold >= 32 doubles.
AK_DEV double block_sum(double v, double* sh) {
    const int tid = threadIdx.x + threadIdx.y * blockDim.x;
    const int lane = tid & 31, wid = tid >> 5;
    const int nw = (blockDim.x * blockDim.y + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();  // protect sh from a previous use
    if (lane == 0) sh[wid] = v;
    __syncthreads();
    double r = 0.0;
    if (wid == 0) {
        r = (lane < nw) ? sh[lane] : 0.0;
        r = warp_sum(r);
    }
    return r;
}

// Grid-wide sum with a fixed summation order: each block deposits its partial, the block
// that takes the last ticket adds the partials in index order and stores the result.
// `nblocks` = total blocks of the launch, `bid` = this block's linear id.
// Must be called by all threads of every block.  Result: *out (+)= sum.
AK_DEV void grid_sum_finish(double block_partial_in_t0, double* partials, unsigned int* ticket, int bid,
                            int nblocks, double* out, double* sh) {
    __shared__ bool is_last;
    const int tid = threadIdx.x + threadIdx.y * blockDim.x;
    const int nthreads = blockDim.x * blockDim.y;
    if (tid == 0) {
        partials[bid] = block_partial_in_t0;
        __threadfence();
        unsigned int t = atomicAdd(ticket, 1u);
        is_last = (t == (unsigned int)(nblocks - 1));
    }
    __syncthreads();
    if (is_last) {
        __threadfence();
        double s = 0.0;
        for (int i = tid; i < nblocks; i += nthreads) s += __ldcg(partials + i);
        // block_sum assumes 1-D indexing through threadIdx.x; handle 2-D blocks via flat id
        const int lane = tid & 31, wid = tid >> 5;
        const int nw = (nthreads + 31) >> 5;
        s = warp_sum(s);
        __syncthreads();
        if (lane == 0) sh[wid] = s;
        __syncthreads();
        if (wid == 0) {
            double r = (lane < nw) ? sh[lane] : 0.0;
            r = warp_sum(r);
            if (lane == 0) {
                *out = r;
                *ticket = 0u;  // self-reset for the next launch on this stream
            }
        }
    }
}

// Three sums at once (pair-wise Gram-Schmidt: <y_a,w>, <y_b,w>, <y_b,y_a>); same ordering rules.
// partials layout: [3 * bid + c].
AK_DEV void grid_sum_finish3(double s0, double s1, double s2, double* partials, unsigned int* ticket, int bid,
                             int nblocks, double* out, double* sh) {
    __shared__ bool is_last3;
    const int tid = threadIdx.x + threadIdx.y * blockDim.x;
    const int nthreads = blockDim.x * blockDim.y;
    if (tid == 0) {
        partials[3 * bid + 0] = s0;
        partials[3 * bid + 1] = s1;
        partials[3 * bid + 2] = s2;
        __threadfence();
        unsigned int t = atomicAdd(ticket, 1u);
        is_last3 = (t == (unsigned int)(nblocks - 1));
    }
    __syncthreads();
    if (is_last3) {
        __threadfence();
        double a[3] = {0.0, 0.0, 0.0};
        for (int i = tid; i < nblocks; i += nthreads) {
            a[0] += __ldcg(partials + 3 * i + 0);
            a[1] += __ldcg(partials + 3 * i + 1);
            a[2] += __ldcg(partials + 3 * i + 2);
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const double r = block_sum(a[c], sh);
            if (tid == 0) out[c] = r;
        }
        if (tid == 0) *ticket = 0u;
    }
}

// Same, but the three totals are handed back to flat thread 0 of the last block (returns true there only),
// so that the caller decides where they go (device memory, or the peers' mailboxes over NVLink).
// `sysfence`: the block also issued stores to peer memory that must be visible before the result is published.
AK_DEV bool grid_reduce3(double s0, double s1, double s2, double* partials, unsigned int* ticket, int bid, int nblocks,
                         double* sh, double (&r)[3], bool sysfence) {
    __shared__ bool is_last_r;
    const int tid = threadIdx.x + threadIdx.y * blockDim.x;
    const int nthreads = blockDim.x * blockDim.y;
    __syncthreads();  // all stores of the block precede thread 0's fence
    if (tid == 0) {
        partials[3 * bid + 0] = s0;
        partials[3 * bid + 1] = s1;
        partials[3 * bid + 2] = s2;
        if (sysfence) __threadfence_system(); else __threadfence();
        unsigned int t = atomicAdd(ticket, 1u);
        is_last_r = (t == (unsigned int)(nblocks - 1));
    }
    __syncthreads();
    if (!is_last_r) return false;
    __threadfence();
    double a[3] = {0.0, 0.0, 0.0};
    for (int i = tid; i < nblocks; i += nthreads) {
        a[0] += __ldcg(partials + 3 * i + 0);
        a[1] += __ldcg(partials + 3 * i + 1);
        a[2] += __ldcg(partials + 3 * i + 2);
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) r[c] = block_sum(a[c], sh);
    if (tid == 0) *ticket = 0u;
    return tid == 0;
}

// ---- mailbox all-reduce over NVLink peer memory ---------------------------------------------------------
// record (slot, src) = 4 doubles {v0, v1, v2, tag}; tag = sequence number of the collective (never reused)
AK_DEV void st_release_sys_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
AK_DEV unsigned long long ld_acquire_sys_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
// ONE thread: deposit this rank's partial sums in every rank's mailbox (own included)
AK_DEV void mail_post(const P2PDev& pd, unsigned long long seq, double v0, double v1, double v2) {
    const int slot = (int)(seq % kMailSlots);
    __threadfence_system();  // halo rows pushed by this grid are visible before the record is
    for (int q = 0; q < pd.nranks; ++q) {
        double* rec = pd.mail_peer[q] + ((size_t)slot * pd.nranks + pd.rank) * 4;
        rec[0] = v0;
        rec[1] = v1;
        rec[2] = v2;
    }
    __threadfence_system();
    for (int q = 0; q < pd.nranks; ++q) {
        double* rec = pd.mail_peer[q] + ((size_t)slot * pd.nranks + pd.rank) * 4;
        st_release_sys_u64(reinterpret_cast<unsigned long long*>(rec + 3), seq);
    }
}
// Whole block: wait until every rank's record `seq` has arrived in the local mailbox and add them in rank
// order (bit-identical on every rank).  Result in out[0..2] for all threads.  `sh4` >= 4 * kMaxPeers doubles.
AK_DEV void mail_wait_sum(const P2PDev& pd, unsigned long long seq, double (&out)[3], double* sh4) {
    const int tid = threadIdx.x + threadIdx.y * blockDim.x;
    const int slot = (int)(seq % kMailSlots);
    if (tid < pd.nranks) {
        const double* rec = pd.mail_local + ((size_t)slot * pd.nranks + tid) * 4;
        const unsigned long long* tag = reinterpret_cast<const unsigned long long*>(rec + 3);
        const long long t0 = clock64();
        while (ld_acquire_sys_u64(tag) != seq) {
            if (clock64() - t0 > pd.spin_cycles) {  // a peer never produced this record: flag it, do not hang
                *pd.err = 1;
                break;
            }
        }
        sh4[4 * tid + 0] = __ldcv(rec + 0);
        sh4[4 * tid + 1] = __ldcv(rec + 1);
        sh4[4 * tid + 2] = __ldcv(rec + 2);
    }
    __syncthreads();
    out[0] = out[1] = out[2] = 0.0;
    for (int q = 0; q < pd.nranks; ++q) {
        out[0] += sh4[4 * q + 0];
        out[1] += sh4[4 * q + 1];
        out[2] += sh4[4 * q + 2];
    }
    __syncthreads();
}

// h of the second vector of a pair from the raw sums (d1 = <y_a,w>, d2 = <y_b,w>, g = <y_b,y_a>):
// <y_b, w - d1 y_a> = d2 - d1 g.  One definition shared by the vector kernels and the Givens kernel.
AK_DEV double pair_second_h(double d1, double d2, double g) { return __dsub_rn(d2, __dmul_rn(d1, g)); }

}  // namespace ak
