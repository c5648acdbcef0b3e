// Device-side helpers: wide streaming loads/stores, deterministic grid reductions.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

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

// NS sums at once, handed back to flat thread 0 of the last block (returns true there only), so that the caller
// decides where they go (device memory, or the peers' mailboxes over NVLink).  partials layout [NS * bid + c].
// `sysfence`: the block also issued stores to peer memory that must be visible before the result is published.
template <int NS>
AK_DEV bool grid_reduce_n(const double (&v)[NS], double* partials, unsigned int* ticket, int bid, int nblocks,
                          double* sh, double (&r)[NS], bool sysfence) {
    __shared__ bool is_last_r;
    const int tid = threadIdx.x + threadIdx.y * blockDim.x;
    const int nthreads = blockDim.x * blockDim.y;
    double b[NS];
#pragma unroll
    for (int c = 0; c < NS; ++c) b[c] = block_sum(v[c], sh);
    __syncthreads();  // all stores of the block precede thread 0's fence
    if (tid == 0) {
#pragma unroll
        for (int c = 0; c < NS; ++c) partials[NS * bid + c] = b[c];
        if (sysfence) __threadfence_system(); else __threadfence();
        unsigned int t = atomicAdd(ticket, 1u);
        is_last_r = (t == (unsigned int)(nblocks - 1));
    }
    __syncthreads();
    if (!is_last_r) return false;
    __threadfence();
    double a[NS];
#pragma unroll
    for (int c = 0; c < NS; ++c) a[c] = 0.0;
    for (int i = tid; i < nblocks; i += nthreads) {
#pragma unroll
        for (int c = 0; c < NS; ++c) a[c] += __ldcg(partials + NS * i + c);
    }
#pragma unroll
    for (int c = 0; c < NS; ++c) r[c] = block_sum(a[c], sh);
    if (tid == 0) *ticket = 0u;
    return tid == 0;
}

// ---- mailbox all-reduce over NVLink peer memory ---------------------------------------------------------
// record (slot, src) = kMailRec doubles {kBlkSums sums, tag, pad}; tag = sequence number of the collective
AK_DEV void st_release_sys_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
AK_DEV unsigned long long ld_acquire_sys_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
// ONE thread: deposit this rank's `ns` partial sums in every rank's mailbox (own included)
AK_DEV void mail_post(const P2PDev& pd, unsigned long long seq, const double* vals, int ns) {
    const int slot = (int)(seq % kMailSlots);
    __threadfence_system();  // ghost rows pushed by this grid are visible before the record is
    for (int q = 0; q < pd.nranks; ++q) {
        double* rec = pd.mail_peer[q] + ((size_t)slot * pd.nranks + pd.rank) * kMailRec;
        for (int c = 0; c < ns; ++c) rec[c] = vals[c];
    }
    __threadfence_system();
    for (int q = 0; q < pd.nranks; ++q) {
        double* rec = pd.mail_peer[q] + ((size_t)slot * pd.nranks + pd.rank) * kMailRec;
        st_release_sys_u64(reinterpret_cast<unsigned long long*>(rec + kBlkSums), seq);
    }
}
// Whole block: wait until every rank's record `seq` has arrived in the local mailbox and add them in rank
// order (bit-identical on every rank).  Result in out_s[0..ns) (SHARED memory, visible to the whole block on return).
// `shm` >= kMaxPeers * kBlkSums doubles.
// A record that never arrives (a peer fell out of step) raises the mapped error flag AND the device `stop` flag of the
// solve (`stop_w`, may be null), so that every later kernel of the solve is a no-op instead of spinning again.
AK_DEV void mail_wait_sum(const P2PDev& pd, unsigned long long seq, double* out_s, int ns, double* shm,
                          int* stop_w = nullptr) {
    const int tid = threadIdx.x + threadIdx.y * blockDim.x;
    const int slot = (int)(seq % kMailSlots);
    if (tid < pd.nranks) {
        const double* rec = pd.mail_local + ((size_t)slot * pd.nranks + tid) * kMailRec;
        const unsigned long long* tag = reinterpret_cast<const unsigned long long*>(rec + kBlkSums);
        const long long t0 = clock64();
        while (ld_acquire_sys_u64(tag) != seq) {
            if (clock64() - t0 > pd.spin_cycles) {  // a peer never produced this record: flag it, do not hang
                *pd.err = 1;
                if (stop_w != nullptr) *stop_w = 1;
                break;
            }
        }
        for (int c = 0; c < ns; ++c) shm[kBlkSums * tid + c] = __ldcv(rec + c);
    }
    __syncthreads();
    if (tid < ns) {
        double a = 0.0;
        for (int q = 0; q < pd.nranks; ++q) a += shm[kBlkSums * q + tid];
        out_s[tid] = a;
    }
    __syncthreads();
}

// Coefficients of one block of the blocked Gram-Schmidt sweep.
//   t[b]            = <S_b, w>            raw projections of this pass (b < m), S_b = stored vector of basis vector v_b
//   gram[b * kBlkMax + a] = <S_b, S_a>    for a < b in the same block — measured once, by the final pass of the
//                                         iteration that finished S_b, and cached (they never change afterwards)
//   rho[b]          : S_b = rho[b] v_b    (un-normalised basis: V_k is the finished w of iteration k, rho_k = ||w||;
//                                         gmres! step 8, V[k+1] = w / Hbis, is never materialised); nullptr: rho = 1
// Modified Gram-Schmidt coefficients:  h_b = <v_b, w - sum_{a<b} h_a v_a> = <v_b,w> - sum_{a<b} h_a <v_b,v_a>, and
// c_b = h_b / rho[b] is what multiplies the stored vector in the update w -= sum_b c_b S_b.
// One definition shared by the vector kernels and the Givens kernel so that both see the same bits.
// Executed by ONE WARP: the scalings <v_b,v_a> = <S_b,S_a> / rho_b / rho_a (two IEEE divisions per entry, up to 28
// entries) are independent and go to the lanes; only the forward substitution (m(m-1)/2 multiply-subtract pairs in the
// serial order) runs on lane 0.  A single thread doing all of it — what every thread of every pass used to do at kernel
// start — took ~12 000 cycles per block: 6 us in front of each of the ~200 passes of a step, 19 us per Givens kernel.
// Results: s_h[b] = h_b, s_c[b] = c_b (zero for b >= m), in shared memory; the caller synchronises the block.
AK_DEV void block_coefficients_warp(int lane, const double* t, const double* gram, const double* rho, int m,
                                    double* s_g /* kBlkMax * kBlkMax */, double* s_h /* kBlkMax */,
                                    double* s_c /* kBlkMax */) {
    for (int q = lane; q < m * kBlkMax; q += 32) {
        const int b = q / kBlkMax, a = q % kBlkMax;
        if (a < b) {
            double g = gram[q];
            if (rho) g = __ddiv_rn(__ddiv_rn(g, rho[b]), rho[a]);
            s_g[q] = g;
        }
    }
    if (lane < kBlkMax) s_h[lane] = lane < m ? (rho ? __ddiv_rn(t[lane], rho[lane]) : t[lane]) : 0.0;
    __syncwarp();
    if (lane == 0) {
        for (int b = 1; b < m; ++b) {
            double acc = s_h[b];
            for (int a = 0; a < b; ++a) acc = __dsub_rn(acc, __dmul_rn(s_h[a], s_g[b * kBlkMax + a]));
            s_h[b] = acc;
        }
    }
    __syncwarp();
    if (lane < kBlkMax) s_c[lane] = lane < m ? (rho ? __ddiv_rn(s_h[lane], rho[lane]) : s_h[lane]) : 0.0;
    __syncwarp();
}

// Division by a loop-invariant divisor d with r = RN(1/d):  q = RN(a r); q' = RN(q + r (a - q d)).
// With the remainder formed exactly by FMA this is the correctly rounded a/d (Markstein), i.e. bit-identical
// to the IEEE division the Julia source performs, at 3 flops instead of the ~12-instruction div.rn.f64
// sequence (checked against a/d on 9e8 random operands incl. every dx^2 = 1/(N+1)^2, N < 1000: 0 mismatches).
// The proof excludes divisors whose significand is all ones; those take the true division.
struct Divisor {
    double d, r;
    int slow;
};
AK_DEV bool all_ones_significand(double d) {
    return (__double_as_longlong(d) & 0x000FFFFFFFFFFFFFll) == 0x000FFFFFFFFFFFFFll;
}
AK_DEV Divisor make_divisor(double d) {
    Divisor v;
    v.d = d;
    v.r = __ddiv_rn(1.0, d);
    v.slow = all_ones_significand(d) || !(fabs(d) > 1e-290 && fabs(d) < 1e290);
    return v;
}
// The same constants formed on the host (IEEE division: the same bits as __ddiv_rn), so that the loop-invariant grid
// spacings cost the kernels nothing: a division per thread was ~30 % of the instructions of the one-chunk 1-D kernels.
inline Divisor make_divisor_host(double d) {
    Divisor v;
    v.d = d;
    v.r = 1.0 / d;
    long long bits;
    memcpy(&bits, &d, sizeof(bits));
    const double ad = d < 0 ? -d : d;
    v.slow = ((bits & 0x000FFFFFFFFFFFFFll) == 0x000FFFFFFFFFFFFFll) || !(ad > 1e-290 && ad < 1e290);
    return v;
}
AK_DEV double div_by(double a, const Divisor& v) {
    if (v.slow) return __ddiv_rn(a, v.d);
    const double q = __dmul_rn(a, v.r);
    const double rem = __fma_rn(-q, v.d, a);
    return __fma_rn(rem, v.r, q);
}

// second difference in the reference's association: ((e - 2c) + w) / d2
AK_DEV double second_diff(double e, double c, double w, const Divisor& d2) {
    return div_by(__dadd_rn(__dsub_rn(e, __dmul_rn(2.0, c)), w), d2);
}

}  // namespace ak
