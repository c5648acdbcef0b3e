// One sweep per GMRES iteration (fuse = AK_FUSE_SWEEP): k_sweep for the 2-D five-point problems, k_sweep1d for the 1-D
// three-point problems and DG.
//
// Replaces, per iteration k of Krylov.jl's gmres! (called at src/Ariadne.jl:338), the whole list
//   kaxpy!/kdot x k, knorm, V[k+1] = w / Hbis, mul!(w, J, V[k+1])          
// by ONE pass over the Krylov basis:
//   z   = W * (1/rho_{k-1}) - sum_{j<k} c_j S_j        (the Gram-Schmidt update; S_j = rho_j v_j are the stored,
//                                                       un-normalised basis vectors, W = J S_{k-1} the raw tangent)
//   ||z||^2, g_j = <S_j, z>                            (rho_k and row k of the Gram matrix: the loss of orthogonality)
//   y   = J z                                          (the tangent of the NEXT iteration, examples/bratu.jl:14-24 in
//                                                       2-D, examples/heat_2D.jl:53-60 behind G_Euler!)
//   t_j = <S_j, y> (j < k), <z, y>                     (all projections of the next iteration)
// Every basis vector is read ONCE per iteration (the blocked sweeps read it twice): 8n(k + 4) bytes instead of
// 8n(2.25k + 6).  The modified-Gram-Schmidt coefficients follow from the raw projections by forward substitution with
// the cached Gram matrix (k_gmres_sweep_scalar, krylov.cu): h_j = <v_j,w> - sum_{a<j} h_a <v_j,v_a>, the same identity
// the blocked sweeps use inside a block, with the block = the whole restart cycle.
//
// The stencil needs z on the neighbouring rows and columns, so a block recomputes the update on a one-row / two-column
// rim of its part of the grid.  Data movement is TMA: the rows of S_0..S_{k-1}, W and lambda e^u are streamed into a ring
// of shared-memory slots with cp.async.bulk (an mbarrier pair per slot); the eight warps of a block are independent
// (each owns up to 30 columns, x-neighbours by shuffle), take turns at issuing the copies, keep the sums in registers
// and the previous row of every S_j as the delay line the projections of y need.  Strips of neighbouring blocks advance
// side by side, so the rim columns they share come out of L2 for the second reader.  How the kernel got from 56 % to
// 90 % of the copy bandwidth: profiles/r02_sweep_tuning.md.
#include <math.h>
#include <stdlib.h>

#include <atomic>

#include "ak_internal.h"
#include "common.cuh"
#include "sweep.h"

namespace ak {

constexpr int kSwThreads = 256;                  // 8 warps; every warp both streams (in turn) and computes
constexpr int kSwWarps = kSwThreads / 32;
constexpr int kSwTW = 30;                        // columns a warp owns at most (32 lanes = 30 + one rim lane per side)
constexpr int kSwHalo = 2;                       // rim columns staged per side (two, so that every bulk copy is 16-byte aligned)
constexpr int kSwTXI = kSwWarps * kSwTW;         // 240: widest strip a block owns
constexpr int kSwTX = kSwTXI + 2 * kSwHalo;      // 244 doubles per staged vector row (1952 bytes: a multiple of 16)
constexpr int kSwMaxSlots = 8;
constexpr int kSwMaxSeg = 8;                     // segments (strip, row range) one block may be given
// shared-memory map (bytes): barriers | coefficients, flag | segment table | reduction scratch | slots
constexpr int kSwOffCoef = 128;
constexpr int kSwOffSeg = 512;
constexpr int kSwOffRed = kSwOffSeg + 512;
constexpr int kSwOffSlots = ((kSwOffRed + kSwWarps * kSwSums * 8 + 127) / 128) * 128;
constexpr int kSwSmemMax = 227 * 1024;

enum { SW_OP_BRATU = 0, SW_OP_HEAT = 1 };

struct SweepArgs {
    int64_t nx, ny;
    int32_t k;          // basis vectors subtracted and projected on
    int32_t nslot;      // ring depth
    int32_t wrap_x;     // periodic in x
    int32_t coef_from_u;
    int32_t txi;        // columns a strip owns (even, <= kSwTXI)
    int32_t tw;         // columns a warp owns: ceil(txi / 8)
    int32_t team_strips;  // > 0: block b owns strip b % team_strips, row band b / team_strips of team_bands (neighbouring
    int32_t team_bands;   //      strips run side by side, so their shared rim sectors hit L2); 0: even split of (strip, row)
    int32_t slack;        // a staged row is issued `nslot - slack` steps ahead of its use, `slack` steps after its slot's last use
    const double* S[kSwKMax];
    const double* S_lo[kSwKMax];  // ghost row y = -1 of S_j (nullptr: zero)
    const double* S_hi[kSwKMax];  // ghost row y = ny
    const double* zin;            // W (or S_0 for the opening sweep of a cycle)
    const double* zin_lo;
    const double* zin_hi;
    const double* coef;           // lambda e^u (or u): Bratu
    double* zout;                 // S_k (nullptr: not stored)
    double* yout;                 // J z
    const double* cvec;           // device: c_j, j < k
    const double* in_scale;       // device scalar multiplying zin (nullptr: 1)
    Divisor dx2d, dy2d;
    double a, c1, lambda;
    double* sums_out;             // kSwSums doubles (layout: sweep.h)
    double* partials;
    unsigned int* ticket;
    const int* stop;
    // multi-GPU (peer memory): boundary rows of z and y go straight into the neighbours' ghost rows
    double* push_z_down;          // neighbour below: its ghost row y = ny of zout's slot
    double* push_z_up;            // neighbour above: its ghost row y = -1
    double* push_y_down;
    double* push_y_up;
    P2PDev pd;
    unsigned long long seq_out;   // != 0: the sums travel through the ranks' sweep mailboxes
    double* swmail_peer[kMaxPeers];
    // 1-D problems (k_sweep1d): nx = unknowns, ny = 1; S_lo / S_hi / zin_lo / zin_hi point at `halo` ghost values
    int32_t halo;                 // rim values staged per side: 2 (three-point stencils), 4 (DG: one element)
    int32_t wrap;                 // periodic (DG on one GPU): the rim of the first / last chunk is the other end of the vector
    int32_t seg_first, seg_last;  // heat 1-D: this rank holds the global first / last point (bc! acts there); 0 otherwise
    double D[4][4];               // DG: LGL derivative matrix, 2/h, (h/2) w_edge
    double jac;
    Divisor mwd;
};

// ---- mbarrier / bulk-copy PTX ------------------------------------------------------------------------------------
AK_DEV uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
AK_DEV void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
AK_DEV void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
AK_DEV void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
AK_DEV bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
// bounded wait: a pipeline that never completes is a bug; trap instead of hanging the GPU (cold path, not inlined)
__device__ __noinline__ void mbar_wait_slow(uint32_t bar, uint32_t parity) {
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 4000000000ll) __trap();
    }
}
AK_DEV void mbar_wait(uint32_t bar, uint32_t parity) {
    if (!mbar_try_wait(bar, parity)) mbar_wait_slow(bar, parity);
}
AK_DEV void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}

// resident blocks per SM the register budget is sized for (8 warps: 255 / 128 / 80 registers per thread)
constexpr int sw_min_blocks(int kb) { return kb <= 4 ? 3 : (kb <= 12 ? 2 : 1); }

// Deterministic grid reduction of the 2 KB + 2 sums of a sweep and their hand-over: one partial per block, the block
// that takes the last ticket adds them in index order and either stores the totals or posts them to every rank's sweep
// mailbox (peer memory).  Called by all threads of every block.
template <int KB>
AK_DEV void sweep_reduce(const SweepArgs& p, int k, double acc_n, double acc_zy, const double (&acc_g)[KB],
                         const double (&acc_t)[KB], double* s_red, int* s_last_p) {
    constexpr int NS = 2 * KB + 2;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double vals[NS];
    vals[0] = acc_n;
    vals[1] = acc_zy;
#pragma unroll
    for (int j = 0; j < KB; ++j) {
        vals[2 + j] = acc_g[j];
        vals[2 + KB + j] = acc_t[j];
    }
#pragma unroll
    for (int c = 0; c < NS; ++c) {
        const double v = warp_sum(vals[c]);
        if (lane == 0) s_red[warp * NS + c] = v;
    }
    __syncthreads();
    const int nblocks = gridDim.x, bid = blockIdx.x;
    if (tid < NS) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < kSwWarps; ++w) s += s_red[w * NS + tid];
        p.partials[(size_t)bid * NS + tid] = s;
        __threadfence();
    }
    __syncthreads();
    if (tid == 0) {
        if (p.seq_out != 0) __threadfence_system();  // rows pushed to the neighbours precede the record
        const unsigned int tk = atomicAdd(p.ticket, 1u);
        *s_last_p = (tk == (unsigned int)(nblocks - 1));
    }
    __syncthreads();
    if (!*s_last_p) return;
    __threadfence();
    // four threads per sum, each over a contiguous quarter of the blocks; quarters added in order
    {
        const int c = tid >> 2, part = tid & 3;
        double s = 0.0;
        if (c < NS) {
            const int b0 = nblocks * part / 4, b1 = nblocks * (part + 1) / 4;
            for (int b = b0; b < b1; ++b) s += __ldcg(p.partials + (size_t)b * NS + c);
        }
        const int gl = lane & ~3;
        const double q0 = __shfl_sync(0xffffffffu, s, gl), q1 = __shfl_sync(0xffffffffu, s, gl + 1);
        const double q2 = __shfl_sync(0xffffffffu, s, gl + 2), q3 = __shfl_sync(0xffffffffu, s, gl + 3);
        if (c < NS && part == 0) s_red[c] = ((q0 + q1) + q2) + q3;
    }
    __syncthreads();
    if (tid < NS) {
        // layout of sums_out (sweep.h): [0] ||z||^2, [1] <z,y>, [2 + j] g_j, [2 + kSwKMax + j] t_j
        const int dst = tid < 2 ? tid : (tid < 2 + KB ? tid : tid - KB + kSwKMax);
        const bool used = tid < 2 || (tid < 2 + KB ? tid - 2 < k : tid - 2 - KB < k);
        if (used && p.seq_out == 0) p.sums_out[dst] = s_red[tid];
        if (p.seq_out != 0) {
            // every rank's record goes into every rank's sweep mailbox; the scalar kernel adds them in rank order
            const int slot = (int)(p.seq_out % kMailSlots);
            for (int q = 0; q < p.pd.nranks; ++q) {
                double* rec = p.swmail_peer[q] + ((size_t)slot * p.pd.nranks + p.pd.rank) * kSwMailRec;
                rec[dst] = used ? s_red[tid] : 0.0;
            }
            __threadfence_system();
        }
    }
    __syncthreads();
    if (tid == 0) {
        if (p.seq_out != 0) {
            __threadfence_system();
            const int slot = (int)(p.seq_out % kMailSlots);
            for (int q = 0; q < p.pd.nranks; ++q) {
                double* rec = p.swmail_peer[q] + ((size_t)slot * p.pd.nranks + p.pd.rank) * kSwMailRec;
                st_release_sys_u64(reinterpret_cast<unsigned long long*>(rec + kSwSums), p.seq_out);
            }
        }
        *p.ticket = 0u;
    }
}

struct SwSeg {       // one (strip, row range) of a block; rows are consumed in the order of the table
    int32_t c0;      // first staged column (strip start - rim), may be -2
    int32_t r0, r1;  // rows owned
    uint32_t f0;     // index of the segment's first staged row in the block's row sequence
};

// One sweep.  KB: compile-time bound on k (register-resident sums); STENCIL: y = J z is formed and projected.
//
// Rows flow through a ring of `nslot` shared-memory slots, one mbarrier pair per slot: `full` (armed with the byte count,
// completed by the bulk copies) and `empty` (one arrival per warp when it is done with the slot).  The WARPS ARE
// INDEPENDENT: warp w owns `tw` columns of the strip and recomputes z on one rim column per side (x-neighbours by
// shuffle), so no block barrier sits between the phases of a row.  Staging duty rotates: staged row f is issued by warp
// f % 8 a few steps after all warps released its slot, lane v streaming vector v.
// The row loop is written for instruction count: at basis size 20 a lane needs 60 DFMA and 50 LDS per point, and the
// first versions spent 5x that on bookkeeping (per-j predicates, integer divisions for slot and parity, null checks of
// the push pointers, 64-bit index arithmetic) and were issue-bound at 55-70 % of the bandwidth (ncu: 409 instructions
// per warp and row, stall_wait 41 %; profiles/r02_sweep_tuning.md).  Now: slot positions k..KB-1 of every slot are zero
// padding so that all loops over j are unpredicated with immediate offsets, slot / parity / row pointers are running
// values, and the row conditions are three compares on the loop counter.
template <int KB, bool STENCIL, int OP>
__global__ void __launch_bounds__(kSwThreads, sw_min_blocks(KB)) k_sweep(const SweepArgs p) {
    extern __shared__ __align__(128) unsigned char sw_smem[];
    if (p.stop != nullptr && *p.stop != 0) return;
    constexpr int NS = 2 * KB + 2;
    constexpr bool HASCOEF = STENCIL && OP == SW_OP_BRATU;
    // slot layout: S_j at vector position j < KB (k..KB-1 zero), zin at KB, lambda e^u at KB + 1
    constexpr int NVEC = KB + 1 + (HASCOEF ? 1 : 0);
    constexpr int SLOTD = NVEC * kSwTX;  // doubles per slot
    uint64_t* bar_full = reinterpret_cast<uint64_t*>(sw_smem);
    uint64_t* bar_empty = bar_full + kSwMaxSlots;
    double* s_c = reinterpret_cast<double*>(sw_smem + kSwOffCoef);
    int* s_last_p = reinterpret_cast<int*>(sw_smem + kSwOffCoef + kSwKMax * 8);
    int* s_nseg = s_last_p + 1;
    uint32_t* s_nfill = reinterpret_cast<uint32_t*>(s_last_p + 2);
    SwSeg* s_seg = reinterpret_cast<SwSeg*>(sw_smem + kSwOffSeg);
    double* s_red = reinterpret_cast<double*>(sw_smem + kSwOffRed);
    double* slots = reinterpret_cast<double*>(sw_smem + kSwOffSlots);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int k = p.k, nslot = p.nslot;
    const int64_t nx = p.nx, ny = p.ny;
    const int nvec = k + 1 + (HASCOEF ? 1 : 0);  // vectors streamed per row
    const int txi = p.txi, tw = p.tw;
    if (tid == 0) {
        for (int s = 0; s < nslot; ++s) {
            mbar_init(bar_full + s, 1);
            mbar_init(bar_empty + s, kSwWarps);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        // this block's segments
        int ns = 0;
        uint32_t f = 0;
        const int halo_rows = STENCIL ? 2 : 0;
        if (p.team_strips > 0) {
            const int strip = (int)(blockIdx.x % (unsigned)p.team_strips), band = (int)(blockIdx.x / (unsigned)p.team_strips);
            const int64_t r0 = ny * band / p.team_bands, r1 = ny * (band + 1) / p.team_bands;
            if (r1 > r0) {
                s_seg[0] = SwSeg{strip * txi - kSwHalo, (int32_t)r0, (int32_t)r1, 0u};
                f = (uint32_t)(r1 - r0) + halo_rows;
                ns = 1;
            }
        } else {
            const int64_t nstrip = (nx + txi - 1) / txi;
            const int64_t total = nstrip * ny;
            int64_t u = total * (int64_t)blockIdx.x / (int64_t)gridDim.x;
            const int64_t u_end = total * ((int64_t)blockIdx.x + 1) / (int64_t)gridDim.x;
            while (u < u_end && ns < kSwMaxSeg) {
                const int64_t strip = u / ny, r0 = u - strip * ny;
                const int64_t r1 = (r0 + (u_end - u) < ny) ? r0 + (u_end - u) : ny;
                s_seg[ns] = SwSeg{(int32_t)(strip * txi - kSwHalo), (int32_t)r0, (int32_t)r1, f};
                f += (uint32_t)(r1 - r0) + halo_rows;
                u += r1 - r0;
                ++ns;
            }
        }
        *s_nseg = ns;
        *s_nfill = f;
    }
    if (tid < KB) s_c[tid] = tid < k ? p.cvec[tid] : 0.0;
    for (int s = 0; s < nslot; ++s)  // zero padding vectors (never touched by the bulk copies)
        for (int q = tid; q < (KB - k) * kSwTX; q += kSwThreads) slots[(size_t)s * SLOTD + (size_t)k * kSwTX + q] = 0.0;
    const double sin = p.in_scale != nullptr ? *p.in_scale : 1.0;
    __syncthreads();
    const int nseg = *s_nseg;
    const uint32_t nfill = *s_nfill;
    const uint32_t bar_full0 = smem_u32(bar_full), bar_empty0 = smem_u32(bar_empty), slots0 = smem_u32(slots);

    double acc_n = 0.0, acc_zy = 0.0;
    double acc_g[KB], acc_t[KB];
#pragma unroll
    for (int j = 0; j < KB; ++j) acc_g[j] = acc_t[j] = 0.0;

    // ---- staging: copies of staged row f (cold path; lane v streams vector v) ----
    auto issue = [&](uint32_t f) {
        int sg = 0;
        while (sg + 1 < nseg && s_seg[sg + 1].f0 <= f) ++sg;
        const SwSeg g = s_seg[sg];
        const int64_t rbeg = STENCIL ? (int64_t)g.r0 - 1 : (int64_t)g.r0;
        const int64_t row = rbeg + (int64_t)(f - g.f0);
        const int64_t c0 = g.c0;
        const int64_t cend = c0 + txi + 2 * kSwHalo;
        const int64_t cs = c0 < 0 ? 0 : c0, ce = cend < nx ? cend : nx;
        const uint32_t nb = (uint32_t)(ce - cs) * 8u;
        const int soff = (int)(cs - c0);
        const bool wrap_l = p.wrap_x && c0 < 0;      // columns -2, -1 are columns nx-2, nx-1
        const bool wrap_r = p.wrap_x && cend > nx;   // columns nx, nx+1 are columns 0, 1
        const uint32_t vec_bytes = nb + (wrap_l ? 16u : 0u) + (wrap_r ? 16u : 0u);
        const uint32_t slot = f % (uint32_t)nslot;
        const bool inside = row >= 0 && row < ny;
        const bool rowdata = inside || (row < 0 ? p.zin_lo != nullptr : p.zin_hi != nullptr);
        if (f >= (uint32_t)nslot) mbar_wait(bar_empty0 + 8u * slot, ((f / (uint32_t)nslot) - 1u) & 1u);  // every warp released the slot
        const uint32_t bar = bar_full0 + 8u * slot;
        if (lane == 0) {
            const uint32_t nv = rowdata ? (uint32_t)(k + 1) + ((HASCOEF && inside) ? 1u : 0u) : 0u;
            if (nv != 0) mbar_arrive_expect_tx(bar, nv * vec_bytes);
            else mbar_arrive(bar);
        }
        __syncwarp();
        if (lane < nvec && rowdata) {
            const double* src = nullptr;
            if (lane < k) src = row < 0 ? p.S_lo[lane] : (row >= ny ? p.S_hi[lane] : p.S[lane] + row * nx);
            else if (lane == k) src = row < 0 ? p.zin_lo : (row >= ny ? p.zin_hi : p.zin + row * nx);
            else if (HASCOEF && inside) src = p.coef + row * nx;
            if (src != nullptr) {
                const int pos = lane < k ? lane : KB + (lane - k);
                const uint32_t dst = slots0 + 8u * (slot * (uint32_t)SLOTD + (uint32_t)pos * kSwTX);
                bulk_g2s(dst + 8u * soff, src + cs, nb, bar);
                if (wrap_l) bulk_g2s(dst, src + (nx - 2), 16u, bar);
                if (wrap_r) bulk_g2s(dst + 8u * (uint32_t)(nx - c0), src, 16u, bar);
            }
        }
    };
    // staged row f is issued by warp f % 8 when that warp starts its step f - lead; the first rows up front
    const uint32_t lead = (uint32_t)(nslot - p.slack);
    uint32_t fnext = (uint32_t)warp;  // next staged row this warp issues
    for (; fnext < lead && fnext < nfill; fnext += kSwWarps) issue(fnext);

    {
        const Divisor dx2 = p.dx2d, dy2 = p.dy2d;
        const int t = kSwHalo - 1 + tw * warp + lane;  // staged column of this lane: rim, tw owned columns, rim
        const bool pushing = p.push_z_down != nullptr || p.push_z_up != nullptr || p.push_y_down != nullptr || p.push_y_up != nullptr;
        double sd[KB];                                 // the row of every S_j the sums of this step need
#pragma unroll
        for (int j = 0; j < KB; ++j) sd[j] = 0.0;
        uint32_t cnt = 0;                              // staged rows consumed so far
        uint32_t slot = 0, par = 0;                    // = cnt % nslot, (cnt / nslot) & 1
        const double* base = slots + t;                // this lane's column in the current slot
        for (int sg = 0; sg < nseg; ++sg) {
            const SwSeg g = s_seg[sg];
            const int64_t c0 = g.c0, r0 = g.r0, r1 = g.r1;
            const int64_t gc = c0 + t;
            const bool lane_on = lane <= tw + 1 && t < txi + 2 * kSwHalo;
            // columns whose z is meaningful: the grid, plus the periodic images next to it
            const bool zvalid = lane_on && ((gc >= 0 && gc < nx) || (p.wrap_x && gc >= -kSwHalo && gc < nx + kSwHalo));
            const bool interior = lane_on && lane >= 1 && lane <= tw && t >= kSwHalo && t < kSwHalo + txi && gc < nx;
            const int nown = (int)(r1 - r0);
            const int nrows = STENCIL ? nown + 2 : nown;
            const int first_own = STENCIL ? 1 : 0;   // loop index of row r0
            // rim rows beyond a physical boundary carry no data: z = 0 there
            const int zero_top = (STENCIL && r0 == 0 && p.zin_lo == nullptr) ? 0 : -1;
            const int zero_bot = (STENCIL && r1 == ny && p.zin_hi == nullptr) ? nrows - 1 : -1;
            double* zrow = p.zout != nullptr ? p.zout + r0 * nx + gc : nullptr;   // row r0 + (i - first_own)
            double* yrow = STENCIL ? p.yout + r0 * nx + gc : nullptr;            // row r0 + (i - 2)
            double z1 = 0.0, z2 = 0.0, coefd = 0.0;
            for (int i = 0; i < nrows; ++i) {
                if (cnt + lead == fnext) {  // staging duty of this step
                    if (fnext < nfill) issue(fnext);
                    fnext += kSwWarps;
                }
                mbar_wait(bar_full0 + 8u * slot, par);
                const bool own_row = (unsigned)(i - first_own) < (unsigned)nown;
                double zt = __dmul_rn(base[KB * kSwTX], sin);
                if (!STENCIL) {
#pragma unroll
                    for (int j = 0; j < KB; ++j) {
                        sd[j] = base[j * kSwTX];
                        zt = fma(-s_c[j], sd[j], zt);  // same order as successive kaxpy!
                    }
                    if (interior) {
                        acc_n = fma(zt, zt, acc_n);
#pragma unroll
                        for (int j = 0; j < KB; ++j) acc_g[j] = fma(sd[j], zt, acc_g[j]);
                        if (zrow != nullptr) *zrow = zt;
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < KB; ++j) zt = fma(-s_c[j], base[j * kSwTX], zt);
                    if (!zvalid || i == zero_top || i == zero_bot) zt = 0.0;
                    if (own_row && interior && zrow != nullptr) *zrow = zt;
                    // y on row q = r0 + i - 2: z1 = z(q), z2 = z(q - 1), zt = z(q + 1); x-neighbours sit in the neighbouring lanes
                    const double zl = __shfl_up_sync(0xffffffffu, z1, 1);
                    const double zr = __shfl_down_sync(0xffffffffu, z1, 1);
                    if (i >= 2 && interior) {
                        const double xx = second_diff(zr, z1, zl, dx2);
                        const double yy = second_diff(zt, z1, z2, dy2);
                        const double lap = __dadd_rn(xx, yy);
                        double y;
                        if (OP == SW_OP_BRATU) {
                            const double kc = p.coef_from_u ? __dmul_rn(p.lambda, exp(coefd)) : coefd;
                            y = __dadd_rn(lap, __dmul_rn(kc, z1));
                        } else {  // tangent of G_Euler! / G_Trapezoid! around diffusion!: c1 * (a * lap) - v
                            y = __dsub_rn(__dmul_rn(p.c1, __dmul_rn(p.a, lap)), z1);
                        }
                        *yrow = y;
                        acc_n = fma(z1, z1, acc_n);
                        acc_zy = fma(z1, y, acc_zy);
#pragma unroll
                        for (int j = 0; j < KB; ++j) {
                            acc_g[j] = fma(sd[j], z1, acc_g[j]);
                            acc_t[j] = fma(sd[j], y, acc_t[j]);
                        }
                        if (pushing) {  // slabs: the first / last row of y goes to the neighbours (rare rows)
                            const int64_t q = r0 + i - 2;
                            if (q == 0 && p.push_y_down != nullptr) p.push_y_down[gc] = y;
                            if (q == ny - 1 && p.push_y_up != nullptr) p.push_y_up[gc] = y;
                        }
                    }
                    // this row becomes the delay line of the next step
                    if (own_row && interior) {
#pragma unroll
                        for (int j = 0; j < KB; ++j) sd[j] = base[j * kSwTX];
                        if (HASCOEF) coefd = base[(KB + 1) * kSwTX];
                    }
                    z2 = z1;
                    z1 = zt;
                    if (i >= 2) yrow += nx;
                }
                if (pushing && own_row && interior) {
                    const int64_t row = r0 + i - first_own;
                    if (row == 0 && p.push_z_down != nullptr) p.push_z_down[gc] = zt;
                    if (row == ny - 1 && p.push_z_up != nullptr) p.push_z_up[gc] = zt;
                }
                if (own_row && zrow != nullptr) zrow += nx;
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_empty0 + 8u * slot);
                ++cnt;
                ++slot;
                base += SLOTD;
                if (slot == (uint32_t)nslot) { slot = 0; par ^= 1u; base = slots + t; }
            }
        }
    }
    __syncthreads();

    sweep_reduce<KB>(p, k, acc_n, acc_zy, acc_g, acc_t, s_red, s_last_p);
}


// ====================================================================================================================
// 1-D problems: Bratu 1-D (examples/bratu.jl:14-24), heat 1-D with bc! (examples/heat_1D.jl:12-37 behind G_Euler!), DG heat
// (examples/heat_1D_DG.jl:32-36).  Same pass, same ring, same scalar step; the "rows" are consecutive chunks of the
// vector, the stencil couples a point only to its chunk neighbours (rim values staged with the chunk), so z and y = J z
// of a point are formed in the same step and no delay line is needed.  8n(k + 4) bytes per iteration (k + 3 for the
// heat / DG tangents, which read no coefficient vector).
// ====================================================================================================================
enum { SW1_BRATU = 0, SW1_HEAT = 1, SW1_DG = 2 };

template <int KB, bool STENCIL, int OP>
__global__ void __launch_bounds__(kSwThreads, sw_min_blocks(KB)) k_sweep1d(const SweepArgs p) {
    extern __shared__ __align__(128) unsigned char sw_smem[];
    if (p.stop != nullptr && *p.stop != 0) return;
    constexpr bool HASCOEF = STENCIL && OP == SW1_BRATU;
    constexpr int H = OP == SW1_DG ? 4 : 2;            // rim values per side
    constexpr int NVEC = KB + 1 + (HASCOEF ? 1 : 0);
    constexpr int SLOTD = NVEC * kSwTX;
    uint64_t* bar_full = reinterpret_cast<uint64_t*>(sw_smem);
    uint64_t* bar_empty = bar_full + kSwMaxSlots;
    double* s_c = reinterpret_cast<double*>(sw_smem + kSwOffCoef);
    int* s_last_p = reinterpret_cast<int*>(sw_smem + kSwOffCoef + kSwKMax * 8);
    double* s_red = reinterpret_cast<double*>(sw_smem + kSwOffRed);
    double* slots = reinterpret_cast<double*>(sw_smem + kSwOffSlots);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int k = p.k, nslot = p.nslot;
    const int64_t n = p.nx;
    const int nvec = k + 1 + (HASCOEF ? 1 : 0);
    const int txi = p.txi, tw = p.tw;
    if (tid == 0) {
        for (int s = 0; s < nslot; ++s) {
            mbar_init(bar_full + s, 1);
            mbar_init(bar_empty + s, kSwWarps);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (tid < KB) s_c[tid] = tid < k ? p.cvec[tid] : 0.0;
    for (int s = 0; s < nslot; ++s)
        for (int q = tid; q < (KB - k) * kSwTX; q += kSwThreads) slots[(size_t)s * SLOTD + (size_t)k * kSwTX + q] = 0.0;
    const double sin = p.in_scale != nullptr ? *p.in_scale : 1.0;
    __syncthreads();
    const uint32_t bar_full0 = smem_u32(bar_full), bar_empty0 = smem_u32(bar_empty), slots0 = smem_u32(slots);
    // this block's chunks [c_begin, c_end)
    const int64_t nchunk = (n + txi - 1) / txi;
    const int64_t c_begin = nchunk * (int64_t)blockIdx.x / (int64_t)gridDim.x;
    const int64_t c_end = nchunk * ((int64_t)blockIdx.x + 1) / (int64_t)gridDim.x;
    const uint32_t nfill = (uint32_t)(c_end - c_begin);

    double acc_n = 0.0, acc_zy = 0.0;
    double acc_g[KB], acc_t[KB];
#pragma unroll
    for (int j = 0; j < KB; ++j) acc_g[j] = acc_t[j] = 0.0;

    // staging of chunk c_begin + f: values [c txi - H, c txi + txi + H) of every vector; beyond the ends of the vector the
    // rim comes from the ghost values (slabs), from the other end (periodic) or stays unused (physical boundary)
    auto issue = [&](uint32_t f) {
        const int64_t c = c_begin + (int64_t)f;
        const int64_t g0 = c * txi - H, g1 = g0 + txi + 2 * H;
        const int64_t cs = g0 < 0 ? 0 : g0, ce = g1 < n ? g1 : n;
        const uint32_t nb = (uint32_t)(ce - cs) * 8u;
        const bool rim_l = g0 < 0 && (p.wrap || p.zin_lo != nullptr);
        const bool rim_r = g1 > n && (p.wrap || p.zin_hi != nullptr);
        const uint32_t vec_bytes = nb + (rim_l ? 8u * H : 0u) + (rim_r ? 8u * H : 0u);
        const uint32_t slot = f % (uint32_t)nslot;
        if (f >= (uint32_t)nslot) mbar_wait(bar_empty0 + 8u * slot, ((f / (uint32_t)nslot) - 1u) & 1u);
        const uint32_t bar = bar_full0 + 8u * slot;
        if (lane == 0) mbar_arrive_expect_tx(bar, (uint32_t)(k + 1) * vec_bytes + (HASCOEF ? nb : 0u));
        __syncwarp();
        if (lane < nvec) {
            const bool iscoef = HASCOEF && lane == k + 1;
            const double* v = lane < k ? p.S[lane] : (lane == k ? p.zin : p.coef);
            const int pos = lane < k ? lane : KB + (lane - k);
            const uint32_t dst = slots0 + 8u * (slot * (uint32_t)SLOTD + (uint32_t)pos * kSwTX);
            bulk_g2s(dst + 8u * (uint32_t)(cs - g0), v + cs, nb, bar);
            if (!iscoef) {  // (the coefficient vector is only read at owned points)
                if (rim_l) bulk_g2s(dst, p.wrap ? v + (n - H) : (lane < k ? p.S_lo[lane] : p.zin_lo), 8u * H, bar);
                if (rim_r) bulk_g2s(dst + 8u * (uint32_t)(n - g0), p.wrap ? v : (lane < k ? p.S_hi[lane] : p.zin_hi), 8u * H, bar);
            }
        }
    };
    const uint32_t lead = (uint32_t)(nslot - p.slack);
    uint32_t fnext = (uint32_t)warp;
    for (; fnext < lead && fnext < nfill; fnext += kSwWarps) issue(fnext);

    {
        const Divisor dx2 = p.dx2d;
        // staged position of this lane: three-point stencils: rim, tw owned points, rim (H - 1 unused staged values in
        // front); DG: one rim element (4 lanes), tw = 24 owned nodes, one rim element
        const int t = (OP == SW1_DG ? 0 : H - 1) + tw * warp + lane;
        const int rimw = OP == SW1_DG ? 4 : 1;       // rim lanes per side
        const bool pushing = p.push_z_down != nullptr || p.push_z_up != nullptr || p.push_y_down != nullptr || p.push_y_up != nullptr;
        // DG: this lane's row of the derivative matrix
        double d0 = 0.0, d1 = 0.0, d2 = 0.0, d3 = 0.0;
        const int node = lane & 3;
        if (OP == SW1_DG) { d0 = p.D[node][0]; d1 = p.D[node][1]; d2 = p.D[node][2]; d3 = p.D[node][3]; }
        double sd[KB];
        uint32_t slot = 0, par = 0;
        const double* base = slots + t;
        for (uint32_t cnt = 0; cnt < nfill; ++cnt) {
            if (cnt + lead == fnext) {
                if (fnext < nfill) issue(fnext);
                fnext += kSwWarps;
            }
            const int64_t c = c_begin + (int64_t)cnt;
            const int64_t g0 = c * txi - H;
            const int64_t gi = g0 + t;                       // global index of this lane's value
            const int64_t own_end = (c + 1) * txi < n ? (c + 1) * txi : n;
            const bool lane_on = lane < tw + 2 * rimw && t < txi + 2 * H;
            const bool have_l = p.wrap || p.zin_lo != nullptr, have_r = p.wrap || p.zin_hi != nullptr;
            const bool zvalid = lane_on && ((gi >= 0 && gi < n) || (gi < 0 && have_l) || (gi >= n && gi < n + H && have_r));
            const bool interior = lane_on && lane >= rimw && lane < rimw + tw && gi >= c * txi && gi < own_end;
            mbar_wait(bar_full0 + 8u * slot, par);
            double zt = __dmul_rn(base[KB * kSwTX], sin);
#pragma unroll
            for (int j = 0; j < KB; ++j) {
                sd[j] = base[j * kSwTX];
                zt = fma(-s_c[j], sd[j], zt);  // same order as successive kaxpy!
            }
            if (!zvalid) zt = 0.0;
            // bc!(u) of the 1-D heat example: the two global end points are zero (heat_1D.jl:34-37), in the stored vector
            // too (seg_first / seg_last are only set for that problem; the update-only variant serves every problem)
            if ((gi == 0 && p.seg_first) || (gi == n - 1 && p.seg_last)) zt = 0.0;
            double y = 0.0;
            if (STENCIL) {
                if (OP == SW1_DG) {
                    // D1p: t1 = jac * (D u) per element, + (u_next0 - u_3) / mw at the last node
                    const int eb = lane & ~3;
                    const double u0 = __shfl_sync(0xffffffffu, zt, eb), u1 = __shfl_sync(0xffffffffu, zt, eb + 1);
                    const double u2 = __shfl_sync(0xffffffffu, zt, eb + 2), u3 = __shfl_sync(0xffffffffu, zt, eb + 3);
                    const double unext = __shfl_down_sync(0xffffffffu, zt, 1);
                    double sacc = __dmul_rn(d0, u0);
                    sacc = __dadd_rn(sacc, __dmul_rn(d1, u1));
                    sacc = __dadd_rn(sacc, __dmul_rn(d2, u2));
                    sacc = __dadd_rn(sacc, __dmul_rn(d3, u3));
                    double t1 = __dmul_rn(p.jac, sacc);
                    if (node == 3) t1 = __dadd_rn(t1, div_by(__dsub_rn(unext, zt), p.mwd));
                    // D1m: du = jac * (D t1), + (t1_0 - t1_prev3) / mw at the first node
                    const double q0 = __shfl_sync(0xffffffffu, t1, eb), q1 = __shfl_sync(0xffffffffu, t1, eb + 1);
                    const double q2 = __shfl_sync(0xffffffffu, t1, eb + 2), q3 = __shfl_sync(0xffffffffu, t1, eb + 3);
                    const double tprev = __shfl_up_sync(0xffffffffu, t1, 1);
                    double s2 = __dmul_rn(d0, q0);
                    s2 = __dadd_rn(s2, __dmul_rn(d1, q1));
                    s2 = __dadd_rn(s2, __dmul_rn(d2, q2));
                    s2 = __dadd_rn(s2, __dmul_rn(d3, q3));
                    double du = __dmul_rn(p.jac, s2);
                    if (node == 0) du = __dadd_rn(du, div_by(__dsub_rn(t1, tprev), p.mwd));
                    y = __dsub_rn(__dmul_rn(p.c1, du), zt);
                } else {
                    const double zl = __shfl_up_sync(0xffffffffu, zt, 1);
                    const double zr = __shfl_down_sync(0xffffffffu, zt, 1);
                    if (OP == SW1_BRATU) {
                        const double cf = base[(KB + 1) * kSwTX];
                        const double kc = p.coef_from_u ? __dmul_rn(p.lambda, exp(cf)) : cf;
                        y = __dadd_rn(second_diff(zr, zt, zl, dx2), __dmul_rn(kc, zt));
                    } else {
                        // heat_1D.jl:22: du[i] = a * (u[i+1] - 2u[i] + u[i-1]) / dx^2 ; du[1] = du[end] = 0
                        const bool bnd = (gi == 0 && p.seg_first) || (gi == n - 1 && p.seg_last);
                        const double du = bnd ? 0.0 : div_by(__dmul_rn(p.a, __dadd_rn(__dsub_rn(zr, __dmul_rn(2.0, zt)), zl)), dx2);
                        y = __dsub_rn(__dmul_rn(p.c1, du), zt);
                    }
                }
            }
            if (interior) {
                if (p.zout != nullptr) p.zout[gi] = zt;
                acc_n = fma(zt, zt, acc_n);
                if (STENCIL) {
                    p.yout[gi] = y;
                    acc_zy = fma(zt, y, acc_zy);
                }
#pragma unroll
                for (int j = 0; j < KB; ++j) {
                    acc_g[j] = fma(sd[j], zt, acc_g[j]);
                    if (STENCIL) acc_t[j] = fma(sd[j], y, acc_t[j]);
                }
                if (pushing) {  // segments: the first / last H values go to the neighbours' ghost values
                    if (gi < H) {
                        if (p.push_z_down != nullptr) p.push_z_down[gi] = zt;
                        if (STENCIL && p.push_y_down != nullptr) p.push_y_down[gi] = y;
                    }
                    if (gi >= n - H) {
                        if (p.push_z_up != nullptr) p.push_z_up[gi - (n - H)] = zt;
                        if (STENCIL && p.push_y_up != nullptr) p.push_y_up[gi - (n - H)] = y;
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_empty0 + 8u * slot);
            ++slot;
            base += SLOTD;
            if (slot == (uint32_t)nslot) { slot = 0; par ^= 1u; base = slots + t; }
        }
    }
    __syncthreads();
    sweep_reduce<KB>(p, k, acc_n, acc_zy, acc_g, acc_t, s_red, s_last_p);
}

// ---- launch ------------------------------------------------------------------------------------------------------
template <int KB, bool STENCIL, int OP>
static int launch_sweep_t(Ctx* ctx, SweepArgs& a) {
    // the opt-in to > 48 KB of dynamic shared memory is a per-device attribute of the kernel
    static std::atomic<unsigned long long> configured{0ull};
    const unsigned long long devbit = 1ull << (ctx->device & 63);
    if (!(configured.load(std::memory_order_relaxed) & devbit)) {
        AK_CUDA(cudaFuncSetAttribute(k_sweep<KB, STENCIL, OP>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSwSmemMax));
        configured.fetch_or(devbit, std::memory_order_relaxed);
    }
    const int nvec = KB + 1 + ((STENCIL && OP == SW_OP_BRATU) ? 1 : 0);  // slot positions (S_j padded to KB)
    const int slot_bytes = nvec * kSwTX * 8;
    int occ = sw_min_blocks(KB);
    if (const char* e = getenv("AK_SWEEP_OCC")) {  // tuning knob
        const int v = atoi(e);
        if (v >= 1 && v <= 4) occ = v;
    }
    int nslot = 0;
    for (; occ >= 1; --occ) {
        const int budget = kSwSmemMax / occ - 1024;  // the driver reserves 1 KB of shared memory per resident block
        nslot = (budget - kSwOffSlots) / slot_bytes;
        if (nslot >= 3) break;
    }
    if (nslot < 3) {
        set_error("launch_sweep: %d vectors do not fit the shared-memory ring", nvec);
        return AK_ERR_UNSUPPORTED;
    }
    if (nslot > kSwMaxSlots) nslot = kSwMaxSlots;
    if (const char* e = getenv("AK_SWEEP_SLOTS")) {  // tuning knob
        const int v = atoi(e);
        if (v >= 3 && v <= nslot) nslot = v;
    }
    a.nslot = nslot;
    a.slack = nslot >= 4 ? 2 : 1;  // three slots: refill right behind the last reader, or nothing would be in flight
    if (const char* e = getenv("AK_SWEEP_SLACK")) {  // tuning knob
        const int v = atoi(e);
        if (v >= 1 && v < nslot) a.slack = v;
    }
    const size_t smem = (size_t)kSwOffSlots + (size_t)nslot * slot_bytes;
    // Decomposition.  Preferred: `strips` x `bands` = one wave of resident blocks, block b on strip b % strips, so that
    // neighbouring strips advance side by side and the rim sectors they share are fetched from HBM once (L2 hit for the
    // second reader): without it the staged rows read 8 % more than they own.  Otherwise: an even split of the
    // (strip, row) space in strip-major order with the widest strips.
    int64_t grid = (int64_t)ctx->num_sms * occ;
    a.txi = kSwTXI;
    a.team_strips = a.team_bands = 0;
    static const int team_env = [] { const char* e = getenv("AK_SWEEP_TEAM"); return e ? atoi(e) : -1; }();  // tuning knob: 0 / 1
    // measured per class at 8192^2 (profiles/r02_sweep_tuning.md): side-by-side strips win wherever two or three blocks
    // share an SM and at KB = 20; at KB = 16 (one block per SM, six slots) the 8 % wider strips of the even split do
    const bool no_team = team_env == 0 || (team_env < 0 && KB == 16);
    if (!no_team) {
        const int64_t smin = (a.nx + kSwTXI - 1) / kSwTXI, smax = (a.nx + 159) / 160;
        for (int64_t st = smin; st <= smax && st <= grid; ++st) {
            if (grid % st != 0) continue;
            const int64_t bands = grid / st;
            if (a.ny < 16 * bands) break;  // too few rows per band: the rim rows would dominate
            int64_t txi = (a.nx + st - 1) / st;
            txi += txi & 1;
            if (txi > kSwTXI || (st - 1) * txi >= a.nx) continue;
            a.txi = (int32_t)txi;
            a.team_strips = (int32_t)st;
            a.team_bands = (int32_t)bands;
            break;
        }
    }
    a.tw = (a.txi + kSwWarps - 1) / kSwWarps;
    if (a.team_strips == 0) {
        const int64_t nstrip = (a.nx + a.txi - 1) / a.txi;
        const int64_t total = nstrip * a.ny;
        const int64_t by_work = (total + 7) / 8;  // at least ~8 rows per block
        if (grid > by_work) grid = by_work;
        if (nstrip > 6 * grid) grid = (nstrip + 5) / 6;  // a block's share spans at most kSwMaxSeg strips
        if (grid < 1) grid = 1;
    }
    ProfScope prof(ctx, PK_SWEEP);
    k_sweep<KB, STENCIL, OP><<<(int)grid, kSwThreads, smem, ctx->stream>>>(a);
    ctx->launches++;
    AK_CUDA(cudaGetLastError());
    return AK_OK;
}

template <int KB>
static int launch_sweep_kb(Ctx* ctx, SweepArgs& a, bool stencil, int op) {
    if (!stencil) return launch_sweep_t<KB, false, SW_OP_BRATU>(ctx, a);
    if (op == SW_OP_BRATU) return launch_sweep_t<KB, true, SW_OP_BRATU>(ctx, a);
    return launch_sweep_t<KB, true, SW_OP_HEAT>(ctx, a);
}

template <int KB, bool STENCIL, int OP>
static int launch_sweep1d_t(Ctx* ctx, SweepArgs& a) {
    static std::atomic<unsigned long long> configured{0ull};
    const unsigned long long devbit = 1ull << (ctx->device & 63);
    if (!(configured.load(std::memory_order_relaxed) & devbit)) {
        AK_CUDA(cudaFuncSetAttribute(k_sweep1d<KB, STENCIL, OP>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSwSmemMax));
        configured.fetch_or(devbit, std::memory_order_relaxed);
    }
    const int nvec = KB + 1 + ((STENCIL && OP == SW1_BRATU) ? 1 : 0);
    const int slot_bytes = nvec * kSwTX * 8;
    int occ = sw_min_blocks(KB);
    int nslot = 0;
    for (; occ >= 1; --occ) {
        const int budget = kSwSmemMax / occ - 1024;
        nslot = (budget - kSwOffSlots) / slot_bytes;
        if (nslot >= 3) break;
    }
    if (nslot < 3) {
        set_error("launch_sweep: %d vectors do not fit the shared-memory ring", nvec);
        return AK_ERR_UNSUPPORTED;
    }
    if (nslot > kSwMaxSlots) nslot = kSwMaxSlots;
    a.nslot = nslot;
    a.slack = nslot >= 4 ? 2 : 1;
    a.txi = OP == SW1_DG ? 192 : kSwTXI;  // DG: whole elements, 6 per warp
    a.tw = a.txi / kSwWarps;
    const size_t smem = (size_t)kSwOffSlots + (size_t)nslot * slot_bytes;
    const int64_t nchunk = (a.nx + a.txi - 1) / a.txi;
    int64_t grid = (int64_t)ctx->num_sms * occ;
    const int64_t by_work = (nchunk + 7) / 8;  // at least ~8 chunks per block
    if (grid > by_work) grid = by_work;
    if (grid < 1) grid = 1;
    ProfScope prof(ctx, PK_SWEEP);
    k_sweep1d<KB, STENCIL, OP><<<(int)grid, kSwThreads, smem, ctx->stream>>>(a);
    ctx->launches++;
    AK_CUDA(cudaGetLastError());
    return AK_OK;
}

template <int KB>
static int launch_sweep1d_kb(Ctx* ctx, SweepArgs& a, bool stencil, int op) {
    if (!stencil) return launch_sweep1d_t<KB, false, SW1_HEAT>(ctx, a);  // (the update alone does not depend on the problem)
    if (op == SW1_BRATU) return launch_sweep1d_t<KB, true, SW1_BRATU>(ctx, a);
    if (op == SW1_HEAT) return launch_sweep1d_t<KB, true, SW1_HEAT>(ctx, a);
    return launch_sweep1d_t<KB, true, SW1_DG>(ctx, a);
}

static inline bool is_sweep_2d(const ak_problem* p) { return p->kind == AK_BRATU2D || p->kind == AK_HEAT2D; }

bool sweep_supported(const Ctx* ctx, const ak_problem* p, const double* u) {
    const bool d2 = is_sweep_2d(p);
    const bool d1 = p->kind == AK_BRATU1D || p->kind == AK_HEAT1D || p->kind == AK_HEAT1D_DG;
    if (!d1 && !d2) return false;
    if (p->jvp_mode != AK_JVP_ANALYTIC || p->scheme == AK_MIDPOINT) return false;
    if (d2 && (p->nx < 4 || p->nx % 2 != 0 || p->ny < 1)) return false;  // bulk copies move 16-byte units
    if (d2 && (p->nx > (1 << 20) || p->ny > (1ll << 30))) return false;  // (segment table and partials buffer are sized for this)
    if (d1) {
        if (p->kind == AK_HEAT1D_DG ? (p->nx % 4 != 0 || p->nx < 8) : (p->nx % 2 != 0 || p->nx < 4)) return false;
        // periodic_bc! of the 1-D heat example copies the end points into one another in place: not reproduced here
        if (p->kind == AK_HEAT1D && p->bc == AK_BC_PERIODIC) return false;
    }
    if (p->kind == AK_BRATU2D || p->kind == AK_BRATU1D) {  // lambda e^u (or u itself) is streamed with bulk copies like the basis
        const double* cf = p->coef != nullptr ? p->coef : u;
        if (cf == nullptr || (reinterpret_cast<uintptr_t>(cf) & 15u)) return false;
    }
    if (ctx->nranks > 1) {
        // slabs / segments: the neighbours' boundary rows / values arrive through peer memory
        if (!ctx->p2p_on || (d2 ? p->nx : 8) > ctx->p2p_halo_cap) return false;
    }
    return true;
}

int launch_sweep(Ctx* ctx, const ak_problem* prob, const double* u, const SweepCall& c) {
    AK_REQUIRE(c.k >= 0 && c.k <= kSwKMax, "launch_sweep: k out of range");
    AK_REQUIRE(c.zin != nullptr && c.sums != nullptr, "launch_sweep: NULL operand");
    AK_REQUIRE(!c.stencil || c.yout != nullptr, "launch_sweep: the tangent needs a destination");
    const bool d2 = is_sweep_2d(prob);
    SweepArgs a{};
    a.nx = prob->nx;
    a.ny = d2 ? prob->ny : 1;
    a.k = c.k;
    a.wrap_x = d2 && (prob->bc == AK_BC_PERIODIC);
    bool ok16 = (reinterpret_cast<uintptr_t>(c.zin) & 15u) == 0;
    for (int j = 0; j < c.k; ++j) {
        a.S[j] = c.S[j];
        a.S_lo[j] = c.S_lo ? c.S_lo[j] : nullptr;
        a.S_hi[j] = c.S_hi ? c.S_hi[j] : nullptr;
        ok16 = ok16 && (reinterpret_cast<uintptr_t>(c.S[j]) & 15u) == 0;
    }
    a.zin = c.zin;
    a.zin_lo = c.zin_lo;
    a.zin_hi = c.zin_hi;
    if (d2 && ctx->nranks == 1 && prob->bc == AK_BC_PERIODIC) {  // one GPU: the ghost rows are the opposite rows
        for (int j = 0; j < c.k; ++j) { a.S_lo[j] = c.S[j] + (prob->ny - 1) * prob->nx; a.S_hi[j] = c.S[j]; }
        a.zin_lo = c.zin + (prob->ny - 1) * prob->nx;
        a.zin_hi = c.zin;
    }
    AK_REQUIRE(ok16, "launch_sweep: vectors must be 16-byte aligned");
    a.zout = c.zout;
    a.yout = c.stencil ? c.yout : nullptr;
    a.cvec = c.cvec;
    a.in_scale = c.in_scale;
    a.dx2d = make_divisor_host(prob->dx * prob->dx);
    a.dy2d = make_divisor_host(prob->dy * prob->dy);
    a.a = prob->a;
    a.c1 = (prob->scheme == AK_TRAPEZOID) ? prob->dt / 2.0 : prob->dt;
    a.lambda = prob->lambda;
    if (prob->kind == AK_BRATU2D || prob->kind == AK_BRATU1D) {
        a.coef = prob->coef ? prob->coef : u;
        a.coef_from_u = prob->coef ? 0 : 1;
        AK_REQUIRE(!c.stencil || (a.coef != nullptr && (reinterpret_cast<uintptr_t>(a.coef) & 15u) == 0),
                   "launch_sweep: lambda e^u / u must be a 16-byte aligned device vector");
    }
    a.sums_out = c.sums;
    a.partials = ctx->partials;
    a.ticket = ctx->ticket;
    a.stop = c.stop;
    a.push_z_down = c.push_z_down;
    a.push_z_up = c.push_z_up;
    a.push_y_down = c.stencil ? c.push_y_down : nullptr;
    a.push_y_up = c.stencil ? c.push_y_up : nullptr;
    a.seq_out = c.seq_out;
    if (c.seq_out != 0) {
        a.pd = ctx->p2p_dev();
        for (int q = 0; q < ctx->nranks && q < kMaxPeers; ++q) a.swmail_peer[q] = ctx->p2p_swmail_of(q);
    }
    const int k = c.k;
    int rc;
    if (d2) {
        const int op = prob->kind == AK_BRATU2D ? SW_OP_BRATU : SW_OP_HEAT;
        if (k <= 4) rc = launch_sweep_kb<4>(ctx, a, c.stencil, op);
        else if (k <= 8) rc = launch_sweep_kb<8>(ctx, a, c.stencil, op);
        else if (k <= 12) rc = launch_sweep_kb<12>(ctx, a, c.stencil, op);
        else if (k <= 16) rc = launch_sweep_kb<16>(ctx, a, c.stencil, op);
        else if (k <= 20) rc = launch_sweep_kb<20>(ctx, a, c.stencil, op);
        else rc = launch_sweep_kb<24>(ctx, a, c.stencil, op);
    } else {
        const int op = prob->kind == AK_BRATU1D ? SW1_BRATU : (prob->kind == AK_HEAT1D ? SW1_HEAT : SW1_DG);
        a.halo = op == SW1_DG ? 4 : 2;
        a.wrap = (op == SW1_DG && ctx->nranks == 1) ? 1 : 0;  // the DG mesh is periodic; on segments the ends are ghost values
        a.seg_first = op == SW1_HEAT && ctx->rank == 0;
        a.seg_last = op == SW1_HEAT && ctx->rank == ctx->nranks - 1;
        if (op == SW1_DG) {
            const double h = prob->dx, s5 = sqrt(5.0);
            const double da = (5.0 + 5.0 * s5) / 4.0, db = (5.0 - 5.0 * s5) / 4.0;
            const double dc = (1.0 + s5) / 4.0, dd = s5 / 2.0, de = (s5 - 1.0) / 4.0;
            const double M[4][4] = {{-3.0, da, db, 0.5}, {-dc, 0.0, dd, -de}, {de, -dd, 0.0, dc}, {-0.5, -db, -da, 3.0}};
            for (int i = 0; i < 4; ++i)
                for (int j = 0; j < 4; ++j) a.D[i][j] = M[i][j];
            a.jac = 2.0 / h;
            a.mwd = make_divisor_host((h / 2.0) * (1.0 / 6.0));
        }
        if (k <= 4) rc = launch_sweep1d_kb<4>(ctx, a, c.stencil, op);
        else if (k <= 8) rc = launch_sweep1d_kb<8>(ctx, a, c.stencil, op);
        else if (k <= 12) rc = launch_sweep1d_kb<12>(ctx, a, c.stencil, op);
        else if (k <= 16) rc = launch_sweep1d_kb<16>(ctx, a, c.stencil, op);
        else if (k <= 20) rc = launch_sweep1d_kb<20>(ctx, a, c.stencil, op);
        else rc = launch_sweep1d_kb<24>(ctx, a, c.stencil, op);
    }
    AK_TRY(rc);
    // NCCL fallback of the reduction is not offered: slabs take this path only with peer memory (sweep_supported)
    return AK_OK;
}

}  // namespace ak
