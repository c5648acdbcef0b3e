// One sweep per GMRES iteration (fuse = AK_FUSE_SWEEP) for the 2-D five-point problems.
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
// The stencil needs z on the neighbouring rows and columns, so a CTA recomputes the update on a one-row / two-column
// rim of its segment.  Data movement is TMA: a producer warp streams the rows of S_0..S_{k-1}, W and lambda e^u into a
// ring of shared-memory slots with cp.async.bulk (completion on an mbarrier per slot); 256 consumer threads own one
// column each, keep the sums in registers and the previous row of every S_j as the delay line the projections of y
// need.  The persistent grid splits the (strip, row) space evenly, so the rim costs 2 rows per ~1800.
#include <stdlib.h>

#include <atomic>

#include "ak_internal.h"
#include "common.cuh"
#include "sweep.h"

namespace ak {

constexpr int kSwTX = 256;                       // columns staged per strip (2 + 252 + 2)
constexpr int kSwHalo = 2;                       // rim columns per side (two, so that every bulk copy is 16-byte aligned)
constexpr int kSwTXI = kSwTX - 2 * kSwHalo;      // columns a strip owns
constexpr int kSwConsumers = 256;
constexpr int kSwThreads = kSwConsumers;         // every warp both streams and computes
constexpr int kSwWarps = kSwThreads / 32;
constexpr int kSwMaxSlots = 8;
// shared-memory map (bytes): barriers | coefficients | three rows of z | reduction scratch | slots
constexpr int kSwOffCoef = 128;
constexpr int kSwOffZ = 512;
constexpr int kSwOffRed = kSwOffZ + 3 * kSwTX * 8;
constexpr int kSwOffSlots = ((kSwOffRed + kSwWarps * kSwSums * 8 + 127) / 128) * 128;
constexpr int kSwSmemMax = 227 * 1024;

enum { SW_OP_BRATU = 0, SW_OP_HEAT = 1 };

struct SweepArgs {
    int64_t nx, ny;
    int32_t k;          // basis vectors subtracted and projected on
    int32_t nslot;      // ring depth
    int32_t wrap_x;     // periodic in x
    int32_t coef_from_u;
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
};

// ---- mbarrier / bulk-copy PTX ------------------------------------------------------------------------------------
AK_DEV uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
AK_DEV void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
AK_DEV void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
AK_DEV void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
AK_DEV bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// bounded wait: a pipeline that never completes is a bug; trap instead of hanging the GPU
AK_DEV void mbar_wait(uint64_t* bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 4000000000ll) __trap();
    }
}
AK_DEV void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// resident blocks per SM the register budget is sized for (8 warps: 255 / 128 / 80 registers per thread)
constexpr int sw_min_blocks(int kb) { return kb <= 4 ? 3 : (kb <= 8 ? 2 : 1); }

// One sweep.  KB: compile-time bound on k (register-resident sums); STENCIL: y = J z is formed and projected.
template <int KB, bool STENCIL, int OP>
__global__ void __launch_bounds__(kSwThreads, sw_min_blocks(KB)) k_sweep(const SweepArgs p) {
    extern __shared__ __align__(128) unsigned char sw_smem[];
    if (p.stop != nullptr && *p.stop != 0) return;
    constexpr int NS = 2 * KB + 2;
    constexpr bool HASCOEF = STENCIL && OP == SW_OP_BRATU;
    uint64_t* bar_full = reinterpret_cast<uint64_t*>(sw_smem);
    double* s_c = reinterpret_cast<double*>(sw_smem + kSwOffCoef);
    double* s_z = reinterpret_cast<double*>(sw_smem + kSwOffZ);
    double* s_red = reinterpret_cast<double*>(sw_smem + kSwOffRed);
    double* slots = reinterpret_cast<double*>(sw_smem + kSwOffSlots);
    int* s_last_p = reinterpret_cast<int*>(sw_smem + kSwOffCoef + kSwKMax * 8);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int k = p.k, nslot = p.nslot;
    const int64_t nx = p.nx, ny = p.ny;
    const int nvec = k + 1 + (HASCOEF ? 1 : 0);
    const int slot_doubles = nvec * kSwTX;
    if (tid == 0) {
        for (int s = 0; s < nslot; ++s) {
            mbar_init(bar_full + s, 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (tid < k) s_c[tid] = p.cvec[tid];
    const double sin = p.in_scale != nullptr ? *p.in_scale : 1.0;
    __syncthreads();

    // this CTA's share of the (strip, row) space: [u_begin, u_end) in strip-major order
    const int64_t nstrip = (nx + kSwTXI - 1) / kSwTXI;
    const int64_t total = nstrip * ny;
    const int64_t u_begin = total * (int64_t)blockIdx.x / (int64_t)gridDim.x;
    const int64_t u_end = total * ((int64_t)blockIdx.x + 1) / (int64_t)gridDim.x;

    double acc_n = 0.0, acc_zy = 0.0;
    double acc_g[KB], acc_t[KB];
#pragma unroll
    for (int j = 0; j < KB; ++j) acc_g[j] = acc_t[j] = 0.0;

    // Thread t owns staged column t.  Rows flow through a ring of `nslot` shared-memory slots: the copies of row i + nslot - 1
    // are issued right after the barrier of step i (every warp has then finished reading the slot of row i - 1), by the
    // first lanes of every warp (vector v belongs to warp v % 8, lane v / 8); thread 0 arms the slot's mbarrier.
    {
        const int t = tid;
        const Divisor dx2 = p.dx2d, dy2 = p.dy2d;
        uint32_t cnt = 0;  // rows consumed so far by this CTA (slot = cnt % nslot, mbarrier parity = (cnt / nslot) & 1)
        double sd[KB];     // the row of every S_j the sums of this step need
#pragma unroll
        for (int j = 0; j < KB; ++j) sd[j] = 0.0;
        const int myvec = (lane < 4) ? lane * (kSwThreads / 32) + warp : nvec;  // the vector this lane streams (if < nvec)
        for (int64_t u = u_begin; u < u_end;) {
            const int64_t strip = u / ny, r0 = u - strip * ny;
            const int64_t r1 = (r0 + (u_end - u) < ny) ? r0 + (u_end - u) : ny;
            u += r1 - r0;
            const int64_t c0 = strip * kSwTXI - kSwHalo;
            const int64_t cs = c0 < 0 ? 0 : c0, ce = (c0 + kSwTX < nx) ? c0 + kSwTX : nx;
            const uint32_t nb = (uint32_t)(ce - cs) * 8u;
            const int soff = (int)(cs - c0);
            const bool wrap_l = p.wrap_x && c0 < 0;           // columns -2, -1 are columns nx-2, nx-1
            const bool wrap_r = p.wrap_x && c0 + kSwTX > nx;  // columns nx, nx+1 are columns 0, 1
            const uint32_t vec_bytes = nb + (wrap_l ? 16u : 0u) + (wrap_r ? 16u : 0u);
            const int64_t gc = c0 + t;
            // columns whose z is meaningful: the grid, plus the periodic images next to it
            const bool zvalid = (gc >= 0 && gc < nx) || (p.wrap_x && gc >= -kSwHalo && gc < nx + kSwHalo);
            const bool interior = t >= kSwHalo && t < kSwTX - kSwHalo && gc < nx;
            const int64_t rbeg = STENCIL ? r0 - 1 : r0, rend = STENCIL ? r1 + 1 : r1;
            const int nrows = (int)(rend - rbeg);
            const uint32_t cnt0 = cnt;  // value of cnt at the first row of this segment
            auto issue = [&](int ri) {
                const int64_t row = rbeg + ri;
                const uint32_t f = cnt0 + (uint32_t)ri;
                const int slot = (int)(f % (uint32_t)nslot);
                const bool inside = row >= 0 && row < ny;
                const bool rowdata = inside || (row < 0 ? p.zin_lo != nullptr : p.zin_hi != nullptr);
                if (tid == 0) {
                    const uint32_t nv = rowdata ? (uint32_t)(k + 1) + ((HASCOEF && inside) ? 1u : 0u) : 0u;
                    if (nv != 0) mbar_arrive_expect_tx(bar_full + slot, nv * vec_bytes);
                    else mbar_arrive(bar_full + slot);
                }
                if (myvec < nvec && rowdata) {
                    const double* src = nullptr;
                    if (myvec < k) src = row < 0 ? p.S_lo[myvec] : (row >= ny ? p.S_hi[myvec] : p.S[myvec] + row * nx);
                    else if (myvec == k) src = row < 0 ? p.zin_lo : (row >= ny ? p.zin_hi : p.zin + row * nx);
                    else if (HASCOEF && inside) src = p.coef + row * nx;
                    if (src != nullptr) {
                        double* dst = slots + (size_t)slot * slot_doubles + (size_t)myvec * kSwTX;
                        bulk_g2s(dst + soff, src + cs, nb, bar_full + slot);
                        if (wrap_l) bulk_g2s(dst, src + (nx - 2), 16u, bar_full + slot);
                        if (wrap_r) bulk_g2s(dst + (nx - c0), src, 16u, bar_full + slot);
                    }
                }
            };
            // every warp is done with the previous segment's slots before they are refilled
            __syncthreads();
            for (int ri = 0; ri < nslot - 1 && ri < nrows; ++ri) issue(ri);
            double z1 = 0.0, z2 = 0.0, coefd = 0.0;
            for (int i = 0; i < nrows; ++i, ++cnt) {
                const int64_t row = rbeg + i;
                const int slot = (int)(cnt % (uint32_t)nslot);
                const uint32_t par = (cnt / (uint32_t)nslot) & 1u;
                mbar_wait(bar_full + slot, par);
                const double* base = slots + (size_t)slot * slot_doubles;
                const bool rowdata = (row >= 0 && row < ny) || (row < 0 ? p.zin_lo != nullptr : p.zin_hi != nullptr);
                double zt = 0.0;
                if (!STENCIL) {
                    if (rowdata && interior) {
                        zt = __dmul_rn(base[k * kSwTX + t], sin);
#pragma unroll
                        for (int j = 0; j < KB; ++j)
                            if (j < k) {
                                sd[j] = base[j * kSwTX + t];
                                zt = fma(-s_c[j], sd[j], zt);  // same order as successive kaxpy!
                            }
                        acc_n = fma(zt, zt, acc_n);
#pragma unroll
                        for (int j = 0; j < KB; ++j)
                            if (j < k) acc_g[j] = fma(sd[j], zt, acc_g[j]);
                        if (p.zout != nullptr) p.zout[row * nx + gc] = zt;
                        if (row == 0 && p.push_z_down != nullptr) p.push_z_down[gc] = zt;
                        if (row == ny - 1 && p.push_z_up != nullptr) p.push_z_up[gc] = zt;
                    }
                    __syncthreads();  // every warp has read the slot of row i - 1 ... and of this row
                    if (i + nslot - 1 < nrows) issue(i + nslot - 1);
                } else {
                    if (rowdata && zvalid) {
                        zt = __dmul_rn(base[k * kSwTX + t], sin);
#pragma unroll
                        for (int j = 0; j < KB; ++j)
                            if (j < k) zt = fma(-s_c[j], base[j * kSwTX + t], zt);
                    }
                    const bool own_row = row >= r0 && row < r1;
                    if (own_row && interior) {
                        if (p.zout != nullptr) p.zout[row * nx + gc] = zt;
                        if (row == 0 && p.push_z_down != nullptr) p.push_z_down[gc] = zt;
                        if (row == ny - 1 && p.push_z_up != nullptr) p.push_z_up[gc] = zt;
                    }
                    s_z[(i % 3) * kSwTX + t] = zt;
                    __syncthreads();  // z of this row is visible; every warp has finished step i - 1
                    if (i + nslot - 1 < nrows) issue(i + nslot - 1);  // into the slot of row i - 1
                    if (i >= 2 && interior) {
                        // y on row q = row - 1: z1 = z(q), z2 = z(q - 1), zt = z(q + 1); x-neighbours from shared memory
                        const int64_t q = row - 1;
                        const double* zrow = s_z + ((i + 2) % 3) * kSwTX;
                        const double zl = zrow[t - 1], zr = zrow[t + 1];
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
                        p.yout[q * nx + gc] = y;
                        if (q == 0 && p.push_y_down != nullptr) p.push_y_down[gc] = y;
                        if (q == ny - 1 && p.push_y_up != nullptr) p.push_y_up[gc] = y;
                        acc_n = fma(z1, z1, acc_n);
                        acc_zy = fma(z1, y, acc_zy);
#pragma unroll
                        for (int j = 0; j < KB; ++j)
                            if (j < k) {
                                acc_g[j] = fma(sd[j], z1, acc_g[j]);
                                acc_t[j] = fma(sd[j], y, acc_t[j]);
                            }
                    }
                    // this row becomes the delay line of the next step
                    if (own_row && interior) {
#pragma unroll
                        for (int j = 0; j < KB; ++j)
                            if (j < k) sd[j] = base[j * kSwTX + t];
                        if (HASCOEF) coefd = base[(k + 1) * kSwTX + t];
                    }
                    z2 = z1;
                    z1 = zt;
                }
            }
        }
    }

    // ---------------- deterministic grid reduction of the NS sums ----------------
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

// ---- launch ------------------------------------------------------------------------------------------------------
template <int KB, bool STENCIL, int OP>
static int launch_sweep_t(Ctx* ctx, SweepArgs& a) {
    static std::atomic<int> configured{0};
    if (!configured.load(std::memory_order_relaxed)) {
        AK_CUDA(cudaFuncSetAttribute(k_sweep<KB, STENCIL, OP>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSwSmemMax));
        configured.store(1, std::memory_order_relaxed);
    }
    const int nvec = a.k + 1 + ((STENCIL && OP == SW_OP_BRATU) ? 1 : 0);
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
    const size_t smem = (size_t)kSwOffSlots + (size_t)nslot * slot_bytes;
    const int64_t nstrip = (a.nx + kSwTXI - 1) / kSwTXI;
    const int64_t total = nstrip * a.ny;
    int64_t grid = (int64_t)ctx->num_sms * occ;
    const int64_t by_work = (total + 7) / 8;  // at least ~8 rows per block
    if (grid > by_work) grid = by_work;
    if (grid < 1) grid = 1;
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

bool sweep_supported(const Ctx* ctx, const ak_problem* p, const double* u) {
    if (!(p->kind == AK_BRATU2D || p->kind == AK_HEAT2D)) return false;
    if (p->jvp_mode != AK_JVP_ANALYTIC || p->scheme == AK_MIDPOINT) return false;
    if (p->nx < 4 || p->nx % 2 != 0 || p->ny < 1) return false;  // bulk copies move 16-byte units
    if (p->kind == AK_BRATU2D) {  // lambda e^u (or u itself) is streamed with bulk copies like the basis
        const double* cf = p->coef != nullptr ? p->coef : u;
        if (cf == nullptr || (reinterpret_cast<uintptr_t>(cf) & 15u)) return false;
    }
    if (ctx->nranks > 1) {
        // slabs: the neighbours' rows arrive through peer memory (ak_comm_enable_p2p sized the ghost rows)
        if (!ctx->p2p_on || p->nx > ctx->p2p_halo_cap) return false;
    }
    return true;
}

int launch_sweep(Ctx* ctx, const ak_problem* prob, const double* u, const SweepCall& c) {
    AK_REQUIRE(c.k >= 0 && c.k <= kSwKMax, "launch_sweep: k out of range");
    AK_REQUIRE(c.zin != nullptr && c.sums != nullptr, "launch_sweep: NULL operand");
    AK_REQUIRE(!c.stencil || c.yout != nullptr, "launch_sweep: the tangent needs a destination");
    SweepArgs a{};
    a.nx = prob->nx;
    a.ny = prob->ny;
    a.k = c.k;
    a.wrap_x = (prob->bc == AK_BC_PERIODIC);
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
    if (ctx->nranks == 1 && prob->bc == AK_BC_PERIODIC) {  // one GPU: the ghost rows are the opposite rows
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
    const int op = prob->kind == AK_BRATU2D ? SW_OP_BRATU : SW_OP_HEAT;
    if (op == SW_OP_BRATU) {
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
    if (k <= 4) rc = launch_sweep_kb<4>(ctx, a, c.stencil, op);
    else if (k <= 8) rc = launch_sweep_kb<8>(ctx, a, c.stencil, op);
    else if (k <= 12) rc = launch_sweep_kb<12>(ctx, a, c.stencil, op);
    else if (k <= 16) rc = launch_sweep_kb<16>(ctx, a, c.stencil, op);
    else if (k <= 20) rc = launch_sweep_kb<20>(ctx, a, c.stencil, op);
    else rc = launch_sweep_kb<24>(ctx, a, c.stencil, op);
    AK_TRY(rc);
    // NCCL fallback of the reduction is not offered: slabs take this path only with peer memory (sweep_supported)
    return AK_OK;
}

}  // namespace ak
