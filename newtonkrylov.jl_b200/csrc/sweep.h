// Interface of the one-sweep-per-iteration GMRES kernels (sweep.cu) towards the Krylov driver (krylov.cu).
#pragma once
#include "ak_internal.h"

namespace ak {

constexpr int kSwKMax = 24;               // most basis vectors one sweep subtracts / projects on (restart cycles of <= 24)
constexpr int kSwSums = 2 * kSwKMax + 2;  // [0] ||z||^2, [1] <z, y>, [2 + j] g_j = <S_j, z>, [2 + kSwKMax + j] t_j = <S_j, y>
constexpr int kSwMailRec = kSwSums + 2;   // doubles per sweep-mailbox record: the sums, tag, pad
// ghost-row slots of the peer-memory path, per rank: one per basis vector, two for the W buffers
constexpr int kSwGhostSlots = kSwKMax + 3;

struct SweepCall {
    int k = 0;                            // basis vectors S[0..k)
    const double* const* S = nullptr;     // host array of device pointers
    const double* const* S_lo = nullptr;  // ghost rows y = -1 / y = ny of every S_j (slabs; nullptr: physical boundary)
    const double* const* S_hi = nullptr;
    const double* zin = nullptr;          // the vector to orthogonalise (raw tangent W, or S_0 when k == 0)
    const double* zin_lo = nullptr;
    const double* zin_hi = nullptr;
    double* zout = nullptr;               // S_k (nullptr: not stored)
    bool stencil = false;                 // form y = J z and its projections
    double* yout = nullptr;
    const double* cvec = nullptr;         // device: update multipliers c_j
    const double* in_scale = nullptr;     // device scalar multiplying zin (nullptr: 1)
    double* sums = nullptr;               // device: kSwSums doubles
    const int* stop = nullptr;
    // peer memory: where the boundary rows of zout / yout go, and the record the sums are posted under
    double* push_z_down = nullptr;
    double* push_z_up = nullptr;
    double* push_y_down = nullptr;
    double* push_y_up = nullptr;
    unsigned long long seq_out = 0;
};

// can this problem take the sweep kernels on this context (analytic tangents of the 2-D and 1-D stencils and of DG, even
// row length, slabs / segments only with peer memory)
bool sweep_supported(const Ctx* ctx, const ak_problem* p, const double* u);
int launch_sweep(Ctx* ctx, const ak_problem* prob, const double* u, const SweepCall& c);
// peer memory: first / last row of the slab `v` into the neighbours' ghost rows of sweep slot `slot` (context.cu)
// (2-D slabs: count = nx; 1-D segments: the first / last `count` values)
int sweep_push_rows(Ctx* ctx, const double* v, int64_t n, int count, bool periodic, int slot);

}  // namespace ak
