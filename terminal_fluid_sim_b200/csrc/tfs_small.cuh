// tfs_small.cuh -- FluidSim::next_step (simulator.rs:59-65) for terminal-sized grids in ONE launch of ONE CTA.
//
// The reference's own bench runs 100 x 100 (benches/benchmarks/sim_bench.rs:6) and the application a grid the size
// of the terminal (80 x 48, 48 x 42: SURVEY 8d config 1).  Such a grid - u, v, p, smoke and the flag bytes, 17 B per
// cell - fits the shared memory of one SM, and the tiled kernels spend their time on launches and on the fill /
// drain of a pipeline built for millions of cells.  Here a single 1024-thread CTA
//   1. loads u, v, smoke and flags into shared memory (p starts at 0: simulator.rs:154), adding gravity*dt to v on
//      updatable cells on the way (add_gravity, :137-149);
//   2. runs the K projection iterations entirely in shared memory.  Exact mode: the anti-diagonal wavefront
//      t = i + j + 2k (every (i, j, k) on one wavefront touches disjoint faces, and ordering by t keeps the
//      lexicographic order per location), one __syncthreads per wavefront - W + H + 2K of them instead of kernel
//      launches.  A warp takes one iteration k of the wavefront, its lanes walk the diagonal: the flat index moves by
//      H - 1 per lane, an odd stride for even H, so the shared-memory accesses are conflict-free.  Red-black mode: one
//      __syncthreads per colour;
//   3. advects u, v, smoke from the shared-memory copies of the old fields (move_velocity, :192-298, generic sampler:
//      every clamp of sample_vector) and writes the new fields and p to global memory.
// Same cell update (cell_update) and sampler (advect_cell) as the simple global-memory kernels: bit-identical.
#pragma once
#include "tfs_kernels.cuh"

namespace tfs {

constexpr int kSmallThreads = 1024;
constexpr long long kSmallMaxCells = 12800;  // 17 B per cell of dynamic shared memory: 217.6 KB of the 227 KB a CTA may use
constexpr long long kSmallMaxCellsExact = 8192;  // exact order: above this the systolic kernel on several SMs is as fast

struct SmallParams {
    const float *u, *v, *s;      // fields before the step
    const uint8_t *flags;
    float *nu, *nv, *ns, *p;     // fields after the step (p in place: it is overwritten)
    int W, H, K, redblack;
    float omega, pc, gdt, dt;
};

__global__ void __launch_bounds__(kSmallThreads, 1) k_step_small(SmallParams P) {
    extern __shared__ __align__(16) unsigned char small_raw[];
    const int W = P.W, H = P.H, N = W * H;
    float *su = reinterpret_cast<float *>(small_raw), *sv = su + N, *sp = sv + N, *ss = sp + N;
    uint8_t *sf = reinterpret_cast<uint8_t *>(ss + N);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = kSmallThreads / 32;

    for (int idx = tid; idx < N; idx += kSmallThreads) {
        const unsigned f = P.flags[idx];
        float vv = P.v[idx];
        if (f & TFS_F_LIVE) vv = fadd(vv, P.gdt);  // simulator.rs:147
        su[idx] = P.u[idx]; sv[idx] = vv; ss[idx] = P.s[idx]; sp[idx] = 0.0f; sf[idx] = (uint8_t)f;
    }
    __syncthreads();

    if (W >= 3 && H >= 3 && P.K > 0) {
        if (!P.redblack) {
            const int dmax = W + H - 4;  // interior cells: 1 <= i <= W-2, 1 <= j <= H-2, diagonals d = i + j in [2, dmax]
            for (int t = 2; t <= dmax + 2 * (P.K - 1); ++t) {
                const int k_lo = max(0, (t - dmax + 1) / 2), k_hi = min(P.K - 1, (t - 2) / 2);
                for (int k = k_lo + warp; k <= k_hi; k += nwarps) {
                    const int d = t - 2 * k;
                    const int i_lo = max(1, d - (H - 2)), i_hi = min(W - 2, d - 1);
                    for (int i = i_lo + lane; i <= i_hi; i += 32) {
                        const int idx = i * H + (d - i);
                        const unsigned f = sf[idx];
                        if (f & TFS_F_LIVE) {
                            float uC = su[idx], uR = su[idx + H], vC = sv[idx], vT = sv[idx + 1], pp = sp[idx];
                            cell_update(f, P.omega, P.pc, uC, uR, vC, vT, pp);
                            su[idx] = uC; su[idx + H] = uR; sv[idx] = vC; sv[idx + 1] = vT; sp[idx] = pp;
                        }
                    }
                }
                __syncthreads();
            }
        } else {
            for (int it = 0; it < P.K; ++it)
                for (int colour = 0; colour < 2; ++colour) {
                    for (int i = 1 + warp; i <= W - 2; i += nwarps)
                        for (int j = 2 * lane + ((i + colour) & 1); j <= H - 2; j += 64) {
                            if (j < 1) continue;
                            const int idx = i * H + j;
                            const unsigned f = sf[idx];
                            if (f & TFS_F_LIVE) {
                                float uC = su[idx], uR = su[idx + H], vC = sv[idx], vT = sv[idx + 1], pp = sp[idx];
                                cell_update(f, P.omega, P.pc, uC, uR, vC, vT, pp);
                                su[idx] = uC; su[idx + H] = uR; sv[idx] = vC; sv[idx + 1] = vT; sp[idx] = pp;
                            }
                        }
                    __syncthreads();
                }
        }
    }

    // move_velocity: every cell of the new buffers is written (non-updatable cells copy the old value, :193-195)
    const Grid g{(long long)W, (long long)H, 0LL, (long long)W};
    const float Wf = (float)W, Hf = (float)H;
    for (int idx = tid; idx < N; idx += kSmallThreads) {
        const float uo = su[idx], vo = sv[idx], so = ss[idx];
        float un = uo, vn = vo, sn = so;
        if (sf[idx] & TFS_F_LIVE) advect_cell<false>(su, sv, ss, g, (long long)(idx / H), (long long)(idx % H), P.dt, Wf, Hf, uo, vo, un, vn, sn);
        P.nu[idx] = un; P.nv[idx] = vn; P.ns[idx] = sn; P.p[idx] = sp[idx];
    }
}

inline bool small_step_usable(long long W, long long H) { return W * H <= kSmallMaxCells; }
inline size_t small_step_smem(long long W, long long H) { return (size_t)(W * H) * 17 + 16; }

}  // namespace tfs
