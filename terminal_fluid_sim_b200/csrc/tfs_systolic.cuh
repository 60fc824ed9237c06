// tfs_systolic.cuh -- make_incompressible (simulator.rs:151-190) as a temporally blocked,
// register-resident systolic wavefront.  Bit-identical to the reference's lexicographic
// Gauss-Seidel sweep; add_gravity (simulator.rs:137-149) is fused into the first pass.
//
// Schedule (executable specification + proof of equivalence: oracle/systolic_model.py, which
// tests/test_systolic_model.py checks bit-for-bit against the CPU oracle):
//   * one PASS fuses KB projection iterations over the whole grid; passes ping-pong between
//     two (u, v, p) buffer sets, so HBM sees 25/KB bytes per cell-iteration instead of 25.
//   * a BAND is 32 consecutive values of the skewed coordinate q = i + k; lane L of band b owns
//     q = b*32 + L - 1 and at step t the band updates, for every k < KB, cell
//         column i = q - k,   row j = t - L - k - 1.
//     Every operand of that update was produced one step earlier by the same lane or by lane
//     L-1 (__shfl_up), pressure and flags two steps earlier by lane L-1.
//   * one CTA runs a band; its NW warps split the KB iterations into k-groups of KG (warp g owns
//     k in [g*KG, (g+1)*KG)), so a step is only KG updates long.  The hand-over from the last
//     iteration of group g to the first of group g+1 goes through shared memory (one float4 per
//     lane, one __syncthreads per step).  Lane 31 of every warp streams its outputs to the same
//     k-group of the next band through a record buffer in global memory (cp.async.cg prefetch
//     ring + progress counter); pass tb+1 waits on per-band row-progress counters of pass tb.
//   * k = 0 operands are staged global -> shared with cp.async in skewed 32 x S tiles
//     (coalesced along y, conflict-free diagonal reads); final values are staged back the same
//     way.  Tasks (pass, band) are handed out by an atomic ticket in dependency order, so a
//     waiting CTA only ever waits on CTAs that are already running: no deadlock.
#pragma once
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>

#include "tfs_common.cuh"

namespace tfs {

constexpr int kMaxTemporalBlock = 32;

struct SystolicWorkspace {
    void *buf = nullptr;  // records + counters
    size_t bytes = 0;
    int num_sms = 0;
    int last_kg = 0, last_nw = 0, last_passes = 0;  // shape of the last generic-kernel launch
};

struct SystolicArgs {
    float **u, **v, **p, **u2, **v2, **p2;  // handle's buffers; swapped when the pass count is odd
    const uint8_t *flags;
    long long W, H;
    int iterations, temporal_block;
    float omega, pc, gdt;
    bool fuse_gravity;
    cudaStream_t stream;
};

struct SysParams {
    float *buf[2][3];  // [set][u, v, p]; pass tb reads set tb&1, writes set (tb+1)&1
    const uint8_t *flags;
    long long W, H, ncells;
    int n_bands, n_passes, k_last;  // k_last = active iterations of the last pass
    int T;                          // steps per task
    int n_rec;                      // records per band
    int apply_gravity;
    float omega, pc, gdt;
    float4 *bnd;    // [2][n_bands][NW][n_rec][KG] records {uRp, vCf, pout, flags word}
    int *bnd_prog;  // [n_passes][n_bands] records published (all k-groups)
    int *out_prog;  // [n_passes][n_bands] rows complete
    int *ticket;
    long long *trace;  // [NW][8] cycle sums (TFS_SYS_TRACE builds only)
};

TFS_DEVINL int ld_volatile(const int *p) { return *reinterpret_cast<const volatile int *>(p); }
TFS_DEVINL void st_volatile(int *p, int v) { *reinterpret_cast<volatile int *>(p) = v; }
// acquire/release fence at gpu scope (also invalidates this SM's L1, so that L1-allocating
// loads issued afterwards observe what the releasing thread wrote)
TFS_DEVINL void fence_gpu() { asm volatile("fence.acq_rel.gpu;\n" ::: "memory"); }
// store predicated on a runtime condition without a branch (keeps the warp converged)
TFS_DEVINL void st_global_v4_if(bool pred, float4 *ptr, float4 v) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %0, 0;\n\t@p st.global.v4.f32 [%1], {%2, %3, %4, %5};\n\t}\n" ::"r"((int)pred),
        "l"(ptr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
        : "memory");
}

TFS_DEVINL void cp_async4(void *smem_dst, const void *gmem_src) {
    unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(d), "l"(gmem_src) : "memory");
}
TFS_DEVINL void cp_async16_cg(void *smem_dst, const void *gmem_src) {
    unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gmem_src) : "memory");
}
TFS_DEVINL void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
TFS_DEVINL void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

// The cell update with the flags already reduced to the 4 neighbour bits (0 => no update).
// Three warp-uniform variants, chosen by votes over the warp's 32 x KG cells of a step:
//   MODE 0  every cell is a plain interior fluid cell (nb == 15): straight-line arithmetic;
//   MODE 1  every cell is plain or not updatable (nb in {0, 15}), e.g. border rows / phantom lanes:
//           the same arithmetic, results kept only where nb == 15;
//   MODE 2  some cell has a solid neighbour: general update (n in 1..3, 0/1 face multipliers).
template <int MODE>
TFS_DEVINL void cell_update4(unsigned nb, float omega, float pc, float &uC, float &uR, float &vC, float &vT,
                             float &p) {
    if (MODE <= 1) {  // n = 4, all multipliers 1.0; x/4 == x*0.25 exactly (also for subnormals)
        const float div = fsub(fadd(fsub(uR, uC), vT), vC);
        const float corr = fmul(omega, fmul(-div, 0.25f));
        const float a = fsub(uC, corr), r = fadd(uR, corr), c = fsub(vC, corr), d = fadd(vT, corr);
        const float pp = fadd(p, fmul(pc, corr));
        if (MODE == 0) {
            uC = a; uR = r; vC = c; vT = d; p = pp;
        } else {
            const bool on = nb != 0u;
            uC = on ? a : uC; uR = on ? r : uR; vC = on ? c : vC; vT = on ? d : vT; p = on ? pp : p;
        }
        return;
    }
    // general: branch-free decode of n, 1/n and the 0/1 face multipliers (simulator.rs:165-170)
    const int n = __popc(nb);
    const float nf = (float)n;
    // RN(1/n): exact power of two for n = 1, 2, 4 (exponent arithmetic), 0x3eaaaaab for n = 3
    const float rcp_p2 = __int_as_float(0x3f800000 - ((n >> 1) << 23));
    const float rcp = (n == 3) ? 0.333333343267440796f : rcp_p2;
    const float sL = (nb & TFS_F_L) ? 1.0f : 0.0f;
    const float sR = (nb & TFS_F_R) ? 1.0f : 0.0f;
    const float sB = (nb & TFS_F_B) ? 1.0f : 0.0f;
    const float sT = (nb & TFS_F_T) ? 1.0f : 0.0f;
    const float div = fsub(fadd(fsub(uR, uC), vT), vC);        // :176-178
    const float corr = fmul(omega, div_small(-div, nf, rcp));  // :180 (n == 0: garbage, discarded below)
    const float a = __fmaf_rn(-corr, sL, uC);                  // :181  exact products, see tfs_common.cuh
    const float r = __fmaf_rn(corr, sR, uR);                   // :182
    const float c = __fmaf_rn(-corr, sB, vC);                  // :184
    const float d = __fmaf_rn(corr, sT, vT);                   // :185
    const float pp = fadd(p, fmul(pc, corr));                  // :186
    const bool on = nb != 0u;                                  // :161, :172
    uC = on ? a : uC; uR = on ? r : uR; vC = on ? c : vC; vT = on ? d : vT; p = on ? pp : p;
}

// The general update with its per-cell constants taken from a 16-entry table in shared memory, indexed by the four
// neighbour bits: entry nb = two float4 {RN(1/n), -n, sL, sR} {sB, sT, -, -} (n = popcount, s* = 1.0 / +0.0).  This
// replaces popc, int-to-float, the 1/n selection and four 0/1 selects per cell by two LDS.128, and the five "keep the
// old value when the cell is not updated" selects by ONE: entry 0 holds {1, -1, +0, -0} {+0, -0} and the correction of
// such a cell is forced to 1.0, so that every face update adds -0.0 (fma(-1, +0, x) and fma(1, -0, x)), which is the
// identity for every x including both zeros and NaN.  Same arithmetic as cell_update4<2> for every other entry.
// ~25 instructions per cell instead of ~41.  Used by the red-black kernel (config 5, 2 % obstacles: 660 -> 586 ms per
// projection).  In the exact block-ring kernel it gained 1 % at config 5 and LOST 6-8 % on obstacle-free grids (the
// plain path was scheduled differently around it), so that kernel keeps cell_update4<2>.
TFS_DEVINL void nb_lut_fill(float4 *lut, int tid, int nthreads) {
    for (int nb = tid; nb < 16; nb += nthreads) {
        const int n = __popc(nb);
        const float rcp = (n == 3) ? 0.333333343267440796f : ((n == 4) ? 0.25f : ((n == 2) ? 0.5f : 1.0f));
        if (nb == 0) {
            lut[0] = make_float4(1.0f, -1.0f, 0.0f, -0.0f);
            lut[1] = make_float4(0.0f, -0.0f, 0.0f, 0.0f);
        } else {
            lut[2 * nb] = make_float4(rcp, -(float)n, (nb & TFS_F_L) ? 1.0f : 0.0f, (nb & TFS_F_R) ? 1.0f : 0.0f);
            lut[2 * nb + 1] = make_float4((nb & TFS_F_B) ? 1.0f : 0.0f, (nb & TFS_F_T) ? 1.0f : 0.0f, 0.0f, 0.0f);
        }
    }
}
TFS_DEVINL void cell_update_lut(const float4 *__restrict__ lut, unsigned nb, float omega, float pc, float &uC, float &uR, float &vC,
                                float &vT, float &p) {
    const float4 A = lut[2 * nb];
    const float2 B = *reinterpret_cast<const float2 *>(&lut[2 * nb + 1]);
    const float div = fsub(fadd(fsub(uR, uC), vT), vC);        // :176-178
    const float x = -div;
    // div_small(x, n, RN(1/n)) (tfs_common.cuh): exact IEEE x / n
    const float q0 = __fmul_rn(x, A.x);
    const float rr = __fmaf_rn(A.y, q0, x);
    float q = __fmaf_rn(rr, A.x, q0);
    q = (fabsf(q0) <= 3.402823466e38f && q0 != 0.0f) ? q : q0;
    const float corr = fmul(omega, q);                         // :180
    const bool on = nb != 0u;                                  // :161, :172
    const float ce = on ? corr : 1.0f;
    uC = __fmaf_rn(-ce, A.z, uC);                              // :181
    uR = __fmaf_rn(ce, A.w, uR);                               // :182
    vC = __fmaf_rn(-ce, B.x, vC);                              // :184
    vT = __fmaf_rn(ce, B.y, vT);                               // :185
    const float pp = fadd(p, fmul(pc, corr));                  // :186
    p = on ? pp : p;
}

// optional per-segment cycle accounting of the compute warps (-DTFS_SYS_TRACE; debugging aid only)
#ifdef TFS_SYS_TRACE
#define TFS_TR_DECL long long tr_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0}; long long tr_t = clock64();
#define TFS_TR(i) { long long now_ = clock64(); tr_acc[i] += now_ - tr_t; tr_t = now_; }
#else
#define TFS_TR_DECL
#define TFS_TR(i)
#endif

// named barriers (id 0 is __syncthreads)
// (ids are immediates so that ptxas reserves only the barriers actually used)
template <int ID, int COUNT>
TFS_DEVINL void bar_sync_imm() { asm volatile("bar.sync %0, %1;\n" ::"n"(ID), "n"(COUNT) : "memory"); }
template <int ID, int COUNT>
TFS_DEVINL void bar_arrive_imm() { asm volatile("bar.arrive %0, %1;\n" ::"n"(ID), "n"(COUNT) : "memory"); }
template <int ID0, int COUNT>
TFS_DEVINL void bar_sync2(int parity) {  // double-buffered barrier pair ID0 / ID0+1
    if (parity & 1) bar_sync_imm<ID0 + 1, COUNT>(); else bar_sync_imm<ID0, COUNT>();
}
template <int ID0, int COUNT>
TFS_DEVINL void bar_arrive2(int parity) {
    if (parity & 1) bar_arrive_imm<ID0 + 1, COUNT>(); else bar_arrive_imm<ID0, COUNT>();
}
enum { BAR_STEP = 1, BAR_READY0 = 2, BAR_FREE0 = 4 };  // READY/FREE are double-buffered: +0 / +1

// Warp roles: warps 0..NW-1 are the k-groups (compute); warp NW is the IO warp, which owns every
// global-memory wait: it polls the progress counters, stages the k = 0 operands and the left
// band's records one chunk (S steps) ahead with cp.async, flushes the finished chunk's final
// values and publishes progress behind ONE gpu-scope fence per chunk.  Compute warps never spin
// and never fence; they meet the IO warp only at chunk boundaries through named barriers.
template <int KG, int NW, int S, int D>
__global__ void __launch_bounds__(32 * (NW + 1)) k_project_systolic(SysParams P) {
    constexpr int KB = KG * NW;
    constexpr int PITCH = S + 1;               // words per lane row of a staged f32 tile
    constexpr int FW = ((S + 4 + 3) / 4) | 1;  // words per lane row of the staged flag bytes
    constexpr int NCT = 32 * NW;               // compute threads
    constexpr int NALL = 32 * (NW + 1);
    static_assert(KG >= 1 && KG <= 8, "a k-group's flags are packed 4 bits per iteration into 32 bits");
    static_assert((S * KG) % 2 == 0, "record chunks are copied 16 B per lane");

    extern __shared__ __align__(16) unsigned char smem_raw[];
    float *s_in = reinterpret_cast<float *>(smem_raw);                         // [2][3][32][PITCH]
    unsigned *s_fl = reinterpret_cast<unsigned *>(s_in + 2 * 3 * 32 * PITCH);  // [2][32][FW]
    float *s_out = reinterpret_cast<float *>(s_fl + 2 * 32 * FW);              // [2][3][32][PITCH]
    float4 *s_grp = reinterpret_cast<float4 *>(s_out + 2 * 3 * 32 * PITCH);    // [2][NW][32]
    float4 *s_ring = s_grp + 2 * NW * 32;                                      // [2][NW][S][KG]
    __shared__ int s_task;

    const int lane = threadIdx.x & 31, g = threadIdx.x >> 5;
    const bool io_warp = g == NW;
    const long long W = P.W, H = P.H;

    for (;;) {
        __syncthreads();  // previous task fully retired (smem reuse, s_task)
        if (threadIdx.x == 0) s_task = atomicAdd(P.ticket, 1);
        __syncthreads();
        const int task = s_task;
        if (task >= P.n_passes * P.n_bands) break;
        const int tb = task / P.n_bands, b = task % P.n_bands;
        const int k_act = (tb == P.n_passes - 1) ? P.k_last : KB;
        const bool first_pass = tb == 0;
        const long long q0 = (long long)b * 32 - 1;  // skewed coordinate of lane 0
        const bool has_left = b > 0;
        const size_t rec_stride = (size_t)P.n_rec * KG;  // float4 per (band, k-group)
        const int n_chunks = (P.T + S - 1) / S;

        if (io_warp) {
            // ================================ IO warp ================================================
            const float *__restrict__ gin_u = P.buf[tb & 1][0];
            const float *__restrict__ gin_v = P.buf[tb & 1][1];
            const float *__restrict__ gin_p = P.buf[tb & 1][2];
            float *__restrict__ gout_u = P.buf[(tb + 1) & 1][0];
            float *__restrict__ gout_v = P.buf[(tb + 1) & 1][1];
            float *__restrict__ gout_p = P.buf[(tb + 1) & 1][2];
            const float4 *rec_in = P.bnd + ((size_t)(tb & 1) * P.n_bands + (b - 1)) * NW * rec_stride;
            const int *left_prog = P.bnd_prog + (size_t)tb * P.n_bands + (b - 1);
            int *my_prog = P.bnd_prog + (size_t)tb * P.n_bands + b;
            int *my_out_prog = P.out_prog + (size_t)tb * P.n_bands + b;
            const int *dep_out0 = P.out_prog + (size_t)(tb - 1) * P.n_bands + b;
            const int *dep_out1 = P.out_prog + (size_t)(tb - 1) * P.n_bands + min(b + 1, P.n_bands - 1);
            int known_left = 0, known_rows = first_pass ? (int)H : 0;

            // spin until rows [0, need_rows) of the previous pass are final in bands b, b+1 and the
            // left band has published records [0, need_rec); returns after the counters were seen.
            auto wait_deps = [&](int need_rows, int need_rec) {
                if (need_rows > (int)H) need_rows = (int)H;
                if (need_rec > P.n_rec) need_rec = P.n_rec;
                while (known_rows < need_rows) {
                    int k0 = 0;
                    if (lane == 0) k0 = min(ld_volatile(dep_out0), ld_volatile(dep_out1));
                    known_rows = __shfl_sync(0xffffffffu, k0, 0);
                }
                while (has_left && known_left < need_rec) {
                    int k0 = 0;
                    if (lane == 0) k0 = ld_volatile(left_prog);
                    known_left = __shfl_sync(0xffffffffu, k0, 0);
                }
            };
            // chunk c holds the k = 0 operands of steps tau in [c*S, (c+1)*S): for lane l
            //   U[l][s] = u[q_l + 1, tau - l - 1],  V[l][s] = v[q_l, tau - l],  P[l][s] = p[q_l, tau - l - 1],
            //   flag bytes of rows tau - l - 1 .. (+S); plus the left band's records of those steps.
            auto issue_chunk = [&](int c) {
                const int bufi = c & 1;
                float *su = s_in + (bufi * 3 + 0) * 32 * PITCH;
                float *sv = s_in + (bufi * 3 + 1) * 32 * PITCH;
                float *sp = s_in + (bufi * 3 + 2) * 32 * PITCH;
                unsigned *sf = s_fl + bufi * 32 * FW;
                const long long nmax = P.ncells - 1;
#pragma unroll 4
                for (int e = lane; e < 32 * S; e += 32) {
                    const int cc = e / S, r = e % S;
                    const long long col = q0 + cc;
                    const long long row = (long long)c * S + r - cc - 1;  // tau - l - 1
                    long long iu = (col + 1) * H + row, iv = col * H + row + 1, ip = col * H + row;
                    iu = iu < 0 ? 0 : (iu > nmax ? nmax : iu);  // out-of-range operands are masked at use
                    iv = iv < 0 ? 0 : (iv > nmax ? nmax : iv);
                    ip = ip < 0 ? 0 : (ip > nmax ? nmax : ip);
                    cp_async4(su + cc * PITCH + r, gin_u + iu);
                    cp_async4(sv + cc * PITCH + r, gin_v + iv);
                    if (!first_pass) cp_async4(sp + cc * PITCH + r, gin_p + ip);
                }
                // flag bytes: aligned words covering rows [row0, row0 + S] of each column
                for (int e = lane; e < 32 * FW; e += 32) {
                    const int cc = e / FW, wd = e % FW;
                    const long long col = q0 + cc;
                    const long long row0 = (long long)c * S - cc - 1;
                    long long a = ((col * H + row0) >> 2) + wd;   // arithmetic shift: floor for negatives
                    const long long amax = (P.ncells + 56) >> 2;  // the arrays carry 64 bytes of padding
                    a = a < 0 ? 0 : (a > amax ? amax : a);
                    cp_async4(sf + cc * FW + wd, reinterpret_cast<const unsigned *>(P.flags) + a);
                }
                // records [c*S, (c+1)*S) of every k-group (record r is consumed at the end of step r)
                if (has_left) {
                    float4 *dst = s_ring + (size_t)bufi * NW * S * KG;
                    for (int e = lane; e < NW * S * KG; e += 32) {
                        const int gg = e / (S * KG), x = e % (S * KG);
                        long long r4 = (long long)c * S * KG + x;  // float4 index inside the group's stream
                        if (r4 >= (long long)rec_stride) r4 = (long long)rec_stride - 1;  // past the end: masked at use
                        cp_async16_cg(dst + e, rec_in + (size_t)gg * rec_stride + r4);
                    }
                }
                cp_async_commit();
            };
            // final values of chunk c (steps [c*S, (c+1)*S) and < T): lane l, step t <-> cell
            //   column q_l - (KB-1), row t - l - KB
            auto flush_chunk = [&](int c) {
                const float *so = s_out + (size_t)(c & 1) * 3 * 32 * PITCH;
#pragma unroll 4
                for (int e = lane; e < 32 * S; e += 32) {
                    const int cc = e / S, s = e % S;
                    const int t = c * S + s;
                    const long long col = q0 + cc - (KB - 1);
                    const long long row = (long long)t - cc - KB;
                    if (t < P.T && col >= 0 && col < W && row >= 0 && row < H) {
                        const long long idx = col * H + row;
                        gout_u[idx] = so[(0 * 32 + cc) * PITCH + s];
                        gout_v[idx] = so[(1 * 32 + cc) * PITCH + s];
                        gout_p[idx] = so[(2 * 32 + cc) * PITCH + s];
                    }
                }
            };
            auto publish = [&](int c_done) {
                // after chunk c_done: lane 31 has retired row t - 32 - KB + 1 with t = (c_done+1)*S - 1,
                // and records up to step t (record index t - 32) are written
                const int t = min((c_done + 1) * S, P.T) - 1;
                if (lane == 0) {
                    int rows = t - 32 - KB + 2;
                    rows = rows < 0 ? 0 : (rows > (int)H ? (int)H : rows);
                    int recs = t - 32 + 1;
                    recs = recs < 0 ? 0 : (recs > P.n_rec ? P.n_rec : recs);
                    if (c_done == n_chunks - 1) { rows = (int)H; recs = P.n_rec; }
                    st_volatile(my_out_prog, rows);
                    st_volatile(my_prog, recs);
                }
            };

            TFS_TR_DECL
            wait_deps(S + 1, S);
            fence_gpu();
            issue_chunk(0);
            cp_async_wait<0>();
            bar_arrive2<BAR_READY0, NALL>(0);
            TFS_TR(0)
            for (int c = 0; c < n_chunks; ++c) {
                if (c >= 1) {
                    bar_sync2<BAR_FREE0, NALL>(c - 1);  // compute finished chunk c-1
                    __syncwarp();
                    TFS_TR(1)
                    flush_chunk(c - 1);
                    TFS_TR(2)
                    fence_gpu();  // release the flushed rows and the records the compute warps wrote
                    publish(c - 1);  // promptly: dependants must never wait on OUR dependencies
                    TFS_TR(3)
                }
                if (c + 1 < n_chunks) {
                    wait_deps((c + 2) * S + 1, (c + 2) * S);
                    TFS_TR(4)
                    fence_gpu();  // acquire what the counters announced
                    TFS_TR(5)
                    issue_chunk(c + 1);
                    TFS_TR(6)
                    cp_async_wait<0>();
                    bar_arrive2<BAR_READY0, NALL>(c + 1);
                    TFS_TR(7)
                }
            }
            bar_sync2<BAR_FREE0, NALL>(n_chunks - 1);
            flush_chunk(n_chunks - 1);
            fence_gpu();
            publish(n_chunks - 1);
#ifdef TFS_SYS_TRACE
            if (lane == 0 && P.trace) {
                for (int i = 0; i < 8; ++i) atomicAdd((unsigned long long *)&P.trace[(NW * 8 + i)], (unsigned long long)tr_acc[i]);
            }
#endif
        } else {
            // ================================ compute warps ========================================
            const bool grav = first_pass && P.apply_gravity;
            const long long q = q0 + lane;
            float4 *rec_out = P.bnd + (((size_t)(tb & 1) * P.n_bands + b) * NW + g) * rec_stride;

            float uL[KG], uRo[KG], vB[KG], vTo[KG], pin[KG], pd[KG];
#pragma unroll
            for (int k = 0; k < KG; ++k) uL[k] = uRo[k] = vB[k] = vTo[k] = pin[k] = pd[k] = 0.0f;
            unsigned fl = 0u, fld = 0u;  // KG nibbles: flags used this step / arriving for the step after next

            // staged k = 0 operands of step tau, loaded early (software pipelining) into plain values
            struct K0 { float su, sv, sp; unsigned f0, f1; };
            auto load_k0 = [&](int tau) -> K0 {
                const int c = tau / S, s = tau % S, bufi = c & 1;
                K0 o;
                o.su = s_in[((bufi * 3 + 0) * 32 + lane) * PITCH + s];
                o.sv = s_in[((bufi * 3 + 1) * 32 + lane) * PITCH + s];
                o.sp = s_in[((bufi * 3 + 2) * 32 + lane) * PITCH + s];
                const long long row0 = (long long)c * S - lane - 1;
                const int off = (int)((q * H + row0) & 3) + s;  // byte offset inside the staged aligned words
                const unsigned char *fb = reinterpret_cast<const unsigned char *>(s_fl + (bufi * 32 + lane) * FW);
                o.f0 = fb[off];
                o.f1 = fb[off + 1];
                return o;
            };
            auto apply_k0 = [&](int tau, const K0 &o) {
                const long long row = (long long)tau - lane - 1;
                const bool col_ok = q >= 0 && q < W;
                const bool row_ok = row >= 0 && row < H;
                const bool rowp_ok = row + 1 >= 0 && row + 1 < H;
                const unsigned f0 = (col_ok && row_ok) ? o.f0 : 0u;
                const unsigned f1 = (col_ok && rowp_ok) ? o.f1 : 0u;
                uRo[0] = (q + 1 >= 0 && q + 1 < W && row_ok) ? o.su : 0.0f;
                float vv = (col_ok && rowp_ok) ? o.sv : 0.0f;
                if (grav && (f1 & TFS_F_LIVE)) vv = fadd(vv, P.gdt);  // simulator.rs:147
                vTo[0] = vv;
                pin[0] = (!first_pass && col_ok && row_ok) ? o.sp : 0.0f;
                fl = (fl & ~0xFu) | (f0 & 15u);
            };
            constexpr unsigned FULL = (KG == 8) ? 0xFFFFFFFFu : ((1u << (4 * KG)) - 1u);
            const bool group_active = (g + 1) * KG <= k_act;  // every iteration of this k-group runs in this pass

            TFS_TR_DECL
            for (int c = 0; c < n_chunks; ++c) {
                TFS_TR(0)
                bar_sync2<BAR_READY0, NALL>(c);  // stage + records of chunk c have landed
                TFS_TR(6)
                if (g == 0) apply_k0(c * S, load_k0(c * S));
                const float4 *ring = s_ring + ((size_t)(c & 1) * NW + g) * S * KG;
                float *so = s_out + (size_t)(c & 1) * 3 * 32 * PITCH;
                const int s_end = min(S, P.T - c * S);
                for (int s = 0; s < s_end; ++s) {
                    const int t = c * S + s;
                    TFS_TR(0)
                    const bool use_left = has_left && t < P.n_rec;
                    // ---- early loads: left band's record (for lane 0) and next step's k = 0 operands
                    float4 lr[KG];
#pragma unroll
                    for (int kk = 0; kk < KG; ++kk) {
                        lr[kk] = ring[s * KG + kk];  // broadcast read; only lane 0 uses it (and only if use_left)
                    }
                    const bool have_next = (g == 0) && (s + 1 < s_end);
                    K0 nk0 = {0.f, 0.f, 0.f, 0u, 0u};
                    if (have_next) nk0 = load_k0(t + 1);
                    // ---- KG independent cell updates ---------------------------------------------------
                    TFS_TR(1)
                    const unsigned fl_used = fl;
                    float a[KG], r[KG], cv[KG], d[KG], pp[KG];
#pragma unroll
                    for (int kk = 0; kk < KG; ++kk) { a[kk] = uL[kk]; r[kk] = uRo[kk]; cv[kk] = vB[kk]; d[kk] = vTo[kk]; pp[kk] = pin[kk]; }
                    // effective flags of this step: iterations beyond the pass's active count do nothing
                    unsigned fl_eff = fl_used & FULL;
                    if (!group_active) {
#pragma unroll
                        for (int kk = 0; kk < KG; ++kk)
                            if (g * KG + kk >= k_act) fl_eff &= ~(15u << (4 * kk));
                    }
                    // per-nibble "is 0 or 15": x ^ (x >> 1 ...) trick: a nibble is uniform iff all 4 bits equal
                    const unsigned lo = fl_eff & (FULL / 15u);                  // bit 0 of every nibble
                    const unsigned spread = lo * 15u;                           // 0xF where bit 0 is set
                    const bool plain_or_off = fl_eff == spread;                 // every nibble is 0x0 or 0xF
                    const bool all_plain = fl_eff == FULL;
                    if (__all_sync(0xffffffffu, all_plain)) {
#pragma unroll
                        for (int kk = 0; kk < KG; ++kk)
                            cell_update4<0>(15u, P.omega, P.pc, a[kk], r[kk], cv[kk], d[kk], pp[kk]);
                    } else if (__all_sync(0xffffffffu, plain_or_off)) {
#pragma unroll
                        for (int kk = 0; kk < KG; ++kk)
                            cell_update4<1>((fl_eff >> (4 * kk)) & 15u, P.omega, P.pc, a[kk], r[kk], cv[kk], d[kk], pp[kk]);
                    } else {
#pragma unroll
                        for (int kk = 0; kk < KG; ++kk)
                            cell_update4<2>((fl_eff >> (4 * kk)) & 15u, P.omega, P.pc, a[kk], r[kk], cv[kk], d[kk], pp[kk]);
                    }
                    // a = uCf (u of the cell after iteration k), r = uRp, cv = vCf, d = vTp, pp = pout
                    // ---- lane L-1's outputs ---------------------------------------------------------------
                    TFS_TR(2)
                    float n_uL[KG], s_v[KG], s_p[KG];
#pragma unroll
                    for (int kk = 0; kk < KG; ++kk) {
                        n_uL[kk] = __shfl_up_sync(0xffffffffu, r[kk], 1);
                        s_v[kk] = __shfl_up_sync(0xffffffffu, cv[kk], 1);
                        s_p[kk] = __shfl_up_sync(0xffffffffu, pp[kk], 1);
                    }
                    unsigned s_flv = __shfl_up_sync(0xffffffffu, fl_used, 1);
                    if (lane == 0) {  // "lane -1" is lane 31 of the left band (zeros where there is none)
#pragma unroll
                        for (int kk = 0; kk < KG; ++kk) {
                            n_uL[kk] = use_left ? lr[kk].x : 0.0f;
                            s_v[kk] = use_left ? lr[kk].y : 0.0f;
                            s_p[kk] = use_left ? lr[kk].z : 0.0f;
                        }
                        s_flv = use_left ? __float_as_uint(lr[0].w) : 0u;
                    }
                    TFS_TR(3)
                    // ---- hand-over to the next k-group: {uCf, lane L-1's vCf, lane L-1's pout, its flags nibble}
                    s_grp[((t & 1) * NW + g) * 32 + lane] =
                        make_float4(a[KG - 1], s_v[KG - 1], s_p[KG - 1], __uint_as_float((s_flv >> (4 * (KG - 1))) & 15u));
                    if (g == NW - 1) {  // iteration KB-1 retired: final values of the pass -> output stage
                        so[(0 * 32 + lane) * PITCH + s] = a[KG - 1];
                        so[(1 * 32 + lane) * PITCH + s] = cv[KG - 1];
                        so[(2 * 32 + lane) * PITCH + s] = pp[KG - 1];
                    }
                    // ---- record of lane 31 for the same k-group of the next band (consumed at the end of ITS step t-32)
                    {
                        const int r_out = t - 32;
                        if (lane == 31 && r_out >= 0 && r_out < P.n_rec) {
                            float4 *rec_dst = rec_out + (size_t)r_out * KG;
#pragma unroll
                            for (int kk = 0; kk < KG; ++kk)
                                rec_dst[kk] = make_float4(r[kk], cv[kk], pp[kk], kk == 0 ? __uint_as_float(fl_used) : 0.0f);
                        }
                    }
                    // ---- operands of step t+1 (in-group part) ------------------------------------------------
#pragma unroll
                    for (int kk = KG - 1; kk >= 0; --kk) {
                        vB[kk] = d[kk];
                        uL[kk] = n_uL[kk];
                        if (kk >= 1) {
                            uRo[kk] = a[kk - 1];
                            vTo[kk] = s_v[kk - 1];
                            pin[kk] = pd[kk];  // filled at the end of step t-1 (two-step path)
                            pd[kk] = s_p[kk - 1];
                        }
                    }
                    fl = fld;           // flags: two-step path; slot 0 is set below
                    fld = s_flv << 4;   // (used at t by lane L-1, needed at t+2 by lane L one slot higher)
                    TFS_TR(4)
                    bar_sync_imm<BAR_STEP, NCT>();
                    TFS_TR(5)
                    if (g > 0) {
                        const float4 gi = s_grp[((t & 1) * NW + (g - 1)) * 32 + lane];
                        uRo[0] = gi.x;
                        vTo[0] = gi.y;
                        pin[0] = pd[0];
                        pd[0] = gi.z;
                        fld |= __float_as_uint(gi.w);  // the nibble for slot 0, two steps from now
                    } else if (have_next) {
                        apply_k0(t + 1, nk0);
                    }
                }
                bar_arrive2<BAR_FREE0, NALL>(c);  // stage, ring and output slots of chunk c are done
            }
#ifdef TFS_SYS_TRACE
            if (lane == 0 && P.trace) {
                for (int i = 0; i < 8; ++i) atomicAdd((unsigned long long *)&P.trace[(g * 8 + i)], (unsigned long long)tr_acc[i]);
            }
#endif
        }
    }
}

// ---- host side ------------------------------------------------------------------------------------

template <int KG, int NW, int S, int D>
struct SysLaunch {
    static constexpr int KB = KG * NW;
    static size_t smem_bytes() {
        constexpr int PITCH = S + 1, FW = ((S + 4 + 3) / 4) | 1;
        return sizeof(float) * (size_t)(2 * 3 * 32 * PITCH) + sizeof(unsigned) * (size_t)(2 * 32 * FW) +
               sizeof(float) * (size_t)(2 * 3 * 32 * PITCH) + sizeof(float4) * (size_t)(2 * NW * 32) +
               sizeof(float4) * (size_t)(2 * NW * S * KG);
    }
    static cudaError_t run(const SystolicArgs &a, SystolicWorkspace &ws, uint64_t *launched, const char **msg) {
        cudaError_t e;
        int dev = 0;
        if ((e = cudaGetDevice(&dev)) != cudaSuccess) return e;
        if (ws.num_sms == 0) {
            if ((e = cudaDeviceGetAttribute(&ws.num_sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return e;
        }
        const int K = a.iterations;
        SysParams P{};
        P.W = a.W; P.H = a.H; P.ncells = a.W * a.H;
        P.n_passes = (K + KB - 1) / KB;
        P.k_last = K - (P.n_passes - 1) * KB;
        P.n_bands = (int)((a.W + KB + 31) / 32);
        P.T = (int)a.H + 32 + KB;
        P.n_rec = (int)a.H + KB;
        P.apply_gravity = a.fuse_gravity ? 1 : 0;
        P.omega = a.omega; P.pc = a.pc; P.gdt = a.gdt;
        P.flags = a.flags;
        P.buf[0][0] = *a.u; P.buf[0][1] = *a.v; P.buf[0][2] = *a.p;
        P.buf[1][0] = *a.u2; P.buf[1][1] = *a.v2; P.buf[1][2] = *a.p2;
        // workspace: records [2][n_bands][NW][n_rec][KG] float4, then counters
        const size_t rec_bytes = (size_t)2 * P.n_bands * NW * P.n_rec * KG * sizeof(float4);
        const size_t n_prog = (size_t)P.n_passes * P.n_bands;
        const size_t n_out = (size_t)P.n_passes * P.n_bands;
        const size_t n_cnt = n_prog + n_out + 16;
        const size_t need = rec_bytes + n_cnt * sizeof(int);
        if (ws.bytes < need) {
            if (ws.buf) cudaFree(ws.buf);
            ws.buf = nullptr;
            ws.bytes = 0;
            if ((e = cudaMalloc(&ws.buf, need)) != cudaSuccess) { *msg = "workspace allocation"; return e; }
            ws.bytes = need;
        }
        P.bnd = reinterpret_cast<float4 *>(ws.buf);
        int *cnt = reinterpret_cast<int *>(reinterpret_cast<char *>(ws.buf) + rec_bytes);
        P.bnd_prog = cnt;
        P.out_prog = cnt + n_prog;
        P.ticket = cnt + n_prog + n_out;
        P.trace = nullptr;
#ifdef TFS_SYS_TRACE
        static long long *d_trace = nullptr;
        if (!d_trace) cudaMalloc((void **)&d_trace, 64 * 8 * sizeof(long long));
        cudaMemsetAsync(d_trace, 0, 64 * 8 * sizeof(long long), a.stream);
        P.trace = d_trace;
#endif
        if ((e = cudaMemsetAsync(cnt, 0, n_cnt * sizeof(int), a.stream)) != cudaSuccess) return e;
        auto kern = k_project_systolic<KG, NW, S, D>;
        const size_t smem = smem_bytes();
        static bool attr_set[64] = {false};  // function attributes are per device
        static int occ_dev[64] = {0};
        const int di = dev & 63;
        if (!attr_set[di]) {
            if ((e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) {
                *msg = "cudaFuncSetAttribute(max dynamic smem)";
                return e;
            }
            if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_dev[di], kern, 32 * (NW + 1), smem)) != cudaSuccess) return e;
            attr_set[di] = true;
        }
        const int occ = occ_dev[di];
        if (occ < 1) { *msg = "kernel does not fit on an SM"; return cudaErrorLaunchOutOfResources; }
        const long long n_tasks = (long long)P.n_passes * P.n_bands;
        const int grid = (int)std::min<long long>(n_tasks, (long long)occ * ws.num_sms);
        kern<<<grid, 32 * (NW + 1), smem, a.stream>>>(P);
        if ((e = cudaGetLastError()) != cudaSuccess) { *msg = "kernel launch"; return e; }
        *launched += 1;
        ws.last_kg = KG; ws.last_nw = NW; ws.last_passes = P.n_passes;
#ifdef TFS_SYS_TRACE
        {
            long long h[64 * 8];
            cudaStreamSynchronize(a.stream);
            cudaMemcpy(h, P.trace, sizeof(h), cudaMemcpyDeviceToHost);
            const double nsteps = (double)P.T * n_tasks;
            const char *names[8] = {"loop/other", "early loads", "updates", "shuffles+patch", "stores+shift", "STEP barrier", "READY barrier", "-"};
            {
                const char *ion[8] = {"prologue", "wait FREE", "flush", "fence+publish", "wait deps", "acq fence", "issue", "cp wait"};
                const double nchunks = (double)((P.T + S - 1) / S) * n_tasks;
                fprintf(stderr, "[trace] IO warp cycles/chunk:");
                for (int i = 0; i < 8; ++i) fprintf(stderr, " %s=%.0f", ion[i], h[NW * 8 + i] / nchunks);
                fprintf(stderr, "\n");
            }
            for (int gq = 0; gq < 1; ++gq) {
                fprintf(stderr, "[trace] warp %d cycles/step:", gq);
                for (int i = 0; i < 7; ++i) fprintf(stderr, " %s=%.0f", names[i], h[gq * 8 + i] / nsteps);
                fprintf(stderr, "\n");
            }
        }
#endif
        if (P.n_passes & 1) {  // results are in set 1
            std::swap(*a.u, *a.u2);
            std::swap(*a.v, *a.v2);
            std::swap(*a.p, *a.p2);
        }
        return cudaSuccess;
    }
};

inline bool systolic_available() { return true; }

// (KG, NW) shapes; KB = KG*NW iterations are fused per pass.  v3: instantiated for the generic kernel below
// (any H), v4: for the block-ring kernel of tfs_systolic4.cuh (H % 4 == 0).  A shape is only ever picked for
// the kernel family that has it, so the family never changes silently with the iteration count.
struct SysShape { int kg, nw; bool v3, v4; };
inline const SysShape *systolic_shapes(int *n) {
    static const SysShape shapes[] = {{4, 4, true, true},  {5, 4, false, true}, {6, 4, false, true}, {4, 8, true, true},
                                      {4, 3, true, true},  {2, 5, true, true},  {4, 2, true, true},  {2, 8, true, true},
                                      {2, 4, true, true},  {2, 2, true, true},  {5, 2, true, true},  {1, 8, true, true},
                                      {8, 2, false, true}, {8, 4, false, true}, {8, 1, false, true}};
    *n = (int)(sizeof(shapes) / sizeof(shapes[0]));
    return shapes;
}

// choose the shape: an explicit temporal_block picks the first shape with that KB (else the
// largest KB below it); auto minimises passes * (KB + overhead) so the last pass wastes little.
inline SysShape systolic_pick(int iterations, int requested, bool v4_only) {
    int n = 0;
    const SysShape *sh = systolic_shapes(&n);
    auto has = [&](const SysShape &x) { return v4_only ? x.v4 : x.v3; };
    const Tunables &tn = tunables();
    if (tn.sys_kg && tn.sys_nw) {
        for (int i = 0; i < n; ++i)
            if (sh[i].kg == tn.sys_kg && sh[i].nw == tn.sys_nw && has(sh[i])) return sh[i];
    }
    if (requested > 0) {
        SysShape best = {2, 2, true, true};
        int best_kb = 0;
        for (int i = 0; i < n; ++i) {
            const int kb = sh[i].kg * sh[i].nw;
            if (has(sh[i]) && kb <= requested && kb > best_kb) { best_kb = kb; best = sh[i]; }
        }
        return best;
    }
    SysShape best = sh[0];
    double best_cost = 1e30;
    for (int i = 0; i < n; ++i) {
        const int kb = sh[i].kg * sh[i].nw;
        if (!has(sh[i]) || kb > tn.sys_max_kb || sh[i].nw > 4) continue;  // measured on B200: 3 CTAs/SM shapes win or tie at every K tried
        const int passes = (iterations + kb - 1) / kb;
        const double cost = passes * (kb + 16.0);  // measured on B200: a pass costs ~16 iterations' worth of fixed latency
        if (cost < best_cost - 1e-9) { best_cost = cost; best = sh[i]; }
    }
    return best;
}

inline cudaError_t systolic_project(const SystolicArgs &a, SystolicWorkspace &ws, uint64_t *launched, const char **msg) {
    ws.last_kg = ws.last_nw = ws.last_passes = 0;
    const SysShape sh = systolic_pick(a.iterations, a.temporal_block, false);
#define TFS_SYS_CASE(KG_, NW_) \
    if (sh.kg == KG_ && sh.nw == NW_) return SysLaunch<KG_, NW_, 16, 6>::run(a, ws, launched, msg);
    TFS_SYS_CASE(4, 4)
    TFS_SYS_CASE(2, 8)
    TFS_SYS_CASE(4, 2)
    TFS_SYS_CASE(2, 4)
    TFS_SYS_CASE(4, 3)
    TFS_SYS_CASE(2, 2)
    TFS_SYS_CASE(5, 2)
    TFS_SYS_CASE(2, 5)
    TFS_SYS_CASE(4, 8)
    TFS_SYS_CASE(1, 8)
#undef TFS_SYS_CASE
    *msg = "internal: systolic_pick returned a shape without a generic-kernel instantiation";
    return cudaErrorInvalidValue;
}

inline void systolic_release(SystolicWorkspace &ws) {
    if (ws.buf) cudaFree(ws.buf);
    ws.buf = nullptr;
    ws.bytes = 0;
}

}  // namespace tfs
