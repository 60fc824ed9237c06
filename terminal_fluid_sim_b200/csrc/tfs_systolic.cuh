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
#include <cstdlib>

#include "tfs_common.cuh"

namespace tfs {

constexpr int kMaxTemporalBlock = 32;

struct SystolicWorkspace {
    void *buf = nullptr;  // records + counters
    size_t bytes = 0;
    int num_sms = 0;
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
    int *bnd_prog;  // [n_passes][n_bands][NW] records published
    int *out_prog;  // [n_passes][n_bands] rows complete
    int *ticket;
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
// `all_interior` is warp-uniform (a vote): the common case of 32 plain fluid cells takes a
// straight-line path; otherwise every lane runs the general update and selects.
TFS_DEVINL void cell_update4(unsigned nb, bool all_interior, float omega, float pc, float &uC, float &uR,
                             float &vC, float &vT, float &p) {
    if (all_interior) {  // n = 4, all multipliers 1.0; x/4 == x*0.25 exactly
        const float div = fsub(fadd(fsub(uR, uC), vT), vC);
        const float corr = fmul(omega, fmul(-div, 0.25f));
        uC = fsub(uC, corr);
        uR = fadd(uR, corr);
        vC = fsub(vC, corr);
        vT = fadd(vT, corr);
        p = fadd(p, fmul(pc, corr));
        return;
    }
    float a = uC, r = uR, c = vC, d = vT, pp = p;
    cell_update((nb ? nb : 15u) | TFS_F_LIVE, omega, pc, a, r, c, d, pp);  // branch-free; discarded if nb == 0
    const bool on = nb != 0u;
    uC = on ? a : uC;
    uR = on ? r : uR;
    vC = on ? c : vC;
    vT = on ? d : vT;
    p = on ? pp : p;
}

template <int KG, int NW, int S, int D>
__global__ void __launch_bounds__(32 * NW) k_project_systolic(SysParams P) {
    constexpr int KB = KG * NW;
    constexpr int PITCH = S + 1;               // words per lane row of a staged f32 tile
    constexpr int FW = ((S + 4 + 3) / 4) | 1;  // words per lane row of the staged flag bytes
    constexpr int G = 4;                       // records per progress publication
    constexpr int NT = 32 * NW;
    static_assert(S > D + 1, "staging groups must be older than the record ring depth");
    static_assert(KG >= 1 && KG <= 8, "a k-group's flags are packed 4 bits per iteration into 32 bits");

    extern __shared__ __align__(16) unsigned char smem_raw[];
    float *s_in = reinterpret_cast<float *>(smem_raw);                      // [2][3][32][PITCH]
    unsigned *s_fl = reinterpret_cast<unsigned *>(s_in + 2 * 3 * 32 * PITCH);  // [2][32][FW]
    float *s_out = reinterpret_cast<float *>(s_fl + 2 * 32 * FW);           // [3][32][PITCH]
    float4 *s_grp = reinterpret_cast<float4 *>(s_out + 3 * 32 * PITCH);   // [2][NW][32] (all sizes above are multiples of 16 B)
    float4 *s_ring = s_grp + 2 * NW * 32;                                   // [NW][D+1][KG]
    float4 *s_zero = s_ring + NW * (D + 1) * KG;                            // [KG] zeros
    __shared__ int s_task;

    const int lane = threadIdx.x & 31, g = threadIdx.x >> 5;
    const long long W = P.W, H = P.H;
    float4 *my_ring = s_ring + g * (D + 1) * KG;
    if (threadIdx.x < KG) s_zero[threadIdx.x] = make_float4(0.f, 0.f, 0.f, 0.f);

    for (;;) {
        __syncthreads();  // previous task fully retired (smem reuse, s_task)
        if (threadIdx.x == 0) s_task = atomicAdd(P.ticket, 1);
        __syncthreads();
        const int task = s_task;
        if (task >= P.n_passes * P.n_bands) break;
        const int tb = task / P.n_bands, b = task % P.n_bands;
        const int k_act = (tb == P.n_passes - 1) ? P.k_last : KB;
        const float *__restrict__ gin_u = P.buf[tb & 1][0];
        const float *__restrict__ gin_v = P.buf[tb & 1][1];
        const float *__restrict__ gin_p = P.buf[tb & 1][2];
        float *__restrict__ gout_u = P.buf[(tb + 1) & 1][0];
        float *__restrict__ gout_v = P.buf[(tb + 1) & 1][1];
        float *__restrict__ gout_p = P.buf[(tb + 1) & 1][2];
        const bool first_pass = tb == 0;
        const bool grav = first_pass && P.apply_gravity;
        const long long q = (long long)b * 32 + lane - 1;  // skewed coordinate of this lane
        const long long q0 = (long long)b * 32 - 1;        // ... of lane 0
        // record streams of this k-group
        const bool has_left = b > 0;
        const size_t rec_stride = (size_t)P.n_rec * KG;
        const float4 *rec_in = P.bnd + (((size_t)(tb & 1) * P.n_bands + (b - 1)) * NW + g) * rec_stride;
        float4 *rec_out = P.bnd + (((size_t)(tb & 1) * P.n_bands + b) * NW + g) * rec_stride;
        const int *left_prog = P.bnd_prog + ((size_t)tb * P.n_bands + (b - 1)) * NW + g;
        int *my_prog = P.bnd_prog + ((size_t)tb * P.n_bands + b) * NW + g;
        int *my_out_prog = P.out_prog + (size_t)tb * P.n_bands + b;
        const int *dep_out0 = P.out_prog + (size_t)(tb - 1) * P.n_bands + b;
        const int *dep_out1 = P.out_prog + (size_t)(tb - 1) * P.n_bands + min(b + 1, P.n_bands - 1);
        int known_left = 0, known_rows = first_pass ? (int)H : 0;

        // rows [0, need) of the previous pass must be final in bands b and b+1 (CTA-uniform call)
        auto wait_rows = [&](int need) {
            if (need > (int)H) need = (int)H;
            if (known_rows >= need) return;
            int k0 = 0;
            do {
                if (lane == 0) k0 = min(ld_volatile(dep_out0), ld_volatile(dep_out1));
                k0 = __shfl_sync(0xffffffffu, k0, 0);
            } while (k0 < need);
            known_rows = k0;
            fence_gpu();  // acquire: the (L1-allocating) stage loads below see the producers' stores
        };
        // chunk c holds the k = 0 operands of steps tau in [c*S, (c+1)*S): for lane l
        //   U[l][s] = u[q_l + 1, tau - l - 1],  V[l][s] = v[q_l, tau - l],  P[l][s] = p[q_l, tau - l - 1],
        //   flag bytes of rows tau - l - 1 .. (+S).  All NT threads share the copy work.
        auto issue_chunk = [&](int c) {
            const int bufi = c & 1;
            float *su = s_in + (bufi * 3 + 0) * 32 * PITCH;
            float *sv = s_in + (bufi * 3 + 1) * 32 * PITCH;
            float *sp = s_in + (bufi * 3 + 2) * 32 * PITCH;
            unsigned *sf = s_fl + bufi * 32 * FW;
            const long long nmax = P.ncells - 1;
            for (int e = threadIdx.x; e < 32 * S; e += NT) {
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
            for (int e = threadIdx.x; e < 32 * FW; e += NT) {
                const int cc = e / FW, wd = e % FW;
                const long long col = q0 + cc;
                const long long row0 = (long long)c * S - cc - 1;
                long long a = ((col * H + row0) >> 2) + wd;   // arithmetic shift: floor for negatives
                const long long amax = (P.ncells + 56) >> 2;  // the arrays carry 64 bytes of padding
                a = a < 0 ? 0 : (a > amax ? amax : a);
                cp_async4(sf + cc * FW + wd, reinterpret_cast<const unsigned *>(P.flags) + a);
            }
            cp_async_commit();
        };
        // final values of steps [t_first, t_last] (one chunk's worth of slots): lane l, step t <->
        // cell column q_l - (KB-1), row t - l - KB.  All NT threads share the stores.
        auto flush_out = [&](int t_first, int t_last) {
            for (int e = threadIdx.x; e < 32 * S; e += NT) {
                const int cc = e / S, s = e % S;
                const int t = t_first + s;
                if (t > t_last) continue;
                const long long col = q0 + cc - (KB - 1);
                const long long row = (long long)t - cc - KB;
                if (col >= 0 && col < W && row >= 0 && row < H) {
                    const long long idx = col * H + row;
                    gout_u[idx] = s_out[(0 * 32 + cc) * PITCH + s];
                    gout_v[idx] = s_out[(1 * 32 + cc) * PITCH + s];
                    gout_p[idx] = s_out[(2 * 32 + cc) * PITCH + s];
                }
            }
        };

        // ---- pipeline registers of this k-group ------------------------------------------------------
        float uL[KG], uRo[KG], vB[KG], vTo[KG], pin[KG], pd[KG];
#pragma unroll
        for (int k = 0; k < KG; ++k) uL[k] = uRo[k] = vB[k] = vTo[k] = pin[k] = pd[k] = 0.0f;
        unsigned fl = 0u, fld = 0u;  // KG nibbles: flags used this step / arriving for the step after next

        // k = 0 operands of step tau from the staged chunk tau / S (k-group 0 only)
        auto prep_k0 = [&](int tau) {
            const int c = tau / S, s = tau % S, bufi = c & 1;
            const long long row = (long long)tau - lane - 1;
            const bool col_ok = q >= 0 && q < W;
            const bool row_ok = row >= 0 && row < H;
            const bool rowp_ok = row + 1 >= 0 && row + 1 < H;
            const float su = s_in[((bufi * 3 + 0) * 32 + lane) * PITCH + s];
            const float sv = s_in[((bufi * 3 + 1) * 32 + lane) * PITCH + s];
            const float sp = s_in[((bufi * 3 + 2) * 32 + lane) * PITCH + s];
            const long long row0 = (long long)c * S - lane - 1;
            const int off = (int)((q * H + row0) & 3) + s;  // byte offset inside the staged aligned words
            const unsigned char *fb = reinterpret_cast<const unsigned char *>(s_fl + (bufi * 32 + lane) * FW);
            const unsigned f0 = (col_ok && row_ok) ? fb[off] : 0u;
            const unsigned f1 = (col_ok && rowp_ok) ? fb[off + 1] : 0u;
            uRo[0] = (q + 1 >= 0 && q + 1 < W && row_ok) ? su : 0.0f;
            float vv = (col_ok && rowp_ok) ? sv : 0.0f;
            if (grav && (f1 & TFS_F_LIVE)) vv = fadd(vv, P.gdt);  // simulator.rs:147
            vTo[0] = vv;
            pin[0] = (!first_pass && col_ok && row_ok) ? sp : 0.0f;
            fl = (fl & ~0xFu) | (f0 & 15u);
        };
        auto poll_left = [&](int r) {
            if (known_left > r) return;
            int k0 = 0;
            do {
                if (lane == 0) k0 = ld_volatile(left_prog);
                k0 = __shfl_sync(0xffffffffu, k0, 0);
            } while (k0 <= r);
            known_left = k0;
            fence_gpu();
        };

        // ---- prologue ---------------------------------------------------------------------------------
        wait_rows(S + 1);
        issue_chunk(0);
        wait_rows(2 * S + 1);
        issue_chunk(1);
        cp_async_wait<1>();  // chunk 0 landed (this thread's part)
        if (has_left) {
            for (int r = 0; r < D; ++r) {  // record r is consumed at the end of step r
                if (r < P.n_rec) {
                    poll_left(r);
                    if (lane < KG) cp_async16_cg(my_ring + (r % (D + 1)) * KG + lane, rec_in + (size_t)r * KG + lane);
                }
                cp_async_commit();
            }
        }
        __syncthreads();  // chunk 0 visible to k-group 0
        if (g == 0) prep_k0(0);

        // ---- main loop ----------------------------------------------------------------------------------
        for (int t = 0; t < P.T; ++t) {
            if (has_left) {
                const int r = t + D;
                if (r < P.n_rec) {
                    poll_left(r);
                    if (lane < KG) cp_async16_cg(my_ring + (r % (D + 1)) * KG + lane, rec_in + (size_t)r * KG + lane);
                }
                cp_async_commit();
                cp_async_wait<D>();  // record t has landed in the ring
                __syncwarp();
            }
            const float4 *lrec = (has_left && t < P.n_rec) ? (my_ring + (t % (D + 1)) * KG) : s_zero;
            const int r_out = t - 32;  // the next band consumes this step's record at the end of ITS step r_out
            const bool g_out = (lane == 31) && r_out >= 0 && r_out < P.n_rec;
            float4 *rec_dst = rec_out + (size_t)(r_out < 0 ? 0 : r_out) * KG;
            const unsigned fl_used = fl;
            float4 grp_out = make_float4(0.f, 0.f, 0.f, 0.f);
            // KG cell updates, kk descending so that every next-step operand is written in place
#pragma unroll
            for (int kk = KG - 1; kk >= 0; --kk) {
                const int k = g * KG + kk;
                unsigned nb = (fl_used >> (4 * kk)) & 15u;
                if (k >= k_act) nb = 0u;
                const bool all_interior = __all_sync(0xffffffffu, nb == 15u);
                float a = uL[kk], r = uRo[kk], c = vB[kk], d = vTo[kk], pp = pin[kk];
                cell_update4(nb, all_interior, P.omega, P.pc, a, r, c, d, pp);
                // a = uCf (u of the cell after iteration k), r = uRp, c = vCf, d = vTp, pp = pout
                st_global_v4_if(g_out, rec_dst + kk, make_float4(r, c, pp, kk == 0 ? __uint_as_float(fl_used) : 0.0f));
                const float4 lr = lrec[kk];  // what "lane -1" (lane 31 of the left band) produced; broadcast read
                float n_uL = __shfl_up_sync(0xffffffffu, r, 1);
                float s_v = __shfl_up_sync(0xffffffffu, c, 1);
                float s_p = __shfl_up_sync(0xffffffffu, pp, 1);
                if (lane == 0) { n_uL = lr.x; s_v = lr.y; s_p = lr.z; }
                vB[kk] = d;
                uL[kk] = n_uL;
                if (kk + 1 < KG) {
                    uRo[kk + 1] = a;
                    vTo[kk + 1] = s_v;
                    pin[kk + 1] = pd[kk + 1];  // filled at the end of step t-1 (two-step path)
                    pd[kk + 1] = s_p;
                } else {
                    grp_out.x = a; grp_out.y = s_v; grp_out.z = s_p;
                    if (g == NW - 1) {  // iteration KB-1 retired: final values of the pass -> output stage
                        const int slot = t % S;
                        s_out[(0 * 32 + lane) * PITCH + slot] = a;
                        s_out[(1 * 32 + lane) * PITCH + slot] = c;
                        s_out[(2 * 32 + lane) * PITCH + slot] = pp;
                    }
                }
            }
            {
                // flags: two-step path (used at t by lane L-1, needed at t+2 by lane L one slot higher)
                unsigned s_flv = __shfl_up_sync(0xffffffffu, fl_used, 1);
                if (lane == 0) s_flv = __float_as_uint(lrec[0].w);
                grp_out.w = __uint_as_float((s_flv >> (4 * (KG - 1))) & 15u);
                fl = fld;            // filled at the end of step t-1; slot 0 is set below / by prep_k0
                fld = s_flv << 4;
            }
            // hand-over to the next k-group: {uCf, lane L-1's vCf, lane L-1's pout, lane L-1's flags}
            s_grp[((t & 1) * NW + g) * 32 + lane] = grp_out;
            if (g_out && (((r_out + 1) % G) == 0 || r_out + 1 == P.n_rec)) {
                fence_gpu();
                st_volatile(my_prog, r_out + 1);
            }
            __syncthreads();
            if (g > 0) {
                const float4 gi = s_grp[((t & 1) * NW + (g - 1)) * 32 + lane];
                uRo[0] = gi.x;
                vTo[0] = gi.y;
                pin[0] = pd[0];
                pd[0] = gi.z;
                fld |= __float_as_uint(gi.w);  // the nibble for slot 0, two steps from now
            }
            // ---- chunk boundary ---------------------------------------------------------------------------
            if ((t + 1) % S == 0) {
                const int c_next = (t + 1) / S;  // chunk that serves steps t+1 .. t+S
                flush_out(t + 1 - S, t);
                if (!has_left) cp_async_wait<0>();  // chunk c_next landed (with a ring: older than the ring depth)
                fence_gpu();                         // release the flushed rows
                __syncthreads();
                if (threadIdx.x == 0) {
                    // rows complete for every lane of the band: lane 31 has just retired row t - 32 - KB + 1
                    int rows = t - 32 - KB + 2;
                    if (rows > 0) st_volatile(my_out_prog, rows > (int)H ? (int)H : rows);
                }
                if ((long long)(c_next + 1) * S < (long long)P.T) {
                    wait_rows((c_next + 2) * S + 1);
                    issue_chunk(c_next + 1);
                } else {
                    cp_async_commit();
                }
            }
            if (g == 0 && t + 1 < P.T) prep_k0(t + 1);
        }
        // ---- epilogue: last partial chunk, final progress ----------------------------------------------------
        cp_async_wait<0>();
        __syncthreads();
        if (P.T % S != 0) flush_out((P.T / S) * S, P.T - 1);
        fence_gpu();
        __syncthreads();
        if (threadIdx.x == 0) st_volatile(my_out_prog, (int)H);
        if (lane == 31) st_volatile(my_prog, P.n_rec);
    }
}

// ---- host side ------------------------------------------------------------------------------------

template <int KG, int NW, int S, int D>
struct SysLaunch {
    static constexpr int KB = KG * NW;
    static size_t smem_bytes() {
        constexpr int PITCH = S + 1, FW = ((S + 4 + 3) / 4) | 1;
        return sizeof(float) * (size_t)(2 * 3 * 32 * PITCH) + sizeof(unsigned) * (size_t)(2 * 32 * FW) +
               sizeof(float) * (size_t)(3 * 32 * PITCH) + sizeof(float4) * (size_t)(2 * NW * 32) +
               sizeof(float4) * (size_t)(NW * (D + 1) * KG) + sizeof(float4) * (size_t)KG;
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
        const size_t n_prog = (size_t)P.n_passes * P.n_bands * NW;
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
        if ((e = cudaMemsetAsync(cnt, 0, n_cnt * sizeof(int), a.stream)) != cudaSuccess) return e;
        auto kern = k_project_systolic<KG, NW, S, D>;
        const size_t smem = smem_bytes();
        static bool attr_set = false;
        static int occ = 0;
        if (!attr_set) {
            if ((e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) {
                *msg = "cudaFuncSetAttribute(max dynamic smem)";
                return e;
            }
            if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 32 * NW, smem)) != cudaSuccess) return e;
            attr_set = true;
        }
        if (occ < 1) { *msg = "kernel does not fit on an SM"; return cudaErrorLaunchOutOfResources; }
        const long long n_tasks = (long long)P.n_passes * P.n_bands;
        const int grid = (int)std::min<long long>(n_tasks, (long long)occ * ws.num_sms);
        kern<<<grid, 32 * NW, smem, a.stream>>>(P);
        if ((e = cudaGetLastError()) != cudaSuccess) { *msg = "kernel launch"; return e; }
        *launched += 1;
        if (P.n_passes & 1) {  // results are in set 1
            std::swap(*a.u, *a.u2);
            std::swap(*a.v, *a.v2);
            std::swap(*a.p, *a.p2);
        }
        return cudaSuccess;
    }
};

inline bool systolic_available() { return true; }

// (KG, NW) shapes that are instantiated; KB = KG*NW iterations are fused per pass.
struct SysShape { int kg, nw; };
inline const SysShape *systolic_shapes(int *n) {
    static const SysShape shapes[] = {{4, 4}, {2, 8}, {4, 2}, {2, 4}, {4, 3}, {2, 2}, {5, 2}, {2, 5}, {4, 8}, {1, 8}};
    *n = (int)(sizeof(shapes) / sizeof(shapes[0]));
    return shapes;
}

// choose the shape: an explicit temporal_block picks the first shape with that KB (else the
// largest KB below it); auto minimises passes * (KB + overhead) so the last pass wastes little.
inline SysShape systolic_pick(int iterations, int requested) {
    int n = 0;
    const SysShape *sh = systolic_shapes(&n);
    const char *env_kg = getenv("TFS_SYS_KG"), *env_nw = getenv("TFS_SYS_NW");
    if (env_kg && env_nw) {
        for (int i = 0; i < n; ++i)
            if (sh[i].kg == atoi(env_kg) && sh[i].nw == atoi(env_nw)) return sh[i];
    }
    if (requested > 0) {
        SysShape best = {2, 2};
        int best_kb = 0;
        for (int i = 0; i < n; ++i) {
            const int kb = sh[i].kg * sh[i].nw;
            if (kb <= requested && kb > best_kb) { best_kb = kb; best = sh[i]; }
        }
        return best;
    }
    SysShape best = sh[0];
    double best_cost = 1e30;
    for (int i = 0; i < n; ++i) {
        const int kb = sh[i].kg * sh[i].nw;
        if (kb > 16) continue;
        const int passes = (iterations + kb - 1) / kb;
        const double cost = passes * (kb + 2.0);
        if (cost < best_cost - 1e-9) { best_cost = cost; best = sh[i]; }
    }
    return best;
}

inline cudaError_t systolic_project(const SystolicArgs &a, SystolicWorkspace &ws, uint64_t *launched, const char **msg) {
    const SysShape sh = systolic_pick(a.iterations, a.temporal_block);
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
    *msg = "no kernel instantiated for this shape";
    return cudaErrorInvalidValue;
}

inline void systolic_release(SystolicWorkspace &ws) {
    if (ws.buf) cudaFree(ws.buf);
    ws.buf = nullptr;
    ws.bytes = 0;
}

}  // namespace tfs
