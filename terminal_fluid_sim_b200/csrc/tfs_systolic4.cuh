// tfs_systolic4.cuh -- the production variant of the systolic wavefront projection (H % 4 == 0).
//
// Same schedule and arithmetic as tfs_systolic.cuh (see its header and oracle/systolic_model.py);
// what changes is how data moves between HBM and the pipeline:
//   * warp roles: NW compute warps (k-groups), one LOADER warp, one STORER warp, one FEEDER warp.  Every
//     wait on global memory (progress counters of other CTAs), every fence and every bulk copy lives in
//     the loader and the storer; compute warps only touch registers, shared memory and shared-memory
//     mbarriers, and all of them execute the same instructions: the feeder warp reads the k = 0 operands
//     out of the input ring for the first k-group and moves the last k-group's results into the output
//     ring, both through the k-groups' own hand-over slots and the same per-step named barrier.
//   * k = 0 operands come from rectangular blocks of 32 columns x RB rows (RB = 16: 4-slot ring, the
//     shipped form; RB = 32: 3 slots), copied global -> shared with 16-byte cp.async.cg (bypasses L1: no
//     stale lines); lane l reads row tau - l - 1, i.e. bank (tau - l - 1) & 31: conflict-free.
//   * final values are written into a ring of the same geometry and stored with 128-bit coalesced
//     stores once the last lane has retired the block's last row.
//   * the left band's records are prefetched per interval of G steps (not per chunk), published by
//     the producer's storer warp every G steps, so a band trails its left neighbour by ~32 + 2G
//     steps; all hand-offs inside the CTA are mbarrier full/empty pairs.
//   * tasks are issued in dependency-respecting "start time" order (host-sorted list).
#pragma once
#include <utility>
#include <vector>

#include "tfs_systolic.cuh"

namespace tfs {

TFS_DEVINL unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
TFS_DEVINL void mbar_init(unsigned long long *b, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(b)), "r"(count) : "memory");
}
TFS_DEVINL void mbar_arrive(unsigned long long *b) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}\n" ::"r"(smem_u32(b)) : "memory");
}
// arrive once all cp.async issued so far by this thread have landed (counts as one expected arrival)
TFS_DEVINL void mbar_arrive_on_cp_async(unsigned long long *b) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];\n" ::"r"(smem_u32(b)) : "memory");
}
// Blocking wait on phase `parity`.  `hint_ns` is the suspend-time hint of try_wait: a waiting warp sleeps in
// hardware instead of re-issuing the poll (ncu showed ~100 spin instructions per step competing with the
// compute warps for issue slots before this was added).
template <unsigned HINT_NS = 400>
TFS_DEVINL void mbar_wait(unsigned long long *b, unsigned parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tLAB_WAIT:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t@p bra DONE;\n\tbra LAB_WAIT;\n\tDONE:\n\t}\n" ::"r"(
            smem_u32(b)),
        "r"(parity), "r"(HINT_NS)
        : "memory");
}

// x-slab placement of a handle (one process / GPU per slab).  A single-GPU handle is the slab
// [0, W) with no neighbours.  Progress counters are 64-bit (seq << 32 | progress): `seq` grows by one
// per projection, so counters never need a reset (a neighbour may write its mailbox entry before this
// rank has even launched) and a stale value of an earlier projection always compares as "nothing yet".
struct Sys4Slab {
    long long xs0 = 0;     // global column of storage column 0
    long long ncols = 0;   // stored columns
    long long c_lo = 0;    // first owned column: final values of columns >= c_lo are stored locally ...
    int b0 = 0;            // first global band of this rank
    int nb_local = 0;      // number of bands this rank runs
    int multi = 0;         // 1: neighbours exist -> system-scope fences
    long long seq = 1;
    float *peerL_buf[2][3] = {{nullptr, nullptr, nullptr}, {nullptr, nullptr, nullptr}};  // ... and columns <= c_lo also here
    long long peerL_xs0 = 0;
    float4 *rec_peerR[2] = {nullptr, nullptr};  // right rank's record slot 0 (per pass parity)
    long long *bnd_prog_peerR = nullptr;        // right rank's bnd_prog mailbox (its index 0), row stride cnt_strideR
    int cnt_strideR = 0;
    long long *out_prog_peerL = nullptr;        // left rank's out_prog mailbox (its index nbL+1), row stride cnt_strideL
    int cnt_strideL = 0;
    // fine-grained hand-offs with the neighbours' STREAMS (null / 0: the host orders these with whole-kernel flags)
    const long long *wait_right_halo = nullptr;  // my flag the right rank raises after pushing its post-advection halo:
    long long wait_right_halo_value = 0;         //   the last local band waits for it before its first load of pass 0
    long long *halo_ready_peerL = nullptr;       // left rank's flag: raised (to seq) once my band 0 has stored, in the LAST pass,
    int halo_cols = 0;                           //   its last owned columns and its halo columns [c_lo, c_lo + halo_cols)
};

struct Sys4Params {
    SysParams base;
    const int2 *tasks;  // (pass, GLOBAL band) in issue order, local bands only
    int n_tasks;
    Sys4Slab slab;
    long long *bnd_prog64;  // [n_passes][nb_local + 2]: index 0 = mailbox written by the left rank
    long long *out_prog64;  // [n_passes][nb_local + 2]: index nb_local + 1 = mailbox written by the right rank
    int cstride;            // row stride of the two counter arrays (>= nb_local + 2)
#ifdef TFS_TASK_TRACE
    unsigned long long *ttrace;  // [n_tasks][4]: globaltimer at task start / end, SM id, first compute step time
#endif
};

TFS_DEVINL long long ld_volatile64(const long long *p) { return *reinterpret_cast<const volatile long long *>(p); }
TFS_DEVINL void st_volatile64(long long *p, long long v) { *reinterpret_cast<volatile long long *>(p) = v; }
TFS_DEVINL void fence_sys() { asm volatile("fence.acq_rel.sys;\n" ::: "memory"); }
TFS_DEVINL int prog_of(long long v, long long seq) {
    const long long e = v >> 32;
    return e < seq ? 0 : (e > seq ? 0x7fffffff : (int)(v & 0xffffffffLL));
}
TFS_DEVINL long long prog_make(long long seq, int x) { return (seq << 32) | (long long)(unsigned)x; }

// RB = rows per block of the input / output rings: 32 (3 slots, 96 ring rows) or 16 (4 slots, 64 ring rows:
// 24 KB less shared memory per CTA, which is what lets a third CTA fit on an SM).
template <int KG, int NW, int G, int RB>
struct Sys4Smem {
    static constexpr int NS_IN = (RB == 32) ? 3 : 4, NS_OUT = NS_IN, NR = 3, NPH = 4;
    static constexpr int RIN = RB * NS_IN, ROUT = RB * NS_OUT;
    // row rings: [column][ring row]; the ring length is a multiple of 32, so bank = row & 31
    float in_u[32][RIN];
    float in_v[32][RIN];
    float in_p[32][RIN];
    float out_u[32][ROUT];
    float out_v[32][ROUT];
    float out_p[32][ROUT];
    float4 grp[2][NW + 1][32];  // slot 0: the feeder warp's k = 0 operands; slot g + 1: hand-over of k-group g
    float4 rec[NR][NW][G][KG];
    unsigned char in_f[32][RIN + 4];  // +4: odd word pitch, fewer bank conflicts on the diagonal byte reads
    unsigned long long in_full[NS_IN], in_empty[NS_IN];
    unsigned long long rec_full[NR], rec_empty[NR];
    unsigned long long rec_written[NPH], pub_done[NPH];
    unsigned long long out_empty[NS_OUT];
    int task;
};

template <int KG, int NW, int G, int RB>
__global__ void __launch_bounds__(32 * (NW + 4), (RB == 16 && NW <= 4 && KG <= 5) ? 3 : (NW <= 4 ? 2 : 1)) k_project_systolic4(Sys4Params P4) {
    using SM = Sys4Smem<KG, NW, G, RB>;
    static_assert(RB == 16 || RB == 32, "ring blocks are 16 or 32 rows");
    static_assert(RB % G == 0, "a ring block is a whole number of record intervals");
    constexpr int KB = KG * NW;
    constexpr int NCT = 32 * NW;
    constexpr int NS_IN = SM::NS_IN, NS_OUT = SM::NS_OUT, NR = SM::NR, NPH = SM::NPH;
    static_assert(KG >= 1 && KG <= 8, "a k-group's flags are packed 4 bits per iteration into 32 bits");
    const SysParams &P = P4.base;

    extern __shared__ __align__(128) unsigned char smem_raw[];
    SM &sm = *reinterpret_cast<SM *>(smem_raw);

    const int lane = threadIdx.x & 31, g = threadIdx.x >> 5;
    const bool is_loader = g == NW, is_storer = g == NW + 1, is_feeder = g == NW + 2;
    // An eighth, idle warp: the same kernel launched with NW + 3 warps per CTA is 2-5 % slower at every size (one band
    // alone 111 -> 105 ms for 262144 steps, 4096^2 K=100 6.77 -> 6.45 ms, 32768^2 K=200 674 -> 662 ms), whichever of the
    // warps NW .. NW+3 is the idle one (all eight assignments of loader / storer / feeder / idle tried gave 6.45-6.49
    // and 660-664 ms): the warp count of the CTA matters, not which scheduler hosts which helper.
    if (g == NW + 3) return;
    const long long W = P.W, H = P.H;
    const int n_tasks = P4.n_tasks;
    const Sys4Slab &SL = P4.slab;
    const int cstride = P4.cstride;

#ifdef TFS_TASK_TRACE
    int tt_prev = -1;
#endif
    for (;;) {
        __syncthreads();  // previous task fully retired (smem and mbarrier reuse)
        if (threadIdx.x == 0) {
#ifdef TFS_TASK_TRACE
            {
                unsigned long long now; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
                if (tt_prev >= 0) P4.ttrace[4 * tt_prev + 1] = now;
            }
#endif
            sm.task = atomicAdd(P.ticket, 1);
            for (int i = 0; i < NS_IN; ++i) { mbar_init(&sm.in_full[i], 32); mbar_init(&sm.in_empty[i], 1); }
            for (int i = 0; i < NR; ++i) { mbar_init(&sm.rec_full[i], 32); mbar_init(&sm.rec_empty[i], NW); }
            for (int i = 0; i < NPH; ++i) { mbar_init(&sm.rec_written[i], NW + 1); mbar_init(&sm.pub_done[i], 1); }  // k-groups + feeder
            for (int i = 0; i < NS_OUT; ++i) mbar_init(&sm.out_empty[i], 1);
            asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
        }
        __syncthreads();
        const int ticket = sm.task;
        if (ticket >= n_tasks) break;
#ifdef TFS_TASK_TRACE
        if (threadIdx.x == 0) {
            unsigned long long now; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            unsigned smid; asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
            P4.ttrace[4 * ticket] = now; P4.ttrace[4 * ticket + 2] = smid;
            tt_prev = ticket;
        }
#endif
        const int2 tk = P4.tasks[ticket];
        const int tb = tk.x, b = tk.y, lb = tk.y - SL.b0;  // pass, global band, local band
        const int par = tb & 1;
        const bool last_local = lb == SL.nb_local - 1;
        // only the bands at a slab edge talk to another GPU: everything else keeps the cheaper gpu-scope fences
        const bool sysfence = SL.multi && (lb == 0 || last_local);
        const int k_act = (tb == P.n_passes - 1) ? P.k_last : KB;
        const bool first_pass = tb == 0;
        const long long q0 = (long long)b * 32 - 1;  // skewed coordinate of lane 0
        const bool has_left = b > 0;
        const size_t rec_stride = (size_t)P.n_rec * KG;  // float4 per (band, k-group)
        const int T = P.T;
        const int n_int = (T + G - 1) / G;                 // record intervals
        const int n_blk_in = (int)((H + RB - 1) / RB);      // field blocks that hold real rows
        const int m_last = (int)((H - 1) / RB);             // last output block

        if (is_loader) {
            // ===================================== LOADER ===========================================
            const float *__restrict__ gin_u = P.buf[tb & 1][0];
            const float *__restrict__ gin_v = P.buf[tb & 1][1];
            const float *__restrict__ gin_p = P.buf[tb & 1][2];
            // record slot lb holds what the left band wrote (slot 0: the left rank's last band)
            const float4 *rec_in = P.bnd + ((size_t)par * (SL.nb_local + 1) + lb) * NW * rec_stride;
            const long long *left_prog = P4.bnd_prog64 + (size_t)tb * cstride + lb;
            const long long *dep_out0 = P4.out_prog64 + (size_t)(tb - 1) * cstride + lb + 1;
            const long long *dep_out1 = P4.out_prog64 + (size_t)(tb - 1) * cstride + ((b + 1 < P.n_bands) ? lb + 2 : lb + 1);
            int known_left = 0, known_rows = first_pass ? (int)H : 0;
            const long long seq = SL.seq;
            const float *bu = gin_u - SL.xs0 * H, *bv = gin_v - SL.xs0 * H, *bp = gin_p - SL.xs0 * H;  // global-column bases
            const unsigned char *bf = P.flags - SL.xs0 * H;
            const long long cs0 = SL.xs0, cs1 = SL.xs0 + SL.ncols;  // stored columns

            if (first_pass && last_local && SL.wait_right_halo) {
                // column c_hi (u of the right rank's first column, after ITS advection of the previous step) is read below;
                // the same flag says the right rank's previous kernel is over, so its record slot 0 and mailboxes are free
                if (lane == 0)
                    for (SpinGuard guard; ld_volatile64(SL.wait_right_halo) < SL.wait_right_halo_value; guard.tick()) __nanosleep(1000);
                __syncwarp();
                fence_sys();
            }
            auto load_block = [&](int n) {
                // rows [RB*n, RB*n + RB) of columns q0.. (u: q0+1..) -> slot n % NS_IN
                const int slot = n % NS_IN;
                mbar_wait<4000>(&sm.in_empty[slot], ((n / NS_IN) & 1) ^ 1);
                int need = RB * n + RB;
                if (need > (int)H) need = (int)H;
                if (known_rows < need) {
                    do {
                        int k0 = 0;
                        if (lane == 0) k0 = min(prog_of(ld_volatile64(dep_out0), seq), prog_of(ld_volatile64(dep_out1), seq));
                        known_rows = __shfl_sync(0xffffffffu, k0, 0);
                        if (known_rows < need) __nanosleep(500);
                    } while (known_rows < need);
                    if (sysfence) fence_sys(); else fence_gpu();  // acquire
                }
                if (n < n_blk_in) {
                    const long long row0 = (long long)RB * n;
                    constexpr int PPC = RB / 4;            // 16-byte pieces per column
                    constexpr int CPI = 32 / PPC;          // columns per iteration
                    const int part = lane % PPC;
#pragma unroll
                    for (int it = 0; it < 32 / CPI; ++it) {
                        const int cc = it * CPI + lane / PPC;
                        const long long col = q0 + cc;
                        if (col + 1 >= 0 && col + 1 < W && col + 1 >= cs0 && col + 1 < cs1)
                            cp_async16_cg(&sm.in_u[cc][slot * RB + part * 4], bu + (col + 1) * H + row0 + part * 4);
                        if (col >= 0 && col < W && col >= cs0 && col < cs1) {
                            cp_async16_cg(&sm.in_v[cc][slot * RB + part * 4], bv + col * H + row0 + part * 4);
                            if (!first_pass) cp_async16_cg(&sm.in_p[cc][slot * RB + part * 4], bp + col * H + row0 + part * 4);
                            cp_async4(&sm.in_f[cc][slot * RB + part * 4], bf + col * H + row0 + part * 4);  // static data: L1 is fine
                        }
                    }
                }
                mbar_arrive_on_cp_async(&sm.in_full[slot]);
            };
            auto load_records = [&](int j) {
                const int slot = j % NR;
                mbar_wait<2000>(&sm.rec_empty[slot], ((j / NR) & 1) ^ 1);
                if (has_left) {
                    int need = (j + 1) * G;
                    if (need > P.n_rec) need = P.n_rec;
                    if (known_left < need) {
                        do {
                            int k0 = 0;
                            if (lane == 0) k0 = prog_of(ld_volatile64(left_prog), seq);
                            known_left = __shfl_sync(0xffffffffu, k0, 0);
                            if (known_left < need) __nanosleep(300);
                        } while (known_left < need);
                        if (sysfence) fence_sys(); else fence_gpu();  // acquire
                    }
                    float4 *dst = &sm.rec[slot][0][0][0];
                    for (int e = lane; e < NW * G * KG; e += 32) {
                        const int gg = e / (G * KG), x = e % (G * KG);
                        const long long r4 = (long long)j * G * KG + x;  // float4 index inside the group's stream
                        if (r4 < (long long)rec_stride) cp_async16_cg(dst + e, rec_in + (size_t)gg * rec_stride + r4);
                        else dst[e] = make_float4(0.f, 0.f, 0.f, 0.f);  // past the stream: "lane -1" produces zeros
                    }
                } else {
                    float4 *dst = &sm.rec[slot][0][0][0];
                    for (int e = lane; e < NW * G * KG; e += 32) dst[e] = make_float4(0.f, 0.f, 0.f, 0.f);  // no left band
                }
                __threadfence_block();  // the plain zero stores above must be visible before the slot is handed over
                __syncwarp();
                mbar_arrive_on_cp_async(&sm.rec_full[slot]);
            };

            const int n_blk_needed = (T - 1) / RB + 1;  // blocks the compute warps will wait for
            TFS_TR_DECL
            load_block(0);
            TFS_TR(0)
            for (int j = 0; j < n_int; ++j) {
                load_records(j);
                TFS_TR(1)
                if ((j * G) % RB == 0) {
                    const int n = (j * G) / RB + 1;  // one block ahead of the compute warps
                    if (n < n_blk_needed) load_block(n);
                    TFS_TR(2)
                }
            }
#ifdef TFS_SYS_TRACE
            if (lane == 0 && P.trace) {
                for (int i = 0; i < 8; ++i) atomicAdd((unsigned long long *)&P.trace[(NW * 8 + i)], (unsigned long long)tr_acc[i]);
            }
#endif
        } else if (is_storer) {
            // ===================================== STORER ===========================================
            float *__restrict__ gout_u = P.buf[(tb + 1) & 1][0];
            float *__restrict__ gout_v = P.buf[(tb + 1) & 1][1];
            float *__restrict__ gout_p = P.buf[(tb + 1) & 1][2];
            long long *my_prog = P4.bnd_prog64 + (size_t)tb * cstride + lb + 1;
            long long *my_out_prog = P4.out_prog64 + (size_t)tb * cstride + lb + 1;
            // boundary bands also publish into the neighbour rank's mailboxes
            long long *peer_prog = (last_local && SL.bnd_prog_peerR) ? SL.bnd_prog_peerR + (size_t)tb * SL.cnt_strideR : nullptr;
            long long *peer_out_prog = (lb == 0 && SL.out_prog_peerL) ? SL.out_prog_peerL + (size_t)tb * SL.cnt_strideL : nullptr;
            const long long seq = SL.seq;
            float *lu = gout_u - SL.xs0 * H, *lv = gout_v - SL.xs0 * H, *lp = gout_p - SL.xs0 * H;  // global-column bases
            float *pu = nullptr, *pv = nullptr, *pp_ = nullptr;  // the left rank's copies (columns <= c_lo)
            if (lb == 0 && SL.peerL_buf[(tb + 1) & 1][0]) {
                pu = SL.peerL_buf[(tb + 1) & 1][0] - SL.peerL_xs0 * H;
                pv = SL.peerL_buf[(tb + 1) & 1][1] - SL.peerL_xs0 * H;
                pp_ = SL.peerL_buf[(tb + 1) & 1][2] - SL.peerL_xs0 * H;
            }
            const long long c_lo = SL.c_lo;
            // last pass: band 0 also delivers the left rank's right-halo columns (u, v of [c_lo, c_lo + halo_cols))
            const bool halo_pass = pu && SL.halo_ready_peerL && tb == P.n_passes - 1;
            const long long c_peer_hi = halo_pass ? c_lo + SL.halo_cols - 1 : c_lo;  // last column copied to the left rank
            int m_next = 0, rows_done = 0;
            // block m (rows [32m, 32m+32)) is complete once lane 31 retired its last row
            auto done_step = [&](int m) { return min(RB * m + (RB - 1) + 31 + KB, T - 1); };
            auto store_block = [&](int m) {
                const int slot = m % NS_OUT;
                const long long row0 = (long long)RB * m;
                constexpr int PPC = RB / 4, CPI = 32 / PPC;
                const int part = lane % PPC;
                if (row0 + part * 4 < H) {  // H % 4 == 0: a float4 is entirely inside or outside the column
#pragma unroll
                    for (int it = 0; it < 32 / CPI; ++it) {
                        const int cc = it * CPI + lane / PPC;
                        const long long col = q0 + cc - (KB - 1);
                        if (col >= 0 && col < W) {
                            const long long idx = col * H + row0 + part * 4;
                            const float4 vu = *reinterpret_cast<const float4 *>(&sm.out_u[cc][slot * RB + part * 4]);
                            const float4 vv = *reinterpret_cast<const float4 *>(&sm.out_v[cc][slot * RB + part * 4]);
                            const float4 vp = *reinterpret_cast<const float4 *>(&sm.out_p[cc][slot * RB + part * 4]);
                            if (col >= c_lo) {
                                *reinterpret_cast<float4 *>(lu + idx) = vu;
                                *reinterpret_cast<float4 *>(lv + idx) = vv;
                                *reinterpret_cast<float4 *>(lp + idx) = vp;
                            }
                            if (pu && col <= c_peer_hi) {  // owned by (or a halo column of) the left rank
                                *reinterpret_cast<float4 *>(pu + idx) = vu;
                                *reinterpret_cast<float4 *>(pv + idx) = vv;
                                if (col <= c_lo) *reinterpret_cast<float4 *>(pp_ + idx) = vp;
                            }
                        }
                    }
                }
            };
            TFS_TR_DECL
            for (int j = 0; j < n_int; ++j) {
                TFS_TR(0)
                mbar_wait<2000>(&sm.rec_written[j % NPH], (j / NPH) & 1);  // steps < (j+1)*G retired by every k-group
                TFS_TR(1)
                const int t_done = min((j + 1) * G, T) - 1;
                int m_first = m_next;
                while (m_next <= m_last && done_step(m_next) <= t_done) { store_block(m_next); ++m_next; }
                __syncwarp();
                TFS_TR(2)
                if (sysfence) fence_sys(); else fence_gpu();  // release: block stores above + the records the compute warps wrote
                TFS_TR(3)
                if (lane == 0) {
                    int recs = t_done - 32 + 1;
                    recs = recs < 0 ? 0 : (recs > P.n_rec ? P.n_rec : recs);
                    if (j == n_int - 1) recs = P.n_rec;
                    st_volatile64(my_prog, prog_make(seq, recs));
                    if (peer_prog) st_volatile64(peer_prog, prog_make(seq, recs));
                    if (m_next != m_first) {
                        rows_done = min(RB * m_next, (int)H);
                        st_volatile64(my_out_prog, prog_make(seq, rows_done));
                        if (peer_out_prog) st_volatile64(peer_out_prog, prog_make(seq, rows_done));
                    }
                    for (int m = m_first; m < m_next; ++m) mbar_arrive(&sm.out_empty[m % NS_OUT]);
                    mbar_arrive(&sm.pub_done[j % NPH]);
                    // everything this band owes the left rank for this step is stored and fenced (system scope: lb == 0)
                    if (halo_pass && j == n_int - 1) st_volatile64(SL.halo_ready_peerL, seq);
                }
            }
#ifdef TFS_SYS_TRACE
            if (lane == 0 && P.trace) {
                for (int i = 0; i < 8; ++i) atomicAdd((unsigned long long *)&P.trace[((NW + 1) * 8 + i)], (unsigned long long)tr_acc[i]);
            }
#endif
        } else if (is_feeder) {
            // ===================================== FEEDER ===========================================
            // The k = 0 operands of the first k-group, delivered in the k-groups' own hand-over format (slot 0 of
            // grp) so that all compute warps run the same code: with the ring reads, bounds tests and the gravity
            // add in its step, the first compute warp executed 178 instructions per step against 128 of the others,
            // which waited for it at the step barrier 17 % of their time (ncu source page, round 2).  Iteration tf
            // runs alongside compute step tf and delivers u, v of step tf + 1 and p, flags of step tf + 2 (the
            // two-step path of the compute loop); iterations -2 and -1 prime the first k-group.
            // Operands of step tau: u[q+1, rho], p[q, rho], flags[q, rho], v[q, rho+1], flags[q, rho+1] with
            // rho = tau - lane - 1.
            const bool grav = first_pass && P.apply_gravity;
            const long long q = q0 + lane;
            const bool col_ok = q >= 0 && q < W;
            const bool colp_ok = q + 1 >= 0 && q + 1 < W;
            constexpr int RIN = SM::RIN;
            const int Hi = (int)H;
            int rr = ((-2 - lane) % RIN + RIN) % RIN;  // ring position of row tf - lane at tf = -2
            int rr1 = (rr + 1 == RIN) ? 0 : rr + 1;
            // input blocks: block n is first touched (by lane 0, row RB*n) in iteration RB*n - 1; rows in use are
            // [tf-31, tf+1], so block n-(NS_IN-1) is then behind lane 31 and goes back to the loader
            int in_slot = 0, in_par = 0, in_n = 0;         // next block to acquire
            int rel_slot = 0;                              // slot of block in_n - (NS_IN - 1)
            auto acquire_in_block = [&]() {
                mbar_wait(&sm.in_full[in_slot], in_par);
                if (in_n >= NS_IN - 1) {
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&sm.in_empty[rel_slot]);
                    rel_slot = (rel_slot + 1 == NS_IN) ? 0 : rel_slot + 1;
                }
                ++in_n;
                if (++in_slot == NS_IN) { in_slot = 0; in_par ^= 1; }
            };
            acquire_in_block();
            // drain side: output-ring and interval bookkeeping (the storer waits for the k-groups AND this warp)
            constexpr int ROUT = SM::ROUT;
            int ro = ((-lane - KB) % ROUT + ROUT) % ROUT;   // ring position of rho_out = t - lane - KB at t = 0
            int out_slot = 0, out_par = 0;                  // slot / parity of the next output block that reuses a slot
            int pub_slot = 0, pub_par = 0, jd = 0, sgd = 0; // rec_written / pub_done slot of the interval, interval index, step in it
            for (int tf = -2; tf < T; ++tf) {
                if (tf + 1 < T) {
                    if (tf + 1 > 0 && ((tf + 1) % RB) == 0) acquire_in_block();
                    const int row = tf - lane;  // rho of step tf + 1; rho of step tf + 2 is row + 1
                    const bool row_ok = (unsigned)row < (unsigned)Hi;
                    const bool rowp_ok = (unsigned)(row + 1) < (unsigned)Hi;
                    const float su = sm.in_u[lane][rr];
                    const float sv = sm.in_v[lane][rr1];
                    const float sp = sm.in_p[lane][rr1];
                    const unsigned f1 = (col_ok && rowp_ok) ? (unsigned)sm.in_f[lane][rr1] : 0u;
                    float vv = (col_ok && rowp_ok) ? sv : 0.0f;
                    if (grav && (f1 & TFS_F_LIVE)) vv = fadd(vv, P.gdt);  // simulator.rs:147
                    float4 o;
                    o.x = (colp_ok && row_ok) ? su : 0.0f;
                    o.y = vv;
                    o.z = (!first_pass && col_ok && rowp_ok) ? sp : 0.0f;
                    o.w = __uint_as_float(f1 & 15u);
                    sm.grp[tf & 1][0][lane] = o;
                    rr = rr1;
                    rr1 = (rr1 + 1 == RIN) ? 0 : rr1 + 1;
                }
                bar_sync_imm<BAR_STEP, NCT + 32>();
                if (tf >= 0) {
                    // ---- drain: iteration KB-1 of step tf retired -> final values into the output ring
                    const int tk = tf - KB;  // lane 0's output row
                    if (tk >= 0 && (tk % RB) == 0 && tk / RB >= NS_OUT) {  // lane 0 enters a block whose slot was used before:
                        mbar_wait(&sm.out_empty[out_slot], out_par);      // the storer must have emptied it
                        if (++out_slot == NS_OUT) { out_slot = 0; out_par ^= 1; }
                    }
                    const float4 o = sm.grp[tf & 1][NW][lane];
                    if (tf - lane - KB >= 0) {
                        sm.out_u[lane][ro] = o.x;
                        sm.out_v[lane][ro] = o.y;
                        sm.out_p[lane][ro] = o.z;
                    }
                    ro = (ro + 1 == ROUT) ? 0 : ro + 1;
                    if (sgd == G - 1 || tf == T - 1) {  // end of a record interval: this warp's part of "steps retired"
                        __syncwarp();
                        if (lane == 0) {
                            if (jd >= NPH) mbar_wait(&sm.pub_done[pub_slot], pub_par ^ 1);
                            mbar_arrive(&sm.rec_written[pub_slot]);
                        }
                        ++jd;
                        if (++pub_slot == NPH) { pub_slot = 0; pub_par ^= 1; }
                        sgd = 0;
                    } else {
                        ++sgd;
                    }
                }
            }
        } else {
            // ================================== compute warps =========================================
            // records go to slot lb + 1, or straight into slot 0 of the right rank for the last local band
            float4 *rec_out = (last_local && SL.rec_peerR[par])
                                  ? SL.rec_peerR[par] + (size_t)g * rec_stride
                                  : P.bnd + (((size_t)par * (SL.nb_local + 1) + lb + 1) * NW + g) * rec_stride;

            float uL[KG], uRo[KG], vB[KG], vTo[KG], pin[KG], pd[KG];
#pragma unroll
            for (int k = 0; k < KG; ++k) uL[k] = uRo[k] = vB[k] = vTo[k] = pin[k] = pd[k] = 0.0f;
            unsigned fl = 0u, fld = 0u;  // KG nibbles: flags used this step / arriving for the step after next
            constexpr unsigned FULL = (KG == 8) ? 0xFFFFFFFFu : ((1u << (4 * KG)) - 1u);
            const bool group_active = (g + 1) * KG <= k_act;  // every iteration of this k-group runs in this pass
            // record-interval bookkeeping (all k-groups) and output-ring bookkeeping (last k-group)
            int rec_slot = 0, rec_par = 0;                  // slot / parity of the interval being consumed
            int pub_slot = 0, pub_par = 0;                  // rec_written / pub_done slot of this interval
            int j = 0;                                      // interval index
            const bool last_group = g == NW - 1;
            const float4 *rec_base = &sm.rec[0][g][0][0];
            constexpr int REC_SLOT_STRIDE = NW * G * KG;    // float4 per ring slot

            // two priming rounds for the first k-group: p, flags of step 0, then u, v of step 0 and p, flags of step 1
#pragma unroll
            for (int pr = 0; pr < 2; ++pr) {
                bar_sync_imm<BAR_STEP, NCT + 32>();
                if (g == 0) {
                    const float4 gi = sm.grp[pr][0][lane];
                    uRo[0] = gi.x;
                    vTo[0] = gi.y;
                    pin[0] = pd[0];
                    pd[0] = gi.z;
                    fld |= __float_as_uint(gi.w);
                    if (pr == 0) { fl = fld; fld = 0u; }
                }
            }
            TFS_TR_DECL
            for (int t = 0, sg = 0; t < T; ++t) {
                TFS_TR(0)
                if (sg == 0) {
                    mbar_wait(&sm.rec_full[rec_slot], rec_par);
                }
                TFS_TR(6)
                // ---- early loads: left band's record (for lane 0)
                float4 lr[KG];
                {
                    const float4 *rp = rec_base + rec_slot * REC_SLOT_STRIDE + sg * KG;
#pragma unroll
                    for (int kk = 0; kk < KG; ++kk) lr[kk] = rp[kk];  // broadcast read
                }
                TFS_TR(1)
                // ---- KG independent cell updates ---------------------------------------------------
                const unsigned fl_used = fl;
                float a[KG], r[KG], cv[KG], d[KG], pp[KG];
#pragma unroll
                for (int kk = 0; kk < KG; ++kk) { a[kk] = uL[kk]; r[kk] = uRo[kk]; cv[kk] = vB[kk]; d[kk] = vTo[kk]; pp[kk] = pin[kk]; }
                unsigned fl_eff = fl_used & FULL;
                if (!group_active) {
#pragma unroll
                    for (int kk = 0; kk < KG; ++kk)
                        if (g * KG + kk >= k_act) fl_eff &= ~(15u << (4 * kk));
                }
                const unsigned spread = (fl_eff & (FULL / 15u)) * 15u;  // 0xF where bit 0 of the nibble is set
                const bool plain_or_off = fl_eff == spread;            // every nibble is 0x0 or 0xF
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
                TFS_TR(2)
                // ---- lane L-1's outputs ("lane -1" = lane 31 of the left band, zeros where there is none)
                float n_uL[KG], s_v[KG], s_p[KG];
#pragma unroll
                for (int kk = 0; kk < KG; ++kk) {
                    n_uL[kk] = __shfl_up_sync(0xffffffffu, r[kk], 1);
                    s_v[kk] = __shfl_up_sync(0xffffffffu, cv[kk], 1);
                    s_p[kk] = __shfl_up_sync(0xffffffffu, pp[kk], 1);
                }
                unsigned s_flv = __shfl_up_sync(0xffffffffu, fl_used, 1);
#pragma unroll
                for (int kk = 0; kk < KG; ++kk) {  // (the loader zero-fills the record slot where there is no left band)
                    n_uL[kk] = (lane == 0) ? lr[kk].x : n_uL[kk];
                    s_v[kk] = (lane == 0) ? lr[kk].y : s_v[kk];
                    s_p[kk] = (lane == 0) ? lr[kk].z : s_p[kk];
                }
                s_flv = (lane == 0) ? __float_as_uint(lr[0].w) : s_flv;
                TFS_TR(3)
                // ---- hand-over to the next k-group: {uCf, lane L-1's vCf, lane L-1's pout, its flags nibble}; the last
                // k-group hands its own lane's final u, v, p of cell (q-(KB-1), t-lane-KB) to the feeder warp, which
                // moves them into the output ring (three stores and the ring bookkeeping less in this warp's step)
                sm.grp[t & 1][g + 1][lane] = make_float4(a[KG - 1], last_group ? cv[KG - 1] : s_v[KG - 1], last_group ? pp[KG - 1] : s_p[KG - 1],
                                                         __uint_as_float((s_flv >> (4 * (KG - 1))) & 15u));
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
                // ---- end of a record interval: hand the slot back, tell the storer the records are written
                if (sg == G - 1 || t == T - 1) {
                    __syncwarp();
                    if (lane == 0) {
                        mbar_arrive(&sm.rec_empty[rec_slot]);
                        if (j >= NPH) mbar_wait(&sm.pub_done[pub_slot], pub_par ^ 1);
                        mbar_arrive(&sm.rec_written[pub_slot]);
                    }
                    ++j;
                    if (++rec_slot == NR) { rec_slot = 0; rec_par ^= 1; }
                    if (++pub_slot == NPH) { pub_slot = 0; pub_par ^= 1; }
                    sg = 0;
                } else {
                    ++sg;
                }
                TFS_TR(4)
                bar_sync_imm<BAR_STEP, NCT + 32>();
                TFS_TR(5)
                {  // from the previous k-group, or from the feeder warp (g == 0)
                    const float4 gi = sm.grp[t & 1][g][lane];
                    uRo[0] = gi.x;
                    vTo[0] = gi.y;
                    pin[0] = pd[0];
                    pd[0] = gi.z;
                    fld |= __float_as_uint(gi.w);  // the nibble for slot 0, two steps from now
                }
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

// Everything a slab handle provides from its arena (tfs_api.cu); nullptr in SystolicArgs = single GPU.
struct Sys4External {
    float4 *bnd = nullptr;          // record slots
    size_t bnd_cap_f4 = 0;          // capacity in float4
    long long *cnt = nullptr;       // [bnd_prog: max_passes*cstride][out_prog: max_passes*cstride]
    int *ticket = nullptr;
    int max_passes = 0, cstride = 0;
    long long xs0 = 0, ncols = 0, c_lo = 0;
    int b0 = 0, b0_next = -1;       // first band of this rank / of the right rank (-1: this is the last rank)
    long long seq = 1;
    // left neighbour (all null / 0 if none)
    float *peerL_buf[2][3] = {{nullptr, nullptr, nullptr}, {nullptr, nullptr, nullptr}};
    long long peerL_xs0 = 0;
    long long *peerL_cnt = nullptr; int peerL_cstride = 0, peerL_b0 = 0, peerL_max_passes = 0;
    // right neighbour
    float4 *peerR_bnd = nullptr;
    long long *peerR_cnt = nullptr; int peerR_cstride = 0, peerR_b0 = 0, peerR_b0_next = -1;
    // in-kernel hand-offs (see Sys4Slab); halo_cols > 0 asks for them, *inkernel_halo reports whether the chosen
    // shape can deliver the halo from band 0 alone (KB + halo_cols - 1 <= 32)
    const long long *wait_right_halo = nullptr; long long wait_right_halo_value = 0;
    long long *halo_ready_peerL = nullptr; int halo_cols = 0;
    bool *inkernel_halo = nullptr;
};

struct Systolic4Workspace {
    int2 *d_tasks = nullptr;
    int2 *h_tasks = nullptr;  // pinned, so the upload never blocks the host
    size_t n_tasks_cap = 0;
    long long key_W = -1, key_H = -1;
    int key_passes = -1, key_bands = -1, key_b0 = -1, key_nbl = -1, key_kb = -1;
    long long seq = 0;
    int last_kg = 0, last_nw = 0, last_passes = 0;  // shape of the last launch (tfs_get_last_projection)
};

// Allocate the task list up front (x-slab handles: tfs_create).  cudaMalloc waits for the device to go idle, which
// inside a step would deadlock two slab handles that share ONE device (the other rank's kernels spin on flags this
// rank has not raised yet).
inline cudaError_t systolic4_reserve(Systolic4Workspace &ws4, size_t n_tasks) {
    if (ws4.n_tasks_cap >= n_tasks) return cudaSuccess;
    if (ws4.d_tasks) cudaFree(ws4.d_tasks);
    if (ws4.h_tasks) cudaFreeHost(ws4.h_tasks);
    ws4.d_tasks = nullptr; ws4.h_tasks = nullptr; ws4.n_tasks_cap = 0;
    cudaError_t e;
    if ((e = cudaMalloc((void **)&ws4.d_tasks, n_tasks * sizeof(int2))) != cudaSuccess) return e;
    if ((e = cudaMallocHost((void **)&ws4.h_tasks, n_tasks * sizeof(int2))) != cudaSuccess) return e;
    ws4.n_tasks_cap = n_tasks;
    return cudaSuccess;
}

template <int KG, int NW, int G, int RB>
struct Sys4Launch {
    static constexpr int KB = KG * NW;
    static cudaError_t run(const SystolicArgs &a, const Sys4External *ext, SystolicWorkspace &ws, Systolic4Workspace &ws4,
                           uint64_t *launched, const char **msg) {
        cudaError_t e;
        int dev = 0;
        if ((e = cudaGetDevice(&dev)) != cudaSuccess) return e;
        if (ws.num_sms == 0) {
            if ((e = cudaDeviceGetAttribute(&ws.num_sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return e;
        }
        const int K = a.iterations;
        Sys4Params P4{};
        SysParams &P = P4.base;
        Sys4Slab &SL = P4.slab;
        P.W = a.W; P.H = a.H; P.ncells = a.W * a.H;
        P.n_passes = (K + KB - 1) / KB;
        P.k_last = K - (P.n_passes - 1) * KB;
        P.n_bands = (int)((a.W + KB + 31) / 32);  // global
        P.T = (int)a.H + 32 + KB;
        P.n_rec = (int)a.H + KB;
        P.apply_gravity = a.fuse_gravity ? 1 : 0;
        P.omega = a.omega; P.pc = a.pc; P.gdt = a.gdt;
        P.flags = a.flags;
        P.buf[0][0] = *a.u; P.buf[0][1] = *a.v; P.buf[0][2] = *a.p;
        P.buf[1][0] = *a.u2; P.buf[1][1] = *a.v2; P.buf[1][2] = *a.p2;
        P.trace = nullptr;
        const size_t band_stride = (size_t)NW * P.n_rec * KG;  // float4 per record slot
        if (!ext) {
            SL.xs0 = 0; SL.ncols = a.W; SL.c_lo = 0; SL.b0 = 0; SL.nb_local = P.n_bands; SL.multi = 0;
            SL.seq = ++ws4.seq;
            P4.cstride = SL.nb_local + 2;
            const size_t rec_bytes = (size_t)2 * (SL.nb_local + 1) * band_stride * sizeof(float4);
            const size_t n_cnt = (size_t)2 * P.n_passes * P4.cstride;
            const size_t need = rec_bytes + n_cnt * sizeof(long long) + 64;
            if (ws.bytes < need) {
                if (ws.buf) cudaFree(ws.buf);
                ws.buf = nullptr;
                ws.bytes = 0;
                if ((e = cudaMalloc(&ws.buf, need)) != cudaSuccess) { *msg = "workspace allocation"; return e; }
                ws.bytes = need;
            }
            P.bnd = reinterpret_cast<float4 *>(ws.buf);
            P4.bnd_prog64 = reinterpret_cast<long long *>(reinterpret_cast<char *>(ws.buf) + rec_bytes);
            // The counters' position depends on (W, H, shape, passes) and the buffer is shared with the record
            // streams and with the generic kernel, so whatever lies there now may be stale records of another
            // layout: zero them on every launch (a few hundred KB at most; the slab path keeps fixed offsets and
            // relies on `seq` instead, because a neighbour may publish before this rank launches).
            if ((e = cudaMemsetAsync(P4.bnd_prog64, 0, n_cnt * sizeof(long long) + 64, a.stream)) != cudaSuccess) return e;
            P4.out_prog64 = P4.bnd_prog64 + (size_t)P.n_passes * P4.cstride;
            P.ticket = reinterpret_cast<int *>(P4.out_prog64 + (size_t)P.n_passes * P4.cstride);
        } else {
            SL.xs0 = ext->xs0; SL.ncols = ext->ncols; SL.c_lo = ext->c_lo; SL.b0 = ext->b0;
            SL.nb_local = (ext->b0_next >= 0 ? ext->b0_next : P.n_bands) - ext->b0;
            SL.multi = 1;
            SL.seq = ext->seq;
            if (SL.nb_local < 1) { *msg = "a slab must own at least one band"; return cudaErrorInvalidValue; }
            if (P.n_passes > ext->max_passes || SL.nb_local + 2 > ext->cstride ||
                (size_t)2 * (SL.nb_local + 1) * band_stride > ext->bnd_cap_f4) {
                *msg = "slab workspace too small for this iteration count";
                return cudaErrorInvalidValue;
            }
            P4.cstride = ext->cstride;
            P.bnd = ext->bnd;
            P4.bnd_prog64 = ext->cnt;
            P4.out_prog64 = ext->cnt + (size_t)ext->max_passes * ext->cstride;
            P.ticket = ext->ticket;
            for (int i = 0; i < 2; ++i)
                for (int f = 0; f < 3; ++f) SL.peerL_buf[i][f] = ext->peerL_buf[i][f];
            SL.peerL_xs0 = ext->peerL_xs0;
            if (ext->peerL_cnt) {  // the left rank's out_prog mailbox: its index nbL + 1
                const int nbL = ext->b0 - ext->peerL_b0;
                SL.out_prog_peerL = ext->peerL_cnt + (size_t)ext->peerL_max_passes * ext->peerL_cstride + (nbL + 1);
                SL.cnt_strideL = ext->peerL_cstride;
            }
            SL.wait_right_halo = ext->wait_right_halo;
            SL.wait_right_halo_value = ext->wait_right_halo_value;
            // band b0 stores the columns [c_lo - KB + 1, c_lo + 33 - KB): with this shape, do they include the halo?
            // (a property of the shape alone, so every rank reaches the same answer)
            const bool halo_fits = ext->halo_cols > 0 && KB + ext->halo_cols - 1 <= 32;
            if (ext->inkernel_halo) *ext->inkernel_halo = halo_fits;
            if (halo_fits && ext->halo_ready_peerL) {
                SL.halo_ready_peerL = ext->halo_ready_peerL;
                SL.halo_cols = ext->halo_cols;
            }
            if (ext->peerR_bnd) {
                const int nbR = (ext->peerR_b0_next >= 0 ? ext->peerR_b0_next : P.n_bands) - ext->peerR_b0;
                SL.rec_peerR[0] = ext->peerR_bnd;
                SL.rec_peerR[1] = ext->peerR_bnd + (size_t)(nbR + 1) * band_stride;
                SL.bnd_prog_peerR = ext->peerR_cnt;  // its index 0
                SL.cnt_strideR = ext->peerR_cstride;
            }
        }
        if ((e = cudaMemsetAsync(P.ticket, 0, sizeof(int), a.stream)) != cudaSuccess) return e;
        // task list (local bands) in "start time" order: a band trails its left neighbour by ~lag_b steps
        // and the same band of the previous pass by ~lag_t steps; every dependency has a strictly smaller
        // key, on this rank and on its neighbours alike.
        const long long n_tasks = (long long)P.n_passes * SL.nb_local;
        if (ws4.key_W != a.W || ws4.key_H != a.H || ws4.key_passes != P.n_passes || ws4.key_bands != P.n_bands ||
            ws4.key_b0 != SL.b0 || ws4.key_nbl != SL.nb_local || ws4.key_kb != KB) {
            if ((e = systolic4_reserve(ws4, (size_t)n_tasks)) != cudaSuccess) { *msg = "task list allocation"; return e; }
            // Issue order.  "diag": key = start time (a band trails its left neighbour by ~lag_b steps, the
            // previous pass by ~lag_t): all passes advance together, good when every task is resident.
            // "lex": pass-major.  With more tasks than CTA slots the diagonal order makes the band chain of
            // pass 0 wait behind later passes of earlier bands, which serialises the GPUs of a slab run
            // (rank r+1 cannot start before rank r's LAST band runs); pass-major lets pass 0 sweep first.
            int lag_b = 32 + 2 * G, lag_t = 2 * lag_b + 64 + KB;  // lag_t > lag_b: (b+1, tb-1) sorts before (b, tb)
            bool hybrid = false;  // pass 0 first (feeds the next rank's bands), later passes in start-time order
            {
                const char ord = tunables().sys_order;
                const long long slots = (long long)ws.num_sms * 3;  // resident CTAs of the shipped shapes
                // measured on 8 B200 at 32768^2, K=200: pass-major 190 ms/step, hybrid 505, start-time 524
                const bool lex = ord ? (ord == 'l') : (ext != nullptr && n_tasks > slots);
                hybrid = ord == 'h';
                if (lex) lag_t = 1 << 24;
            }
            std::vector<std::pair<long long, int2>> v;
            v.reserve((size_t)n_tasks);
            for (int tb = 0; tb < P.n_passes; ++tb)
                for (int b = SL.b0; b < SL.b0 + SL.nb_local; ++b)
                    v.push_back({(long long)b * lag_b + (long long)tb * lag_t + ((hybrid && tb > 0) ? (long long)SL.nb_local * lag_b : 0LL),
                                 make_int2(tb, b)});
            std::stable_sort(v.begin(), v.end(), [](const auto &x, const auto &y) { return x.first < y.first; });
            for (size_t i = 0; i < v.size(); ++i) ws4.h_tasks[i] = v[i].second;
            if ((e = cudaMemcpyAsync(ws4.d_tasks, ws4.h_tasks, (size_t)n_tasks * sizeof(int2), cudaMemcpyHostToDevice, a.stream)) != cudaSuccess) return e;
            ws4.key_W = a.W; ws4.key_H = a.H; ws4.key_passes = P.n_passes; ws4.key_bands = P.n_bands;
            ws4.key_b0 = SL.b0; ws4.key_nbl = SL.nb_local; ws4.key_kb = KB;
        }
        P4.tasks = ws4.d_tasks;
        P4.n_tasks = (int)n_tasks;
        auto kern = k_project_systolic4<KG, NW, G, RB>;
        const size_t smem = sizeof(Sys4Smem<KG, NW, G, RB>) + 128;
        static bool attr_set[64] = {false};  // function attributes are per device
        static int occ_dev[64] = {0};
        const int di = dev & 63;
        if (!attr_set[di]) {
            if ((e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) {
                *msg = "cudaFuncSetAttribute(max dynamic smem)";
                return e;
            }
            if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_dev[di], kern, 32 * (NW + 4), smem)) != cudaSuccess) return e;
            attr_set[di] = true;
        }
        const int occ = occ_dev[di];
        if (occ < 1) { *msg = "kernel does not fit on an SM"; return cudaErrorLaunchOutOfResources; }
        int grid = (int)std::min<long long>(n_tasks, (long long)occ * ws.num_sms);
        if (const int ge = tunables().sys_grid) {  // tuning aid: -1 ("even") = whole rounds, or a CTA count
            if (ge < 0) {
                const long long rounds = (n_tasks + grid - 1) / grid;
                grid = (int)((n_tasks + rounds - 1) / rounds);
            } else grid = std::min(grid, ge);
        }
#ifdef TFS_SYS_TRACE
        static long long *d_trace = nullptr;
        if (!d_trace) cudaMalloc((void **)&d_trace, 64 * 8 * sizeof(long long));
        cudaMemsetAsync(d_trace, 0, 64 * 8 * sizeof(long long), a.stream);
        P.trace = d_trace;
#endif
#ifdef TFS_TASK_TRACE
        // debug build only: per-task start / end times, dumped to $TFS_TASK_TRACE_PREFIX.b<first band>.bin after every launch
        static unsigned long long *d_tt[64] = {nullptr};
        static size_t d_tt_cap[64] = {0};
        if (d_tt_cap[di] < (size_t)n_tasks) {
            if (d_tt[di]) cudaFree(d_tt[di]);
            cudaMalloc((void **)&d_tt[di], (size_t)n_tasks * 4 * sizeof(unsigned long long));
            d_tt_cap[di] = (size_t)n_tasks;
        }
        cudaMemsetAsync(d_tt[di], 0, (size_t)n_tasks * 4 * sizeof(unsigned long long), a.stream);
        P4.ttrace = d_tt[di];
#endif
        kern<<<grid, 32 * (NW + 4), smem, a.stream>>>(P4);
        if ((e = cudaGetLastError()) != cudaSuccess) { *msg = "kernel launch"; return e; }
        *launched += 1;
        ws4.last_kg = KG; ws4.last_nw = NW; ws4.last_passes = P.n_passes;
#ifdef TFS_TASK_TRACE
        if (const char *pre = tunables().task_trace_prefix) {
            cudaStreamSynchronize(a.stream);
            std::vector<unsigned long long> h((size_t)n_tasks * 4);
            cudaMemcpy(h.data(), d_tt[di], h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
            char path[512];
            snprintf(path, sizeof(path), "%s.b%d.bin", pre, SL.b0);
            if (FILE *f = fopen(path, "wb")) {
                const long long hdr[8] = {n_tasks, P.n_passes, SL.nb_local, SL.b0, grid, P.T, KB, 0};
                fwrite(hdr, sizeof(hdr), 1, f);
                fwrite(ws4.h_tasks, sizeof(int2), (size_t)n_tasks, f);
                fwrite(h.data(), sizeof(unsigned long long), h.size(), f);
                fclose(f);
            }
        }
#endif
#ifdef TFS_SYS_TRACE
        {
            long long h[64 * 8];
            cudaStreamSynchronize(a.stream);
            cudaMemcpy(h, P.trace, sizeof(h), cudaMemcpyDeviceToHost);
            const double nsteps = (double)P.T * n_tasks, nint = (double)((P.T + G - 1) / G) * n_tasks;
            const char *names[8] = {"loop", "early loads", "updates", "shuffles+patch", "stores+shift", "STEP bar issue", "rec_full wait", "-"};
            for (int gq = 0; gq < NW; gq += (NW > 1 ? NW - 1 : 1)) {
                fprintf(stderr, "[trace4] warp %d cycles/step:", gq);
                for (int i = 0; i < 7; ++i) fprintf(stderr, " %s=%.0f", names[i], h[gq * 8 + i] / nsteps);
                fprintf(stderr, "\n");
            }
            fprintf(stderr, "[trace4] loader cycles/interval: prologue=%.0f records=%.0f block=%.0f\n", h[NW * 8 + 0] / nint, h[NW * 8 + 1] / nint, h[NW * 8 + 2] / nint);
            fprintf(stderr, "[trace4] storer cycles/interval: loop=%.0f wait_written=%.0f store=%.0f fence=%.0f\n", h[(NW + 1) * 8 + 0] / nint,
                    h[(NW + 1) * 8 + 1] / nint, h[(NW + 1) * 8 + 2] / nint, h[(NW + 1) * 8 + 3] / nint);
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

inline void systolic4_release(Systolic4Workspace &ws4) {
    if (ws4.d_tasks) cudaFree(ws4.d_tasks);
    if (ws4.h_tasks) cudaFreeHost(ws4.h_tasks);
    ws4.d_tasks = nullptr;
    ws4.h_tasks = nullptr;
    ws4.n_tasks_cap = 0;
    ws4.key_W = ws4.key_H = -1;
}

// The block-ring kernel needs H % 4 == 0 (16-byte column alignment); otherwise the generic cp.async-4
// kernel of tfs_systolic.cuh runs (single GPU only).  The choice depends on H alone, never on the iteration
// count: every shape systolic_pick can return for v4_only has a case below.
inline bool systolic4_usable(const SystolicArgs &a) { return (a.H % 4) == 0 && !tunables().sys_v3; }

// Force the lazy loader to load every block-ring instantiation now (see preload_step_kernels in tfs_api.cu).
inline void systolic4_preload() {
    cudaFuncAttributes fa;
#define TFS_SYS4_PRELOAD(KG_, NW_) cudaFuncGetAttributes(&fa, k_project_systolic4<KG_, NW_, 8, 16>);
    TFS_SYS4_PRELOAD(4, 4) TFS_SYS4_PRELOAD(5, 4) TFS_SYS4_PRELOAD(6, 4) TFS_SYS4_PRELOAD(4, 8) TFS_SYS4_PRELOAD(4, 3)
    TFS_SYS4_PRELOAD(2, 5) TFS_SYS4_PRELOAD(4, 2) TFS_SYS4_PRELOAD(2, 8) TFS_SYS4_PRELOAD(2, 4) TFS_SYS4_PRELOAD(2, 2)
    TFS_SYS4_PRELOAD(5, 2) TFS_SYS4_PRELOAD(1, 8) TFS_SYS4_PRELOAD(8, 2) TFS_SYS4_PRELOAD(8, 4) TFS_SYS4_PRELOAD(8, 1)
#undef TFS_SYS4_PRELOAD
    cudaFuncGetAttributes(&fa, k_project_systolic4<4, 4, 4, 32>);
    cudaFuncGetAttributes(&fa, k_project_systolic4<4, 4, 8, 32>);
    cudaGetLastError();
}

inline cudaError_t systolic_project_auto(const SystolicArgs &a, const Sys4External *ext, SystolicWorkspace &ws,
                                         Systolic4Workspace &ws4, uint64_t *launched, const char **msg) {
    ws4.last_kg = ws4.last_nw = ws4.last_passes = 0;
    if (!systolic4_usable(a)) {
        if (ext) { *msg = "x-slab handles need H % 4 == 0"; return cudaErrorInvalidValue; }
        return systolic_project(a, ws, launched, msg);
    }
    const SysShape sh = systolic_pick(a.iterations, a.temporal_block, true);
    const int G_ = tunables().sys_g, RB_ = tunables().sys_rb;
    // the 32-row-ring / G = 4 variants are tuning aids, built for the default shape only
    if (sh.kg == 4 && sh.nw == 4 && (RB_ != 16 || G_ == 4)) {
        if (G_ == 4) return Sys4Launch<4, 4, 4, 32>::run(a, ext, ws, ws4, launched, msg);
        return Sys4Launch<4, 4, 8, 32>::run(a, ext, ws, ws4, launched, msg);
    }
#define TFS_SYS4_CASE(KG_, NW_) \
    if (sh.kg == KG_ && sh.nw == NW_) return Sys4Launch<KG_, NW_, 8, 16>::run(a, ext, ws, ws4, launched, msg);
    TFS_SYS4_CASE(4, 4)
    TFS_SYS4_CASE(5, 4)
    TFS_SYS4_CASE(6, 4)
    TFS_SYS4_CASE(4, 8)
    TFS_SYS4_CASE(4, 3)
    TFS_SYS4_CASE(2, 5)
    TFS_SYS4_CASE(4, 2)
    TFS_SYS4_CASE(2, 8)
    TFS_SYS4_CASE(2, 4)
    TFS_SYS4_CASE(2, 2)
    TFS_SYS4_CASE(5, 2)
    TFS_SYS4_CASE(1, 8)
    TFS_SYS4_CASE(8, 2)
    TFS_SYS4_CASE(8, 4)
    TFS_SYS4_CASE(8, 1)
#undef TFS_SYS4_CASE
    *msg = "internal: systolic_pick returned a shape without a block-ring instantiation";
    return cudaErrorInvalidValue;
}

}  // namespace tfs
