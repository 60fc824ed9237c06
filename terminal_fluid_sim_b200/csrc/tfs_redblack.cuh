// tfs_redblack.cuh -- make_incompressible in red-black order (TFS_MODE_RED_BLACK), temporally blocked.
//
// Same cell update as simulator.rs:161-186, but the sweep order is "all cells with (i + j) even, then all
// with (i + j) odd" per iteration (cells of one colour touch disjoint faces).  The result is bit-identical
// to the oracle's own statement of that order (orc_make_incompressible_redblack); against the reference's
// lexicographic order it is compared under the measured tolerance stated in tests/test_gpu_parity.py.
//
// One PASS fuses KB iterations (2*KB half-sweeps = "stages") and is one kernel launch; passes ping-pong
// between the two (u, v, p) buffer sets like the systolic kernel, so HBM sees ~(13 r + 12 w)/KB bytes per
// cell-iteration instead of >= 42 for one launch per colour.
//
//   * a WARP owns a strip of 64 rows; lane l holds the row pair (j0 + 2l, j0 + 2l + 1) -- one cell of each
//     colour per column -- so every lane has exactly one active cell per column and half-sweep, and which
//     of the two it is is known at compile time (column parity);
//   * the warp marches in +x holding a window of NS = 2*KB columns in REGISTERS.  At march step x stage s
//     (half-sweep s of the pass) updates column x-1-s: its four neighbours were brought to half-sweep s-1 by
//     stage s-1 earlier in the same step (column x-s), two steps ago (column x-2-s) or one step ago (the
//     rows above / below, same column) -- a software pipeline over the half-sweeps with no synchronisation
//     between warps at all.  Column x-NS leaves the window final after step x;
//   * vertical neighbours inside the strip are a register of the same lane (pair partner) or of lane l+1
//     (two shuffles per upper-row update: fetch v[top], hand the updated value back);
//   * strips and column chunks OVERLAP by NS rows / columns and recompute the overlap: a wrong value at the
//     edge of what was loaded moves inwards by at most one cell per half-sweep, so after NS stages the
//     inner 64 - 2*NS rows and Lx columns are exact.  No halo exchange, no cross-CTA dependency, and the
//     same kernel serves x-slabs (the slab's stored halo columns are the chunk overlap; the slab keeps NS + 2 of
//     them because its outermost stored column has a wrong flag byte, see kHalo in tfs_api.cu).
//   * warp-uniform fast update when every cell of a column's 64 rows is plain interior fluid (11 flops/cell).
//
// KB = 3 (kRbKB): 6 window columns, 52 exact rows per strip, 80 registers -> 6 CTAs (24 warps) per SM and 36 KB of
// unrolled march loop.  Measured on one B200, ms per projection at 1024^2 K=40 / 4096^2 K=100 / 16384^2 K=50 /
// 32768^2 K=200: KB = 4 (8 columns, 48 rows, 96 registers, 5 CTAs, 57 KB of loop) 0.374 / 4.55 / 28.5 / 571;
// KB = 3: 0.343 / 4.17 / 23.7 / 495; KB = 2 (64 registers, 8 CTAs, twice the passes) 0.332 / 5.47 / 31.2 / 491;
// KB = 3 squeezed to 72 registers for 7 CTAs 0.370 / 4.66 / 25.6 / 522.  The kernel is bound by dependent-issue
// latency and instruction fetch, not by HBM: more resident warps and a smaller loop win over fewer passes.
// Measured and rejected at KB = 4 (4096^2 K=100 / 32768^2 2 % obstacles K=200): a straight-line step for an all-plain
// window 4.73 / 745, an instantiation without the per-stage "stage active" test 4.85 / 684 against 4.63 / 660; the
// general update out of line (__noinline__) 6.9 / 948.
#pragma once
#include "tfs_systolic.cuh"  // cell_update4

namespace tfs {

struct RbParams {
    const float *in_u, *in_v, *in_p;  // set read by this pass (storage pointers: column xs0 first)
    float *out_u, *out_v, *out_p;     // set written by this pass
    const uint8_t *flags;
    int W, H;            // global grid
    int xs0, ncols;      // stored columns [xs0, xs0 + ncols)
    int c_lo, c_hi;      // owned columns: the ones this launch writes
    int n_strips;        // strips of V = 64 - 2*NS rows
    int chunk0, lx;      // first column chunk (global index), chunk length (multiple of NS)
    int ns_act;          // active stages of this pass (2 * iterations fused in it)
    int pf;              // L2 prefetch distance in columns (0 = none)
    int step_sync;       // 1: the CTA's warps meet at a barrier after every march step
    int apply_gravity;   // first pass: v += gdt on updatable cells while loading (simulator.rs:137-149)
    float omega, pc, gdt;
};

// CG: field loads bypass L1 (ld.global.cg) - x-slab handles, whose halo columns are rewritten between launches by
// ANOTHER stream (the neighbour's push): a line kept in L1 from an earlier launch would be stale when two slabs share one
// device.  A single-GPU handle reads through L1 (ld.global.nc), where the rows shared by adjacent strips hit.
template <int KB, bool FIRST, bool CG>
__global__ void __launch_bounds__(128, KB <= 2 ? 8 : (KB <= 3 ? 6 : 5)) k_project_rb(RbParams P) {
    constexpr int NS = 2 * KB;       // stages per pass = columns in the register window = overlap
    constexpr int V = 64 - 2 * NS;   // rows of a strip that are exact after NS stages
    static_assert(V >= 8, "too many fused iterations for a 64-row strip");
    __shared__ float4 lut[32];  // cell_update_lut's table (tfs_systolic.cuh)
    nb_lut_fill(lut, threadIdx.x, blockDim.x);
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int strip = blockIdx.x * (blockDim.x >> 5) + warp;
    const int H = P.H, xs0 = P.xs0, ncols = P.ncols;
    const int rA = strip * V - NS + 2 * lane;             // lower row of this lane's pair (even)
    // H is even: the pair is inside or outside together.  (A warp past the last strip marches along with all rows
    // outside: it has to keep meeting the CTA's per-step barrier below.)
    const bool row_ok = rA >= 0 && rA < H && strip < P.n_strips;
    const bool store_lane = row_ok && lane >= KB && lane < 32 - KB;
    const int gx0 = (P.chunk0 + (int)blockIdx.y) * P.lx;  // first column this chunk finalises
    const int vx0 = max(gx0, P.c_lo), vx1 = min(gx0 + P.lx, P.c_hi);
    const int ns_act = P.ns_act;
    const float omega = P.omega, pc = P.pc;

    // register window: slot c holds column i with i % NS == c (rows A = lower, B = upper)
    float uA[NS], uB[NS], vA[NS], vB[NS], pA[NS], pB[NS];
    unsigned fl[NS];      // flag bytes: A | B << 8
    unsigned plain = 0u;  // bit c: all 64 cells of slot c are plain interior fluid (warp-uniform)
#pragma unroll
    for (int c = 0; c < NS; ++c) { uA[c] = uB[c] = vA[c] = vB[c] = pA[c] = pB[c] = 0.0f; fl[c] = 0u; }

    // this lane's element offset inside a stored column, and the running offset of the column being loaded
    const size_t row_off = (size_t)(row_ok ? rA : 0);
    struct Col { float2 u, v, p; unsigned f; };
    auto load_col = [&](int i) -> Col {
        Col r;
        r.u = r.v = r.p = make_float2(0.0f, 0.0f);
        r.f = 0u;
        const int li = i - xs0;
        if (row_ok && (unsigned)li < (unsigned)ncols) {  // stored columns lie inside the grid
            const size_t o = (size_t)li * (size_t)H + row_off;
            r.u = CG ? __ldcg(reinterpret_cast<const float2 *>(P.in_u + o)) : *reinterpret_cast<const float2 *>(P.in_u + o);
            r.v = CG ? __ldcg(reinterpret_cast<const float2 *>(P.in_v + o)) : *reinterpret_cast<const float2 *>(P.in_v + o);
            if (!FIRST) r.p = CG ? __ldcg(reinterpret_cast<const float2 *>(P.in_p + o)) : *reinterpret_cast<const float2 *>(P.in_p + o);
            r.f = __ldg(reinterpret_cast<const unsigned short *>(P.flags + o));
            if (FIRST && P.apply_gravity) {
                if (r.f & TFS_F_LIVE) r.v.x = fadd(r.v.x, P.gdt);
                if ((r.f >> 8) & TFS_F_LIVE) r.v.y = fadd(r.v.y, P.gdt);
            }
        }
        return r;
    };

    // Deep prefetch without registers: the loads above are issued one march step (~200 issue slots) before their use,
    // which leaves too few bytes in flight per SM to cover the DRAM latency (ncu: long_scoreboard 3.3 cycles per issued
    // instruction).  Lanes 0-9 each own one 128-byte line of the warp's segment (0-2: u, 3-5: v, 6-8: p, 9: flags) and
    // ask the L2 for it P.pf columns ahead: one pointer increment and one prefetch.global.L2 per step.
    const int PF = P.pf;
    const char *pf_ptr = nullptr;
    size_t pf_stride = 0;
    if (PF > 0 && lane < 10) {
        const int f = lane / 3, part = lane - 3 * f;
        if (!(FIRST && f == 2)) {
            const int r0 = min(max(strip * V - NS + (f == 3 ? 0 : part * 32), 0), H - 1);
            const char *base = f == 0 ? (const char *)P.in_u : f == 1 ? (const char *)P.in_v : f == 2 ? (const char *)P.in_p : (const char *)P.flags;
            const size_t esz = f == 3 ? 1 : 4;
            pf_stride = (size_t)H * esz;
            pf_ptr = base + (size_t)r0 * esz + (size_t)(gx0 - NS - xs0) * pf_stride;  // column gx0 - NS (may lie outside: checked below)
        }
    }
    auto prefetch_col = [&](int i) {  // pf_ptr points at column i
        if (pf_ptr && (unsigned)(i - xs0) < (unsigned)ncols) asm volatile("prefetch.global.L2 [%0];" ::"l"(pf_ptr));
        pf_ptr += pf_stride;          // (a null pointer stays harmless: stride 0)
    };

    int x = gx0 - NS;          // newest column held (the "pending" one): stage s works on column x-1-s
    if (PF > 0) {
#pragma unroll 1
        for (int d = 0; d <= PF; ++d) prefetch_col(x + d);
    }
    Col pend = load_col(x);    // its u is the right face of column x-1; v, p, flags wait for the next step
    const int n_it = (P.lx + 2 * NS) / NS;
    for (int it = 0; it < n_it; ++it) {
#pragma unroll
        for (int m = 0; m < NS; ++m, ++x) {  // x % NS == m, x % 2 == m % 2
            const Col nxt = load_col(x + 1);  // consumed at the end of this step
            if (PF > 0) prefetch_col(x + 1 + PF);
            // colour of half-sweep s = s & 1; the lower cell (even row) of column i = x-1-s is active when
            // (i & 1) == (s & 1), i.e. when m is odd -- a compile-time fact of the unrolled step
            constexpr int dummy = 0; (void)dummy;
            const bool lower = (m & 1) != 0;
            {
#pragma unroll
                for (int s = 0; s < NS; ++s) {
                    if (s < ns_act) {
                        const int c = (m - 1 - s + 2 * NS) % NS;  // slot of column x-1-s
                        const int cn = (c + 1) % NS;              // slot of column x-s (s == 0: the pending column)
                        const bool is_plain = (plain >> c) & 1u;
                        if (lower) {
                            float &uR = (s == 0) ? pend.u.x : uA[cn];
                            if (is_plain) cell_update4<0>(15u, omega, pc, uA[c], uR, vA[c], vB[c], pA[c]);
                            else cell_update_lut(lut, fl[c] & 15u, omega, pc, uA[c], uR, vA[c], vB[c], pA[c]);
                        } else {
                            float &uR = (s == 0) ? pend.u.y : uB[cn];
                            float vT = __shfl_down_sync(0xffffffffu, vA[c], 1);  // lane 31: its own value, inside the overlap
                            if (is_plain) cell_update4<0>(15u, omega, pc, uB[c], uR, vB[c], vT, pB[c]);
                            else cell_update_lut(lut, (fl[c] >> 8) & 15u, omega, pc, uB[c], uR, vB[c], vT, pB[c]);
                            vT = __shfl_up_sync(0xffffffffu, vT, 1);
                            if (lane > 0) vA[c] = vT;
                        }
                    }
                }
            }
            // column x-NS (slot m) has seen all NS stages: store its exact part, then the slot takes column x
            {
                const int i = x - NS;
                if (store_lane && i >= vx0 && i < vx1) {
                    const size_t o = (size_t)(i - xs0) * (size_t)H + row_off;
                    *reinterpret_cast<float2 *>(P.out_u + o) = make_float2(uA[m], uB[m]);
                    *reinterpret_cast<float2 *>(P.out_v + o) = make_float2(vA[m], vB[m]);
                    *reinterpret_cast<float2 *>(P.out_p + o) = make_float2(pA[m], pB[m]);
                }
                uA[m] = pend.u.x; uB[m] = pend.u.y; vA[m] = pend.v.x; vB[m] = pend.v.y; pA[m] = pend.p.x; pB[m] = pend.p.y;
                fl[m] = pend.f;
                const bool pl = __all_sync(0xffffffffu, pend.f == 0x1F1Fu);
                plain = (plain & ~(1u << m)) | ((pl ? 1u : 0u) << m);
                pend = nxt;
            }
            // Keep the CTA's warps in the same unrolled step: the march loop is ~58 KB of straight-line code and the
            // SM's warps share a 32 KB instruction cache level; warps drifting apart made instruction fetch the
            // second largest stall (ncu: no_instruction 1.8 cycles per issued instruction at 16384^2).
            if (P.step_sync) __syncthreads();
        }
    }
}

// ---- host side -------------------------------------------------------------------------------------
constexpr int kRbKB = 3;  // iterations fused per pass: 6 window columns, 52 exact rows per 64-row strip

struct RbLaunchArgs {
    float *set[2][3];      // [set][u, v, p]; pass t reads set t & 1 and writes the other
    const uint8_t *flags;
    long long W, H, xs0, ncols, c_lo, c_hi;
    float omega, pc, gdt;
    bool apply_gravity;
    bool slab;             // x-slab handle: loads bypass L1
    int chunk_len;         // 0 = default
};

inline bool redblack_fused_usable(long long W, long long H, long long ncols) {
    return (H % 2) == 0 && W < (1LL << 30) && H < (1LL << 30) && ncols < (1LL << 30);
}

// one pass (iterations [kb*t, kb*t + k_act)) of the fused red-black projection
inline cudaError_t redblack_pass(const RbLaunchArgs &a, int t, int k_act, cudaStream_t stream) {
    constexpr int KB = kRbKB, NS = 2 * KB, V = 64 - 2 * NS;
    RbParams P{};
    const int in = t & 1, out = in ^ 1;
    P.in_u = a.set[in][0]; P.in_v = a.set[in][1]; P.in_p = a.set[in][2];
    P.out_u = a.set[out][0]; P.out_v = a.set[out][1]; P.out_p = a.set[out][2];
    P.flags = a.flags;
    P.W = (int)a.W; P.H = (int)a.H; P.xs0 = (int)a.xs0; P.ncols = (int)a.ncols; P.c_lo = (int)a.c_lo; P.c_hi = (int)a.c_hi;
    P.n_strips = (int)((a.H + V - 1) / V);
    // Chunk length: every warp marches lx + 2*NS columns to finalise lx of them, so long chunks waste less, but a grid
    // that yields fewer warps than the GPU holds (148 SMs x 24 at KB = 3) runs them at single-warp latency: take the
    // shortest chunk that still fits in one wave, 126 columns (the multiple of NS below 128) at most.
    int lx = 128 / NS * NS;
    if (a.chunk_len > 0) lx = (a.chunk_len + NS - 1) / NS * NS;
    else
        for (int c = 2 * NS; c < 128; c *= 2)
            if ((long long)P.n_strips * ((a.c_hi - a.c_lo + c - 1) / c) <= 148LL * 4 * (KB <= 2 ? 8 : (KB <= 3 ? 6 : 5))) { lx = c; break; }
    P.lx = lx;
    P.chunk0 = (int)(a.c_lo / P.lx);
    const int n_chunks = (int)((a.c_hi + P.lx - 1) / P.lx) - P.chunk0;
    P.ns_act = 2 * k_act;
    P.apply_gravity = (t == 0 && a.apply_gravity) ? 1 : 0;
    P.omega = a.omega; P.pc = a.pc; P.gdt = a.gdt;
    P.pf = tunables().rb_pf;
    P.step_sync = tunables().rb_sync;
    // 4 warps (adjacent strips) per CTA: their overlapping rows hit L1; 1 or 2 warps per CTA balance a one-wave launch
    // better but measured 1-5 % slower at 4096^2 and 16384^2
    const int wpc = tunables().rb_wpc == 1 ? 1 : (tunables().rb_wpc == 2 ? 2 : 4);
    dim3 grid((unsigned)((P.n_strips + wpc - 1) / wpc), (unsigned)n_chunks);
    const unsigned bs = 32u * wpc;
    const bool first = t == 0;
    if (a.slab) { if (first) k_project_rb<KB, true, true><<<grid, bs, 0, stream>>>(P); else k_project_rb<KB, false, true><<<grid, bs, 0, stream>>>(P); }
    else { if (first) k_project_rb<KB, true, false><<<grid, bs, 0, stream>>>(P); else k_project_rb<KB, false, false><<<grid, bs, 0, stream>>>(P); }
    return cudaGetLastError();
}

}  // namespace tfs
