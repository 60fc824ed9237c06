// tfs_kernels.cuh -- streaming kernels of the simulator step: cell flags, init/edit, gravity,
// the simple (one launch per wavefront) exact projection, red-black projection, semi-Lagrangian
// advection and the reductions.  The temporally blocked systolic projection lives in
// tfs_systolic.cuh.
//
// Layout (simulator.rs:7): idx = x*H + y, y contiguous.  Every kernel takes the GLOBAL grid
// size (W, H); `x0` is the global x of storage column 0 and `ncols` the number of stored
// columns (a single-GPU handle has x0 = 0, ncols = W; an x-slab stores its columns plus halos).
#pragma once
#include "tfs_common.cuh"

namespace tfs {

struct Grid {
    long long W, H;      // global size
    long long x0;        // global x of storage column 0
    long long ncols;     // stored columns
};

// ---------------------------------------------------------------------------------------------
// cell flags from the block mask (simulator.rs:161, :165-168).  Integer work, bit-exact.
// One thread per cell of the stored range [c_begin, c_end) x [0, H).
__global__ void k_cell_flags(const uint8_t *__restrict__ m, uint8_t *__restrict__ flags, Grid g,
                             long long c_begin, long long c_end) {
    const long long H = g.H;
    long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    long long n = (c_end - c_begin) * H;
    if (t >= n) return;
    long long lc = c_begin + t / H, y = t % H;
    long long x = g.x0 + lc;
    long long idx = lc * H + y;
    unsigned f = 0;
    if (!m[idx] && !is_border(x, y, g.W, H)) {
        // an updatable cell is never on the border, so all four neighbours exist globally;
        // in a slab the neighbour columns must be stored (halo >= 1), else treated as fluid=0
        f = TFS_F_LIVE;
        if (lc - 1 >= 0 && !m[idx - H]) f |= TFS_F_L;
        if (lc + 1 < g.ncols && !m[idx + H]) f |= TFS_F_R;
        if (!m[idx - 1]) f |= TFS_F_B;
        if (!m[idx + 1]) f |= TFS_F_T;
    }
    flags[idx] = (uint8_t)f;
}

// ---------------------------------------------------------------------------------------------
// init / edit (simulator.rs:73-119, :67-71)
__global__ void k_fill_f32(float *__restrict__ a, long long n, float x) {
    long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    long long stride = (long long)gridDim.x * blockDim.x;
    for (; t < n; t += stride) a[t] = x;
}
__global__ void k_fill_u8(uint8_t *__restrict__ a, long long n, uint8_t x) {
    long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    long long stride = (long long)gridDim.x * blockDim.x;
    for (; t < n; t += stride) a[t] = x;
}
// a[idx] = x for GLOBAL flat indices in [lo, hi) that fall inside the stored column range
__global__ void k_fill_range_f32(float *__restrict__ a, Grid g, long long lo, long long hi, float x) {
    long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    long long gidx = lo + t;
    if (gidx >= hi) return;
    long long lidx = gidx - g.x0 * g.H;
    if (lidx < 0 || lidx >= g.ncols * g.H) return;
    a[lidx] = x;
}
__global__ void k_fill_range_u8(uint8_t *__restrict__ a, Grid g, long long lo, long long hi, uint8_t x) {
    long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    long long gidx = lo + t;
    if (gidx >= hi) return;
    long long lidx = gidx - g.x0 * g.H;
    if (lidx < 0 || lidx >= g.ncols * g.H) return;
    a[lidx] = x;
}

// ---------------------------------------------------------------------------------------------
// add_gravity (simulator.rs:137-149): v[idx] = v[idx] + (gravity*dt) on updatable cells.
__global__ void k_gravity(float *__restrict__ v, const uint8_t *__restrict__ flags, long long n, float gdt) {
    long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    long long stride = (long long)gridDim.x * blockDim.x;
    for (; t < n; t += stride)
        if (flags[t] & TFS_F_LIVE) v[t] = fadd(v[t], gdt);
}

// ---------------------------------------------------------------------------------------------
// make_incompressible, exact order, simple form: one launch per wavefront t = i + j + 2k
// (SURVEY 8a-3).  blockIdx.y = k, threads along the anti-diagonal d = t - 2k.  Every member of
// a wavefront touches disjoint locations, so no synchronisation is needed inside a launch.
__global__ void k_project_wavefront(float *__restrict__ u, float *__restrict__ v, float *__restrict__ p,
                                    const uint8_t *__restrict__ flags, Grid g, long long t, int k_lo,
                                    float omega, float pc) {
    const long long W = g.W, H = g.H;
    const int k = k_lo + blockIdx.y;
    const long long d = t - 2LL * k;
    if (d < 2 || d > W + H - 4) return;  // interior cells have 1 <= i <= W-2, 1 <= j <= H-2
    long long i_lo = d - (H - 2);
    if (i_lo < 1) i_lo = 1;
    long long i_hi = d - 1;
    if (i_hi > W - 2) i_hi = W - 2;
    long long i = i_lo + blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i > i_hi) return;
    long long j = d - i;
    long long idx = i * H + j;
    unsigned f = flags[idx];
    if (!(f & TFS_F_LIVE)) return;
    float uC = u[idx], uR = u[idx + H], vC = v[idx], vT = v[idx + 1], pp = p[idx];
    cell_update(f, omega, pc, uC, uR, vC, vT, pp);
    u[idx] = uC;
    u[idx + H] = uR;
    v[idx] = vC;
    v[idx + 1] = vT;
    p[idx] = pp;
}

// ---------------------------------------------------------------------------------------------
// make_incompressible, red-black order (TFS_MODE_RED_BLACK): one launch per colour.  Cells of
// one colour ((i+j)&1) touch disjoint faces.  Thread = (column i, pair index jj), j = 2*jj + par.
__global__ void k_project_redblack(float *__restrict__ u, float *__restrict__ v, float *__restrict__ p,
                                   const uint8_t *__restrict__ flags, Grid g, int colour, float omega,
                                   float pc) {
    const long long H = g.H;
    const long long half = (H + 1) >> 1;
    long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t >= g.ncols * half) return;
    long long lc = t / half, jj = t % half;
    long long i = g.x0 + lc;
    long long j = 2 * jj + ((i + colour) & 1);
    if (j >= H) return;
    long long idx = lc * H + j;
    unsigned f = flags[idx];
    if (!(f & TFS_F_LIVE)) return;
    float uC = u[idx], uR = u[idx + H], vC = v[idx], vT = v[idx + 1], pp = p[idx];
    cell_update(f, omega, pc, uC, uR, vC, vT, pp);
    u[idx] = uC;
    u[idx + H] = uR;
    v[idx] = vC;
    v[idx + 1] = vT;
    p[idx] = pp;
}

// ---------------------------------------------------------------------------------------------
// move_velocity (simulator.rs:192-298): semi-Lagrangian advection of u, v and smoke.
// Texture-free bilinear sampling; coordinates are GLOBAL f32.

enum { FT_U = 0, FT_V = 1, FT_S = 2 };

// sample_vector (simulator.rs:269-298).  lx0/ncols translate global columns to storage.
// NC: the field is global memory read through the non-coherent path (__ldg); false = plain loads (the field may then
// live in shared memory: k_step_small)
template <int FT, bool NC = true>
TFS_DEVINL float sample_field(const float *__restrict__ f, float x, float y, const Grid &g, float Wf,
                              float Hf) {
    const float dx = (FT == FT_U) ? 0.0f : 0.5f;   // :277-281
    const float dy = (FT == FT_V) ? 0.0f : 0.5f;
    x = fmaxf(fminf(x, Wf), 1.0f);                 // :271
    y = fmaxf(fminf(y, Hf), 1.0f);                 // :272
    const float xs = fsub(x, dx), ys = fsub(y, dy);
    // xs, ys >= 0.5 here, so floor == trunc and the saturating `as usize` is a plain convert
    long long xl = (long long)__float2uint_rz(xs);   // :283
    if (xl > g.W - 1) xl = g.W - 1;
    long long xr = xl + 1;                           // :284
    if (xr > g.W - 1) xr = g.W - 1;
    const float tx = fsub(xs, (float)xl);            // :285
    long long yb = (long long)__float2uint_rz(ys);   // :287
    if (yb > g.H - 1) yb = g.H - 1;
    long long yt = yb + 1;                           // :288
    if (yt > g.H - 1) yt = g.H - 1;
    const float ty = fsub(ys, (float)yb);            // :289
    const float sx = fsub(1.0f, tx), sy = fsub(1.0f, ty);  // :291-292
    // storage columns (clamped for memory safety; halo sufficiency is checked by the host)
    long long cl = xl - g.x0, cr = xr - g.x0;
    cl = cl < 0 ? 0 : (cl > g.ncols - 1 ? g.ncols - 1 : cl);
    cr = cr < 0 ? 0 : (cr > g.ncols - 1 ? g.ncols - 1 : cr);
    const float f00 = NC ? __ldg(f + cl * g.H + yb) : f[cl * g.H + yb], f10 = NC ? __ldg(f + cr * g.H + yb) : f[cr * g.H + yb];
    const float f11 = NC ? __ldg(f + cr * g.H + yt) : f[cr * g.H + yt], f01 = NC ? __ldg(f + cl * g.H + yt) : f[cl * g.H + yt];
    // :294-297, left to right
    return fadd(fadd(fadd(fmul(fmul(sx, sy), f00), fmul(fmul(tx, sy), f10)), fmul(fmul(tx, ty), f11)),
                fmul(fmul(sx, ty), f01));
}

// one cell of move_velocity; (lc, j) = storage column / row of an UPDATABLE cell
template <bool NC = true>
TFS_DEVINL void advect_cell(const float *__restrict__ u, const float *__restrict__ v,
                            const float *__restrict__ s, const Grid &g, long long lc, long long j,
                            float dt, float Wf, float Hf, float uC, float vC, float &nu, float &nv,
                            float &ns) {
    const long long H = g.H;
    const long long idx = lc * H + j;
    const float fi = (float)(g.x0 + lc), fj = (float)j;
    // u component (:207-213), avg_vertical (:245-255): ((((0+a)+b)+c)+d)*0.25
    const float vL = NC ? __ldg(v + idx - H) : v[idx - H], vLT = NC ? __ldg(v + idx - H + 1) : v[idx - H + 1], vT = NC ? __ldg(v + idx + 1) : v[idx + 1];
    const float avg_v = fmul(fadd(fadd(fadd(fadd(0.0f, vL), vLT), vT), vC), 0.25f);
    float xp = fsub(fi, fmul(uC, dt));
    float yp = fsub(fadd(fj, 0.5f), fmul(avg_v, dt));
    nu = sample_field<FT_U, NC>(u, xp, yp, g, Wf, Hf);
    // v component (:216-224), avg_horizontal (:257-267)
    const float uB = NC ? __ldg(u + idx - 1) : u[idx - 1], uRB = NC ? __ldg(u + idx + H - 1) : u[idx + H - 1], uR = NC ? __ldg(u + idx + H) : u[idx + H];
    const float avg_u = fmul(fadd(fadd(fadd(fadd(0.0f, uB), uRB), uR), uC), 0.25f);
    xp = fsub(fadd(fi, 0.5f), fmul(avg_u, dt));
    yp = fsub(fj, fmul(vC, dt));
    nv = sample_field<FT_V, NC>(v, xp, yp, g, Wf, Hf);
    // smoke (:227-237); the guard at :227 is always true for an updatable cell
    const float cv = fmul(fadd(vC, vT), 0.5f);
    const float cu = fmul(fadd(uC, uR), 0.5f);
    xp = fsub(fadd(fi, 0.5f), fmul(cu, dt));
    yp = fsub(fadd(fj, 0.5f), fmul(cv, dt));
    ns = sample_field<FT_S, NC>(s, xp, yp, g, Wf, Hf);
}

// VEC = 4: one thread per 4 consecutive rows of a column, 128-bit loads/stores (H % 4 == 0);
// VEC = 1: scalar fallback for any H.  Writes EVERY cell of columns [c_begin, c_end) of the new
// buffers (non-updatable cells copy the old value: the reference clones the Vecs, :193-195).
template <int VEC>
__global__ void __launch_bounds__(256)
k_advect(const float *__restrict__ u, const float *__restrict__ v, const float *__restrict__ s,
         const uint8_t *__restrict__ flags, float *__restrict__ nu, float *__restrict__ nv,
         float *__restrict__ ns, Grid g, long long c_begin, long long c_end, float dt) {
    const long long H = g.H;
    const long long per_col = H / VEC;
    long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t >= (c_end - c_begin) * per_col) return;
    const long long lc = c_begin + t / per_col;
    const long long j0 = (t % per_col) * VEC;
    const long long idx0 = lc * H + j0;
    const float Wf = (float)g.W, Hf = (float)g.H;
    if (VEC == 4) {
        const float4 u4 = *reinterpret_cast<const float4 *>(u + idx0);
        const float4 v4 = *reinterpret_cast<const float4 *>(v + idx0);
        const float4 s4 = *reinterpret_cast<const float4 *>(s + idx0);
        const uchar4 f4 = *reinterpret_cast<const uchar4 *>(flags + idx0);
        float uo[4] = {u4.x, u4.y, u4.z, u4.w}, vo[4] = {v4.x, v4.y, v4.z, v4.w};
        float so[4] = {s4.x, s4.y, s4.z, s4.w};
        const unsigned ff[4] = {f4.x, f4.y, f4.z, f4.w};
        float un[4], vn[4], sn[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            un[r] = uo[r]; vn[r] = vo[r]; sn[r] = so[r];
            if (ff[r] & TFS_F_LIVE)
                advect_cell(u, v, s, g, lc, j0 + r, dt, Wf, Hf, uo[r], vo[r], un[r], vn[r], sn[r]);
        }
        *reinterpret_cast<float4 *>(nu + idx0) = make_float4(un[0], un[1], un[2], un[3]);
        *reinterpret_cast<float4 *>(nv + idx0) = make_float4(vn[0], vn[1], vn[2], vn[3]);
        *reinterpret_cast<float4 *>(ns + idx0) = make_float4(sn[0], sn[1], sn[2], sn[3]);
    } else {
        const float uo = u[idx0], vo = v[idx0], so = s[idx0];
        float un = uo, vn = vo, sn = so;
        if (flags[idx0] & TFS_F_LIVE) advect_cell(u, v, s, g, lc, j0, dt, Wf, Hf, uo, vo, un, vn, sn);
        nu[idx0] = un; nv[idx0] = vn; ns[idx0] = sn;
    }
}

// ---------------------------------------------------------------------------------------------
// move_velocity, fast variant for single-GPU handles whose grid fits 32-bit indexing (N < 2^31):
// one thread per cell, lanes along y, so every bilinear gather of a warp lands in one or two
// 128-byte lines; all index arithmetic is 32-bit.  Same expression order as advect_cell above.
// a + b and a - b issued as FFMA with a RUNTIME multiplier `one` (== 1.0f, a kernel argument, so
// ptxas cannot fold it back into FADD): a*1 is exact, hence fma(a, 1, b) == RN(a + b) bit for bit.
// On sm_100 FADD shares the ALU pipe with the integer / min-max / compare work that dominates this
// kernel (ncu: ALU 60 % vs FMA 30 % busy); moving most adds to the FMA pipe balances the two.
TFS_DEVINL float fadd_f(float a, float b, float one) { return __fmaf_rn(a, one, b); }
TFS_DEVINL float fsub_f(float a, float b, float one) { return __fmaf_rn(b, -one, a); }

// floor of 0 <= x < 2^22 without the conversion pipe: F2I / I2F issue at 1/8 and 1/4 of the FP32 rate
// on sm_100 (scripts/micro/cvt_rate.cu: 8 and 4 cycles per warp instruction), and a bilinear sample needs
// two of each.  RZ(x + 2^23) = 2^23 + floor(x) exactly (floats in [2^23, 2^24) are the integers), so the
// integer is the low mantissa bits and the float floor is one exact subtraction.  Same values as
// `x as usize` / `xl as f32` (simulator.rs:283-289) on that domain.
#define TFS_MAGIC 8388608.0f
TFS_DEVINL float floor_magic(float x) { return __fadd_rz(x, TFS_MAGIC); }                  // 2^23 + floor(x)
TFS_DEVINL unsigned magic_uint(float y) { return __float_as_uint(y) - 0x4B000000u; }       // floor(x) as integer
TFS_DEVINL float magic_float(float y, float one) { return __fmaf_rn(y, one, -TFS_MAGIC); }  // floor(x) as f32 (exact)

// SLAB: `f` is a global-column base (storage pointer - xs0*H); columns outside the stored range
// [cs0, cs1] are clamped for memory safety and reported through *oob (halo too small for this dt).
// field load: x-slab kernels bypass L1 (ld.global.cg).  Their halo columns are rewritten between launches by another
// stream (the neighbour's push); with two slabs on ONE device an L1 line from an earlier launch could be stale.
// (Single GPU: a plain load of a const __restrict__ pointer - the compiler picks the read-only path itself; forcing it
// with __ldg cost 8 instructions and 2 % in the packed kernel.)
template <bool SLAB>
TFS_DEVINL float ldf(const float *p) { return SLAB ? __ldcg(p) : *p; }

// What an x-slab's advection can reach beyond its own stored columns [cs0, cs1]: the neighbours' memory, read
// directly over NVLink (the back-trace distance |vel|*dt is unbounded, simulator.rs:211-212, so a halo of
// fixed width is only the fast path).  L / R are the peers' pre-advection u, v, s rebased to GLOBAL columns; [l0, l1] and
// [r0, r1] are the columns that are safe to read there (tfs_api.cu, step_slab); anything further sets *oob.
struct SlabFar {
    const float *L[3] = {nullptr, nullptr, nullptr}, *R[3] = {nullptr, nullptr, nullptr};
    unsigned cs0 = 0u, cs1 = 0u, l0 = 1u, l1 = 0u, r0 = 1u, r1 = 0u;
    int *oob = nullptr;
};
// one corner of a bilinear sample whose column may lie outside the stored range (rare: out of line)
__device__ __noinline__ float far_load(const SlabFar &F, int ft, const float *__restrict__ local, unsigned col, unsigned row, unsigned H) {
    if (col >= F.cs0 && col <= F.cs1) return __ldcg(local + (col * H + row));
    if (F.L[ft] && col >= F.l0 && col <= F.l1) return __ldcg(F.L[ft] + (col * H + row));
    if (F.R[ft] && col >= F.r0 && col <= F.r1) return __ldcg(F.R[ft] + (col * H + row));
    *F.oob = 1;  // benign race: every writer stores 1
    return __ldcg(local + (min(max(col, F.cs0), F.cs1) * H + row));
}

template <int FT, bool SLAB = false>
TFS_DEVINL float sample_field32(const float *__restrict__ f, float x, float y, int W, int H, float Wf, float Hf,
                                float one, const SlabFar *far = nullptr) {
    const float dx = (FT == FT_U) ? 0.0f : 0.5f;   // simulator.rs:277-281
    const float dy = (FT == FT_V) ? 0.0f : 0.5f;
    x = fmaxf(fminf(x, Wf), 1.0f);                 // :271
    y = fmaxf(fminf(y, Hf), 1.0f);                 // :272
    const float xs = (FT == FT_U) ? x : fsub_f(x, dx, one), ys = (FT == FT_V) ? y : fsub_f(y, dy, one);  // x - 0.0 == x
    // xs, ys in [0.5, 2^22): floor by the magic constant; the min(.., W-1) / min(.., H-1) clamps of :283 / :287
    // are applied to the float image first (order-preserving), so index and fraction see the same value
    const float xm = fminf(floor_magic(xs), TFS_MAGIC - 1.0f + Wf), ym = fminf(floor_magic(ys), TFS_MAGIC - 1.0f + Hf);   // 2^23 + (W-1), exact
    const unsigned xl = magic_uint(xm), yb = magic_uint(ym);           // :283, :287
    const float tx = fsub_f(xs, magic_float(xm, one), one);            // :285
    const float ty = fsub_f(ys, magic_float(ym, one), one);            // :289
    const float sx = fsub_f(1.0f, tx, one), sy = fsub_f(1.0f, ty, one);  // :291-292
    // xr = min(xl+1, W-1), yt = min(yb+1, H-1) (:284, :288) as index offsets: +H / +1 unless clamped
    const unsigned dxo = (xl < (unsigned)(W - 1)) ? (unsigned)H : 0u;
    const unsigned dyo = (yb < (unsigned)(H - 1)) ? 1u : 0u;
    float f00, f10, f11, f01;
    const unsigned xr = xl + (dxo ? 1u : 0u);
    if (SLAB && (xl < far->cs0 || xr > far->cs1)) {  // beyond the stored columns: the neighbours' memory
        f00 = far_load(*far, FT, f, xl, yb, (unsigned)H); f10 = far_load(*far, FT, f, xr, yb, (unsigned)H);
        f11 = far_load(*far, FT, f, xr, yb + dyo, (unsigned)H); f01 = far_load(*far, FT, f, xl, yb + dyo, (unsigned)H);
    } else {
        const unsigned i00 = xl * (unsigned)H + yb, i10 = i00 + dxo;
        f00 = ldf<SLAB>(f + i00); f10 = ldf<SLAB>(f + i10);
        f11 = ldf<SLAB>(f + i10 + dyo); f01 = ldf<SLAB>(f + i00 + dyo);
    }
    return fadd_f(fadd_f(fadd(fmul(fmul(sx, sy), f00), fmul(fmul(tx, sy), f10)), fmul(fmul(tx, ty), f11), one),
                  fmul(fmul(sx, ty), f01), one);   // :294-297
}

// sample_vector for a point with 1 <= x < W-1 and 1 <= y < H-1: none of the clamps of simulator.rs:271-288
// can bind (x, y in range; xl + 1 <= W-1; yb + 1 <= H-1), so they are not issued.  Bit-identical to
// sample_field32 on that domain; the caller checks the domain with a warp vote.
template <int FT>
TFS_DEVINL float sample_interior32(const float *__restrict__ f, float x, float y, int H, float one) {
    const float xs = (FT == FT_U) ? x : fsub_f(x, 0.5f, one), ys = (FT == FT_V) ? y : fsub_f(y, 0.5f, one);
    const float xm = floor_magic(xs), ym = floor_magic(ys);
    const unsigned xl = magic_uint(xm), yb = magic_uint(ym);
    const float tx = fsub_f(xs, magic_float(xm, one), one), ty = fsub_f(ys, magic_float(ym, one), one);
    const float sx = fsub_f(1.0f, tx, one), sy = fsub_f(1.0f, ty, one);
    const float *p0 = f + (xl * (unsigned)H + yb);  // (xl, yb); the other three corners are +H, +H+1, +1
    const float *p1 = p0 + H;
    const float f00 = __ldg(p0), f10 = __ldg(p1), f11 = __ldg(p1 + 1), f01 = __ldg(p0 + 1);
    return fadd_f(fadd_f(fadd(fmul(fmul(sx, sy), f00), fmul(fmul(tx, sy), f10)), fmul(fmul(tx, ty), f11), one),
                  fmul(fmul(sx, ty), f01), one);   // :294-297
}

// Ask the L2 for the block's whole tile up front: rows [r0, r0 + rows) of the columns [i0, i0 + ncol]
// of u and v, [i0, i0 + ncol) of s and flags, one bulk prefetch (cp.async.bulk.prefetch.L2, a TMA
// operation without destination) per column and array, issued by the first 4*ncol + 2 threads.  The
// march below otherwise exposes two dependent DRAM latencies per column with far too few bytes in
// flight per SM to saturate HBM; with the tile already on its way the loads
// hit the L2 instead.  Needs H % 4 == 0 and r0 % 4 == 0 (16-byte aligned column segments).
TFS_DEVINL void prefetch_tile_l2(const float *u, const float *v, const float *s, const uint8_t *flags, int W, int H, int i0,
                                 int ncol, int r0, int rows) {
    const int t = threadIdx.x, nuv = ncol + 1;
    const void *src = nullptr;
    unsigned bytes = (unsigned)rows * 4u;
    if (t < 2 * nuv) {
        const int c = i0 + (t >> 1);
        if (c < W) src = ((t & 1) ? v : u) + ((long long)c * H + r0);
    } else if (t < 2 * nuv + 2 * ncol) {
        const int q = t - 2 * nuv, c = i0 + (q >> 1);
        if (c < W) {
            if (q & 1) { src = flags + ((long long)c * H + r0); bytes = (unsigned)rows; }
            else src = s + ((long long)c * H + r0);
        }
    }
    if (src) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}

// blockDim.x threads along y (rows y0 .. y0+blockDim.x), CPB columns per block marched in +x so the
// column i+1 values loaded for cell (i, j) are reused as "own" values of cell (i+1, j).
// SLAB: all pointers are global-column bases, only columns [c_begin, c_end) are advected, the
// stored columns are [cs0, cs1].
// L2PF: see prefetch_tile_l2.
template <int CPB, bool SLAB = false, bool L2PF = false>
__global__ void __launch_bounds__(256)
k_advect32(const float *__restrict__ u, const float *__restrict__ v, const float *__restrict__ s,
           const uint8_t *__restrict__ flags, float *__restrict__ nu, float *__restrict__ nv,
           float *__restrict__ ns, int W, int H, float dt, float one, float4 dims, int c_begin = 0, int c_end = 0x7fffffff,
           const __grid_constant__ SlabFar far = SlabFar()) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i0 = c_begin + blockIdx.y * CPB;
    if (L2PF) prefetch_tile_l2(u, v, s, flags, W, H, i0, CPB, blockIdx.x * blockDim.x, min((int)blockDim.x, H - (int)(blockIdx.x * blockDim.x)));
    if (j >= H) return;
    // dims = {W, H, W-1, H-1} as floats, converted on the host: ptxas otherwise re-converts them every column,
    // and I2F issues at a quarter of the FP32 rate (scripts/micro/cvt_rate.cu)
    const float Wf = dims.x, Hf = dims.y, Wm1 = dims.z, Hm1 = dims.w;
    const float fj = (float)j, fjh = fadd(fj, 0.5f);
    float fi = (float)i0;
    const int jm = j > 0 ? j - 1 : 0, jp = j < H - 1 ? j + 1 : j;   // clamped only for memory safety (border rows never use them)
    unsigned idx = (unsigned)i0 * (unsigned)H + (unsigned)j;
    // values of column i (own) and column i-1 (left)
    float uC = ldf<SLAB>(u + idx), vC = ldf<SLAB>(v + idx), uB = ldf<SLAB>(u + (idx - j + jm)), vT = ldf<SLAB>(v + (idx - j + jp));
    float vL = 0.0f, vLT = 0.0f;
    if (i0 > 0) { vL = ldf<SLAB>(v + (idx - H)); vLT = ldf<SLAB>(v + (idx - H - j + jp)); }
#pragma unroll
    for (int c = 0; c < CPB; ++c) {
        const int i = i0 + c;
        if (i >= W || i >= c_end) break;
        // column i+1 (right); clamped at the right border for memory safety only
        const unsigned idr = (i + 1 < W) ? idx + (unsigned)H : idx;
        const float uR = ldf<SLAB>(u + idr), uRB = ldf<SLAB>(u + (idr - j + jm)), vR = ldf<SLAB>(v + idr), vRT = ldf<SLAB>(v + (idr - j + jp));
        const float sC = ldf<SLAB>(s + idx);
        const unsigned f = flags[idx];
        float un = uC, vn = vC, sn = sC;
        const bool live = (f & TFS_F_LIVE) != 0u;
        const float fih = fadd(fi, 0.5f);
        // back-traced positions of the three samples (:207-236).  (The leading "0 +" of Iterator::sum only
        // affects the sign of a zero sum, which is then scaled and subtracted from a coordinate >= 1: it
        // cannot reach the result, so it is not issued.)
        const float avg_v = fmul(fadd_f(fadd(fadd(vL, vLT), vT), vC, one), 0.25f);    // :245-255
        const float avg_u = fmul(fadd_f(fadd(fadd(uB, uRB), uR), uC, one), 0.25f);    // :257-267
        const float cv = fmul(fadd(vC, vT), 0.5f), cu = fmul(fadd(uC, uR), 0.5f);     // :228-233
        const float xu = fsub_f(fi, fmul(uC, dt), one), yu = fsub_f(fjh, fmul(avg_v, dt), one);
        const float xv = fsub_f(fih, fmul(avg_u, dt), one), yv = fsub_f(fj, fmul(vC, dt), one);
        const float xsm = fsub_f(fih, fmul(cu, dt), one), ysm = fsub_f(fjh, fmul(cv, dt), one);
        // all three points strictly inside [1, W-1) x [1, H-1)?  (NaNs fail the test and take the general path)
        const float xmin = fminf(fminf(xu, xv), xsm), xmax = fmaxf(fmaxf(xu, xv), xsm);
        const float ymin = fminf(fminf(yu, yv), ysm), ymax = fmaxf(fmaxf(yu, yv), ysm);
        const bool inside = !SLAB && xmin >= 1.0f && xmax < Wm1 && ymin >= 1.0f && ymax < Hm1 &&
                            xu == xu && xv == xv && xsm == xsm && yu == yu && yv == yv && ysm == ysm;
        if (__all_sync(__activemask(), inside || !live)) {
            if (live) {
                un = sample_interior32<FT_U>(u, xu, yu, H, one);
                vn = sample_interior32<FT_V>(v, xv, yv, H, one);
                sn = sample_interior32<FT_S>(s, xsm, ysm, H, one);
            }
        } else if (live) {
            un = sample_field32<FT_U, SLAB>(u, xu, yu, W, H, Wf, Hf, one, &far);
            vn = sample_field32<FT_V, SLAB>(v, xv, yv, W, H, Wf, Hf, one, &far);
            sn = sample_field32<FT_S, SLAB>(s, xsm, ysm, W, H, Wf, Hf, one, &far);
        }
        nu[idx] = un; nv[idx] = vn; ns[idx] = sn;
        // march: column i+1 becomes the own column, column i the left one
        vL = vC; vLT = vT;
        uC = uR; uB = uRB; vC = vR; vT = vRT;
        idx = idr;
        fi = fadd_f(fi, 1.0f, one);   // == (float)(i + 1): integers below 2^24 are exact
    }
}

// ---------------------------------------------------------------------------------------------
// move_velocity, packed variant: one thread advects TWO cells of a column (rows j and j + 128) and
// carries every f32 quantity of the pair in one 64-bit register pair, so each arithmetic step of
// simulator.rs:207-297 is ONE Blackwell packed instruction (FFMA2 via fma.rn.f32x2, two independent
// IEEE round-to-nearest FMAs) for both cells.  k_advect32 is issue-bound (ncu: ~190 warp instructions
// per cell, issue slots 72 % busy); packing halves the floating-point share of that.
// Bit-exactness: every operation is issued as an FMA that is EXACT in its extra operand:
//   a + b = fma(a, 1, b)      a - b = fma(b, -1, a)      a * b = fma(a, b, -0.0)
// (a*1 and b*-1 are exact; x + (-0.0) == x for every x including both zeros).  The constants 1, -1
// and -0.0 arrive as kernel arguments so ptxas cannot turn an FMA back into mul/add: it contracts even
// explicitly rounded mul.rn.f32x2 + add.rn.f32x2 pairs into one FFMA2 (observed, ptxas 12.9, --fmad
// false), which would fuse a*b+c and break parity with the reference.
typedef unsigned long long f2;
TFS_DEVINL f2 f2_pack(float lo, float hi) { f2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
TFS_DEVINL float f2_lo(f2 a) { float lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a)); return lo; }
TFS_DEVINL float f2_hi(f2 a) { float lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a)); return hi; }
TFS_DEVINL f2 f2_fma(f2 a, f2 b, f2 c) { f2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
TFS_DEVINL f2 f2_add_rz(f2 a, f2 b) { f2 r; asm("add.rz.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
struct F2K { f2 one, mone, nzero; };  // (1, 1), (-1, -1), (-0.0, -0.0): runtime values, see above
TFS_DEVINL f2 add2(f2 a, f2 b, const F2K &k) { return f2_fma(a, k.one, b); }
TFS_DEVINL f2 sub2(f2 a, f2 b, const F2K &k) { return f2_fma(b, k.mone, a); }
TFS_DEVINL f2 mul2(f2 a, f2 b, const F2K &k) { return f2_fma(a, b, k.nzero); }

// sample_vector on the clamp-free interior domain (see sample_interior32) for the cell pair.  Both
// points must be inside [1, W-1) x [1, H-1) (also for a cell that is not live, whose result the
// caller discards), so all eight loads are unconditional.
template <int FT, bool SLAB = false>
TFS_DEVINL f2 sample_interior_pair(const float *__restrict__ f, f2 x, f2 y, int H, const F2K &k) {
    const f2 mhalf = f2_pack(-0.5f, -0.5f);
    const f2 xs = (FT == FT_U) ? x : f2_fma(x, k.one, mhalf), ys = (FT == FT_V) ? y : f2_fma(y, k.one, mhalf);  // :277-281
    const f2 magic = f2_pack(TFS_MAGIC, TFS_MAGIC), mmagic = f2_pack(-TFS_MAGIC, -TFS_MAGIC);
    const f2 xm = f2_add_rz(xs, magic), ym = f2_add_rz(ys, magic);     // 2^23 + floor, see floor_magic
    const unsigned xlA = (unsigned)xm - 0x4B000000u, xlB = (unsigned)(xm >> 32) - 0x4B000000u;                // :283
    const unsigned ybA = (unsigned)ym - 0x4B000000u, ybB = (unsigned)(ym >> 32) - 0x4B000000u;                // :287
    const f2 tx = sub2(xs, f2_fma(xm, k.one, mmagic), k), ty = sub2(ys, f2_fma(ym, k.one, mmagic), k);        // :285, :289
    const f2 sx = f2_fma(tx, k.mone, k.one), sy = f2_fma(ty, k.mone, k.one);                                 // :291-292
    const float *a0 = f + (xlA * (unsigned)H + ybA), *a1 = a0 + H;   // (xl, yb); the other corners are +H, +H+1, +1
    const float *b0 = f + (xlB * (unsigned)H + ybB), *b1 = b0 + H;
    const f2 f00 = f2_pack(ldf<SLAB>(a0), ldf<SLAB>(b0)), f10 = f2_pack(ldf<SLAB>(a1), ldf<SLAB>(b1));
    const f2 f11 = f2_pack(ldf<SLAB>(a1 + 1), ldf<SLAB>(b1 + 1)), f01 = f2_pack(ldf<SLAB>(a0 + 1), ldf<SLAB>(b0 + 1));
    const f2 t0 = mul2(mul2(sx, sy, k), f00, k), t1 = mul2(mul2(tx, sy, k), f10, k);
    const f2 t2 = mul2(mul2(tx, ty, k), f11, k), t3 = mul2(mul2(sx, ty, k), f01, k);
    return add2(add2(add2(t0, t1, k), t2, k), t3, k);   // :294-297
}

// all six floats in [lo, hi), tested on their bit patterns: for 0 < lo <= x the f32 order equals the
// unsigned order of the bits; negatives (sign bit set) and NaNs (above +inf) exceed every hi.
TFS_DEVINL bool in_range6(f2 a, f2 b, f2 c, unsigned lo_bits, unsigned hi_bits) {
    const unsigned a0 = (unsigned)a, a1 = (unsigned)(a >> 32), b0 = (unsigned)b, b1 = (unsigned)(b >> 32);
    const unsigned c0 = (unsigned)c, c1 = (unsigned)(c >> 32);
    const unsigned mn = min(min(min(a0, b0), c0), min(min(a1, b1), c1));
    const unsigned mx = max(max(max(a0, b0), c0), max(max(a1, b1), c1));
    return (mn >= lo_bits) & (mx < hi_bits);
}

// the rare general path of one cell (clamps of simulator.rs:271-288 may bind), out of line to keep the
// unrolled main loop small
struct Uvs { float u, v, s; };
template <bool SLAB>
__device__ __noinline__ Uvs advect_general32(const float *__restrict__ u, const float *__restrict__ v,
                                             const float *__restrict__ s, int W, int H, float one, float xu, float yu,
                                             float xv, float yv, float xsm, float ysm, const SlabFar *far) {
    const float Wf = (float)W, Hf = (float)H;
    Uvs r;
    r.u = sample_field32<FT_U, SLAB>(u, xu, yu, W, H, Wf, Hf, one, far);
    r.v = sample_field32<FT_V, SLAB>(v, xv, yv, W, H, Wf, Hf, one, far);
    r.s = sample_field32<FT_S, SLAB>(s, xsm, ysm, W, H, Wf, Hf, one, far);
    return r;
}

// 128 threads: thread t of block (bx, by) owns rows jA = 256*bx + t and jB = jA + 128 of the columns
// [CPB*by, CPB*by + CPB), marched in +x like k_advect32.  Needs H % 256 == 0 and W*H < 2^31.
// A warp takes the clamp-free path when all 6 x 64 back-traced coordinates of its cells lie inside
// [1, W-1) x [1, H-1) - cells that are not live included, so the warps holding the rows 0 / H-1 and
// the columns 0 / W-1 always take the general path (about 3 % of the warps at 4096^2).
// SLAB: the pointers are global-column bases (storage pointer - xs0*H), the columns [c_begin, c_end) are advected and
// the clamp-free path additionally requires every sample column to be stored (x in [cs0 + 1, cs1)).
template <int CPB, int MINB, bool L2PF, bool SLAB = false>
__global__ void __launch_bounds__(128, MINB)
k_advect32_pair(const float *__restrict__ u, const float *__restrict__ v, const float *__restrict__ s,
                const uint8_t *__restrict__ flags, float *__restrict__ nu, float *__restrict__ nv,
                float *__restrict__ ns, int W, int H, f2 dt2, float one, F2K k, int c_begin = 0, int c_end = 0x7fffffff,
                const __grid_constant__ SlabFar far = SlabFar()) {
    constexpr int RB = 128;  // row distance of the two cells
    const int jA = blockIdx.x * (2 * RB) + threadIdx.x, jB = jA + RB;
    const int i0 = SLAB ? c_begin + blockIdx.y * CPB : blockIdx.y * CPB;
    const int c_stop = SLAB ? min(W, c_end) : W;
    if (L2PF) prefetch_tile_l2(u, v, s, flags, SLAB ? min(W, (int)far.cs1 + 1) : W, H, i0, SLAB ? min(CPB, c_stop - i0) : CPB, blockIdx.x * (2 * RB), 2 * RB);
    // x of every sample in [xlo, xhi): no clamp of simulator.rs:271-288 binds and (SLAB) both sample columns are stored
    const unsigned xlo_bits = SLAB ? __float_as_uint((float)max(1, (int)far.cs0 + 1)) : 0x3f800000u;
    const unsigned xhi_bits = __float_as_uint((float)(SLAB ? min(W - 1, (int)far.cs1) : W - 1));
    const unsigned one_bits = 0x3f800000u, hm1_bits = __float_as_uint((float)(H - 1));
    const f2 half2 = f2_pack(0.5f, 0.5f), quarter2 = f2_pack(0.25f, 0.25f);
    const f2 fj = f2_pack((float)jA, (float)jB), fjh = add2(fj, half2, k);
    // row offsets of the (j-1) and (j+1) neighbours relative to row jA; clamped only for memory safety
    const int dAm = jA > 0 ? -1 : 0, dBp = RB + (jB < H - 1 ? 1 : 0);
    long long idx = (long long)i0 * H + jA;
    const float *pu = u + idx, *pv = v + idx;
    f2 uC = f2_pack(ldf<SLAB>(pu), ldf<SLAB>(pu + RB)), vC = f2_pack(ldf<SLAB>(pv), ldf<SLAB>(pv + RB));
    f2 uB = f2_pack(ldf<SLAB>(pu + dAm), ldf<SLAB>(pu + (RB - 1))), vT = f2_pack(ldf<SLAB>(pv + 1), ldf<SLAB>(pv + dBp));
    f2 vL = 0ull, vLT = 0ull;
    if (i0 > 0) { vL = f2_pack(ldf<SLAB>(pv - H), ldf<SLAB>(pv + (RB - H))); vLT = f2_pack(ldf<SLAB>(pv + (1 - H)), ldf<SLAB>(pv + (dBp - H))); }
    f2 fi = f2_pack((float)i0, (float)i0);
#pragma unroll
    for (int c = 0; c < CPB; ++c) {
        const int i = i0 + c;
        if (i >= c_stop) break;
        const int hr = (i + 1 < W) ? H : 0;   // to column i+1; clamped for memory safety only
        const float *pur = u + idx + hr, *pvr = v + idx + hr, *ps = s + idx;
        const uint8_t *pf = flags + idx;
        const f2 uR = f2_pack(ldf<SLAB>(pur), ldf<SLAB>(pur + RB)), uRB = f2_pack(ldf<SLAB>(pur + dAm), ldf<SLAB>(pur + (RB - 1)));
        const f2 vR = f2_pack(ldf<SLAB>(pvr), ldf<SLAB>(pvr + RB)), vRT = f2_pack(ldf<SLAB>(pvr + 1), ldf<SLAB>(pvr + dBp));
        const float sA = ldf<SLAB>(ps), sB = ldf<SLAB>(ps + RB);
        const bool liveA = (pf[0] & TFS_F_LIVE) != 0u, liveB = (pf[RB] & TFS_F_LIVE) != 0u;
        const f2 fih = add2(fi, half2, k);
        // back-traced positions (:207-236), same association as k_advect32
        const f2 avg_v = mul2(add2(add2(add2(vL, vLT, k), vT, k), vC, k), quarter2, k);   // :245-255
        const f2 avg_u = mul2(add2(add2(add2(uB, uRB, k), uR, k), uC, k), quarter2, k);   // :257-267
        const f2 cv = mul2(add2(vC, vT, k), half2, k), cu = mul2(add2(uC, uR, k), half2, k);   // :228-233
        const f2 xu = sub2(fi, mul2(uC, dt2, k), k), yu = sub2(fjh, mul2(avg_v, dt2, k), k);
        const f2 xv = sub2(fih, mul2(avg_u, dt2, k), k), yv = sub2(fj, mul2(vC, dt2, k), k);
        const f2 xsm = sub2(fih, mul2(cu, dt2, k), k), ysm = sub2(fjh, mul2(cv, dt2, k), k);
        const bool inside = in_range6(xu, xv, xsm, xlo_bits, xhi_bits) & in_range6(yu, yv, ysm, one_bits, hm1_bits);
        float unA = f2_lo(uC), unB = f2_hi(uC), vnA = f2_lo(vC), vnB = f2_hi(vC), snA = sA, snB = sB;
        if (__all_sync(0xffffffffu, inside)) {
            const f2 un = sample_interior_pair<FT_U, SLAB>(u, xu, yu, H, k);
            const f2 vn = sample_interior_pair<FT_V, SLAB>(v, xv, yv, H, k);
            const f2 sn = sample_interior_pair<FT_S, SLAB>(s, xsm, ysm, H, k);
            if (liveA) { unA = f2_lo(un); vnA = f2_lo(vn); snA = f2_lo(sn); }
            if (liveB) { unB = f2_hi(un); vnB = f2_hi(vn); snB = f2_hi(sn); }
        } else {
            if (liveA) {
                const Uvs r = advect_general32<SLAB>(u, v, s, W, H, one, f2_lo(xu), f2_lo(yu), f2_lo(xv), f2_lo(yv), f2_lo(xsm), f2_lo(ysm), &far);
                unA = r.u; vnA = r.v; snA = r.s;
            }
            if (liveB) {
                const Uvs r = advect_general32<SLAB>(u, v, s, W, H, one, f2_hi(xu), f2_hi(yu), f2_hi(xv), f2_hi(yv), f2_hi(xsm), f2_hi(ysm), &far);
                unB = r.u; vnB = r.v; snB = r.s;
            }
        }
        float *qu = nu + idx, *qv = nv + idx, *qs = ns + idx;
        qu[0] = unA; qu[RB] = unB;
        qv[0] = vnA; qv[RB] = vnB;
        qs[0] = snA; qs[RB] = snB;
        // march: column i+1 becomes the own column, column i the left one
        vL = vC; vLT = vT;
        uC = uR; uB = uRB; vC = vR; vT = vRT;
        idx += hr;
        fi = add2(fi, k.one, k);   // exact: integers below 2^24
    }
}

// ---------------------------------------------------------------------------------------------
// x-slab halo exchange: plain peer copies plus sequence flags living in the RECEIVER's memory.
__global__ void k_copy_f4(const float4 *__restrict__ src, float4 *__restrict__ dst, long long n4) {
    long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; t < n4; t += stride) dst[t] = src[t];
}
__global__ void k_flag_set(long long *flag, long long value) {
    __threadfence_system();
    *reinterpret_cast<volatile long long *>(flag) = value;
}
__global__ void k_flag_wait2(const long long *f0, const long long *f1, long long value) {
    // f0 / f1 may be null (no neighbour on that side)
    SpinGuard guard;
    if (f0) while (*reinterpret_cast<const volatile long long *>(f0) < value) { __nanosleep(200); guard.tick(); }
    if (f1) while (*reinterpret_cast<const volatile long long *>(f1) < value) { __nanosleep(200); guard.tick(); }
    __threadfence_system();
}

// ---------------------------------------------------------------------------------------------
// reductions: warp shuffle -> one value per block -> second pass by one block.
TFS_DEVINL float warp_min(float x) {
    for (int o = 16; o; o >>= 1) x = fminf(x, __shfl_xor_sync(0xffffffffu, x, o));
    return x;
}
TFS_DEVINL float warp_max(float x) {
    for (int o = 16; o; o >>= 1) x = fmaxf(x, __shfl_xor_sync(0xffffffffu, x, o));
    return x;
}

// out[2*b] = min, out[2*b+1] = max over the block's share (NaNs are ignored, as by `<` / `>`)
__global__ void k_minmax_partial(const float *__restrict__ a, long long n, float *__restrict__ out) {
    __shared__ float smin[32], smax[32];
    float mn = 3.402823466e38f, mx = -3.402823466e38f;  // f32::MAX / f32::MIN (sim_renderer.rs:24-25)
    long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    long long stride = (long long)gridDim.x * blockDim.x;
    for (; t < n; t += stride) {
        float x = a[t];
        mn = fminf(mn, x);
        mx = fmaxf(mx, x);
    }
    mn = warp_min(mn);
    mx = warp_max(mx);
    int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) { smin[w] = mn; smax[w] = mx; }
    __syncthreads();
    if (w == 0) {
        int nw = blockDim.x >> 5;
        mn = l < nw ? smin[l] : 3.402823466e38f;
        mx = l < nw ? smax[l] : -3.402823466e38f;
        mn = warp_min(mn);
        mx = warp_max(mx);
        if (l == 0) { out[2 * blockIdx.x] = mn; out[2 * blockIdx.x + 1] = mx; }
    }
}
__global__ void k_minmax_final(const float *__restrict__ part, int nblocks, float *__restrict__ out) {
    __shared__ float smin[32], smax[32];
    float mn = 3.402823466e38f, mx = -3.402823466e38f;
    for (int t = threadIdx.x; t < nblocks; t += blockDim.x) {
        mn = fminf(mn, part[2 * t]);
        mx = fmaxf(mx, part[2 * t + 1]);
    }
    mn = warp_min(mn);
    mx = warp_max(mx);
    int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) { smin[w] = mn; smax[w] = mx; }
    __syncthreads();
    if (w == 0) {
        int nw = blockDim.x >> 5;
        mn = l < nw ? smin[l] : 3.402823466e38f;
        mx = l < nw ? smax[l] : -3.402823466e38f;
        mn = warp_min(mn);
        mx = warp_max(mx);
        if (l == 0) { out[0] = mn; out[1] = mx; }
    }
}

// max |divergence| over updatable cells with at least one fluid neighbour (SURVEY a-12);
// partial[b] = block max, combined by k_minmax_final's max lane (min lane unused).
__global__ void k_div_partial(const float *__restrict__ u, const float *__restrict__ v,
                              const uint8_t *__restrict__ flags, Grid g, long long c_begin,
                              long long c_end, float *__restrict__ out) {
    __shared__ float smax[32];
    const long long H = g.H;
    float mx = 0.0f;
    long long t = c_begin * H + blockIdx.x * (long long)blockDim.x + threadIdx.x;
    long long stride = (long long)gridDim.x * blockDim.x;
    for (; t < c_end * H; t += stride) {
        unsigned f = flags[t];
        if ((f & TFS_F_LIVE) && (f & 15u)) {
            float d = fsub(fadd(fsub(u[t + H], u[t]), v[t + 1]), v[t]);
            mx = fmaxf(mx, fabsf(d));
        }
    }
    mx = warp_max(mx);
    int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) smax[w] = mx;
    __syncthreads();
    if (w == 0) {
        int nw = blockDim.x >> 5;
        mx = l < nw ? smax[l] : 0.0f;
        mx = warp_max(mx);
        if (l == 0) { out[2 * blockIdx.x] = 0.0f; out[2 * blockIdx.x + 1] = mx; }
    }
}

// ---------------------------------------------------------------------------------------------
// Order-independent 64-bit checksum of the bit patterns of u, v, p, s over the columns [c_begin, c_end) of the
// storage: every element contributes h = splitmix64(global_index << 32 ^ bits) to a wrapping SUM and
// splitmix64(h) to an XOR, so the pair is exact under any partition of the grid (x-slabs add / xor their
// shares) and sensitive to where a value sits.  out[2f] += sum, out[2f+1] ^= xor for field f in u, v, p, s.
__global__ void __launch_bounds__(256) k_checksum(const float *__restrict__ u, const float *__restrict__ v, const float *__restrict__ p,
                                                  const float *__restrict__ s, Grid g, long long c_begin, long long c_end,
                                                  unsigned long long *__restrict__ out) {
    const float *f[4] = {u, v, p, s};
    unsigned long long sum[4] = {0, 0, 0, 0}, x[4] = {0, 0, 0, 0};
    const long long lo = c_begin * g.H, hi = c_end * g.H, shift = g.x0 * g.H;
    long long t = lo + blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; t < hi; t += stride) {
        const unsigned long long gi = (unsigned long long)(t + shift);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const unsigned long long h = splitmix64((gi << 32) ^ (unsigned long long)__float_as_uint(f[k][t]) ^ ((unsigned long long)k << 62));
            sum[k] += h;
            x[k] ^= splitmix64(h);
        }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        for (int o = 16; o; o >>= 1) {
            sum[k] += __shfl_xor_sync(0xffffffffu, sum[k], o);
            x[k] ^= __shfl_xor_sync(0xffffffffu, x[k], o);
        }
        if ((threadIdx.x & 31) == 0) {
            atomicAdd(out + 2 * k, sum[k]);
            atomicXor(out + 2 * k + 1, x[k]);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// view packing: dst[(x-vx0)*vh + (y-vy0)] = src[(x-g.x0)*H + y]  (per-frame D2H staging)
template <typename T>
__global__ void k_pack_view(const T *__restrict__ src, T *__restrict__ dst, Grid g, long long vx0,
                            long long vy0, long long vw, long long vh) {
    long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t >= vw * vh) return;
    long long x = vx0 + t / vh, y = vy0 + t % vh;
    dst[t] = src[(x - g.x0) * g.H + y];
}

// ---------------------------------------------------------------------------------------------
// renderer colour path (src/ui/sim_renderer.rs:85-151) on the device: the per-frame D2H becomes
// 8 bytes per terminal cell instead of three f32/u8 fields of the visible window.

// Rust `f32 as u8`: truncating, saturating, NaN -> 0  (cvt.rzi.u32.f32 saturates and maps NaN to 0)
TFS_DEVINL unsigned f32_as_u8(float x) { return min(__float2uint_rz(x), 255u); }

// get_linear_gradient (:116-140) then get_color (:142-151) / THEME.sim_blocks (:92); returns
// r | g<<8 | b<<16 | kind<<24, kind 1 = block (the named colour White, not an RGB triple)
TFS_DEVINL uint32_t cell_colour(float value, float smoke, uint8_t block, float mn, float mx) {
    if (block) return 0x01FFFFFFu;
    const float difference = fsub(mx, mn);
    const float normalized = (difference == 0.0f) ? 0.5f : __fdiv_rn(fsub(value, mn), difference);
    const float color_type = floorf(fmul(normalized, 4.0f));
    const float cv = __fdiv_rn(fsub(normalized, fmul(color_type, 0.25f)), 0.25f);
    float r, g, b;
    switch (f32_as_u8(color_type)) {
    case 0: r = 0.0f; g = cv; b = 1.0f; break;
    case 1: r = 0.0f; g = 1.0f; b = fsub(1.0f, cv); break;
    case 2: r = cv; g = 1.0f; b = 0.0f; break;
    case 3: r = 1.0f; g = fsub(1.0f, cv); b = 0.0f; break;
    default: r = 1.0f; g = 0.0f; b = 0.0f; break;
    }
    const unsigned red = f32_as_u8(fmul(255.0f, smoke));
    const unsigned R = f32_as_u8(fmul(255.0f, r)), G = f32_as_u8(fmul(255.0f, g)), B = f32_as_u8(fmul(255.0f, b));
    return (R - min(R, red)) | ((G - min(G, red)) << 8) | ((B - min(B, red)) << 16);  // saturating_sub
}

// One CTA colours a tile of 32 columns x 64 sim rows (= 32 terminal rows, two cells each,
// sim_renderer.rs:45-82): phase 1 reads p/s/m with threads along y (the contiguous direction of
// idx = x*H + y), phase 2 writes with threads along x (the contiguous direction of the terminal
// buffer, index = row*area_w + x; bg colour first = the UPPER sim cell H-1-2*row, then fg).
__global__ void __launch_bounds__(256) k_render_view(const float *__restrict__ p, const float *__restrict__ s,
                                                     const uint8_t *__restrict__ m, Grid g, long long vx0, int area_w, int area_h,
                                                     const float *__restrict__ minmax, uint32_t *__restrict__ out) {
    __shared__ uint32_t tile[64][33];
    const float mn = minmax[0], mx = minmax[1];
    const int tx0 = blockIdx.x * 32, row0 = blockIdx.y * 32;
    for (int id = threadIdx.x; id < 32 * 64; id += 256) {
        const int ly = id & 63, lx = id >> 6;
        const int x = tx0 + lx, trow = row0 + (ly >> 1);
        if (x < area_w && trow < area_h) {
            const long long y = g.H - 1 - 2LL * row0 - ly;
            const long long idx = (vx0 + x - g.x0) * g.H + y;
            tile[ly][lx] = cell_colour(p[idx], s[idx], m[idx], mn, mx);
        }
    }
    __syncthreads();
    for (int id = threadIdx.x; id < 32 * 64; id += 256) {
        const int j = id & 63, r = id >> 6;  // j = 2*lx + half within terminal row r of the tile
        const int x = tx0 + (j >> 1), trow = row0 + r;
        if (x < area_w && trow < area_h) out[((long long)trow * area_w + x) * 2 + (j & 1)] = tile[2 * r + (j & 1)][j >> 1];
    }
}

// ---------------------------------------------------------------------------------------------
// scenes (same integer predicates as oracle/fluid_oracle.c: orc_scene_*)
__global__ void k_scene_disc(uint8_t *__restrict__ m, Grid g, long long cx, long long cy, long long r) {
    long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t >= g.ncols * g.H) return;
    long long i = g.x0 + t / g.H, j = t % g.H;
    long long dx = 2 * i + 1 - 2 * cx, dy = 2 * j + 1 - 2 * cy;
    if (dx * dx + dy * dy < 4 * r * r) m[t] = 1;
}
__global__ void k_scene_lattice(uint8_t *__restrict__ m, Grid g, long long pitch, long long r) {
    long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t >= g.ncols * g.H) return;
    long long i = g.x0 + t / g.H, j = t % g.H;
    long long a = i / pitch, b = j / pitch;
    if (a == 0) return;
    long long cx = pitch / 2 + pitch * a, cy = pitch / 2 + pitch * b;
    long long dx = 2 * i + 1 - 2 * cx, dy = 2 * j + 1 - 2 * cy;
    if (dx * dx + dy * dy < 4 * r * r) m[t] = 1;
}
__global__ void k_scene_random(uint8_t *__restrict__ m, Grid g, uint64_t seed, uint64_t threshold) {
    long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t >= g.ncols * g.H) return;
    long long i = g.x0 + t / g.H, j = t % g.H;
    if (is_border(i, j, g.W, g.H)) return;
    uint64_t gidx = (uint64_t)(i * g.H + j);
    if (splitmix64(seed ^ gidx) < threshold) m[t] = 1;
}
__global__ void k_fill_random_velocity(float *__restrict__ u, float *__restrict__ v, Grid g,
                                       uint64_t seed, float amp) {
    long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t >= g.ncols * g.H) return;
    uint64_t gidx = (uint64_t)(t + g.x0 * g.H);
    uint64_t a = splitmix64(seed ^ (2 * gidx)), b = splitmix64(seed ^ (2 * gidx + 1));
    float fa = fmul((float)(a >> 40), 1.0f / 16777216.0f);
    float fb = fmul((float)(b >> 40), 1.0f / 16777216.0f);
    u[t] = fmul(fsub(fmul(fa, 2.0f), 1.0f), amp);
    v[t] = fmul(fsub(fmul(fb, 2.0f), 1.0f), amp);
}

// exhaustive check of div_small against IEEE division: x over all 2^32 bit patterns
__global__ void k_selftest_division(unsigned long long *mismatches) {
    unsigned long long bad = 0;
    const unsigned long long total = 1ull << 32;
    unsigned long long t = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (; t < total; t += stride) {
        float x = __uint_as_float((unsigned)t);
#pragma unroll
        for (int n = 1; n <= 4; ++n) {
            const float nf = (float)n;
            const float rcp = (n == 3) ? 0.333333343267440796f : ((n == 4) ? 0.25f : ((n == 2) ? 0.5f : 1.0f));
            float ref = __fdiv_rn(x, nf), got = div_small(x, nf, rcp);
            bool same = (__float_as_uint(ref) == __float_as_uint(got)) || (ref != ref && got != got);
            bad += same ? 0 : 1;
        }
    }
    if (bad) atomicAdd(mismatches, bad);
}

}  // namespace tfs
