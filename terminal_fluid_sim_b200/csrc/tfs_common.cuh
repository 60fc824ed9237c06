// tfs_common.cuh -- shared device helpers for the simulator-step kernels (sm_100a).
//
// Bit-exactness rules: Rust evaluates every f32 operation with one rounding
// and never contracts a*b+c.  All arithmetic that must match the reference therefore goes
// through the __f*_rn intrinsics (which nvcc never fuses) and the library is additionally
// compiled with -fmad=false.  FMAs appear only where the product is exact by construction
// (a multiplier of 0.0 or 1.0), in which case fma(a,b,c) == a*b + c bit-for-bit.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#define TFS_DEVINL __device__ __forceinline__

// per-cell flag byte (tfs.h TFS_FIELD_FLAGS)
#define TFS_F_L 1u   // left neighbour is fluid   (simulator.rs:168, "left_is_block" = !block)
#define TFS_F_R 2u   // right neighbour is fluid  (:166)
#define TFS_F_B 4u   // bottom neighbour is fluid (:167)
#define TFS_F_T 8u   // top neighbour is fluid    (:165)
#define TFS_F_LIVE 16u  // !block && !border      (:161)

// Every wait on another GPU (or on another CTA through global memory) is bounded: if the awaited value has not
// arrived after TFS_SPIN_TIMEOUT_NS the kernel traps, so a rank that died or a protocol error surfaces as a
// CUDA error on every rank instead of hanging the box.  Only the slow path (already sleeping) reads the timer.
#ifndef TFS_SPIN_TIMEOUT_NS
#define TFS_SPIN_TIMEOUT_NS 30000000000ULL
#endif
TFS_DEVINL unsigned long long tfs_now_ns() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
struct SpinGuard {
    unsigned long long t0 = 0;
    TFS_DEVINL void tick() {  // call once per unsuccessful poll
        const unsigned long long now = tfs_now_ns();
        if (t0 == 0) t0 = now;
        else if (now - t0 > TFS_SPIN_TIMEOUT_NS) __trap();
    }
};

TFS_DEVINL float fadd(float a, float b) { return __fadd_rn(a, b); }
TFS_DEVINL float fsub(float a, float b) { return __fsub_rn(a, b); }
TFS_DEVINL float fmul(float a, float b) { return __fmul_rn(a, b); }

// index_is_border_with_size (simulator.rs:365-375) on (x, y) instead of the flat index
TFS_DEVINL bool is_border(long long x, long long y, long long W, long long H) {
    return x == 0 || x == W - 1 || y == 0 || y == H - 1;
}

// x / n for n in {1,2,3,4}, correctly rounded, without the slow path of a general divide.
// rcp = RN(1/n).  q0 = RN(x*rcp); r = x - n*q0 exactly (one fma); q = RN(q0 + r*rcp).
// For n = 1, 2, 4 rcp is exact and r == 0 in the normal range; for n = 3 this is the classic
// Markstein correction.  tfs_selftest_division() proves equality with __fdiv_rn for all
// 2^32 inputs x on the device; the same sequence was checked exhaustively on the host
// (gcc -mfma, all 2^32 x, n = 1..4: 0 mismatches).
TFS_DEVINL float div_small(float x, float nf, float rcp) {
    float q0 = __fmul_rn(x, rcp);
    float r = __fmaf_rn(-nf, q0, x);
    float q = __fmaf_rn(r, rcp, q0);
    // |x| = inf or NaN: the product already is the IEEE answer (the fma would make inf - inf).
    // q0 == +-0 (x == +-0 or total underflow): the product is the answer and keeps the sign of
    // x, which 0 + (-0) in the correction would lose.
    return (fabsf(q0) <= 3.402823466e38f && q0 != 0.0f) ? q : q0;
}

// One Gauss-Seidel cell update (simulator.rs:161-186) on register operands.
//   uC = u[idx], uR = u[right], vC = v[idx], vT = v[top], p = p[idx]; f = flag byte of the cell.
// Operand order and rounding follow the reference exactly; s* multipliers are 0.0 / 1.0 so the
// four face updates may use fma (exact product).  p uses mul then add (pc*corr rounds).
TFS_DEVINL void cell_update(unsigned f, float omega, float pc, float &uC, float &uR, float &vC,
                            float &vT, float &p) {
    const unsigned nb = f & 15u;
    if (!(f & TFS_F_LIVE) || nb == 0u) return;  // :161, :172
    const int n = __popc(nb);                   // :169-170 (sum of 0/1 floats is exact)
    const float nf = (float)n;
    const float rcp = (n == 3) ? 0.333333343267440796f : ((n == 4) ? 0.25f : ((n == 2) ? 0.5f : 1.0f));
    const float sL = (f & TFS_F_L) ? 1.0f : 0.0f;
    const float sR = (f & TFS_F_R) ? 1.0f : 0.0f;
    const float sB = (f & TFS_F_B) ? 1.0f : 0.0f;
    const float sT = (f & TFS_F_T) ? 1.0f : 0.0f;
    const float div = fsub(fadd(fsub(uR, uC), vT), vC);  // :176-178
    const float corr = fmul(omega, div_small(-div, nf, rcp));  // :180
    uC = __fmaf_rn(-corr, sL, uC);  // :181  u[idx]   -= corr * sL
    uR = __fmaf_rn(corr, sR, uR);   // :182  u[right] += corr * sR
    vC = __fmaf_rn(-corr, sB, vC);  // :184  v[idx]   -= corr * sB
    vT = __fmaf_rn(corr, sT, vT);   // :185  v[top]   += corr * sT
    p = fadd(p, fmul(pc, corr));    // :186
}

// splitmix64, shared with the oracle's scene generators (integer, bit-exact)
__host__ TFS_DEVINL uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

// Tuning / debugging knobs.  They are environment variables because none is part of
// the product's interface; they are read ONCE per process, on the first call that needs them (tfs_create),
// never on the step path.
struct Tunables {
    int sys_kg = 0, sys_nw = 0;   // TFS_SYS_KG + TFS_SYS_NW: force a systolic shape
    int sys_g = 8, sys_rb = 16;   // TFS_SYS_G = 4 / TFS_SYS_RB = 32: record interval / 32-row ring blocks
    int sys_max_kb = 16;          // TFS_SYS_MAXKB: largest KB the automatic shape choice considers
    bool sys_v3 = false;          // TFS_SYS_V3 = 1: generic kernel even when H % 4 == 0
    char sys_order = 0;           // TFS_SYS_ORDER = d / l / h: task order
    int sys_grid = 0;             // TFS_SYS_GRID = n (cap on resident CTAs), -1 = "even"
    bool small_kernel = true;     // TFS_SMALL = 0: terminal-sized grids take the tiled kernels too
    bool adv_pair = true;         // TFS_ADV_PAIR = 0: one-cell advection kernel
    char advect = 0;              // TFS_ADVECT = vec4 / scalar: generic advection kernels
    int rb_variant = 0;           // TFS_RB_VARIANT = 1: the one-launch-per-colour red-black kernel
    int rb_pf = 4;                // TFS_RB_PF = n: L2 prefetch distance (columns) of the fused red-black kernel
    int rb_sync = 1;              // TFS_RB_SYNC = 0: no per-step barrier in the fused red-black kernel
    int rb_wpc = 4;               // TFS_RB_WPC = 1 / 2 / 4: warps per CTA of the fused red-black kernel
    int rb_chunk = 0;             // TFS_RB_CHUNK = n: columns per chunk of the fused red-black kernel (0 = default)
    const char *task_trace_prefix = nullptr;
};
inline const Tunables &tunables() {
    static const Tunables t = [] {
        Tunables x;
        auto geti = [](const char *n, int d) { const char *e = getenv(n); return e ? atoi(e) : d; };
        x.sys_kg = geti("TFS_SYS_KG", 0); x.sys_nw = geti("TFS_SYS_NW", 0);
        x.sys_g = geti("TFS_SYS_G", 8); x.sys_rb = geti("TFS_SYS_RB", 16);
        x.sys_max_kb = geti("TFS_SYS_MAXKB", 16);
        x.sys_v3 = geti("TFS_SYS_V3", 0) != 0;
        if (const char *e = getenv("TFS_SYS_ORDER")) x.sys_order = e[0];
        if (const char *e = getenv("TFS_SYS_GRID")) x.sys_grid = e[0] == 'e' ? -1 : atoi(e);
        x.adv_pair = geti("TFS_ADV_PAIR", 1) != 0;
        x.small_kernel = geti("TFS_SMALL", 1) != 0;
        if (const char *e = getenv("TFS_ADVECT")) x.advect = e[0];
        x.rb_variant = geti("TFS_RB_VARIANT", 0);
        x.rb_chunk = geti("TFS_RB_CHUNK", 0);
        x.rb_pf = geti("TFS_RB_PF", 4);
        x.rb_wpc = geti("TFS_RB_WPC", 4);
        x.rb_sync = geti("TFS_RB_SYNC", 1);
        x.task_trace_prefix = getenv("TFS_TASK_TRACE_PREFIX");
        return x;
    }();
    return t;
}
