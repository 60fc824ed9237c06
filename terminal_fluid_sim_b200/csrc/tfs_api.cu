// tfs_api.cu -- the C ABI of libtfs_cuda.so (include/tfs.h) and the host-side step driver.
// Single translation unit: kernels are in the .cuh files included below.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <unistd.h>
#include <new>
#include <vector>

#include "../../include/tfs.h"
#include "tfs_kernels.cuh"
#include "tfs_systolic.cuh"
#include "tfs_systolic4.cuh"
#include "tfs_redblack.cuh"
#include "tfs_small.cuh"

#define TFS_VERSION_STRING "tfs-cuda 0.1.0 (sm_100a)"

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define CU(call)                                                                                  \
    do {                                                                                          \
        cudaError_t e_ = (call);                                                                  \
        if (e_ != cudaSuccess)                                                                    \
            return fail(e_ == cudaErrorMemoryAllocation ? TFS_ERR_OOM : TFS_ERR_CUDA, "%s: %s (%s:%d)", \
                        #call, cudaGetErrorString(e_), __FILE__, __LINE__);                       \
    } while (0)

#define NEED(cond, code, msg) \
    do {                      \
        if (!(cond)) return fail(code, "%s", msg); \
    } while (0)

constexpr int kMaxProfSteps = 256;

}  // namespace

struct tfs_sim {
    int device = 0;
    size_t W = 0, H = 0;
    tfs_config cfg{};
    tfs_options opt{};
    cudaStream_t stream = nullptr;
    // fields; *2 are the alternates (advection double buffers, projection ping-pong)
    float *u = nullptr, *v = nullptr, *p = nullptr, *s = nullptr;
    float *u2 = nullptr, *v2 = nullptr, *p2 = nullptr, *s2 = nullptr;
    uint8_t *m = nullptr, *flags = nullptr;
    size_t cap = 0;  // allocated cells per array
    bool flags_dirty = true;
    // scratch
    float *d_part = nullptr;  // reduction partials
    float *h_pin = nullptr;   // pinned, 16 floats
    void *d_stage = nullptr;  // view packing
    size_t stage_bytes = 0;
    tfs::SystolicWorkspace sys;
    tfs::Systolic4Workspace sys4;
    // measurement
    cudaEvent_t t0 = nullptr, t1 = nullptr;
    bool profiling = false;
    cudaEvent_t prof_ev[kMaxProfSteps][4];
    int prof_n = 0;
    bool prof_created = false;
    double acc_ms[3] = {0, 0, 0};
    uint64_t acc_steps = 0;
    uint64_t launches = 0;
    // ---- x-slab placement (world == 1: the slab is the whole grid) ----
    int rank = 0, world = 1;
    long long c_lo = 0, c_hi = 0;  // owned global columns [c_lo, c_hi)
    long long xs0 = 0, ncols = 0;  // stored global columns [xs0, xs0 + ncols) (owned + halos)
    int b0 = 0, b0_next = -1;      // first band of this rank / of the right rank (-1: last rank)
    char *arena = nullptr;         // slab handles keep everything in ONE allocation (one IPC handle)
    struct Layout {
        size_t off_f32[8] = {0}, off_m = 0, off_flags = 0, off_sync = 0, off_cnt = 0, off_ticket = 0, off_bnd = 0, total = 0;
        size_t bnd_cap_f4 = 0;
        int cstride = 0, max_passes = 0;
    } lay;
    int phys[8] = {0, 1, 2, 3, 4, 5, 6, 7};  // logical u v p s u2 v2 p2 s2 -> physical array index
    struct Peer {
        bool present = false, ipc = false;
        char *base = nullptr;
        long long xs0 = 0, ncols = 0, c_lo = 0, c_hi = 0;
        int b0 = 0, b0_next = -1;
        Layout lay;
    } peerL, peerR;
    bool attached = false;
    long long step_seq = 0;        // projections run so far (same on every rank)
    bool inkernel_halo = false;    // the last projection kernel delivered the left rank's halo itself (step_slab)
    int *d_oob = nullptr;          // device flag: an advection sample fell outside the columns this rank can reach
    int *h_oob = nullptr;          // pinned copy, refreshed asynchronously after every slab step
    cudaEvent_t oob_ev = nullptr;  // recorded after that copy
    bool oob_pending = false;
    long long rb_seq = 0;          // red-black passes run so far on a slab handle (same on every rank)
    unsigned long long *d_sum = nullptr;  // 8 x u64 checksum accumulators
    tfs_projection_info last_proj{};

    tfs::Grid grid() const { return tfs::Grid{(long long)W, (long long)H, xs0, ncols}; }
    long long n() const { return ncols * (long long)H; }          // stored cells
    long long gn() const { return (long long)W * (long long)H; }  // cells of the global grid
    bool slab() const { return world > 1; }
};

namespace {

// Halo columns stored on each side of a slab.  Advection reaches ceil(max|vel|*dt) + 2 columns.  The temporally blocked
// red-black pass (tfs_redblack.cuh, NS = 2 * kRbKB stages, at most 8) needs NS + 1 on the LEFT: the first stored column carries a wrong
// flag byte (its left neighbour's mask is not stored), so it is wrong from stage 0, the column d to its right from
// stage d - and the first OWNED column's stored u is also the right face of the column before it, which therefore
// has to be exact through the last stage.  10 keeps the column pitch of the pushes even.
constexpr long long kHalo = 10;
static_assert(2 * tfs::kRbKB + 1 <= kHalo, "the slab halo must cover the red-black pass's overlap");
constexpr int kSlabMaxPasses = 64;    // passes a slab workspace is sized for
constexpr unsigned kBlobMagic = 0x54465331u;  // "TFS1"

struct SlabGeom { long long c_lo, c_hi, xs0, xs1; int b0, b0_next; };

// Bands (32 values of the skewed coordinate, band b covering columns [32b-1, 32b+30] at k = 0) are split
// evenly over the ranks; a rank owns the columns of its bands.  Pure host arithmetic (tfs_slab_bounds).
SlabGeom slab_geom(long long W, int rank, int world) {
    const long long nb0 = (W + 1 + 31) / 32;  // bands that contain real columns (-1 .. W-1)
    auto bstart = [&](int r) { return (long long)r * nb0 / world; };
    SlabGeom g;
    g.b0 = (int)bstart(rank);
    g.b0_next = rank == world - 1 ? -1 : (int)bstart(rank + 1);
    g.c_lo = rank == 0 ? 0 : 32 * bstart(rank) - 1;
    g.c_hi = rank == world - 1 ? W : 32 * bstart(rank + 1) - 1;
    g.xs0 = std::max(0LL, g.c_lo - kHalo);
    g.xs1 = std::min(W, g.c_hi + kHalo);
    return g;
}

size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

tfs_sim::Layout slab_layout(long long W, long long H, long long ncols, int b0, int b0_next) {
    tfs_sim::Layout L;
    const size_t cells = (size_t)(ncols * H) + 64;
    size_t off = 0;
    for (int i = 0; i < 8; ++i) { L.off_f32[i] = off; off += align_up(cells * sizeof(float), 256); }
    L.off_m = off; off += align_up(cells, 256);
    L.off_flags = off; off += align_up(cells, 256);
    L.off_sync = off; off += 256;
    const int nb_max = (int)((b0_next >= 0 ? b0_next : (W + 32 + 31) / 32) - b0);  // KB <= 32
    L.cstride = nb_max + 2;
    L.max_passes = kSlabMaxPasses;
    L.off_cnt = off; off += align_up((size_t)2 * L.max_passes * L.cstride * sizeof(long long), 256);
    L.off_ticket = off; off += 256;
    L.bnd_cap_f4 = (size_t)2 * (nb_max + 1) * 32 * (size_t)(H + 32);
    L.off_bnd = off; off += align_up(L.bnd_cap_f4 * sizeof(float4), 256);
    L.total = off;
    return L;
}

// indices into the sync area (long long each); all are written by a NEIGHBOUR into this rank's memory
enum { SYNC_PUSHA_FROM_L = 0, SYNC_PUSHA_FROM_R, SYNC_PUSHB_FROM_L, SYNC_PUSHB_FROM_R, SYNC_ADVDONE_FROM_L, SYNC_ADVDONE_FROM_R,
       SYNC_PROJDONE_FROM_R, SYNC_RB_FROM_L, SYNC_RB_FROM_R };

struct PeerBlob {
    unsigned magic;
    int rank, world, device;
    long long W, H, xs0, ncols, c_lo, c_hi;
    int b0, b0_next;
    unsigned long long pid;
    unsigned long long raw_ptr;
    cudaIpcMemHandle_t handle;
};
static_assert(sizeof(PeerBlob) <= TFS_PEER_BLOB_BYTES, "blob too large");

inline unsigned blocks_for(long long n, int bs) { return (unsigned)((n + bs - 1) / bs); }

int set_device(const tfs_sim *s) {
    CU(cudaSetDevice(s->device));
    return TFS_OK;
}

float *&logical(tfs_sim *s, int i) {
    float **fp[] = {&s->u, &s->v, &s->p, &s->s, &s->u2, &s->v2, &s->p2, &s->s2};
    return *fp[i];
}
float *peer_logical(const tfs_sim *s, const tfs_sim::Peer &p, int i) {
    return reinterpret_cast<float *>(p.base + p.lay.off_f32[s->phys[i]]);
}
long long *sync_slot(char *base, const tfs_sim::Layout &L, int idx) { return reinterpret_cast<long long *>(base + L.off_sync) + idx; }

void free_fields(tfs_sim *s) {
    if (s->arena) {
        cudaFree(s->arena);
        s->arena = nullptr;
        s->u = s->v = s->p = s->s = s->u2 = s->v2 = s->p2 = s->s2 = nullptr;
        s->m = s->flags = nullptr;
        s->cap = 0;
        return;
    }
    float **fp[] = {&s->u, &s->v, &s->p, &s->s, &s->u2, &s->v2, &s->p2, &s->s2};
    for (auto f : fp) {
        if (*f) cudaFree(*f);
        *f = nullptr;
    }
    if (s->m) cudaFree(s->m);
    if (s->flags) cudaFree(s->flags);
    s->m = s->flags = nullptr;
    s->cap = 0;
}

// allocate all arrays for `cells` cells (+ one padding column worth so vector tails are safe)
int alloc_fields(tfs_sim *s, size_t cells) {
    free_fields(s);
    if (s->slab()) {
        s->lay = slab_layout((long long)s->W, (long long)s->H, s->ncols, s->b0, s->b0_next);
        CU(cudaMalloc((void **)&s->arena, s->lay.total));
        // masks, flags, sync flags, counters and ticket start at zero (fields are filled by reset_fields)
        CU(cudaMemsetAsync(s->arena + s->lay.off_m, 0, s->lay.off_bnd - s->lay.off_m, s->stream));
        for (int i = 0; i < 8; ++i) { s->phys[i] = i; logical(s, i) = reinterpret_cast<float *>(s->arena + s->lay.off_f32[i]); }
        s->m = reinterpret_cast<uint8_t *>(s->arena + s->lay.off_m);
        s->flags = reinterpret_cast<uint8_t *>(s->arena + s->lay.off_flags);
        s->cap = cells;
        return TFS_OK;
    }
    size_t padded = cells + 64;
    float **fp[] = {&s->u, &s->v, &s->p, &s->s, &s->u2, &s->v2, &s->p2, &s->s2};
    for (auto f : fp) CU(cudaMalloc((void **)f, padded * sizeof(float)));
    CU(cudaMalloc((void **)&s->m, padded));
    CU(cudaMalloc((void **)&s->flags, padded));
    CU(cudaMemsetAsync(s->m, 0, padded, s->stream));
    CU(cudaMemsetAsync(s->flags, 0, padded, s->stream));
    s->cap = cells;
    return TFS_OK;
}

// Rust `f32 as usize` (saturating, NaN -> 0)
size_t f32_as_usize(float x) {
    if (!(x > 0.0f)) return 0;
    if (x >= 18446744073709551616.0f) return SIZE_MAX;
    return (size_t)x;
}

// set_horizontal_speed (simulator.rs:79-85) and set_smoke_pipe (:94-111) on the device copy
int apply_wind_and_pipe(tfs_sim *s) {
    const long long H = (long long)s->H, n = s->gn();  // ranges are GLOBAL flat indices; the fill kernels clip to the slab
    tfs::Grid g = s->grid();
    long long lo = H, hi = std::min(2 * H, n);
    if (hi > lo) {
        tfs::k_fill_range_f32<<<blocks_for(hi - lo, 256), 256, 0, s->stream>>>(s->u, g, lo, hi, s->cfg.wind_speed);
        s->launches++;
    }
    hi = std::min(H, n);
    tfs::k_fill_range_f32<<<blocks_for(hi, 256), 256, 0, s->stream>>>(s->s, g, 0, hi, 1.0f);
    s->launches++;
    const float pipe_height = (float)s->H * s->cfg.smoke_size;  // :101
    const float middle = (float)s->H * 0.5f;                    // :102
    long long mn = (long long)std::min<size_t>(f32_as_usize(middle - pipe_height * 0.5f), (size_t)n);
    long long mx = (long long)std::min<size_t>(f32_as_usize(middle + pipe_height * 0.5f), (size_t)n);
    if (mx > mn) {
        tfs::k_fill_range_f32<<<blocks_for(mx - mn, 256), 256, 0, s->stream>>>(s->s, g, mn, mx, 0.0f);
        s->launches++;
    }
    CU(cudaGetLastError());
    return TFS_OK;
}

// create_horizontal_speed / create_smoke_pipe / zeroed v, p (simulator.rs:31-34, :44-48)
int reset_fields(tfs_sim *s) {
    const long long n = s->n();
    const unsigned gb = std::min<unsigned>(blocks_for(n, 256), 148 * 16);
    tfs::k_fill_f32<<<gb, 256, 0, s->stream>>>(s->u, n, 0.0f);
    tfs::k_fill_f32<<<gb, 256, 0, s->stream>>>(s->v, n, 0.0f);
    tfs::k_fill_f32<<<gb, 256, 0, s->stream>>>(s->p, n, 0.0f);
    tfs::k_fill_f32<<<gb, 256, 0, s->stream>>>(s->s, n, 1.0f);
    s->launches += 4;
    CU(cudaGetLastError());
    return apply_wind_and_pipe(s);
}

int refresh_flags(tfs_sim *s) {
    if (!s->flags_dirty) return TFS_OK;
    tfs::Grid g = s->grid();
    tfs::k_cell_flags<<<blocks_for(s->n(), 256), 256, 0, s->stream>>>(s->m, s->flags, g, 0, s->ncols);
    s->launches++;
    CU(cudaGetLastError());
    s->flags_dirty = false;
    return TFS_OK;
}

int check_options(const tfs_options *o) {
    NEED(o->iterations >= 0, TFS_ERR_INVALID_ARG, "options.iterations must be >= 0");
    NEED(o->mode == TFS_MODE_WAVEFRONT_EXACT || o->mode == TFS_MODE_RED_BLACK, TFS_ERR_INVALID_ARG,
         "options.mode is not a tfs_mode");
    NEED(o->kernel >= TFS_KERNEL_AUTO && o->kernel <= TFS_KERNEL_DIAGONAL, TFS_ERR_INVALID_ARG,
         "options.kernel is not a tfs_kernel");
    NEED(o->temporal_block >= 0 && o->temporal_block <= tfs::kMaxTemporalBlock, TFS_ERR_INVALID_ARG,
         "options.temporal_block out of range");
    NEED(o->world >= 1 && o->rank >= 0 && o->rank < o->world, TFS_ERR_INVALID_ARG, "bad rank/world");
    return TFS_OK;
}

// CUDA loads kernels lazily, and loading one may wait for the device to go idle.  An x-slab step enqueues kernels
// that spin on flags raised by ANOTHER handle's kernels; when one host thread drives several handles (SlabGroup,
// or two handles on one device) a first-time launch behind such a spinning kernel would then never return.  So a
// slab handle loads every kernel of the step path when it is created.  (Measured: two handles on one B200 dead-
// locked in their first step without this, and ran with CUDA_MODULE_LOADING=EAGER.)
void preload_step_kernels() {
    cudaFuncAttributes fa;
    const void *fns[] = {(const void *)tfs::k_cell_flags, (const void *)tfs::k_flag_wait2, (const void *)tfs::k_flag_set,
                         (const void *)tfs::k_copy_f4, (const void *)tfs::k_advect32<8, true, false>,
                         (const void *)tfs::k_advect32_pair<4, 6, true, true>,
                         (const void *)tfs::k_project_rb<tfs::kRbKB, true, true>, (const void *)tfs::k_project_rb<tfs::kRbKB, false, true>,
                         (const void *)tfs::k_fill_f32, (const void *)tfs::k_fill_range_f32, (const void *)tfs::k_fill_range_u8,
                         (const void *)tfs::k_checksum, (const void *)tfs::k_minmax_partial, (const void *)tfs::k_minmax_final,
                         (const void *)tfs::k_div_partial, (const void *)tfs::k_pack_view<float>, (const void *)tfs::k_pack_view<uint8_t>,
                         (const void *)tfs::k_render_view, (const void *)tfs::k_scene_disc, (const void *)tfs::k_scene_lattice,
                         (const void *)tfs::k_scene_random, (const void *)tfs::k_fill_random_velocity};
    for (const void *f : fns) cudaFuncGetAttributes(&fa, f);
    tfs::systolic4_preload();
    cudaGetLastError();
}

// what an x-slab handle can run (tfs_create and tfs_set_options)
int check_slab_options(const tfs_options *o) {
    NEED(o->kernel != TFS_KERNEL_DIAGONAL, TFS_ERR_INVALID_ARG,
         "x-slab handles run the systolic (exact) or the fused red-black projection, not the diagonal kernel");
    NEED(o->iterations >= 1, TFS_ERR_INVALID_ARG, "x-slab handles need iterations >= 1");
    return TFS_OK;
}

// ---- phases ----------------------------------------------------------------------------------

int phase_gravity(tfs_sim *s, float dt) {
    int rc = refresh_flags(s);
    if (rc) return rc;
    const float gdt = s->cfg.gravity * dt;  // simulator.rs:147 (one f32 product, then one add per cell)
    const long long n = s->n();
    const unsigned gb = std::min<unsigned>(blocks_for(n, 256), 148 * 16);
    tfs::k_gravity<<<gb, 256, 0, s->stream>>>(s->v, s->flags, n, gdt);
    s->launches++;
    CU(cudaGetLastError());
    return TFS_OK;
}

int project_diagonal(tfs_sim *s, float omega, float pc) {
    const long long W = (long long)s->W, H = (long long)s->H;
    const int K = s->opt.iterations;
    if (W < 3 || H < 3 || K == 0) return TFS_OK;
    tfs::Grid g = s->grid();
    const long long dmax = W + H - 4;
    const long long len = std::min(W, H) - 2;
    const int bs = 128;
    for (long long t = 2; t <= dmax + 2LL * (K - 1); ++t) {
        long long k_lo = std::max(0LL, (t - dmax + 1) / 2);
        long long k_hi = std::min<long long>(K - 1, (t - 2) / 2);
        if (k_hi < k_lo) continue;
        dim3 grid(blocks_for(len, bs), (unsigned)(k_hi - k_lo + 1));
        tfs::k_project_wavefront<<<grid, bs, 0, s->stream>>>(s->u, s->v, s->p, s->flags, g, t, (int)k_lo, omega, pc);
        s->launches++;
    }
    CU(cudaGetLastError());
    return TFS_OK;
}

int project_redblack(tfs_sim *s, float omega, float pc) {
    tfs::Grid g = s->grid();
    const long long half = ((long long)s->H + 1) / 2;
    const unsigned gb = blocks_for((long long)s->W * half, 256);
    for (int it = 0; it < s->opt.iterations; ++it)
        for (int colour = 0; colour < 2; ++colour) {
            tfs::k_project_redblack<<<gb, 256, 0, s->stream>>>(s->u, s->v, s->p, s->flags, g, colour, omega, pc);
            s->launches++;
        }
    CU(cudaGetLastError());
    s->last_proj.kernel = TFS_PROJ_REDBLACK_TWO_LAUNCH;
    s->last_proj.passes = 2 * s->opt.iterations;
    return TFS_OK;
}

// the temporally blocked red-black kernel needs an even H (8-byte aligned row pairs); options.kernel == DIAGONAL
// (or TFS_RB_VARIANT=1) asks for the simple one-launch-per-colour kernel, kept as its on-device cross-check
bool rb_fused(const tfs_sim *s) {
    return tfs::redblack_fused_usable((long long)s->W, (long long)s->H, s->ncols) &&
           (s->slab() || (s->opt.kernel != TFS_KERNEL_DIAGONAL && tunables().rb_variant != 1));
}

int halo_push(tfs_sim *s, const int *which, int nwhich, int flag_from_r_idx, int flag_from_l_idx, long long seq, bool to_left);
int flags_wait(tfs_sim *s, int idx_from_l, int idx_from_r, long long value);

// Red-black order, temporally blocked: one launch per pass of kRbKB iterations (tfs_redblack.cuh).  On an
// x-slab handle every pass is followed by a push of the boundary columns of u and v (the pass's output
// set) into the neighbours' halo columns; the next pass waits for the neighbours' pushes.  `seq` is the
// step number (slabs only): the first push of a step must wait until both neighbours have finished the
// previous step's advection, which still reads the halo columns about to be overwritten.
int project_redblack_fused(tfs_sim *s, float omega, float pc, float gdt, bool fuse_gravity, long long seq) {
    const int K = s->opt.iterations, KB = tfs::kRbKB;
    const int passes = (K + KB - 1) / KB;
    tfs::RbLaunchArgs a{};
    a.set[0][0] = s->u; a.set[0][1] = s->v; a.set[0][2] = s->p;
    a.set[1][0] = s->u2; a.set[1][1] = s->v2; a.set[1][2] = s->p2;
    a.flags = s->flags;
    a.W = (long long)s->W; a.H = (long long)s->H; a.xs0 = s->xs0; a.ncols = s->ncols; a.c_lo = s->c_lo; a.c_hi = s->c_hi;
    a.omega = omega; a.pc = pc; a.gdt = gdt; a.apply_gravity = fuse_gravity;
    a.chunk_len = tunables().rb_chunk;
    a.slab = s->slab();
    for (int t = 0; t < passes; ++t) {
        const int k_act = std::min(KB, K - t * KB);
        cudaError_t e = tfs::redblack_pass(a, t, k_act, s->stream);
        if (e != cudaSuccess) return fail(TFS_ERR_CUDA, "red-black pass %d: %s", t, cudaGetErrorString(e));
        s->launches++;
        if (s->slab()) {
            int rc;
            if (t == 0) { rc = flags_wait(s, SYNC_ADVDONE_FROM_L, SYNC_ADVDONE_FROM_R, seq - 1); if (rc) return rc; }
            const long long rs = ++s->rb_seq;
            const int out_uv[2] = {(t & 1) ? 0 : 4, (t & 1) ? 1 : 5};  // logical indices of the set pass t wrote
            rc = halo_push(s, out_uv, 2, SYNC_RB_FROM_R, SYNC_RB_FROM_L, rs, true);
            if (rc) return rc;
            rc = flags_wait(s, SYNC_RB_FROM_L, SYNC_RB_FROM_R, rs);
            if (rc) return rc;
        }
    }
    if (passes & 1) {  // results are in set 1
        std::swap(s->u, s->u2); std::swap(s->v, s->v2); std::swap(s->p, s->p2);
        std::swap(s->phys[0], s->phys[4]); std::swap(s->phys[1], s->phys[5]); std::swap(s->phys[2], s->phys[6]);
    }
    s->last_proj.kernel = TFS_PROJ_REDBLACK_FUSED;
    s->last_proj.kg = KB; s->last_proj.nw = 4; s->last_proj.passes = passes;
    return TFS_OK;
}

// make_incompressible (simulator.rs:151-190).  `fuse_gravity`: the systolic kernel can apply
// add_gravity while it loads v, saving the separate pass.
int phase_project(tfs_sim *s, float dt, bool fuse_gravity, bool *gravity_done) {
    int rc = refresh_flags(s);
    if (rc) return rc;
    if (gravity_done) *gravity_done = false;
    const float pc = s->cfg.density / dt;  // :155 (inf for dt == 0, like the reference)
    const float omega = s->opt.omega;
    if (s->opt.iterations == 0) {  // only `pressure_grid.fill(0.0)` remains (:154)
        CU(cudaMemsetAsync(s->p, 0, (size_t)s->n() * sizeof(float), s->stream));
        return TFS_OK;
    }
    s->last_proj = tfs_projection_info{};
    if (s->opt.mode == TFS_MODE_RED_BLACK) {
        if (rb_fused(s)) {  // p.fill(0) (:154) is implied: the first pass does not read p
            int rc2 = project_redblack_fused(s, omega, pc, s->cfg.gravity * dt, fuse_gravity, s->step_seq);
            if (rc2 == TFS_OK && gravity_done) *gravity_done = fuse_gravity;
            return rc2;
        }
        CU(cudaMemsetAsync(s->p, 0, (size_t)s->n() * sizeof(float), s->stream));  // :154
        return project_redblack(s, omega, pc);
    }
    int kernel = s->opt.kernel;
    if (kernel == TFS_KERNEL_AUTO) kernel = tfs::systolic_available() ? TFS_KERNEL_SYSTOLIC : TFS_KERNEL_DIAGONAL;
    if (kernel == TFS_KERNEL_SYSTOLIC) {
        NEED(tfs::systolic_available(), TFS_ERR_INVALID_ARG, "systolic kernel not built");
        tfs::SystolicArgs a{};
        a.u = &s->u; a.v = &s->v; a.p = &s->p; a.u2 = &s->u2; a.v2 = &s->v2; a.p2 = &s->p2;
        a.flags = s->flags;
        a.W = (long long)s->W; a.H = (long long)s->H;
        a.iterations = s->opt.iterations; a.temporal_block = s->opt.temporal_block;
        a.omega = omega; a.pc = pc;
        a.gdt = s->cfg.gravity * dt; a.fuse_gravity = fuse_gravity;
        a.stream = s->stream;
        const char *msg = nullptr;
        uint64_t launched = 0;
        tfs::Sys4External ext{};
        const tfs::Sys4External *extp = nullptr;
        if (s->slab()) {
            NEED(s->attached, TFS_ERR_COMM, "x-slab handle: call tfs_peer_attach before stepping");
            ext.bnd = reinterpret_cast<float4 *>(s->arena + s->lay.off_bnd);
            ext.bnd_cap_f4 = s->lay.bnd_cap_f4;
            ext.cnt = reinterpret_cast<long long *>(s->arena + s->lay.off_cnt);
            ext.ticket = reinterpret_cast<int *>(s->arena + s->lay.off_ticket);
            ext.max_passes = s->lay.max_passes;
            ext.cstride = s->lay.cstride;
            ext.xs0 = s->xs0; ext.ncols = s->ncols; ext.c_lo = s->c_lo; ext.b0 = s->b0; ext.b0_next = s->b0_next;
            ext.seq = s->step_seq;
            s->inkernel_halo = false;
            ext.halo_cols = (int)kHalo;
            ext.inkernel_halo = &s->inkernel_halo;
            if (s->peerL.present) {
                ext.halo_ready_peerL = sync_slot(s->peerL.base, s->peerL.lay, SYNC_PROJDONE_FROM_R);
                // set 0 = logical (u, v, p), set 1 = logical (u2, v2, p2) -- the same mapping on every rank
                for (int f = 0; f < 3; ++f) {
                    ext.peerL_buf[0][f] = peer_logical(s, s->peerL, f);
                    ext.peerL_buf[1][f] = peer_logical(s, s->peerL, 4 + f);
                }
                ext.peerL_xs0 = s->peerL.xs0;
                ext.peerL_cnt = reinterpret_cast<long long *>(s->peerL.base + s->peerL.lay.off_cnt);
                ext.peerL_cstride = s->peerL.lay.cstride;
                ext.peerL_max_passes = s->peerL.lay.max_passes;
                ext.peerL_b0 = s->peerL.b0;
            }
            if (s->peerR.present) {
                ext.wait_right_halo = sync_slot(s->arena, s->lay, SYNC_PUSHB_FROM_R);
                ext.wait_right_halo_value = s->step_seq - 1;
                ext.peerR_bnd = reinterpret_cast<float4 *>(s->peerR.base + s->peerR.lay.off_bnd);
                ext.peerR_cnt = reinterpret_cast<long long *>(s->peerR.base + s->peerR.lay.off_cnt);
                ext.peerR_cstride = s->peerR.lay.cstride;
                ext.peerR_b0 = s->peerR.b0;
                ext.peerR_b0_next = s->peerR.b0_next;
            }
            extp = &ext;
        }
        float *u_before = s->u;
        cudaError_t e = tfs::systolic_project_auto(a, extp, s->sys, s->sys4, &launched, &msg);
        if (e == cudaSuccess && s->u != u_before) {  // odd pass count: the ping-pong sets were swapped
            std::swap(s->phys[0], s->phys[4]);
            std::swap(s->phys[1], s->phys[5]);
            std::swap(s->phys[2], s->phys[6]);
        }
        s->launches += launched;
        if (e != cudaSuccess)
            return fail(e == cudaErrorMemoryAllocation ? TFS_ERR_OOM : TFS_ERR_CUDA, "systolic projection: %s (%s)",
                        cudaGetErrorString(e), msg ? msg : "");
        if (s->sys4.last_kg) {
            s->last_proj.kernel = TFS_PROJ_SYSTOLIC_BLOCK_RING;
            s->last_proj.kg = s->sys4.last_kg; s->last_proj.nw = s->sys4.last_nw; s->last_proj.passes = s->sys4.last_passes;
        } else {
            s->last_proj.kernel = TFS_PROJ_SYSTOLIC_GENERIC;
            s->last_proj.kg = s->sys.last_kg; s->last_proj.nw = s->sys.last_nw; s->last_proj.passes = s->sys.last_passes;
        }
        if (gravity_done) *gravity_done = fuse_gravity;
        return TFS_OK;
    }
    CU(cudaMemsetAsync(s->p, 0, (size_t)s->n() * sizeof(float), s->stream));  // :154
    s->last_proj.kernel = TFS_PROJ_DIAGONAL;
    return project_diagonal(s, omega, pc);
}

int phase_advect(tfs_sim *s, float dt) {
    int rc = refresh_flags(s);
    if (rc) return rc;
    tfs::Grid g = s->grid();
    const long long W = (long long)s->W, H = (long long)s->H;
    const char adv_env = tunables().advect;  // tuning aid: 'v' / 's' force the generic kernels
    // 32-bit cell indices, and W, H < 2^22 for the magic-constant floor of the fast kernels (floor_magic)
    const bool fits32 = W * H < (1LL << 31) - (1LL << 20) && W < (1LL << 22) && H < (1LL << 22);
    const float4 dims = make_float4((float)W, (float)H, (float)(W - 1), (float)(H - 1));
    if (s->slab()) {
        // owned columns only; pointers are rebased to global columns, halos supply the neighbours' cells.  Samples that
        // leave the stored columns (|vel|*dt beyond the halo: SURVEY H4) are loaded from the neighbours' memory:
        // their pre-advection u, v, s stay intact until their next projection, which waits for this rank's pushB -
        // the whole kernel of the right neighbour, but only the LAST band of the left one in the exact mode, whose
        // earlier bands write columns below c_lo - 32 (tfs_systolic4.cuh): hence the bounded reach to the left.
        const long long sh = s->xs0 * H;
        tfs::SlabFar far;
        far.cs0 = (unsigned)s->xs0; far.cs1 = (unsigned)(s->xs0 + s->ncols - 1); far.oob = s->d_oob;
        const int fld[3] = {0, 1, 3};  // logical u, v, s
        if (s->peerL.present) {
            for (int f = 0; f < 3; ++f) far.L[f] = peer_logical(s, s->peerL, fld[f]) - s->peerL.xs0 * H;
            const long long reach = s->opt.mode == TFS_MODE_RED_BLACK ? s->peerL.c_lo : std::max(s->peerL.c_lo, s->c_lo - 32);
            far.l0 = (unsigned)reach; far.l1 = (unsigned)(s->peerL.c_hi - 1);
        }
        if (s->peerR.present) {
            for (int f = 0; f < 3; ++f) far.R[f] = peer_logical(s, s->peerR, fld[f]) - s->peerR.xs0 * H;
            far.r0 = (unsigned)s->peerR.c_lo; far.r1 = (unsigned)(s->peerR.c_hi - 1);
        }
        if (tunables().adv_pair && H % 256 == 0) {
            constexpr int CPB = 4;
            auto two = [](float x) { uint32_t b; memcpy(&b, &x, 4); return (tfs::f2)b | ((tfs::f2)b << 32); };
            const tfs::F2K k{two(1.0f), two(-1.0f), two(-0.0f)};
            dim3 grid((unsigned)(H / 256), blocks_for(s->c_hi - s->c_lo, CPB));
            tfs::k_advect32_pair<CPB, 6, true, true><<<grid, 128, 0, s->stream>>>(s->u - sh, s->v - sh, s->s - sh, s->flags - sh, s->u2 - sh,
                                                                                 s->v2 - sh, s->s2 - sh, (int)W, (int)H, two(dt), 1.0f, k,
                                                                                 (int)s->c_lo, (int)s->c_hi, far);
        } else {
            constexpr int CPB = 8;
            dim3 grid(blocks_for(H, 128), blocks_for(s->c_hi - s->c_lo, CPB));
            tfs::k_advect32<CPB, true><<<grid, 128, 0, s->stream>>>(s->u - sh, s->v - sh, s->s - sh, s->flags - sh, s->u2 - sh, s->v2 - sh,
                                                                     s->s2 - sh, (int)W, (int)H, dt, 1.0f, dims, (int)s->c_lo, (int)s->c_hi, far);
        }
    } else if (fits32 && !adv_env) {
        // Two cells per thread with packed f32x2 arithmetic when the rows tile by 256 (k_advect32_pair), else one
        // cell per thread; both ask the L2 for their tile up front when the column segments are 16-byte aligned.
        const bool pair = tunables().adv_pair;   // tuning aid
        if (pair && H % 256 == 0) {
            auto two = [](float x) { uint32_t b; memcpy(&b, &x, 4); return (tfs::f2)b | ((tfs::f2)b << 32); };
            const tfs::F2K k{two(1.0f), two(-1.0f), two(-0.0f)};
            // 4 columns per block and 64 registers (8 CTAs/SM) are the best of the variants measured on B200 at 4096^2
            // (ms per launch): 8 columns 0.137, 72 / 80 registers 0.123 / 0.130, a rolled march over 8 / 16 / 32 columns
            // 0.126 / 0.136 / 0.158, an L1 prefetch of the tile 0.123-0.127, a shared-memory tile kernel (TMA bulk copies, row
            // pairs, 20 % fewer instructions) 0.125-0.167, against 0.119-0.123 for this form
            constexpr int CPB = 4;
            dim3 grid((unsigned)(H / 256), blocks_for(W, CPB));
            tfs::k_advect32_pair<CPB, 8, true><<<grid, 128, 0, s->stream>>>(s->u, s->v, s->s, s->flags, s->u2, s->v2, s->s2, (int)W,
                                                                          (int)H, two(dt), 1.0f, k);
        } else {
            constexpr int CPB = 8;
            dim3 grid(blocks_for(H, 128), blocks_for(W, CPB));
            if (H % 16 == 0)
                tfs::k_advect32<CPB, false, true><<<grid, 128, 0, s->stream>>>(s->u, s->v, s->s, s->flags, s->u2, s->v2, s->s2, (int)W,
                                                                               (int)H, dt, 1.0f, dims);
            else
                tfs::k_advect32<CPB><<<grid, 128, 0, s->stream>>>(s->u, s->v, s->s, s->flags, s->u2, s->v2, s->s2, (int)W, (int)H, dt,
                                                                  1.0f, dims);
        }
    } else if (H % 4 == 0 && adv_env != 's') {
        tfs::k_advect<4><<<blocks_for(W * (H / 4), 256), 256, 0, s->stream>>>(s->u, s->v, s->s, s->flags, s->u2, s->v2,
                                                                               s->s2, g, 0, W, dt);
    } else {
        tfs::k_advect<1><<<blocks_for(W * H, 256), 256, 0, s->stream>>>(s->u, s->v, s->s, s->flags, s->u2, s->v2,
                                                                         s->s2, g, 0, W, dt);
    }
    s->launches++;
    CU(cudaGetLastError());
    std::swap(s->u, s->u2);  // simulator.rs:240-242
    std::swap(s->v, s->v2);
    std::swap(s->s, s->s2);
    std::swap(s->phys[0], s->phys[4]);
    std::swap(s->phys[1], s->phys[5]);
    std::swap(s->phys[3], s->phys[7]);
    return TFS_OK;
}

// ---- x-slab halo exchange ------------------------------------------------------------------------
// Copy `ncopy` owned boundary columns of the logical fields in `which` into the neighbours' halo
// columns (peer stores over NVLink), then raise `flag_idx_*` in the neighbours' sync areas to `seq`.
int halo_push(tfs_sim *s, const int *which, int nwhich, int flag_from_r_idx, int flag_from_l_idx, long long seq, bool to_left) {
    const long long H = (long long)s->H;
    const long long n4 = kHalo * H / 4;
    const unsigned gb = std::min<unsigned>(blocks_for(n4, 256), 148 * 4);
    for (int side = 0; side < 2; ++side) {
        tfs_sim::Peer &p = side == 0 ? s->peerL : s->peerR;
        if (!p.present || (side == 0 && !to_left)) continue;
        // to the left rank: my first kHalo owned columns -> its right halo; to the right rank: my last kHalo -> its left halo
        const long long col = side == 0 ? s->c_lo : s->c_hi - kHalo;
        for (int i = 0; i < nwhich; ++i) {
            const float *src = logical(s, which[i]) + (col - s->xs0) * H;
            float *dst = peer_logical(s, p, which[i]) + (col - p.xs0) * H;
            tfs::k_copy_f4<<<gb, 256, 0, s->stream>>>(reinterpret_cast<const float4 *>(src), reinterpret_cast<float4 *>(dst), n4);
            s->launches++;
        }
        // the left rank sees this as "from R", the right rank as "from L"
        tfs::k_flag_set<<<1, 1, 0, s->stream>>>(sync_slot(p.base, p.lay, side == 0 ? flag_from_r_idx : flag_from_l_idx), seq);
        s->launches++;
    }
    CU(cudaGetLastError());
    return TFS_OK;
}

int flags_wait(tfs_sim *s, int idx_from_l, int idx_from_r, long long value) {
    const long long *fl = (s->peerL.present && idx_from_l >= 0) ? sync_slot(s->arena, s->lay, idx_from_l) : nullptr;  // idx < 0: no wait
    const long long *fr = (s->peerR.present && idx_from_r >= 0) ? sync_slot(s->arena, s->lay, idx_from_r) : nullptr;
    if (!fl && !fr) return TFS_OK;
    tfs::k_flag_wait2<<<1, 1, 0, s->stream>>>(fl, fr, value);
    s->launches++;
    CU(cudaGetLastError());
    return TFS_OK;
}

int flags_signal(tfs_sim *s, int idx_for_left_rank, int idx_for_right_rank, long long value) {
    if (s->peerL.present) { tfs::k_flag_set<<<1, 1, 0, s->stream>>>(sync_slot(s->peerL.base, s->peerL.lay, idx_for_left_rank), value); s->launches++; }
    if (s->peerR.present) { tfs::k_flag_set<<<1, 1, 0, s->stream>>>(sync_slot(s->peerR.base, s->peerR.lay, idx_for_right_rank), value); s->launches++; }
    CU(cudaGetLastError());
    return TFS_OK;
}

// One step of an x-slab handle.  Everything is asynchronous on the stream; cross-rank ordering is by
// sequence flags in device memory (no host synchronisation, so ranks may also share one host thread).
// The wave of the exact sweep reaches rank r+1 later than rank r (one band = ~36 us at 32768 rows), so in
// steady state every rank runs the same schedule shifted by that skew; the hand-offs below are placed so
// that no rank has to wait for a WHOLE kernel of its right neighbour, which would cost the skew each time:
//   P  wait pushB(seq-1) from the left rank only.  The projection kernel itself waits for the right rank's
//      pushB(seq-1) - in its last local band, the only one that reads the right halo (and writes into the
//      right rank's memory) - and, in its last pass, band 0 stores the left rank's halo columns and last
//      owned columns and raises haloready(seq) there (the PROJDONE slot) as soon as they are complete.
//   A1 wait haloready(seq) from the right rank (it also finishes my last KB-1 columns) and advdone(seq-1) from
//      both; push post-projection u, v to the RIGHT rank; raise pushA(seq) there
//   A2 wait pushA(seq) from the left rank; advect the owned columns
//   A3 raise advdone(seq); push post-advection u, v, s halos; raise pushB(seq)
// Shapes whose band 0 does not cover the halo (KB + kHalo - 1 > 32) fall back to whole-kernel flags:
// projdone(seq) after the kernel, pushA in both directions.
int step_slab(tfs_sim *s, float dt, cudaEvent_t *ev) {
    NEED(s->attached, TFS_ERR_COMM, "x-slab handle: call tfs_peer_attach before stepping");
    NEED(s->opt.iterations > 0, TFS_ERR_INVALID_ARG, "x-slab handles need iterations >= 1");
    const bool rb = s->opt.mode == TFS_MODE_RED_BLACK;
    NEED(!rb || rb_fused(s), TFS_ERR_INVALID_ARG, "red-black on x-slabs needs an even height");
    const long long seq = ++s->step_seq;
    // exact mode: the kernel's last band waits for the right rank's pushB itself; red-black reads both halos at once
    int rc = flags_wait(s, SYNC_PUSHB_FROM_L, rb ? SYNC_PUSHB_FROM_R : -1, seq - 1);
    if (rc) return rc;
    if (ev) CU(cudaEventRecord(ev[1], s->stream));
    bool gravity_done = false;
    rc = phase_project(s, dt, true, &gravity_done);
    if (rc) return rc;
    if (!gravity_done) return fail(TFS_ERR_CUDA, "internal: gravity was not applied");
    if (ev) CU(cudaEventRecord(ev[2], s->stream));
    const int uv[2] = {0, 1}, uvs[3] = {0, 1, 3};
    if (rb) {
        // every pass already exchanged the boundary columns of u and v (project_redblack_fused), the last one
        // included, and it waited for advdone(seq-1): the halos are complete
        rc = phase_advect(s, dt);
        if (rc) return rc;
        rc = flags_signal(s, SYNC_ADVDONE_FROM_R, SYNC_ADVDONE_FROM_L, seq);
        if (rc) return rc;
        rc = halo_push(s, uvs, 3, SYNC_PUSHB_FROM_R, SYNC_PUSHB_FROM_L, seq, true);
        if (rc) return rc;
        if (ev) CU(cudaEventRecord(ev[3], s->stream));
        return TFS_OK;
    }
    const bool fine = s->inkernel_halo;
    if (!fine) {
        // The right rank's first band writes the final values of my last KB-1 owned columns (and my first
        // halo column) straight into my memory: my projection result is complete only once ITS kernel is.
        if (s->peerL.present) {
            tfs::k_flag_set<<<1, 1, 0, s->stream>>>(sync_slot(s->peerL.base, s->peerL.lay, SYNC_PROJDONE_FROM_R), seq);
            s->launches++;
        }
        rc = flags_wait(s, -1, SYNC_PROJDONE_FROM_R, seq);
        if (rc) return rc;
    }
    else {
        // haloready(seq): the right rank's band 0 has finished my last KB-1 owned columns (which I am about to
        // push back to it as its left halo) and my right halo.  It trails my own last band by two bands only.
        rc = flags_wait(s, -1, SYNC_PROJDONE_FROM_R, seq);
        if (rc) return rc;
    }
    rc = flags_wait(s, SYNC_ADVDONE_FROM_L, SYNC_ADVDONE_FROM_R, seq - 1);
    if (rc) return rc;
    rc = halo_push(s, uv, 2, SYNC_PUSHA_FROM_R, SYNC_PUSHA_FROM_L, seq, /*to_left=*/!fine);
    if (rc) return rc;
    rc = flags_wait(s, SYNC_PUSHA_FROM_L, fine ? -1 : SYNC_PUSHA_FROM_R, seq);
    if (rc) return rc;
    rc = phase_advect(s, dt);
    if (rc) return rc;
    rc = flags_signal(s, SYNC_ADVDONE_FROM_R, SYNC_ADVDONE_FROM_L, seq);
    if (rc) return rc;
    rc = halo_push(s, uvs, 3, SYNC_PUSHB_FROM_R, SYNC_PUSHB_FROM_L, seq, true);
    if (rc) return rc;
    if (ev) CU(cudaEventRecord(ev[3], s->stream));
    return TFS_OK;
}

int ensure_prof_events(tfs_sim *s) {
    if (s->prof_created) return TFS_OK;
    for (int i = 0; i < kMaxProfSteps; ++i)
        for (int j = 0; j < 4; ++j) CU(cudaEventCreate(&s->prof_ev[i][j]));
    s->prof_created = true;
    return TFS_OK;
}

// fold recorded per-step events into the accumulators (synchronises the stream)
int drain_prof(tfs_sim *s) {
    if (s->prof_n == 0) return TFS_OK;
    CU(cudaStreamSynchronize(s->stream));
    for (int i = 0; i < s->prof_n; ++i)
        for (int j = 0; j < 3; ++j) {
            float ms = 0;
            CU(cudaEventElapsedTime(&ms, s->prof_ev[i][j], s->prof_ev[i][j + 1]));
            s->acc_ms[j] += ms;
        }
    s->acc_steps += (uint64_t)s->prof_n;
    s->prof_n = 0;
    return TFS_OK;
}

int field_ptr(tfs_sim *s, int field, void **ptr, size_t *elem) {
    switch (field) {
    case TFS_FIELD_U: *ptr = s->u; *elem = 4; break;
    case TFS_FIELD_V: *ptr = s->v; *elem = 4; break;
    case TFS_FIELD_P: *ptr = s->p; *elem = 4; break;
    case TFS_FIELD_S: *ptr = s->s; *elem = 4; break;
    case TFS_FIELD_M: *ptr = s->m; *elem = 1; break;
    case TFS_FIELD_FLAGS: *ptr = s->flags; *elem = 1; break;
    default: return fail(TFS_ERR_INVALID_ARG, "field %d is not a tfs_field", field);
    }
    return TFS_OK;
}

// The advection of an x-slab handle flags samples that left the columns it can reach (d_oob).  The flag is
// copied to pinned memory after every step; `wait` blocks on that copy, otherwise only a finished copy is looked at.
int check_oob(tfs_sim *s, bool wait) {
    if (!s->oob_pending) return TFS_OK;
    if (wait) CU(cudaEventSynchronize(s->oob_ev));
    else if (cudaEventQuery(s->oob_ev) != cudaSuccess) { cudaGetLastError(); return TFS_OK; }
    s->oob_pending = false;
    if (*s->h_oob) {
        *s->h_oob = 0;
        cudaMemsetAsync(s->d_oob, 0, sizeof(int), s->stream);
        return fail(TFS_ERR_OUT_OF_RANGE,
                    "an advection back-trace left the columns this slab can reach (|velocity|*dt too large for the %lld halo "
                    "columns and the neighbour slabs); the fields of that step are not valid", kHalo);
    }
    return TFS_OK;
}

int ensure_stage(tfs_sim *s, size_t bytes) {
    if (s->stage_bytes >= bytes) return TFS_OK;
    if (s->d_stage) cudaFree(s->d_stage);
    s->d_stage = nullptr;
    s->stage_bytes = 0;
    CU(cudaMalloc(&s->d_stage, bytes));
    s->stage_bytes = bytes;
    return TFS_OK;
}

}  // namespace

// ================================= C ABI =======================================================
extern "C" {

const char *tfs_version(void) { return TFS_VERSION_STRING; }
const char *tfs_last_error(void) { return g_err; }

int tfs_device_count(int *count) {
    NEED(count, TFS_ERR_INVALID_ARG, "count is NULL");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        *count = 0;
        return fail(TFS_ERR_NO_DEVICE, "cudaGetDeviceCount: %s", cudaGetErrorString(e));
    }
    *count = n;
    return TFS_OK;
}

void tfs_default_config(tfs_config *cfg) {
    if (!cfg) return;
    cfg->gravity = 0.0f;     // config.rs:19
    cfg->wind_speed = 50.0f; // :20
    cfg->smoke_size = 0.25f; // :21
    cfg->density = 1000.0f;  // :22
}

void tfs_default_options(tfs_options *opt) {
    if (!opt) return;
    opt->iterations = 50;  // simulator.rs:153
    opt->omega = 1.9f;     // simulator.rs:152
    opt->mode = TFS_MODE_WAVEFRONT_EXACT;
    opt->device = 0;
    opt->kernel = TFS_KERNEL_AUTO;
    opt->temporal_block = 0;
    opt->rank = 0;
    opt->world = 1;
}

int tfs_create(size_t width, size_t height, const tfs_config *cfg, const tfs_options *opt, tfs_sim **out) {
    NEED(out, TFS_ERR_INVALID_ARG, "out is NULL");
    *out = nullptr;
    NEED(width >= 1 && height >= 1, TFS_ERR_INVALID_ARG,
         "width and height must be >= 1 (the reference underflows usize at simulator.rs:367)");
    NEED(width <= (1u << 24) && height <= (1u << 24), TFS_ERR_INVALID_ARG, "grid side exceeds 2^24 (f32 coordinates)");
    tfs_options o;
    tfs_default_options(&o);
    if (opt) o = *opt;
    int rc = check_options(&o);
    if (rc) return rc;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(TFS_ERR_NO_DEVICE, "no CUDA device (%s); this library has no CPU fallback",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    NEED(o.device >= 0 && o.device < ndev, TFS_ERR_INVALID_ARG, "options.device out of range");
    tfs_sim *s = new (std::nothrow) tfs_sim();
    NEED(s, TFS_ERR_OOM, "host allocation failed");
    s->device = o.device;
    s->opt = o;
    tfs_default_config(&s->cfg);
    if (cfg) s->cfg = *cfg;
    s->W = width;
    s->H = height;
    s->rank = o.rank;
    s->world = o.world;
    {
        const SlabGeom g = slab_geom((long long)width, o.rank, o.world);
        s->c_lo = g.c_lo; s->c_hi = g.c_hi; s->xs0 = g.xs0; s->ncols = g.xs1 - g.xs0; s->b0 = g.b0; s->b0_next = g.b0_next;
    }
    if (o.world > 1) {
        if (height % 4 != 0) { delete s; return fail(TFS_ERR_INVALID_ARG, "x-slab handles need height %% 4 == 0"); }
        if ((long long)width * (long long)height >= (1LL << 31) - (1LL << 20)) { delete s; return fail(TFS_ERR_INVALID_ARG, "x-slab handles need width*height < 2^31"); }
        if (s->b0_next >= 0 && s->b0_next <= s->b0) { delete s; return fail(TFS_ERR_INVALID_ARG, "grid too narrow for this many slabs"); }
        if ((rc = check_slab_options(&o)) != TFS_OK) { delete s; return rc; }
    }
#define CREATE_TRY(x)         \
    do {                      \
        rc = (x);             \
        if (rc) {             \
            tfs_destroy(s);   \
            return rc;        \
        }                     \
    } while (0)
    CREATE_TRY(set_device(s));
    CREATE_TRY([&]() -> int {
        CU(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
        CU(cudaEventCreate(&s->t0));
        CU(cudaEventCreate(&s->t1));
        CU(cudaMalloc((void **)&s->d_part, 2 * 4096 * sizeof(float)));
        CU(cudaMallocHost((void **)&s->h_pin, 16 * sizeof(float)));
        return TFS_OK;
    }());
    CREATE_TRY(alloc_fields(s, (size_t)s->n()));
    CREATE_TRY([&]() -> int {
        CU(cudaMalloc((void **)&s->d_oob, sizeof(int)));
        CU(cudaMemsetAsync(s->d_oob, 0, sizeof(int), s->stream));
        CU(cudaMallocHost((void **)&s->h_oob, sizeof(int)));
        *s->h_oob = 0;
        CU(cudaEventCreateWithFlags(&s->oob_ev, cudaEventDisableTiming));
        return TFS_OK;
    }());
    if (s->slab()) preload_step_kernels();
    if (s->slab()) CREATE_TRY([&]() -> int {  // nothing may be allocated inside a slab step (see systolic4_reserve)
        CU(tfs::systolic4_reserve(s->sys4, (size_t)s->lay.max_passes * (size_t)(s->lay.cstride - 2)));
        CU(cudaMalloc((void **)&s->d_sum, 8 * sizeof(unsigned long long)));
        return TFS_OK;
    }());
    CREATE_TRY(reset_fields(s));
    // make_block_grid (simulator.rs:113-119): the left border column is solid
    tfs::k_fill_range_u8<<<blocks_for((long long)std::min(height, width * height), 256), 256, 0, s->stream>>>(
        s->m, s->grid(), 0, (long long)std::min(height, width * height), 1);
    s->launches++;
    s->flags_dirty = true;
    CREATE_TRY(refresh_flags(s));
    CREATE_TRY([&]() -> int {
        CU(cudaStreamSynchronize(s->stream));
        return TFS_OK;
    }());
#undef CREATE_TRY
    *out = s;
    return TFS_OK;
}

int tfs_destroy(tfs_sim *s) {
    if (!s) return TFS_OK;
    cudaSetDevice(s->device);
    if (s->stream) cudaStreamSynchronize(s->stream);
    free_fields(s);
    tfs::systolic_release(s->sys);
    tfs::systolic4_release(s->sys4);
    if (s->d_part) cudaFree(s->d_part);
    if (s->h_pin) cudaFreeHost(s->h_pin);
    if (s->d_stage) cudaFree(s->d_stage);
    if (s->d_oob) cudaFree(s->d_oob);
    if (s->h_oob) cudaFreeHost(s->h_oob);
    if (s->oob_ev) cudaEventDestroy(s->oob_ev);
    if (s->d_sum) cudaFree(s->d_sum);
    if (s->peerL.present && s->peerL.ipc) cudaIpcCloseMemHandle(s->peerL.base);
    if (s->peerR.present && s->peerR.ipc) cudaIpcCloseMemHandle(s->peerR.base);
    if (s->prof_created)
        for (int i = 0; i < kMaxProfSteps; ++i)
            for (int j = 0; j < 4; ++j) cudaEventDestroy(s->prof_ev[i][j]);
    if (s->t0) cudaEventDestroy(s->t0);
    if (s->t1) cudaEventDestroy(s->t1);
    if (s->stream) cudaStreamDestroy(s->stream);
    delete s;
    return TFS_OK;
}

int tfs_resize(tfs_sim *s, size_t width, size_t height) {
    NEED(s, TFS_ERR_INVALID_ARG, "sim is NULL");
    NEED(width >= 1 && height >= 1, TFS_ERR_INVALID_ARG, "width and height must be >= 1");
    NEED(width <= (1u << 24) && height <= (1u << 24), TFS_ERR_INVALID_ARG, "grid side exceeds 2^24");
    NEED(!s->slab(), TFS_ERR_INVALID_ARG, "resize is not available on x-slab handles (create new handles instead)");
    int rc = set_device(s);
    if (rc) return rc;
    const size_t n_old = s->W * s->H, n_new = width * height, old_h = s->H, old_w = s->W, old_cap = s->cap;
    // Allocate the new arrays first: on failure the handle keeps its old size and fields (the reference's
    // resize cannot leave the simulator unusable either).
    float *old_f[8];
    for (int i = 0; i < 8; ++i) { old_f[i] = logical(s, i); logical(s, i) = nullptr; }
    uint8_t *m_old = s->m, *flags_old = s->flags;
    s->m = s->flags = nullptr;
    const tfs::Grid g_old = s->grid();
    s->W = width;
    s->H = height;
    s->c_lo = 0; s->c_hi = (long long)width; s->xs0 = 0; s->ncols = (long long)width;  // single-GPU handle: the slab is the grid
    rc = alloc_fields(s, n_new);  // zero-fills m
    if (rc) {
        free_fields(s);  // whatever part of the new set was allocated
        for (int i = 0; i < 8; ++i) logical(s, i) = old_f[i];
        s->m = m_old; s->flags = flags_old; s->cap = old_cap;
        s->W = old_w; s->H = old_h;
        s->c_lo = 0; s->c_hi = (long long)old_w; s->xs0 = 0; s->ncols = (long long)old_w;
        return rc;
    }
    // resize_block_grid (simulator.rs:121-135): clear the old left border, Vec::resize keeps the
    // FLAT prefix (no (x,y) re-mapping), then set the first new_h entries.
    tfs::k_fill_range_u8<<<blocks_for((long long)std::min(old_h, n_old), 256), 256, 0, s->stream>>>(
        m_old, g_old, 0, (long long)std::min(old_h, n_old), 0);
    s->launches++;
    CU(cudaMemcpyAsync(s->m, m_old, std::min(n_old, n_new), cudaMemcpyDeviceToDevice, s->stream));
    tfs::k_fill_range_u8<<<blocks_for((long long)std::min(height, n_new), 256), 256, 0, s->stream>>>(
        s->m, s->grid(), 0, (long long)std::min(height, n_new), 1);
    s->launches++;
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(s->stream));
    for (int i = 0; i < 8; ++i) cudaFree(old_f[i]);
    cudaFree(m_old);
    cudaFree(flags_old);
    // the projection workspaces are laid out for the old size
    tfs::systolic_release(s->sys);
    tfs::systolic4_release(s->sys4);
    for (int i = 0; i < 8; ++i) s->phys[i] = i;
    s->flags_dirty = true;
    rc = reset_fields(s);
    if (rc) return rc;
    return refresh_flags(s);
}

int tfs_restart(tfs_sim *s) {
    NEED(s, TFS_ERR_INVALID_ARG, "sim is NULL");
    // resize(self.width, self.height) (simulator.rs:55-57): same size, so the mask prefix copy is
    // the identity and only the left border is rewritten; no reallocation needed.
    int rc = set_device(s);
    if (rc) return rc;
    if (s->oob_pending) CU(cudaEventSynchronize(s->oob_ev));  // a copy of the flag still in flight must not land later
    CU(cudaMemsetAsync(s->d_oob, 0, sizeof(int), s->stream));
    if (s->h_oob) *s->h_oob = 0;
    s->oob_pending = false;
    const long long hh = (long long)std::min(s->H, s->W * s->H);
    tfs::k_fill_range_u8<<<blocks_for(hh, 256), 256, 0, s->stream>>>(s->m, s->grid(), 0, hh, 1);
    s->launches++;
    s->flags_dirty = true;
    rc = reset_fields(s);
    if (rc) return rc;
    return refresh_flags(s);
}

int tfs_set_config(tfs_sim *s, const tfs_config *cfg) {
    NEED(s && cfg, TFS_ERR_INVALID_ARG, "sim or cfg is NULL");
    int rc = set_device(s);
    if (rc) return rc;
    s->cfg = *cfg;
    return apply_wind_and_pipe(s);
}

int tfs_get_config(const tfs_sim *s, tfs_config *cfg) {
    NEED(s && cfg, TFS_ERR_INVALID_ARG, "sim or cfg is NULL");
    *cfg = s->cfg;
    return TFS_OK;
}

int tfs_set_options(tfs_sim *s, const tfs_options *opt) {
    NEED(s && opt, TFS_ERR_INVALID_ARG, "sim or opt is NULL");
    int rc = check_options(opt);
    if (rc) return rc;
    NEED(opt->device == s->device, TFS_ERR_INVALID_ARG, "the device of a handle cannot change");
    NEED(opt->rank == s->opt.rank && opt->world == s->opt.world, TFS_ERR_INVALID_ARG, "rank/world cannot change");
    if (s->slab()) {  // the same constraints tfs_create enforces
        rc = check_slab_options(opt);
        if (rc) return rc;
    }
    s->opt = *opt;
    return TFS_OK;
}

int tfs_get_options(const tfs_sim *s, tfs_options *opt) {
    NEED(s && opt, TFS_ERR_INVALID_ARG, "sim or opt is NULL");
    *opt = s->opt;
    return TFS_OK;
}

int tfs_get_size(const tfs_sim *s, size_t *width, size_t *height) {
    NEED(s && width && height, TFS_ERR_INVALID_ARG, "NULL argument");
    *width = s->W;
    *height = s->H;
    return TFS_OK;
}

int tfs_set_block(tfs_sim *s, size_t x, size_t y, int value) {
    NEED(s, TFS_ERR_INVALID_ARG, "sim is NULL");
    const size_t idx = x * s->H + y;  // calculate_index (simulator.rs:306-308); no per-axis check, like the reference
    if (idx >= s->W * s->H)
        return fail(TFS_ERR_OUT_OF_RANGE, "index out of bounds: the len is %zu but the index is %zu", s->W * s->H, idx);
    int rc = set_device(s);
    if (rc) return rc;
    const long long lidx = (long long)idx - s->xs0 * (long long)s->H;
    if (lidx >= 0 && lidx < s->n()) {  // a slab keeps only its own columns (+ halos); every rank gets every edit
        CU(cudaMemsetAsync(s->m + lidx, value ? 1 : 0, 1, s->stream));
        s->flags_dirty = true;
    }
    return TFS_OK;
}

int tfs_upload_blocks(tfs_sim *s, const uint8_t *mask, size_t offset, size_t n) {
    NEED(s && (mask || n == 0), TFS_ERR_INVALID_ARG, "NULL argument");
    if (offset > s->W * s->H || n > s->W * s->H - offset)
        return fail(TFS_ERR_OUT_OF_RANGE, "range [%zu, %zu) exceeds the grid (%zu cells)", offset, offset + n, s->W * s->H);
    if (n == 0) return TFS_OK;
    int rc = set_device(s);
    if (rc) return rc;
    {
        // clip the GLOBAL range [offset, offset+n) to the stored columns
        const long long lo = std::max((long long)offset, s->xs0 * (long long)s->H);
        const long long hi = std::min((long long)(offset + n), (s->xs0 + s->ncols) * (long long)s->H);
        if (hi > lo) {
            CU(cudaMemcpyAsync(s->m + (lo - s->xs0 * (long long)s->H), mask + (lo - (long long)offset), (size_t)(hi - lo),
                               cudaMemcpyHostToDevice, s->stream));
            CU(cudaStreamSynchronize(s->stream));  // the caller's buffer is not retained past the call
            s->flags_dirty = true;
        }
    }
    return TFS_OK;
}

namespace {
// Terminal-sized grids: the whole step in one launch of one CTA, fields in shared memory (tfs_small.cuh).
// Measured on B200 (ms per step, one-CTA kernel vs tiled kernels): 48 x 42: 0.21 vs 0.35 exact, 0.17 vs 0.49 red-black;
// 80 x 48: 0.27 vs 0.36, 0.21 vs 0.49; 100 x 100: 0.38 vs 0.36, 0.31 vs 0.45 - one SM stops paying off for the exact
// order at about 8000 cells (its wavefront needs W + H + 2K barriers), for red-black only where shared memory ends.
bool small_step(const tfs_sim *s) {
    const long long cells = (long long)s->W * (long long)s->H;
    return !s->slab() && s->opt.kernel == TFS_KERNEL_AUTO && tunables().small_kernel && tfs::small_step_usable((long long)s->W, (long long)s->H) &&
           (s->opt.mode == TFS_MODE_RED_BLACK || cells <= tfs::kSmallMaxCellsExact);
}
int step_small(tfs_sim *s, float dt, cudaEvent_t *ev) {
    static bool attr_done[64] = {false};
    if (!attr_done[s->device & 63]) {
        CU(cudaFuncSetAttribute(tfs::k_step_small, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)tfs::small_step_smem(tfs::kSmallMaxCells, 1)));
        attr_done[s->device & 63] = true;
    }
    tfs::SmallParams P{};
    P.u = s->u; P.v = s->v; P.s = s->s; P.flags = s->flags;
    P.nu = s->u2; P.nv = s->v2; P.ns = s->s2; P.p = s->p;
    P.W = (int)s->W; P.H = (int)s->H; P.K = s->opt.iterations; P.redblack = s->opt.mode == TFS_MODE_RED_BLACK;
    P.omega = s->opt.omega; P.pc = s->cfg.density / dt;  // simulator.rs:155 (inf for dt == 0, like the reference)
    P.gdt = s->cfg.gravity * dt; P.dt = dt;
    if (ev) CU(cudaEventRecord(ev[1], s->stream));
    tfs::k_step_small<<<1, tfs::kSmallThreads, tfs::small_step_smem((long long)s->W, (long long)s->H), s->stream>>>(P);
    s->launches++;
    CU(cudaGetLastError());
    if (ev) { CU(cudaEventRecord(ev[2], s->stream)); CU(cudaEventRecord(ev[3], s->stream)); }
    std::swap(s->u, s->u2); std::swap(s->v, s->v2); std::swap(s->s, s->s2);  // simulator.rs:240-242
    std::swap(s->phys[0], s->phys[4]); std::swap(s->phys[1], s->phys[5]); std::swap(s->phys[3], s->phys[7]);
    s->last_proj = tfs_projection_info{};
    s->last_proj.kernel = TFS_PROJ_SMALL_FUSED;
    s->last_proj.passes = 1;
    return TFS_OK;
}

int step_body(tfs_sim *s, float dt, cudaEvent_t *ev) {
    int rc;
    if (small_step(s)) return step_small(s, dt, ev);
    if (s->slab()) {
        rc = check_oob(s, false);  // a finished earlier step whose advection overran: report it now
        if (rc) return rc;
        rc = step_slab(s, dt, ev);
        if (rc) return rc;
        CU(cudaMemcpyAsync(s->h_oob, s->d_oob, sizeof(int), cudaMemcpyDeviceToHost, s->stream));
        CU(cudaEventRecord(s->oob_ev, s->stream));
        s->oob_pending = true;
        return TFS_OK;
    }
    // gravity folds into the first load of v of the systolic and of the fused red-black projection
    const bool exact = s->opt.mode == TFS_MODE_WAVEFRONT_EXACT;
    const bool fuse = s->opt.iterations > 0 &&
                      (exact ? (tfs::systolic_available() && (s->opt.kernel == TFS_KERNEL_AUTO || s->opt.kernel == TFS_KERNEL_SYSTOLIC))
                             : rb_fused(s));
    if (!fuse) {
        rc = phase_gravity(s, dt);
        if (rc) return rc;
    }
    if (ev) CU(cudaEventRecord(ev[1], s->stream));
    bool gravity_done = false;
    rc = phase_project(s, dt, fuse, &gravity_done);
    if (rc) return rc;
    if (fuse && !gravity_done) return fail(TFS_ERR_CUDA, "internal: gravity was not applied");
    if (ev) CU(cudaEventRecord(ev[2], s->stream));
    rc = phase_advect(s, dt);
    if (rc) return rc;
    if (ev) CU(cudaEventRecord(ev[3], s->stream));
    return TFS_OK;
}
}  // namespace

int tfs_step(tfs_sim *s, float dt) {
    NEED(s, TFS_ERR_INVALID_ARG, "sim is NULL");
    int rc = set_device(s);
    if (rc) return rc;
    rc = refresh_flags(s);
    if (rc) return rc;
    cudaEvent_t *ev = nullptr;
    if (s->profiling) {
        rc = ensure_prof_events(s);
        if (rc) return rc;
        if (s->prof_n == kMaxProfSteps) {
            rc = drain_prof(s);
            if (rc) return rc;
        }
        ev = s->prof_ev[s->prof_n];
        CU(cudaEventRecord(ev[0], s->stream));
    }
    rc = step_body(s, dt, ev);
    if (rc == TFS_OK && ev) s->prof_n++;  // a failed step leaves no half-recorded event set behind
    return rc;
}

int tfs_phase_gravity(tfs_sim *s, float dt) {
    NEED(s, TFS_ERR_INVALID_ARG, "sim is NULL");
    NEED(!s->slab(), TFS_ERR_INVALID_ARG, "x-slab handles step as a whole (tfs_step)");
    int rc = set_device(s);
    return rc ? rc : phase_gravity(s, dt);
}
int tfs_phase_project(tfs_sim *s, float dt) {
    NEED(s, TFS_ERR_INVALID_ARG, "sim is NULL");
    NEED(!s->slab(), TFS_ERR_INVALID_ARG, "x-slab handles step as a whole (tfs_step)");
    int rc = set_device(s);
    return rc ? rc : phase_project(s, dt, false, nullptr);
}
int tfs_phase_advect(tfs_sim *s, float dt) {
    NEED(s, TFS_ERR_INVALID_ARG, "sim is NULL");
    NEED(!s->slab(), TFS_ERR_INVALID_ARG, "x-slab handles step as a whole (tfs_step)");
    int rc = set_device(s);
    return rc ? rc : phase_advect(s, dt);
}

int tfs_sync(tfs_sim *s) {
    NEED(s, TFS_ERR_INVALID_ARG, "sim is NULL");
    int rc = set_device(s);
    if (rc) return rc;
    CU(cudaStreamSynchronize(s->stream));
    return check_oob(s, true);
}

int tfs_stream_idle(tfs_sim *s, int *idle) {
    NEED(s && idle, TFS_ERR_INVALID_ARG, "NULL argument");
    int rc = set_device(s);
    if (rc) return rc;
    const cudaError_t e = cudaStreamQuery(s->stream);
    if (e != cudaSuccess && e != cudaErrorNotReady) return fail(TFS_ERR_CUDA, "cudaStreamQuery: %s", cudaGetErrorString(e));
    *idle = e == cudaSuccess;
    return TFS_OK;
}

int tfs_read_view(tfs_sim *s, int field, size_t x0, size_t y0, size_t w, size_t h, void *dst) {
    NEED(s && dst, TFS_ERR_INVALID_ARG, "NULL argument");
    if (x0 > s->W || w > s->W - x0 || y0 > s->H || h > s->H - y0)
        return fail(TFS_ERR_OUT_OF_RANGE, "view [%zu+%zu) x [%zu+%zu) exceeds the %zu x %zu grid", x0, w, y0, h, s->W, s->H);
    if ((long long)x0 < s->xs0 || (long long)(x0 + w) > s->xs0 + s->ncols)
        return fail(TFS_ERR_OUT_OF_RANGE, "view columns [%zu, %zu) are not stored on this slab ([%lld, %lld))", x0, x0 + w, s->xs0,
                    s->xs0 + s->ncols);
    int rc = set_device(s);
    if (rc) return rc;
    if (field == TFS_FIELD_FLAGS) {
        rc = refresh_flags(s);
        if (rc) return rc;
    }
    void *src = nullptr;
    size_t elem = 0;
    rc = field_ptr(s, field, &src, &elem);
    if (rc) return rc;
    if (w == 0 || h == 0) return TFS_OK;
    if (h == s->H) {  // full-height window: columns are contiguous, no packing needed
        CU(cudaMemcpyAsync(dst, (char *)src + (x0 - (size_t)s->xs0) * s->H * elem, w * h * elem, cudaMemcpyDeviceToHost, s->stream));
    } else {
        rc = ensure_stage(s, w * h * elem);
        if (rc) return rc;
        const long long nv = (long long)(w * h);
        if (elem == 4)
            tfs::k_pack_view<float><<<blocks_for(nv, 256), 256, 0, s->stream>>>((const float *)src, (float *)s->d_stage, s->grid(),
                                                                             (long long)x0, (long long)y0, (long long)w, (long long)h);
        else
            tfs::k_pack_view<uint8_t><<<blocks_for(nv, 256), 256, 0, s->stream>>>((const uint8_t *)src, (uint8_t *)s->d_stage, s->grid(),
                                                                               (long long)x0, (long long)y0, (long long)w, (long long)h);
        s->launches++;
        CU(cudaGetLastError());
        CU(cudaMemcpyAsync(dst, s->d_stage, w * h * elem, cudaMemcpyDeviceToHost, s->stream));
    }
    CU(cudaStreamSynchronize(s->stream));
    return check_oob(s, true);
}

// single GPU: the whole field (n = width*height).  x-slab handle: the OWNED columns only
// (n = (c_hi - c_lo)*height, see tfs_get_slab); the caller concatenates the ranks.
int tfs_read_field(tfs_sim *s, int field, void *dst, size_t n) {
    NEED(s && dst, TFS_ERR_INVALID_ARG, "NULL argument");
    NEED(n == (size_t)(s->c_hi - s->c_lo) * s->H, TFS_ERR_INVALID_ARG, "n must equal (owned columns)*height");
    return tfs_read_view(s, field, (size_t)s->c_lo, 0, (size_t)(s->c_hi - s->c_lo), s->H, dst);
}

int tfs_write_field(tfs_sim *s, int field, const void *src, size_t n) {
    NEED(s && src, TFS_ERR_INVALID_ARG, "NULL argument");
    NEED(n == s->W * s->H, TFS_ERR_INVALID_ARG, "n must equal width*height");
    NEED(field != TFS_FIELD_FLAGS, TFS_ERR_INVALID_ARG, "FLAGS is derived from M and cannot be written");
    int rc = set_device(s);
    if (rc) return rc;
    void *dst = nullptr;
    size_t elem = 0;
    rc = field_ptr(s, field, &dst, &elem);
    if (rc) return rc;
    // `src` is always the GLOBAL field; a slab copies its stored columns (owned + halos), so no exchange is needed
    CU(cudaMemcpyAsync(dst, (const char *)src + (size_t)s->xs0 * s->H * elem, (size_t)s->n() * elem, cudaMemcpyHostToDevice, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    if (field == TFS_FIELD_M) s->flags_dirty = true;
    return TFS_OK;
}

int tfs_pressure_minmax(tfs_sim *s, float *min_p, float *max_p) {
    NEED(s && min_p && max_p, TFS_ERR_INVALID_ARG, "NULL argument");
    int rc = set_device(s);
    if (rc) return rc;
    const long long n = (s->c_hi - s->c_lo) * (long long)s->H;  // owned columns (a slab reports its own share)
    const int nb = (int)std::min<long long>(blocks_for(n, 256 * 8), 2048);
    tfs::k_minmax_partial<<<nb, 256, 0, s->stream>>>(s->p + (s->c_lo - s->xs0) * (long long)s->H, n, s->d_part);
    tfs::k_minmax_final<<<1, 256, 0, s->stream>>>(s->d_part, nb, s->d_part + 2 * 4096 - 2);
    s->launches += 2;
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(s->h_pin, s->d_part + 2 * 4096 - 2, 2 * sizeof(float), cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    *min_p = s->h_pin[0];
    *max_p = s->h_pin[1];
    // sim_renderer.rs:31-37 is `if x < min {..} else if x > max {..}`: an element that lowers the
    // running minimum is not offered to the maximum.  p[0] and p[1] are always 0.0 when H >= 2
    // (border cells are never written), after which the rule equals plain min/max; the only grid
    // where it differs is the single-cell one, where max stays f32::MIN.
    if (s->gn() == 1) *max_p = -3.402823466e38f;
    return TFS_OK;
}

// render_sim's colour loop (src/ui/sim_renderer.rs:23-83) on the device.  dst receives
// area_h*area_w terminal cells, row-major, 8 bytes each: {r,g,b,kind} of the background (sim cell
// (x0+x, H-1-2*row)) then of the foreground (sim cell (x0+x, H-2-2*row)).
int tfs_render_view(tfs_sim *s, size_t x0, size_t area_w, size_t area_h, const float *minmax_in, void *dst, float *min_p,
                    float *max_p) {
    NEED(s && dst, TFS_ERR_INVALID_ARG, "NULL argument");
    if (x0 > s->W || area_w > s->W - x0 || area_h > s->H / 2)  // the reference's `y_index -= 1` would underflow
        return fail(TFS_ERR_OUT_OF_RANGE, "terminal area %zu x %zu at column %zu exceeds the %zu x %zu grid", area_w, area_h, x0, s->W, s->H);
    if ((long long)x0 < s->xs0 || (long long)(x0 + area_w) > s->xs0 + s->ncols)
        return fail(TFS_ERR_OUT_OF_RANGE, "view columns [%zu, %zu) are not stored on this slab ([%lld, %lld))", x0, x0 + area_w, s->xs0,
                    s->xs0 + s->ncols);
    NEED(minmax_in || s->world == 1, TFS_ERR_INVALID_ARG, "an x-slab handle needs the global {min, max} in minmax_in");
    int rc = set_device(s);
    if (rc) return rc;
    float *d_mm = s->d_part + 2 * 4096 - 2;
    if (minmax_in) {
        CU(cudaMemcpyAsync(d_mm, minmax_in, 2 * sizeof(float), cudaMemcpyHostToDevice, s->stream));
    } else {
        const long long n = s->n();
        const int nb = (int)std::min<long long>(blocks_for(n, 256 * 8), 2048);
        tfs::k_minmax_partial<<<nb, 256, 0, s->stream>>>(s->p, n, s->d_part);
        tfs::k_minmax_final<<<1, 256, 0, s->stream>>>(s->d_part, nb, d_mm);
        s->launches += 2;
    }
    const size_t bytes = area_w * area_h * 8;
    if (bytes) {
        rc = ensure_stage(s, bytes);
        if (rc) return rc;
        dim3 grid((unsigned)((area_w + 31) / 32), (unsigned)((area_h + 31) / 32));
        tfs::k_render_view<<<grid, 256, 0, s->stream>>>(s->p, s->s, s->m, s->grid(), (long long)x0, (int)area_w, (int)area_h, d_mm,
                                                        (uint32_t *)s->d_stage);
        s->launches++;
        CU(cudaGetLastError());
        CU(cudaMemcpyAsync(dst, s->d_stage, bytes, cudaMemcpyDeviceToHost, s->stream));
    }
    CU(cudaMemcpyAsync(s->h_pin, d_mm, 2 * sizeof(float), cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    if (min_p) *min_p = s->h_pin[0];
    if (max_p) *max_p = (s->gn() == 1 && !minmax_in) ? -3.402823466e38f : s->h_pin[1];  // see tfs_pressure_minmax
    return TFS_OK;
}

int tfs_divergence_linf(tfs_sim *s, float *out) {
    NEED(s && out, TFS_ERR_INVALID_ARG, "NULL argument");
    int rc = set_device(s);
    if (rc) return rc;
    rc = refresh_flags(s);
    if (rc) return rc;
    const long long n = s->n();
    const int nb = (int)std::min<long long>(blocks_for(n, 256 * 8), 2048);
    tfs::k_div_partial<<<nb, 256, 0, s->stream>>>(s->u, s->v, s->flags, s->grid(), s->c_lo - s->xs0, s->c_hi - s->xs0, s->d_part);
    tfs::k_minmax_final<<<1, 256, 0, s->stream>>>(s->d_part, nb, s->d_part + 2 * 4096 - 2);
    s->launches += 2;
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(s->h_pin, s->d_part + 2 * 4096 - 2, 2 * sizeof(float), cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    *out = s->h_pin[1];
    return TFS_OK;
}

int tfs_scene_disc(tfs_sim *s, long cx, long cy, long r) {
    NEED(s, TFS_ERR_INVALID_ARG, "sim is NULL");
    int rc = set_device(s);
    if (rc) return rc;
    tfs::k_scene_disc<<<blocks_for(s->n(), 256), 256, 0, s->stream>>>(s->m, s->grid(), cx, cy, r);
    s->launches++;
    CU(cudaGetLastError());
    s->flags_dirty = true;
    return TFS_OK;
}
int tfs_scene_lattice(tfs_sim *s, long pitch, long r) {
    NEED(s && pitch > 0, TFS_ERR_INVALID_ARG, "bad argument");
    int rc = set_device(s);
    if (rc) return rc;
    tfs::k_scene_lattice<<<blocks_for(s->n(), 256), 256, 0, s->stream>>>(s->m, s->grid(), pitch, r);
    s->launches++;
    CU(cudaGetLastError());
    s->flags_dirty = true;
    return TFS_OK;
}
int tfs_scene_random(tfs_sim *s, uint64_t seed, uint64_t threshold) {
    NEED(s, TFS_ERR_INVALID_ARG, "sim is NULL");
    int rc = set_device(s);
    if (rc) return rc;
    tfs::k_scene_random<<<blocks_for(s->n(), 256), 256, 0, s->stream>>>(s->m, s->grid(), seed, threshold);
    s->launches++;
    CU(cudaGetLastError());
    s->flags_dirty = true;
    return TFS_OK;
}
int tfs_fill_random_velocity(tfs_sim *s, uint64_t seed, float amp) {
    NEED(s, TFS_ERR_INVALID_ARG, "sim is NULL");
    int rc = set_device(s);
    if (rc) return rc;
    tfs::k_fill_random_velocity<<<blocks_for(s->n(), 256), 256, 0, s->stream>>>(s->u, s->v, s->grid(), seed, amp);
    s->launches++;
    CU(cudaGetLastError());
    return TFS_OK;
}

int tfs_host_alloc(void **ptr, size_t bytes) {
    NEED(ptr, TFS_ERR_INVALID_ARG, "ptr is NULL");
    CU(cudaMallocHost(ptr, bytes));
    return TFS_OK;
}
int tfs_host_free(void *ptr) {
    if (ptr) CU(cudaFreeHost(ptr));
    return TFS_OK;
}

int tfs_timer_begin(tfs_sim *s) {
    NEED(s, TFS_ERR_INVALID_ARG, "sim is NULL");
    int rc = set_device(s);
    if (rc) return rc;
    CU(cudaEventRecord(s->t0, s->stream));
    return TFS_OK;
}
int tfs_timer_end(tfs_sim *s, float *elapsed_ms) {
    NEED(s && elapsed_ms, TFS_ERR_INVALID_ARG, "NULL argument");
    int rc = set_device(s);
    if (rc) return rc;
    CU(cudaEventRecord(s->t1, s->stream));
    CU(cudaEventSynchronize(s->t1));
    CU(cudaEventElapsedTime(elapsed_ms, s->t0, s->t1));
    return TFS_OK;
}

int tfs_set_profiling(tfs_sim *s, int enabled) {
    NEED(s, TFS_ERR_INVALID_ARG, "sim is NULL");
    int rc = set_device(s);
    if (rc) return rc;
    rc = drain_prof(s);
    if (rc) return rc;
    s->profiling = enabled != 0;
    s->acc_ms[0] = s->acc_ms[1] = s->acc_ms[2] = 0;
    s->acc_steps = 0;
    return TFS_OK;
}
int tfs_get_phase_ms(tfs_sim *s, float *gravity_ms, float *project_ms, float *advect_ms, uint64_t *steps) {
    NEED(s, TFS_ERR_INVALID_ARG, "sim is NULL");
    int rc = set_device(s);
    if (rc) return rc;
    rc = drain_prof(s);
    if (rc) return rc;
    if (gravity_ms) *gravity_ms = (float)s->acc_ms[0];
    if (project_ms) *project_ms = (float)s->acc_ms[1];
    if (advect_ms) *advect_ms = (float)s->acc_ms[2];
    if (steps) *steps = s->acc_steps;
    return TFS_OK;
}

int tfs_launch_count(const tfs_sim *s, uint64_t *launches) {
    NEED(s && launches, TFS_ERR_INVALID_ARG, "NULL argument");
    *launches = s->launches;
    return TFS_OK;
}

int tfs_get_last_projection(const tfs_sim *s, tfs_projection_info *info) {
    NEED(s && info, TFS_ERR_INVALID_ARG, "NULL argument");
    *info = s->last_proj;
    return TFS_OK;
}

int tfs_checksum(tfs_sim *s, uint64_t out[8]) {
    NEED(s && out, TFS_ERR_INVALID_ARG, "NULL argument");
    int rc = set_device(s);
    if (rc) return rc;
    if (!s->d_sum) CU(cudaMalloc((void **)&s->d_sum, 8 * sizeof(unsigned long long)));
    CU(cudaMemsetAsync(s->d_sum, 0, 8 * sizeof(unsigned long long), s->stream));
    const long long n = (s->c_hi - s->c_lo) * (long long)s->H;
    const unsigned gb = (unsigned)std::max<long long>(1, std::min<long long>(blocks_for(n, 256), 148 * 16));
    tfs::k_checksum<<<gb, 256, 0, s->stream>>>(s->u, s->v, s->p, s->s, s->grid(), s->c_lo - s->xs0, s->c_hi - s->xs0, s->d_sum);
    s->launches++;
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(out, s->d_sum, 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    return check_oob(s, true);
}

int tfs_selftest_division(int device, uint64_t *mismatches) {
    NEED(mismatches, TFS_ERR_INVALID_ARG, "mismatches is NULL");
    CU(cudaSetDevice(device));
    unsigned long long *d = nullptr;
    CU(cudaMalloc((void **)&d, sizeof(*d)));
    CU(cudaMemset(d, 0, sizeof(*d)));
    tfs::k_selftest_division<<<148 * 8, 256>>>(d);
    CU(cudaGetLastError());
    unsigned long long h = 0;
    CU(cudaMemcpy(&h, d, sizeof(h), cudaMemcpyDeviceToHost));
    cudaFree(d);
    *mismatches = h;
    return TFS_OK;
}

int tfs_peer_export(tfs_sim *s, void *blob) {
    NEED(s && blob, TFS_ERR_INVALID_ARG, "NULL argument");
    NEED(s->slab(), TFS_ERR_INVALID_ARG, "not an x-slab handle (options.world == 1)");
    int rc = set_device(s);
    if (rc) return rc;
    PeerBlob b{};
    b.magic = kBlobMagic;
    b.rank = s->rank; b.world = s->world; b.device = s->device;
    b.W = (long long)s->W; b.H = (long long)s->H; b.xs0 = s->xs0; b.ncols = s->ncols; b.c_lo = s->c_lo; b.c_hi = s->c_hi;
    b.b0 = s->b0; b.b0_next = s->b0_next;
    b.pid = (unsigned long long)getpid();
    b.raw_ptr = (unsigned long long)(uintptr_t)s->arena;
    CU(cudaIpcGetMemHandle(&b.handle, s->arena));
    memset(blob, 0, TFS_PEER_BLOB_BYTES);
    memcpy(blob, &b, sizeof(b));
    return TFS_OK;
}

namespace {
int attach_one(tfs_sim *s, tfs_sim::Peer &p, const void *blob, int expect_rank) {
    PeerBlob b;
    memcpy(&b, blob, sizeof(b));
    NEED(b.magic == kBlobMagic, TFS_ERR_COMM, "peer blob: bad magic");
    NEED(b.world == s->world && b.rank == expect_rank, TFS_ERR_COMM, "peer blob: wrong rank/world");
    NEED(b.W == (long long)s->W && b.H == (long long)s->H, TFS_ERR_COMM, "peer blob: grid size differs");
    p.xs0 = b.xs0; p.ncols = b.ncols; p.c_lo = b.c_lo; p.c_hi = b.c_hi; p.b0 = b.b0; p.b0_next = b.b0_next;
    p.lay = slab_layout(b.W, b.H, b.ncols, b.b0, b.b0_next);
    if (b.pid == (unsigned long long)getpid()) {
        // same process (one host thread driving several GPUs): plain peer access, no IPC
        if (b.device != s->device) {
            int can = 0;
            CU(cudaDeviceCanAccessPeer(&can, s->device, b.device));
            NEED(can, TFS_ERR_COMM, "devices cannot access each other's memory");
            cudaError_t e = cudaDeviceEnablePeerAccess(b.device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
                return fail(TFS_ERR_COMM, "cudaDeviceEnablePeerAccess: %s", cudaGetErrorString(e));
            cudaGetLastError();
        }
        p.base = reinterpret_cast<char *>((uintptr_t)b.raw_ptr);
        p.ipc = false;
    } else {
        void *ptr = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&ptr, b.handle, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) return fail(TFS_ERR_COMM, "cudaIpcOpenMemHandle: %s", cudaGetErrorString(e));
        p.base = reinterpret_cast<char *>(ptr);
        p.ipc = true;
    }
    p.present = true;
    return TFS_OK;
}
}  // namespace

int tfs_peer_attach(tfs_sim *s, const void *left_blob, const void *right_blob) {
    NEED(s, TFS_ERR_INVALID_ARG, "sim is NULL");
    NEED(s->slab(), TFS_ERR_INVALID_ARG, "not an x-slab handle (options.world == 1)");
    NEED((left_blob != nullptr) == (s->rank > 0), TFS_ERR_INVALID_ARG, "left blob: needed exactly when rank > 0");
    NEED((right_blob != nullptr) == (s->rank < s->world - 1), TFS_ERR_INVALID_ARG, "right blob: needed exactly when rank < world-1");
    int rc = set_device(s);
    if (rc) return rc;
    if (left_blob) { rc = attach_one(s, s->peerL, left_blob, s->rank - 1); if (rc) return rc; }
    if (right_blob) { rc = attach_one(s, s->peerR, right_blob, s->rank + 1); if (rc) return rc; }
    s->attached = true;
    return TFS_OK;
}

int tfs_slab_bounds(size_t width, int rank, int world, size_t *c_lo, size_t *c_hi) {
    NEED(c_lo && c_hi, TFS_ERR_INVALID_ARG, "NULL argument");
    NEED(world >= 1 && rank >= 0 && rank < world && width >= 1, TFS_ERR_INVALID_ARG, "bad rank/world/width");
    const SlabGeom g = slab_geom((long long)width, rank, world);
    *c_lo = (size_t)g.c_lo;
    *c_hi = (size_t)g.c_hi;
    return TFS_OK;
}

int tfs_get_slab(const tfs_sim *s, size_t *c_lo, size_t *c_hi, size_t *stored_lo, size_t *stored_hi) {
    NEED(s && c_lo && c_hi, TFS_ERR_INVALID_ARG, "NULL argument");
    *c_lo = (size_t)s->c_lo;
    *c_hi = (size_t)s->c_hi;
    if (stored_lo) *stored_lo = (size_t)s->xs0;
    if (stored_hi) *stored_hi = (size_t)(s->xs0 + s->ncols);
    return TFS_OK;
}

}  // extern "C"
