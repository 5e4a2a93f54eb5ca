// Shared device code of the solve (K7-K11) and marginalisation (K12) kernels.
// K7-K11: one CTA per sliding-window problem runs the whole trust-region solve on chip:
// residual / Jacobian evaluation, robust corrector, block-sparse J^T J with warp-shuffle segmented
// reductions, scalar-landmark Schur complement, in-shared-memory Cholesky of the reduced system,
// traditional dogleg step, Ceres-1.14 accept / reject / radius logic, candidate evaluation.
// The problem wiring follows xrslam/src/xrslam/estimation/solver.cpp:84-190; the loop restates the
// published Ceres 1.14 trust-region minimiser (not under /root/reference).
#pragma once
#include "backend.cuh"
#include "factors.cuh"
#include <cfloat>

namespace xvbe {
using namespace xv;

constexpr int ST = 512;                    // threads per CTA
constexpr int NW = ST / 32;
constexpr int HPACK = kNP * (kNP + 1) / 2;          // sliding-window instance (Smem)
constexpr int kPreScratch = 1024;          // doubles of scratch per preintegration factor
constexpr int kSpec = 4;                   // candidates of consecutive rejected steps evaluated in one pass (evaluate_cost_multi)

// Ceres 1.14 Solver::Options defaults (Appendix C of SURVEY.md)
constexpr double kInitialRadius = 1e4, kMinRadius = 1e-32, kMinRelDecrease = 1e-3;
constexpr double kFuncTol = 1e-6, kGradTol = 1e-10, kParamTol = 1e-8;
constexpr double kMinDiag = 1e-6, kMaxDiag = 1e32, kMinMu = 1e-8, kMaxMu = 1.0, kMuInc = 10.0;
constexpr int kMaxInvalid = 5;

// in-kernel section timing (clock64 of thread 0 of window problems), read back through xrvio_debug_kprof.  A diagnostic:
// compiled in only with -DXRVIO_KPROF (XRVIO_KPROF=1 python xrslam_b200/build.py --force; tools/phase_probe.py needs that
// build), the product kernel carries neither the clock reads nor the atomics.
__device__ unsigned long long g_kprof[32];
#ifdef XRVIO_KPROF
struct KProf {
    long long last; bool on;
    __device__ void start(bool enable) { on = enable && threadIdx.x == 0; if (on) last = clock64(); }
    __device__ void mark(int i) { if (on) { long long t = clock64(); atomicAdd(&g_kprof[i], (unsigned long long)(t - last)); last = t; } }
    __device__ void count(int i) { if (on) atomicAdd(&g_kprof[i], 1ull); }
};
#else
struct KProf {
    static constexpr bool on = false;
    __device__ __forceinline__ void start(bool) {}
    __device__ __forceinline__ void mark(int) {}
    __device__ __forceinline__ void count(int) {}
};
#endif

// Address-space facts the compiler cannot see once a function is not inlined into the kernel (the large-window instance):
// without them every access through `s` or H becomes a generic LD.E / ST.E / CAS-loop atomic.
#if defined(XRVIO_DEBUG_ASSUME) && !defined(CUEMU)
// diagnostic build (XRVIO_DEBUG_ASSUME=1 python xrslam_b200/build.py --force): the assumptions become device-side assertions, so a
// wrong address space traps instead of being undefined behaviour
#include <cassert>
#define XV_ASSUME_SHARED(p) assert(__isShared(p))
#define XV_ASSUME_GLOBAL(p) assert(__isGlobal(p))
#elif !defined(CUEMU)
#define XV_ASSUME_SHARED(p) __builtin_assume(__isShared(p))
#define XV_ASSUME_GLOBAL(p) __builtin_assume(__isGlobal(p))
#else
#define XV_ASSUME_SHARED(p) ((void)0)
#define XV_ASSUME_GLOBAL(p) ((void)0)
#endif

// H of the reduced camera system, packed lower-triangular (row i at i(i+1)/2): inside the CTA's shared memory for sliding
// windows (<= kMaxActive frames), in global memory (L2-resident scratch) for large windows (backend_solve_large.cu).
template <int HP, bool SHARED> struct HStore;
template <int HP> struct HStore<HP, true> {
    double a[HP];
    __device__ __forceinline__ double &operator[](int i) { return a[i]; }
    __device__ __forceinline__ const double &operator[](int i) const { return a[i]; }
};
template <int HP> struct HStore<HP, false> {
    double *a;
    __device__ __forceinline__ double *base() const {
        XV_ASSUME_GLOBAL(a);                   // lets the compiler emit LDG / STG / RED.F64 instead of generic accesses
        return a;
    }
    __device__ __forceinline__ double &operator[](int i) { return base()[i]; }
    __device__ __forceinline__ const double &operator[](int i) const { return base()[i]; }
};

template <int KA, bool SHARED> struct SmemT {
    static constexpr int kA = KA;              // free frames per problem
    static constexpr int NPv = 15 * KA;        // pose / motion columns
    static constexpr int HP = NPv * (NPv + 1) / 2;
    static constexpr bool kSharedH = SHARED;
    HStore<HP, SHARED> H;                  // packed lower-triangular, row i at i(i+1)/2
    double xs[kMaxFrames * 16];            // current iterate
    double cs[kMaxFrames * 16];            // candidate
    double csx[(kSpec - 1) * KA * 16];     // active frames of the speculative candidates 1 .. kSpec-1 (their constant frames are those of xs)
    double us[kMaxFrames * 16];            // user memory (last state written back)
    double g[NPv], scale[NPv], D[NPv], y[NPv], bred[NPv];
    double du[NPv], dw[NPv];               // unscaled steepest-descent / Gauss-Newton directions of the current linearisation
    double colp[6 * NPv];                  // the six scaled columns of the current Cholesky panel (gauss_newton)
    double linv[SHARED ? (NPv / 6) * 36 : 1];   // inverses of the 6 x 6 diagonal blocks of L (sliding-window instance: the substitutions
                                           // multiply by them instead of solving each triangle on one thread)
    double red[NW];
    double red4[2 * kSpec * NW];           // partial sums of evaluate_cost_multi
    double sp[kSpec * 8];                  // per candidate of a batch: radius, c1, c2, step norm, model cost change, valid, cost, |dx|^2
    double jrinv[KA * 9];
    unsigned char col_free[NPv];
    unsigned char pose_free[KA], motion_free[KA];
    double sc[16];                         // broadcast scalars
    int flag;
};
using Smem = SmemT<kMaxActive, true>;

__device__ __forceinline__ int hidx(int i, int j) { return i * (i + 1) / 2 + j; }   // i >= j

template <class SM> __device__ __forceinline__ double block_sum(double v, SM &s) {
    XV_ASSUME_SHARED(&s);
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) s.red[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
    for (int k = 0; k < NW; ++k) t += s.red[k];
    return t;
}
template <class SM> __device__ __forceinline__ double block_max(double v, SM &s) {
    XV_ASSUME_SHARED(&s);
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    __syncthreads();
    if ((threadIdx.x & 31) == 0) s.red[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
    for (int k = 0; k < NW; ++k) t = fmax(t, s.red[k]);
    return t;
}

// D(8x8) += A(8x4) B(4x8), FP64.  Fragments (PTX ISA, mma.m8n8k4 .f64): with g = lane >> 2, t = lane & 3 a lane holds
// A[g][t], B[t][g] and D[g][2t], D[g][2t+1].
__device__ __forceinline__ void dmma_m8n8k4(double &d0, double &d1, double a, double b) {
#ifdef CUEMU
    // test-only SIMT emulator: the same contraction spelled with shuffles, so the fragment layout is what gets tested
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    for (int k = 0; k < 4; ++k) {
        const double ak = __shfl_sync(0xffffffffu, a, g * 4 + k);
        const double b0 = __shfl_sync(0xffffffffu, b, (2 * t) * 4 + k);
        const double b1 = __shfl_sync(0xffffffffu, b, (2 * t + 1) * 4 + k);
        d0 += ak * b0;
        d1 += ak * b1;
    }
#else
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
#endif
}

// Segmented suffix sum over lanes with sorted (contiguous) keys; segment heads end up with the sum.  Which partner values a lane
// takes depends on the keys alone, so the five key comparisons are done once per batch of observations (seg_mask) and every
// one of the ~90 sums of the batch is five value shuffles with predicated adds.
__device__ __forceinline__ unsigned seg_mask(int key, int lane) {
    unsigned m = 0;
#pragma unroll
    for (int b = 0; b < 5; ++b) {
        const int o = 1 << b;
        const int k = __shfl_down_sync(0xffffffffu, key, o);
        if (lane + o < 32 && k == key) m |= 1u << b;
    }
    return m;
}
__device__ __forceinline__ double seg_sum(double v, unsigned take) {
#pragma unroll
    for (int b = 0; b < 5; ++b) {
        const double t = __shfl_down_sync(0xffffffffu, v, 1 << b);
        if ((take >> b) & 1u) v += t;
    }
    return v;
}

struct Prob {
    SolveDesc d;
    double *states_g;              // user memory in global
    const unsigned char *flags_g;
    double *invd_user;             // user memory (global)
    const RpeObs *rpe;
    const RotObs *rot;
    const PreRef *pre;
    const double *pre_blocks;
    const double *pre_out;         // output blocks of the preintegration kernel still on the device (PreRef::block < 0: block ~block there)
    const int *marg_frames;
    const double *marg_lin, *marg_S, *marg_info, *marg_A0;
    double *rpe_tb;                // per observation: tangent basis b1, b2 of the target bearing (6 doubles)
    double *marg_St;               // transposed copy of the non-zero rows of S (scratch), St[c * rank + r]
    int marg_n, marg_rank;
    // scratch (global)
    double *inv_x, *inv_c, *h_ll, *g_l, *sc_l, *D_l, *y_l, *ht_l, *W, *Hcopy, *pre_s, *m_dx, *m_r, *m_v, *m_T, *du_l, *dw_l, *pre_r;
    size_t Mp;                     // doubles between the inverse-depth arrays of consecutive candidates (P.inv_c + b * Mp)
    int M, NPc;                    // free landmarks, pose/motion columns = 15 * n_active
    // J^T J by groups (sliding-window instance): the weighted Jacobian rows of every reprojection residual, transposed
    // ([26][NJ]: Jt row 0, Jr row 0, Jt row 1, Jr row 1 (6 each), r0, r1), and the runs of equal (target, reference) in the sorted list
    double *JB; int *grp_start; int n_grp, NJ;
    // W = H_pl (pose rows x landmarks) is stored TILE-MAJOR: tile t holds the landmarks [t 2^W_sh, (t + 1) 2^W_sh) of all W_rows pose
    // rows, rows W_pitch doubles apart (the padding the shared-memory copy of a tile needs against bank conflicts) -- so a whole tile is
    // ONE contiguous cp.async.bulk in the tensor-pipe Schur complement (solve_large.cuh; one copy per row made the copy engine the
    // bottleneck: 181 issues per tile).  Wp = M rounded up to 32 landmarks; widx() is the only place that knows the layout.
    int Wp, W_sh, W_pitch, W_rows;
};

__device__ __forceinline__ const double *pre_block(const Prob &P, const PreRef &pf) {
    return pf.block >= 0 ? P.pre_blocks + (size_t)pf.block * kPreBlock : P.pre_out + (size_t)(~pf.block) * kPreBlock;
}

template <bool SHARED> struct WLayout { static constexpr int sh = SHARED ? 4 : 5, pitch = SHARED ? 20 : 36; };   // = WStageT<...>::kL / kPitch
__device__ __forceinline__ size_t widx(const Prob &P, int row, int l) {
    return ((size_t)(l >> P.W_sh) * P.W_rows + row) * P.W_pitch + (l & ((1 << P.W_sh) - 1));
}
__host__ __device__ constexpr size_t w_doubles(int n_active, size_t Wp, bool shared) {      // storage of W incl. the row padding
    return (size_t)n_active * 6 * (shared ? Wp / 16 * 20 : Wp / 32 * 36);
}

// One-time per-problem precomputation: tangent bases of all reprojection observations.
static __device__ void prepare_tangents(Prob &P, double *tb) {
    P.rpe_tb = tb;
    for (int o = threadIdx.x; o < P.d.n_rpe; o += ST) {
        const RpeObs &ob = P.rpe[P.d.rpe_off + o];
        const M3 T = local_tangent(ld3(ob.zt));
        double *t = tb + (size_t)o * 6;
        t[0] = T(0, 0); t[1] = T(1, 0); t[2] = T(2, 0); t[3] = T(0, 1); t[4] = T(1, 1); t[5] = T(2, 1);
    }
    __syncthreads();
}

// One-time per problem: the runs of equal (target, reference) in the (sorted) reprojection list.  gs[g] = first observation of
// run g, gs[n_grp] = n_rpe.  Block-wide stream compaction of the run heads (ballot + warp totals).
// (an even number of doubles: the scratch of the next problem must stay 16-byte aligned for the bulk copies of W)
constexpr size_t jb_doubles(int n_rpe) { return (size_t)26 * (((size_t)n_rpe + 31) / 32 * 32) + ((((size_t)n_rpe + 4) / 2 + 3) & ~(size_t)1); }
template <class SM> static __device__ void prepare_groups(Prob &P, SM &s, double *buf) {
    XV_ASSUME_SHARED(&s);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = P.d.n_rpe;
    P.NJ = (n + 31) / 32 * 32;
    P.JB = buf;
    int *gs = reinterpret_cast<int *>(buf + (size_t)26 * P.NJ);
    P.grp_start = gs;
    int run = 0;
    for (int base = 0; base < n; base += ST) {
        const int o = base + tid;
        bool head = false;
        if (o < n) {
            const RpeObs &a = P.rpe[P.d.rpe_off + o];
            head = o == 0;
            if (o > 0) { const RpeObs &b = P.rpe[P.d.rpe_off + o - 1]; head = a.tgt != b.tgt || a.ref != b.ref; }
        }
        const unsigned bal = __ballot_sync(0xffffffffu, head);
        if (lane == 0) s.red[warp] = (double)__popc(bal);
        __syncthreads();
        int woff = 0, tot = 0;
        for (int k = 0; k < NW; ++k) { const int cnt = (int)s.red[k]; if (k < warp) woff += cnt; tot += cnt; }
        if (head) gs[run + woff + __popc(bal & ((1u << lane) - 1u))] = o;
        run += tot;
        __syncthreads();
    }
    if (tid == 0) gs[run] = n;
    P.n_grp = run;
    __syncthreads();
}

// One-time preparation of the prior for a solve: number of leading non-zero rows (stored by the
// marginalisation kernel behind A0; 0 = unknown) and a transposed copy of those rows.
static __device__ void prepare_prior(Prob &P, const double *arena, double *St) {
    const int n = P.marg_n;
    P.marg_St = St;
    if (P.d.marg_k <= 0) { P.marg_rank = 0; return; }
    int rk = (int)arena[(size_t)16 * P.d.marg_k + 2 * (size_t)n * n + n];
    if (rk <= 0 || rk > n) rk = n;
    P.marg_rank = rk;
    for (int e = threadIdx.x; e < rk * n; e += ST) {
        const int r = e / n, c = e % n;
        St[(size_t)c * rk + r] = P.marg_S[(size_t)r * n + c];
    }
    __syncthreads();
}

// ---- warp-uniform fast path of the J^T J accumulation (large windows) ----
// When the 32 observations of a warp share (target, reference) -- the normal case in a 30-keyframe x 2000-landmark window --
// the 90 block entries (tt 21, g_t 6, tr 36, rr 21, g_r 6) are summed by three reduce-scatters of 32 values (31 shuffle rounds
// each, lane k ends up with entry 32 c + k) instead of 90 segmented sums of 5 rounds; every lane then adds its own entries.
__device__ __forceinline__ void tri_decode(int t, int &i, int &j) {
    i = (t >= 15) ? 5 : (t >= 10) ? 4 : (t >= 6) ? 3 : (t >= 3) ? 2 : (t >= 1) ? 1 : 0;
    j = t - i * (i + 1) / 2;
}
__device__ __forceinline__ double jtj_entry(const RpeEval &e, int idx) {
    int i, j;
    if (idx < 21) { tri_decode(idx, i, j); return e.Jt[i] * e.Jt[j] + e.Jt[6 + i] * e.Jt[6 + j]; }
    if (idx < 27) { i = idx - 21; return e.Jt[i] * e.r[0] + e.Jt[6 + i] * e.r[1]; }
    if (idx < 63) { i = (idx - 27) / 6; j = (idx - 27) % 6; return e.Jt[i] * e.Jr[j] + e.Jt[6 + i] * e.Jr[6 + j]; }
    if (idx < 84) { tri_decode(idx - 63, i, j); return e.Jr[i] * e.Jr[j] + e.Jr[6 + i] * e.Jr[6 + j]; }
    if (idx < 90) { i = idx - 84; return e.Jr[i] * e.r[0] + e.Jr[6 + i] * e.r[1]; }
    return 0.0;
}
// v[0] of lane k becomes the sum over all lanes of v[k]
__device__ __forceinline__ double reduce_scatter32(double (&v)[32], int lane) {
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
        const bool upper = (lane & o) != 0;
#pragma unroll
        for (int i = 0; i < o; ++i) {
            const double send = upper ? v[i] : v[i + o];
            const double keep = upper ? v[i + o] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
        }
    }
    return v[0];
}
template <class SM> __device__ __forceinline__ void jtj_uniform_warp(SM &s, const RpeEval &e, int ts, int rs, int lane) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        double v[32];
#pragma unroll
        for (int k = 0; k < 32; ++k) v[k] = jtj_entry(e, 32 * c + k);
        const double sum = reduce_scatter32(v, lane);
        const int idx = 32 * c + lane;
        int i, j;
        if (idx < 21) { tri_decode(idx, i, j); atomicAdd(&s.H[hidx(15 * ts + i, 15 * ts + j)], sum); }
        else if (idx < 27) atomicAdd(&s.g[15 * ts + idx - 21], sum);
        else if (idx < 63) {
            const int a = 15 * ts + (idx - 27) / 6, b = 15 * rs + (idx - 27) % 6;
            atomicAdd(&s.H[a > b ? hidx(a, b) : hidx(b, a)], sum);
        } else if (idx < 84) { tri_decode(idx - 63, i, j); atomicAdd(&s.H[hidx(15 * rs + i, 15 * rs + j)], sum); }
        else if (idx < 90) atomicAdd(&s.g[15 * rs + idx - 84], sum);
    }
}

// ------------------------------------------------------------------------------------------------
// Cost-only evaluation of up to kSpec candidates at once (29 of the 31 evaluations of a window BA).
// After a rejected step Ceres keeps the linearisation and only halves the radius, and a rejection changes neither the iterate nor
// user memory: the candidates of the next rejections are known in advance (radius / 2, / 4, ...).  The loop in backend_solve.cu
// evaluates a batch of them in one pass and consumes the results one iteration at a time, in order, with the unchanged accept /
// reject / tolerance logic -- a candidate behind an accepted one is simply never looked at.  One pass over b candidates costs far
// less than b passes: an observation is loaded once, the serial sections (dogleg scalars, candidate update, reductions, the
// preintegration chains) run side by side, and the prior's S^T is read once for all of them.
//   candidate b: frames cand_frame(s, b, f), inverse depths P.inv_c + b * P.Mp, frame errors of the prior P.m_dx + b * (marg_n + 32)
//   (m_r / m_v / the spare row are free outside the Jacobian evaluation), preintegration residuals P.pre_r + (b * n_pre + f) * 16.
// Warp roles: the last nb warps take the preintegration factors of one candidate each (three lanes per factor: the rotation,
// position and velocity rows are independent serial chains), the others the reprojection and rotation residuals; warps 0 .. nb-1
// first compute the prior's frame errors.  The prior's S dx + infovec follows the barrier, 4 lanes per row.  (Giving the prior
// one warp of its own measured 2.2x slower -- five serial rounds of 135-long L2-latency-bound dot products.)
// Results: s.sp[b * 8 + 6] = cost of candidate b, s.sp[b * 8 + 7] = |x - candidate_b|^2 (its warp sums come from make_candidates
// in s.red4).  Same arithmetic per residual as evaluate(..., want_jac = true) computes.
// ------------------------------------------------------------------------------------------------
#ifdef CUEMU
#define XV_NOINLINE
#else
#define XV_NOINLINE __noinline__          // its own register allocation: inlined into the trust-region loop it spilled 2 KB
#endif
template <class SM> __device__ __forceinline__ const double *cand_frame(const SM &s, int b, int f, int n_active) {
    return b == 0 ? s.cs + f * 16 : (f < n_active ? s.csx + ((b - 1) * SM::kA + f) * 16 : s.xs + f * 16);
}
template <class SM> __device__ __forceinline__ double *cand_frame_w(SM &s, int b, int f) {      // f < n_active for b > 0
    return b == 0 ? s.cs + f * 16 : s.csx + ((b - 1) * SM::kA + f) * 16;
}
__device__ __forceinline__ void acc_add(double (&acc)[kSpec], int b, double v) {
#pragma unroll
    for (int k = 0; k < kSpec; ++k) if (k == b) acc[k] += v;
}

template <class SM> static __device__ XV_NOINLINE void evaluate_cost_multi(const Prob &P, SM &s, int nb, const RigParams &rg, bool robust) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    XV_ASSUME_SHARED(&s);
    const bool has_pre = P.d.n_pre > 0, has_marg = P.d.marg_k > 0;
    const int na = P.d.n_active;
    const int n_rw = NW - (has_pre ? nb : 0);         // warps 0 .. n_rw - 1 take the reprojection residuals
    double acc[kSpec];
#pragma unroll
    for (int k = 0; k < kSpec; ++k) acc[k] = 0.0;
    if (warp >= n_rw) {
        const int b = warp - n_rw;
        double *rb = P.pre_r + (size_t)b * P.d.n_pre * 16;
        for (int t = lane; t < 3 * P.d.n_pre; t += 32) {
            const int f = t / 3;
            const PreRef pf = P.pre[P.d.pre_off + f];
            const double *pb = pre_block(P, pf);
            const double *si = pf.prior ? (s.us + pf.i * 16) : cand_frame(s, b, pf.i, na);
            const double *sj = cand_frame(s, b, pf.j, na);
            pre_eval(si, sj, s.us + pf.i * 16 + 10, pb, rg, false, rb + f * 16, nullptr, nullptr, t % 3);
        }
        __syncwarp();
        double c = 0.0;
        for (int e = lane; e < 15 * P.d.n_pre; e += 32) {          // whitened residual, one lane per row
            const int f = e / 15, r = e % 15;
            const PreRef pf = P.pre[P.d.pre_off + f];
            const double *Sic = pre_block(P, pf) + PB_SIC;
            const double *sc = rb + f * 16;
            double a = 0.0;
#pragma unroll
            for (int k = 0; k < 15; ++k) a += Sic[r * 15 + k] * sc[k];
            c += 0.5 * a * a;
        }
        acc_add(acc, b, c);
    } else {
        if (has_marg && warp < nb)                                  // error of the prior frames against their linearisation points
            for (int k = lane; k < P.d.marg_k; k += 32) {
                const int f = P.marg_frames[P.d.marg_frames_off + k];
                marg_dx(cand_frame(s, warp, f, na), P.marg_lin + k * 16, P.m_dx + (size_t)warp * (P.marg_n + 32) + 15 * k, nullptr);
            }
        for (int base = warp * 32; base < P.d.n_rpe; base += n_rw * 32) {
            const int o = base + lane;
            if (o < P.d.n_rpe) {
                const RpeObs ob = P.rpe[P.d.rpe_off + o];
                const int ts = ob.tgt, rs = ob.ref, lm = ob.lm;
                const bool prior = ob.prior != 0;
                const V3 zt = ld3(ob.zt), zr = ld3(ob.zr);
                const double *tb = P.rpe_tb + (size_t)o * 6;
#pragma unroll 1
                for (int b = 0; b < nb; ++b) {
                    const double *st = cand_frame(s, b, ts, na);
                    const double *sr = prior ? (s.us + rs * 16) : cand_frame(s, b, rs, na);
                    const double idp = prior ? P.invd_user[lm] : P.inv_c[(size_t)b * P.Mp + lm];
                    RpeEval e;
                    rpe_eval(ld4(st), ld3(st + 4), ld4(sr), ld3(sr + 4), idp, zt, zr, tb, rg, false, false, e);
                    const double sq = e.r[0] * e.r[0] + e.r[1] * e.r[1];
                    acc_add(acc, b, robust ? 0.5 * log(1.0 + sq) : 0.5 * sq);
                }
            }
        }
        for (int o = warp * 32 + lane; o < P.d.n_rot; o += n_rw * 32) {
            const RotObs ob = P.rot[P.d.rot_off + o];
#pragma unroll 1
            for (int b = 0; b < nb; ++b) {
                double r[2], Jq[6];
                rot_eval(ld4(cand_frame(s, b, ob.tgt, na)), ld4(s.us + ob.ref * 16), ld3(ob.zt), ld3(ob.zr), rg, false, r, Jq);
                acc_add(acc, b, 0.5 * log(1.0 + (r[0] * r[0] + r[1] * r[1])));
            }
        }
    }
    if (has_marg) {
        __syncthreads();                                            // m_dx complete
        const int n = P.marg_n, rk = P.marg_rank;
        const size_t dxs = P.marg_n + 32;
        // r = S dx + infovec over the non-zero rows on the transposed copy (coalesced): 4 lanes per row (column slices
        // c = q, q + 4, ...); every element of S^T is loaded once for all candidates
        for (int r0 = 0; r0 < rk; r0 += ST / 4) {
            const int r = r0 + (tid >> 2), q = tid & 3;
            double a[kSpec];
#pragma unroll
            for (int b = 0; b < kSpec; ++b) a[b] = 0.0;
            if (r < rk)
#pragma unroll 2
                for (int cidx = q; cidx < n; cidx += 4) {
                    const double sv = P.marg_St[(size_t)cidx * rk + r];
#pragma unroll
                    for (int b = 0; b < kSpec; ++b) if (b < nb) a[b] += sv * P.m_dx[b * dxs + cidx];
                }
#pragma unroll
            for (int b = 0; b < kSpec; ++b) {
                if (b >= nb) continue;                              // uniform
                double t = a[b];
                t += __shfl_xor_sync(0xffffffffu, t, 1);
                t += __shfl_xor_sync(0xffffffffu, t, 2);
                if (r < rk && q == 0) { t += P.marg_info[r]; acc[b] += 0.5 * t * t; }
            }
        }
        // rows beyond the rank are exactly zero in S; their residual is infovec[r] (zero for our priors)
        for (int r = rk + tid; r < n; r += ST) {
            const double t = P.marg_info[r];
#pragma unroll
            for (int b = 0; b < kSpec; ++b) if (b < nb) acc[b] += 0.5 * t * t;
        }
    }
    // one block reduction for the costs and the step norms (whose warp sums make_candidates left in red4); results go to s.sp
#pragma unroll
    for (int k = 0; k < kSpec; ++k) {
        if (k >= nb) continue;
        double v = acc[k];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) s.red4[k * NW + warp] = v;
    }
    __syncthreads();
    if (tid < 2 * kSpec && (tid % kSpec) < nb) {
        double t = 0.0;
        for (int w = 0; w < NW; ++w) t += s.red4[tid * NW + w];
        s.sp[(tid % kSpec) * 8 + 6 + tid / kSpec] = t;
    }
    __syncthreads();
}

// ------------------------------------------------------------------------------------------------
// Residual (+ Jacobian) evaluation at S (= xs or cs) / inv (= inv_x or inv_c).  Returns the cost.
// With want_jac the unscaled H (pose/motion block, packed in shared memory), g, W, h_ll, g_l are
// rebuilt.  Constant blocks and "user memory" reads come from us / invd_user.
// ------------------------------------------------------------------------------------------------
template <class SM> static __device__ double evaluate(const Prob &P, SM &s, const double *S, const double *inv, const RigParams &rg,
                           bool want_jac, bool robust = true, KProf *kp = nullptr) {
    const int kb = 0;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int M = P.M;
    XV_ASSUME_SHARED(&s); XV_ASSUME_SHARED(S); XV_ASSUME_GLOBAL(inv);   // S is s.xs or s.cs, inv is P.inv_x or P.inv_c
    if (want_jac) {
        for (int i = tid; i < SM::HP; i += ST) s.H[i] = 0.0;
        for (int i = tid; i < SM::NPv; i += ST) s.g[i] = 0.0;
        for (int i = tid; i < M; i += ST) { P.h_ll[i] = 0.0; P.g_l[i] = 0.0; }
        const int nw = (int)w_doubles(P.d.n_active, P.Wp, SM::kSharedH);      // including the padding: it enters the tensor-pipe tiles
        for (int i = tid; i < nw; i += ST) P.W[i] = 0.0;
    }
    __syncthreads();
    if (kp) kp->mark(kb + 0);
    double cost = 0.0;
    // Sliding-window instance: the raw evaluation of the preintegration factors (one lane per factor, a serial chain of ~30 k cycles
    // with Jacobians) gets the last warp and runs under the reprojection residuals of the others; their whitening / J^T J products
    // join the per-run tensor-pipe tasks of phase 2 in one warp-task list.  Large instance: the sequential layout further down.
    const bool split_pre = SM::kSharedH && P.d.n_pre > 0;
    const int n_rw = split_pre ? NW - 1 : NW;
    // whitening and J^T J / gradient products of one preintegration factor by one warp (raw residual and Jacobians are in its scratch)
    auto pre_coop = [&](int f, bool whiten_residual) {
        const PreRef pf = P.pre[P.d.pre_off + f];
        const double *pb = pre_block(P, pf);
        double *sc = P.pre_s + (size_t)f * kPreScratch;
        const double *Sic = pb + PB_SIC;
        if (whiten_residual) {                                // (the sequential layout does this with one thread per row before the barrier)
            if (lane < 15) {
                double a = 0.0;
#pragma unroll
                for (int k = 0; k < 15; ++k) a += Sic[lane * 15 + k] * sc[k];
                sc[916 + lane] = a;
                cost += 0.5 * a * a;
            }
            __syncwarp();
        }
        for (int e = lane; e < 450; e += 32) {              // whiten: Jw = sqrt_inv_cov * J
            const int which = e / 225, idx = e % 225, r = idx / 15, c = idx % 15;
            const double *J = sc + (which ? 241 : 16);
            double a = 0.0;
            for (int k = 0; k < 15; ++k) a += Sic[r * 15 + k] * J[k * 15 + c];
            sc[(which ? 691 : 466) + idx] = a;
        }
        __syncwarp();
        const double *Jwi = sc + 466, *Jwj = sc + 691, *rw = sc + 916;
        const bool has_i = !pf.prior && pf.i < P.d.n_active;
        const bool has_j = pf.j < P.d.n_active;
        const int ci = 15 * pf.i, cj = 15 * pf.j;
        // (i,i) and (j,j) lower triangles, cross block, gradients
        for (int e = lane; e < 120 + 120 + 225 + 30; e += 32) {
            if (e < 240) {
                const bool isj = e >= 120;
                if (isj ? !has_j : !has_i) continue;
                int t = isj ? e - 120 : e;
                int r = (int)((sqrt(8.0 * t + 1.0) - 1.0) * 0.5);
                while (r * (r + 1) / 2 > t) --r;
                while ((r + 1) * (r + 2) / 2 <= t) ++r;
                int c = t - r * (r + 1) / 2;
                const double *J = isj ? Jwj : Jwi;
                const int c0 = isj ? cj : ci;
                if (!s.col_free[c0 + r] || !s.col_free[c0 + c]) continue;
                double a = 0.0;
                for (int k = 0; k < 15; ++k) a += J[k * 15 + r] * J[k * 15 + c];
                atomicAdd(&s.H[hidx(c0 + r, c0 + c)], a);
            } else if (e < 465) {
                if (!has_i || !has_j) continue;
                int t = e - 240, r = t / 15, c = t % 15;     // r: column of Jwi, c: column of Jwj
                if (!s.col_free[ci + r] || !s.col_free[cj + c]) continue;
                double a = 0.0;
                for (int k = 0; k < 15; ++k) a += Jwi[k * 15 + r] * Jwj[k * 15 + c];
                const int A = ci + r, B = cj + c;
                atomicAdd(&s.H[A > B ? hidx(A, B) : hidx(B, A)], a);
            } else {
                int t = e - 465;
                const bool isj = t >= 15;
                if (isj ? !has_j : !has_i) continue;
                int c = isj ? t - 15 : t;
                const double *J = isj ? Jwj : Jwi;
                const int c0 = isj ? cj : ci;
                if (!s.col_free[c0 + c]) continue;
                double a = 0.0;
                for (int k = 0; k < 15; ++k) a += J[k * 15 + c] * rw[k];
                atomicAdd(&s.g[c0 + c], a);
            }
        }
        };
    if (split_pre && warp == NW - 1)
        for (int f = lane; f < P.d.n_pre; f += 32) {
            const PreRef pf = P.pre[P.d.pre_off + f];
            const double *pb = pre_block(P, pf);
            double *sc = P.pre_s + (size_t)f * kPreScratch;
            const double *si = pf.prior ? (s.us + pf.i * 16) : (S + pf.i * 16);
            const double *sj = S + pf.j * 16;
            pre_eval(si, sj, s.us + pf.i * 16 + 10, pb, rg, true, sc, sc + 16, sc + 241);
        }
    // ---------------- reprojection residuals (Cauchy loss) ----------------
    for (int base = warp * 32; warp < n_rw && base < P.d.n_rpe; base += n_rw * 32) {
        const int o = base + lane;
        const bool valid = o < P.d.n_rpe;
        RpeEval e;
        int ts = 0, rs = 0, lm = 0;
        bool prior = true, tf = false, rf = false, lf = false;
        double w = 0.0;
        if (valid) {
            const RpeObs ob = P.rpe[P.d.rpe_off + o];
            ts = ob.tgt; rs = ob.ref; lm = ob.lm; prior = ob.prior != 0;
            const double *st = S + ts * 16;
            const double *sr = prior ? (s.us + rs * 16) : (S + rs * 16);
            const double idp = prior ? P.invd_user[lm] : inv[lm];
            tf = ts < P.d.n_active && s.pose_free[ts];
            rf = !prior && rs < P.d.n_active && s.pose_free[rs];
            lf = !prior && lm < M;
            rpe_eval(ld4(st), ld3(st + 4), ld4(sr), ld3(sr + 4), idp, ld3(ob.zt), ld3(ob.zr), P.rpe_tb + (size_t)o * 6, rg, want_jac,
                     want_jac && !prior, e);
            const double sq = e.r[0] * e.r[0] + e.r[1] * e.r[1];
            if (robust) {
                cost += 0.5 * log(1.0 + sq);                // CauchyLoss(1): rho(s) = log(1 + s)
                w = sqrt(1.0 / (1.0 + sq));                 // corrector: rho'' < 0 -> scale by sqrt(rho')
            } else { cost += 0.5 * sq; w = 1.0; }
        }
        if (!want_jac) continue;                            // warp-uniform
        if (valid) {
            e.r[0] *= w; e.r[1] *= w;
            for (int k = 0; k < 12; ++k) e.Jt[k] *= w;
            if (!prior) { for (int k = 0; k < 12; ++k) e.Jr[k] *= w; e.Jl[0] *= w; e.Jl[1] *= w; }
        }
        if constexpr (SM::kSharedH) {
            // ---- sliding-window instance, phase 1: park the weighted rows (zeroed where the block is constant) and finish the
            // landmark column; the (target, reference) blocks are summed per run of observations in phase 2 below ----
            if (valid) {
                double *jb = P.JB + o;
                const size_t NJ = (size_t)P.NJ;
#pragma unroll
                for (int k = 0; k < 6; ++k) {
                    jb[(size_t)k * NJ] = tf ? e.Jt[k] : 0.0;
                    jb[(size_t)(6 + k) * NJ] = rf ? e.Jr[k] : 0.0;
                    jb[(size_t)(12 + k) * NJ] = tf ? e.Jt[6 + k] : 0.0;
                    jb[(size_t)(18 + k) * NJ] = rf ? e.Jr[6 + k] : 0.0;
                }
                jb[(size_t)24 * NJ] = e.r[0];
                jb[(size_t)25 * NJ] = e.r[1];
                if (lf) {
                    atomicAdd(&P.h_ll[lm], e.Jl[0] * e.Jl[0] + e.Jl[1] * e.Jl[1]);
                    atomicAdd(&P.g_l[lm], e.Jl[0] * e.r[0] + e.Jl[1] * e.r[1]);
                    if (tf)
                        for (int k = 0; k < 6; ++k) P.W[widx(P, ts * 6 + k, lm)] = e.Jl[0] * e.Jt[k] + e.Jl[1] * e.Jt[6 + k];
                    if (rf)
                        for (int k = 0; k < 6; ++k) atomicAdd(&P.W[widx(P, rs * 6 + k, lm)], e.Jl[0] * e.Jr[k] + e.Jl[1] * e.Jr[6 + k]);
                }
            }
            continue;
        }
        bool fast = false;
        if constexpr (!SM::kSharedH) {
            const int ts0 = __shfl_sync(0xffffffffu, ts, 0), rs0 = __shfl_sync(0xffffffffu, rs, 0);
            fast = __all_sync(0xffffffffu, valid && !prior && tf && rf && ts == ts0 && rs == rs0 && ts != rs) != 0;
            if (fast) jtj_uniform_warp(s, e, ts, rs, lane);
        }
        // --- (tgt,tgt) block + g_tgt : segments keyed by tgt ---
        if (!fast) {
            const int key = (valid && tf) ? ts : -1 - lane;
            const int key_prev = __shfl_up_sync(0xffffffffu, key, 1);   // unconditional: every lane must take part
            const bool head = (lane == 0) || (key_prev != key);
            const bool act = valid && tf;
            const unsigned take = seg_mask(key, lane);
#pragma unroll
            for (int i = 0; i < 6; ++i) {
#pragma unroll
                for (int j = 0; j <= i; ++j) {
                    double v = act ? (e.Jt[i] * e.Jt[j] + e.Jt[6 + i] * e.Jt[6 + j]) : 0.0;
                    v = seg_sum(v, take);
                    if (act && head) atomicAdd(&s.H[hidx(15 * ts + i, 15 * ts + j)], v);
                }
                double gv = act ? (e.Jt[i] * e.r[0] + e.Jt[6 + i] * e.r[1]) : 0.0;
                gv = seg_sum(gv, take);
                if (act && head) atomicAdd(&s.g[15 * ts + i], gv);
            }
        }
        // --- (tgt,ref), (ref,ref) blocks + g_ref : segments keyed by (tgt, ref) ---
        const unsigned any_full = __ballot_sync(0xffffffffu, valid && !prior);
        if (any_full) {
            const int key = (valid && !prior) ? ((ts << 8) | rs) : -1 - lane;
            const int key_prev = __shfl_up_sync(0xffffffffu, key, 1);   // unconditional: every lane must take part
            const bool head = (lane == 0) || (key_prev != key);
            const bool act_tr = valid && tf && rf, act_rr = valid && rf;
            const unsigned take = fast ? 0u : seg_mask(key, lane);
            if (!fast) {
#pragma unroll
            for (int i = 0; i < 6; ++i) {
#pragma unroll
                for (int j = 0; j < 6; ++j) {
                    double v = act_tr ? (e.Jt[i] * e.Jr[j] + e.Jt[6 + i] * e.Jr[6 + j]) : 0.0;
                    v = seg_sum(v, take);
                    if (act_tr && head) {
                        const int a = 15 * ts + i, b = 15 * rs + j;
                        atomicAdd(&s.H[a > b ? hidx(a, b) : hidx(b, a)], v);
                    }
                }
#pragma unroll
                for (int j = 0; j <= i; ++j) {
                    double v = act_rr ? (e.Jr[i] * e.Jr[j] + e.Jr[6 + i] * e.Jr[6 + j]) : 0.0;
                    v = seg_sum(v, take);
                    if (act_rr && head) atomicAdd(&s.H[hidx(15 * rs + i, 15 * rs + j)], v);
                }
                double gv = act_rr ? (e.Jr[i] * e.r[0] + e.Jr[6 + i] * e.r[1]) : 0.0;
                gv = seg_sum(gv, take);
                if (act_rr && head) atomicAdd(&s.g[15 * rs + i], gv);
            }
            }
            // landmark column: h_ll, g_l, W (W[slot][k][l], l fastest)
            if (valid && lf) {
                atomicAdd(&P.h_ll[lm], e.Jl[0] * e.Jl[0] + e.Jl[1] * e.Jl[1]);
                atomicAdd(&P.g_l[lm], e.Jl[0] * e.r[0] + e.Jl[1] * e.r[1]);
                if (tf)
                    for (int k = 0; k < 6; ++k) P.W[widx(P, ts * 6 + k, lm)] = e.Jl[0] * e.Jt[k] + e.Jl[1] * e.Jt[6 + k];
                if (rf)
                    for (int k = 0; k < 6; ++k) atomicAdd(&P.W[widx(P, rs * 6 + k, lm)], e.Jl[0] * e.Jr[k] + e.Jl[1] * e.Jr[6 + k]);
            }
        }
    }
    if constexpr (SM::kSharedH) {
        // ---- phase 2: one WARP per run of observations with equal (target, reference).  With X = [Jt | Jr | r] (13 columns, padded
        // to 16) over the 2 n residual rows of the run, X^T X holds everything the run contributes: the 12 x 12 pose blocks and, in
        // row 12, the gradient entries.  It is a true contraction over the residual rows and runs on the FP64 tensor pipe
        // (mma.sync.m8n8k4: A and B fragments are the same two loads per lane and k-step); three 8 x 8 tiles cover the lower
        // triangle.  No shuffles, one shared-memory atomic per entry and run, and a summation order that does not depend on how
        // the observations fall into warps.  (Measured on a B200, kcycles per Jacobian evaluation of a window problem: segmented
        // warp sums 141, one thread per (run, row) with scalar loads from L2 588.) ----
        __syncthreads();
        const size_t NJ = (size_t)P.NJ;
        const int g8 = lane >> 2, t4 = lane & 3;
        const int n_pre_tasks = split_pre ? P.d.n_pre : 0;       // the long tasks first
        for (int task = warp; task < n_pre_tasks + P.n_grp; task += NW) {
            if (task < n_pre_tasks) { pre_coop(task, true); continue; }
            const int grp = task - n_pre_tasks;
            const int o0 = P.grp_start[grp], o1 = P.grp_start[grp + 1];
            const RpeObs &ob0 = P.rpe[P.d.rpe_off + o0];
            const int ts = ob0.tgt, rs = ob0.ref;
            const bool t_act = ts < P.d.n_active && s.pose_free[ts];
            const bool r_act = rs < P.d.n_active && s.pose_free[rs];
            if (!t_act && !r_act) continue;                 // warp-uniform
            double d00[2] = {0.0, 0.0}, d10[2] = {0.0, 0.0}, d11[2] = {0.0, 0.0};
            // column g8 of X is a Jacobian column (< 12); column 8 + g8 is a Jacobian column for g8 < 4, r for g8 == 4, padding beyond
            for (int row = 0; row < 2; ++row) {
                const double *x0 = P.JB + (size_t)(row * 12 + g8) * NJ;
                const double *x1 = g8 < 4 ? P.JB + (size_t)(row * 12 + 8 + g8) * NJ : P.JB + (size_t)(24 + row) * NJ;
                for (int ob = o0; ob < o1; ob += 16) {
                    double a0[4], a1[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {                 // four k-steps: the eight loads are issued before the tensor-pipe work
                        const int o = ob + 4 * u + t4;
                        a0[u] = o < o1 ? x0[o] : 0.0;
                        a1[u] = (o < o1 && g8 <= 4) ? x1[o] : 0.0;
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        dmma_m8n8k4(d00[0], d00[1], a0[u], a0[u]);
                        dmma_m8n8k4(d10[0], d10[1], a1[u], a0[u]);
                        dmma_m8n8k4(d11[0], d11[1], a1[u], a1[u]);
                    }
                }
            }
            // lane (g8, t4) holds entries (8 ti + g8, 8 tj + 2 t4 + i) of tile (ti, tj)
#pragma unroll
            for (int tile = 0; tile < 3; ++tile) {
                const int ti = tile == 0 ? 0 : 1, tj = tile == 2 ? 1 : 0;
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const int R = 8 * ti + g8, C = 8 * tj + 2 * t4 + i;
                    const double v = tile == 0 ? d00[i] : (tile == 1 ? d10[i] : d11[i]);
                    if (C > R || R > 12 || C >= 12) continue;
                    if (R == 12) {                                  // gradient row
                        if (C < 6) { if (t_act) atomicAdd(&s.g[15 * ts + C], v); }
                        else if (r_act) atomicAdd(&s.g[15 * rs + C - 6], v);
                    } else if (R < 6) {
                        if (t_act) atomicAdd(&s.H[hidx(15 * ts + R, 15 * ts + C)], v);
                    } else if (C < 6) {
                        if (t_act && r_act) {
                            const int a = 15 * ts + C, b = 15 * rs + R - 6;
                            atomicAdd(&s.H[a > b ? hidx(a, b) : hidx(b, a)], v);
                        }
                    } else if (r_act) atomicAdd(&s.H[hidx(15 * rs + R - 6, 15 * rs + C - 6)], v);
                }
            }
        }
    }
    if (kp) kp->mark(kb + 1);
    // ---------------- rotation priors (Cauchy loss), rare ----------------
    for (int o = tid; o < P.d.n_rot; o += ST) {
        const RotObs ob = P.rot[P.d.rot_off + o];
        const int ts = ob.tgt;
        double r[2], Jq[6];
        rot_eval(ld4(S + ts * 16), ld4(s.us + ob.ref * 16), ld3(ob.zt), ld3(ob.zr), rg, want_jac, r, Jq);
        const double sq = r[0] * r[0] + r[1] * r[1];
        cost += 0.5 * log(1.0 + sq);
        if (want_jac && ts < P.d.n_active && s.pose_free[ts]) {
            const double w = sqrt(1.0 / (1.0 + sq));
            for (int i = 0; i < 3; ++i) {
                for (int j = 0; j <= i; ++j)
                    atomicAdd(&s.H[hidx(15 * ts + i, 15 * ts + j)], w * w * (Jq[i] * Jq[j] + Jq[3 + i] * Jq[3 + j]));
                atomicAdd(&s.g[15 * ts + i], w * w * (Jq[i] * r[0] + Jq[3 + i] * r[1]));
            }
        }
    }
    if (kp) kp->mark(kb + 2);
    // ---------------- preintegration factors (no loss) ----------------
    // residuals (and raw Jacobians) of all factors are evaluated SIMT by the first n_pre threads; the
    // whitening and J^T J products of the Jacobian pass are then done cooperatively, one warp per factor
    if (!split_pre && tid < P.d.n_pre) {
        const int f = tid;
        const PreRef pf = P.pre[P.d.pre_off + f];
        const double *pb = pre_block(P, pf);
        double *sc = P.pre_s + (size_t)f * kPreScratch;     // [0,15) r  [16,241) Ji  [241,466) Jj  [466,691) Jwi [691,916) Jwj [916,931) rw
        const double *si = pf.prior ? (s.us + pf.i * 16) : (S + pf.i * 16);
        const double *sj = S + pf.j * 16;
        pre_eval(si, sj, s.us + pf.i * 16 + 10, pb, rg, want_jac, sc, sc + 16, sc + 241);
    }
    __syncthreads();
    if (!split_pre && tid < 15 * P.d.n_pre) {          // whitened residual, one thread per row
        const int f = tid / 15, r = tid % 15;
        const PreRef pf = P.pre[P.d.pre_off + f];
        const double *Sic = pre_block(P, pf) + PB_SIC;
        double *sc = P.pre_s + (size_t)f * kPreScratch;
        double a = 0.0;
#pragma unroll
        for (int k = 0; k < 15; ++k) a += Sic[r * 15 + k] * sc[k];
        sc[916 + r] = a;
        cost += 0.5 * a * a;
    }
    if (kp) kp->mark(kb + 3);
    if (want_jac) __syncthreads();
    if (!split_pre)
        for (int f = warp; f < P.d.n_pre; f += NW) pre_coop(f, false);
    __syncthreads();
    if (kp) kp->mark(kb + 4);
    // ---------------- marginalisation prior (no loss) ----------------
    if (P.d.marg_k > 0) {
        const int K = P.d.marg_k, n = P.marg_n;
        if (tid < K) {
            const int f = P.marg_frames[P.d.marg_frames_off + tid];
            M3 Jri;
            marg_dx(S + f * 16, P.marg_lin + tid * 16, P.m_dx + 15 * tid, want_jac ? &Jri : nullptr);
            if (want_jac)
                for (int k = 0; k < 9; ++k) s.jrinv[tid * 9 + k] = Jri.m[k];
        }
        __syncthreads();
        {   // r = S dx + infovec over the non-zero rows: one thread per row on the transposed copy (coalesced)
            const int rk = P.marg_rank;
            // 4 lanes per row (column slices c = q, q+4, ...), two independent accumulators per lane
            for (int r0 = 0; r0 < rk; r0 += ST / 4) {
                const int r = r0 + (tid >> 2), q = tid & 3;
                double a0 = 0.0, a1 = 0.0;
                if (r < rk) {
                    int cidx = q;
                    for (; cidx + 4 < n; cidx += 8) {
                        a0 += P.marg_St[(size_t)cidx * rk + r] * P.m_dx[cidx];
                        a1 += P.marg_St[(size_t)(cidx + 4) * rk + r] * P.m_dx[cidx + 4];
                    }
                    for (; cidx < n; cidx += 4) a0 += P.marg_St[(size_t)cidx * rk + r] * P.m_dx[cidx];
                }
                double a = a0 + a1;
                a += __shfl_xor_sync(0xffffffffu, a, 1);
                a += __shfl_xor_sync(0xffffffffu, a, 2);
                if (r < rk && q == 0) {
                    a += P.marg_info[r];
                    P.m_r[r] = a;
                    cost += 0.5 * a * a;
                }
            }
            // rows beyond the rank are exactly zero in S; their residual is infovec[r] (zero for our priors)
            for (int r = rk + tid; r < n; r += ST) { const double a = P.marg_info[r]; P.m_r[r] = a; cost += 0.5 * a * a; }
        }
        if (want_jac) {
            __syncthreads();
            {   // v = S^T r on the transposed copy: one warp per column, lanes over the (contiguous) non-zero rows
                const int rk = P.marg_rank;
                for (int c = warp; c < n; c += NW) {
                    double a = 0.0;
                    for (int r = lane; r < rk; r += 32) a += P.marg_St[(size_t)c * rk + r] * P.m_r[r];
                    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
                    if (lane == 0) P.m_v[c] = a;
                }
            }
            // H += B^T A0 B (lower triangle) with B = blockdiag(Jr^-1 on the q blocks, I): every 15-block is five 3-blocks, so one
            // thread per 3 x 3 block of A0 applies B on both sides in registers -- one pass over A0, no intermediate product
            {
                const int nb3 = n / 3;
                for (int e = tid; e < nb3 * (nb3 + 1) / 2; e += ST) {
                    int I = (int)((sqrtf(8.0f * (float)e + 1.0f) - 1.0f) * 0.5f);
                    while (I * (I + 1) / 2 > e) --I;
                    while ((I + 1) * (I + 2) / 2 <= e) ++I;
                    const int J = e - I * (I + 1) / 2;
                    const int ki = I / 5, kj = J / 5;
                    const int fi = P.marg_frames[P.d.marg_frames_off + ki], fj = P.marg_frames[P.d.marg_frames_off + kj];
                    if (fi >= P.d.n_active || fj >= P.d.n_active) continue;
                    double t[3][3];
#pragma unroll
                    for (int r = 0; r < 3; ++r) {
                        const double *a0 = P.marg_A0 + (size_t)(3 * I + r) * n + 3 * J;
                        if (J % 5 == 0) {
                            const double *Jr = s.jrinv + 9 * kj;
#pragma unroll
                            for (int c = 0; c < 3; ++c) t[r][c] = a0[0] * Jr[0 * 3 + c] + a0[1] * Jr[1 * 3 + c] + a0[2] * Jr[2 * 3 + c];
                        } else { t[r][0] = a0[0]; t[r][1] = a0[1]; t[r][2] = a0[2]; }
                    }
                    const bool roti = I % 5 == 0;
                    const double *Ji = s.jrinv + 9 * ki;
#pragma unroll
                    for (int oi = 0; oi < 3; ++oi)
#pragma unroll
                        for (int oj = 0; oj < 3; ++oj) {
                            if (I == J && oj > oi) continue;
                            const double a = roti ? Ji[0 * 3 + oi] * t[0][oj] + Ji[1 * 3 + oi] * t[1][oj] + Ji[2 * 3 + oi] * t[2][oj] : t[oi][oj];
                            const int A = 15 * fi + 3 * (I % 5) + oi, B = 15 * fj + 3 * (J % 5) + oj;
                            if (!s.col_free[A] || !s.col_free[B]) continue;
                            s.H[A >= B ? hidx(A, B) : hidx(B, A)] += a;     // distinct (i,j) map to distinct (A,B)
                        }
                }
            }
            __syncthreads();                                        // m_v complete
            for (int c = tid; c < n; c += ST) {
                const int k = c / 15, oc = c % 15;
                double a;
                if (oc < 3) {
                    const double *Jr = s.jrinv + 9 * k;
                    a = Jr[0 * 3 + oc] * P.m_v[15 * k] + Jr[1 * 3 + oc] * P.m_v[15 * k + 1] + Jr[2 * 3 + oc] * P.m_v[15 * k + 2];
                } else a = P.m_v[c];
                const int f = P.marg_frames[P.d.marg_frames_off + k];
                if (f < P.d.n_active && s.col_free[15 * f + oc]) s.g[15 * f + oc] += a;
            }
        }
    }
    if (kp) kp->mark(kb + 5);
    cost = block_sum(cost, s);
    if (kp) kp->mark(kb + 6);
    if (want_jac) {
        // constant columns: unit diagonal, zero gradient (their off-diagonals were never accumulated)
        for (int c = tid; c < P.NPc; c += ST)
            if (!s.col_free[c]) { s.H[hidx(c, c)] = 1.0; s.g[c] = 0.0; }
        __syncthreads();
    }
    return cost;
}


// Eliminate the (scalar) inverse-depth blocks: H_pp -= sum_l w_l w_l^T / ht_l and
// rhs_p -= sum_l w_l g_l / ht_l, with ht_l = P.ht_l (landmarks with a non-finite 1/ht_l are skipped,
// marginalization_factor.h:384-386).  One warp per pair of frame slots, lanes over landmarks.
template <class SM> static __device__ void schur_landmarks(const Prob &P, SM &s, double *rhs) {
    XV_ASSUME_SHARED(&s); XV_ASSUME_SHARED(rhs);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int M = P.M, na = P.d.n_active;
    (void)tid;
    if (M > 0) {
        // S -= sum_l w_l w_l^T / ht_l ; one warp per pair of frame slots, lanes over landmarks
        const int npairs = na * (na + 1) / 2;
        for (int pr = warp; pr < npairs; pr += NW) {
            int a = (int)((sqrt(8.0 * pr + 1.0) - 1.0) * 0.5);
            while (a * (a + 1) / 2 > pr) --a;
            while ((a + 1) * (a + 2) / 2 <= pr) ++a;
            const int b = pr - a * (a + 1) / 2;        // a >= b
            if (!s.pose_free[a] || !s.pose_free[b]) continue;
            double acc[36];
#pragma unroll
            for (int k = 0; k < 36; ++k) acc[k] = 0.0;
            for (int l = lane; l < M; l += 32) {
                double ih = 1.0 / P.ht_l[l];
                if (!isfinite(ih)) ih = 0.0;
                double wa[6], wb[6];
#pragma unroll
                for (int k = 0; k < 6; ++k) { wa[k] = P.W[widx(P, a * 6 + k, l)]; wb[k] = P.W[widx(P, b * 6 + k, l)] * ih; }
#pragma unroll
                for (int i = 0; i < 6; ++i)
#pragma unroll
                    for (int j = 0; j < 6; ++j) acc[i * 6 + j] += wa[i] * wb[j];
            }
#pragma unroll
            for (int k = 0; k < 36; ++k) {
                double v = acc[k];
                for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                acc[k] = v;
            }
            if (lane == 0) {
                for (int i = 0; i < 6; ++i)
                    for (int j = 0; j < 6; ++j) {
                        if (a == b && j > i) continue;
                        s.H[hidx(15 * a + i, 15 * b + j)] -= acc[i * 6 + j];
                    }
            }
        }
        // b_red = gs_p - sum_l w_l gs_l / ht_l
        for (int c = warp; c < na * 6; c += NW) {
            const int a = c / 6, k = c % 6;
            if (!s.pose_free[a]) continue;
            double v = 0.0;
            for (int l = lane; l < M; l += 32) { double ih = 1.0 / P.ht_l[l]; if (!isfinite(ih)) ih = 0.0; v += P.W[widx(P, c, l)] * P.g_l[l] * ih; }
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0) rhs[15 * a + k] -= v;
        }
    }
}

}  // namespace xvbe
