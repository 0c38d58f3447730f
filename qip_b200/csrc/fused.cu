// qip_b200/csrc/fused.cu -- fused multi-gate pass: one HBM round trip for a whole list of gates.
//
// The reference sweeps the full state once per op (qip/ext/kronprod.pyx:157-197; QFFT on n qubits
// is n(n+1)/2 + n/2 sweeps, qip/qfft.py:33-39).  Here the state is cut into tiles of 2^TB
// amplitudes spanned by an arbitrary set of TB index bits (the lowest L of them contiguous, so
// every global access is a run of 2^L amplitudes = 2^L * 16 B for complex128).  A CTA stages one
// tile in shared memory (64 KiB at TB = 12, complex128), runs the gate list on it and writes it
// back in place.  Legal in one pass:
//   * dense 1- and 2-qubit gates whose target bits are tile bits,
//   * diagonal gates and phase gates on ANY bits (bits outside the tile are constant per tile and
//     only select which diagonal entry applies),
//   * control bits anywhere (outside the tile they switch the gate on or off per tile).
//   * runs of >= 3 consecutive diagonal gates are folded on the host into a "stage": per-cell phase
//     tables (tile-lo, tile-hi, groups of <= 7 outside bits) applied in one sweep (a QFT stage is
//     ~n controlled phases).
// Roofline: 2 * sizeof(amp) * 2^nbits bytes per pass against HBM; the tile pipeline alone runs at the copy
// roofline (6.1 TB/s measured), a sweep costs >= 2^TB * 32 B of shared-memory traffic (1024 cycles per tile at
// 128 B/clk) or, for general complex 4x4 blocks, 128 FP64 issue cycles per group; the first two dense sweeps of
// a pass hide behind the tile traffic, each further one adds ~0.25 of a sweep of HBM time.
#include <stdlib.h>
#include <complex>
#include <vector>
#include "common.cuh"
#include "../../include/qip_b200.h"
#include "fused_shared.cuh"

namespace qipb {

// Block pairs (sweep_pair2: two disjoint dense 2-qubit blocks per sweep on 16-amplitude register groups) are compiled
// out by default.  Measured on B200 (profiles/r02_wide_pairs.md): with both phases inlined the WIDE kernel needs more than
// its 168 registers, ptxas then drops the warp-uniform descriptor loads of EVERY sweep of the kernel (coefficients come
// through indexed LDC into vector registers instead of LDCU / UR operands) and spills; the pairs won 3 % on layered passes
// and the kernel as a whole lost 5 %.  -DQIPB_ENABLE_PAIRS=1 builds them (tests/test_fused_emul.py covers the arithmetic).
#ifndef QIPB_ENABLE_PAIRS
#define QIPB_ENABLE_PAIRS 0
#endif
#ifndef QIPB_ENABLE_TRIOS
#define QIPB_ENABLE_TRIOS 1     // sweep_trio: a dense 2-qubit block + a lone dense 1-qubit gate per sweep (fits the register budget)
#endif
#ifndef QIPB_WIDE_MINB_C64
#define QIPB_WIDE_MINB_C64 6  /* complex64 WIDE kernel: 32 KiB tiles, SIX CTAs per SM fit the shared memory -- keep the registers at <= 80 so
                                 that they fit the register file too (at 110 registers only four did: layered 34 q 282 -> 350 ms) */
#endif
#ifndef QIPB_WIDE_MINB
#define QIPB_WIDE_MINB 3      /* complex128 WIDE kernel: 3 CTAs per SM, cap 168 registers (SLIM instantiation: 4, cap 128) */
#endif
#define FUSED_THREADS 256
#define FUSED_MAX_INS 12
#define FUSED_MAX_OPS 96          // device ops per launch (a pass with more is split into several launches)

struct DevGate {
    unsigned char k;        // target bits in total (0..2)
    unsigned char kin;      // how many of them are tile bits
    unsigned char diag;     // 1: diagonal gate; 2: stage (StageInfo overlays m)
    unsigned char nins;     // fixed tile-local positions (in-tile targets + in-tile controls)
    unsigned char tl[2];    // target j (matrix order, 0 = MSB): tile-local position, 0xFF if outside
    unsigned char tg[2];    // target j: position in the state index
    u32 nmask[FUSED_MAX_INS];   // ~((1 << p) - 1) for the fixed positions p, ascending
    u32 in_or;              // tile-local mask of in-tile control bits
    unsigned char post;     // dense 1-qubit gate without controls: the NEXT op is a stage applied in the same sweep
    unsigned char mk;       // dense 2-qubit block: MK_GENERAL / MK_REAL / MK_REALPHASE / MK_MONOMIAL (coefficient layout below)
    unsigned char phmask;   // MK_REALPHASE: columns with a non-trivial phase; MK_MONOMIAL: columns whose coefficient is not 1
    unsigned char perm;     // MK_MONOMIAL: bits 2j..2j+1 = the row that column j maps to
    unsigned char pair;     // post == 4 (leader of a block pair, WIDE kernel; the partner is the NEXT op, post == 5): the
                            // pair's four sorted expansion masks are nmask[2..5]
    u64 out_ctrl;           // state-index mask of controls outside the tile
    double2 m[16];
};

static_assert(sizeof(StageInfo) <= sizeof(double2) * 16, "StageInfo must fit over DevGate::m");

struct FusedArgs {
    int nbits, tb, ngates, lowrun;      // lowrun = number of contiguous low tile bits (0..L-1)
    int prefetch, nstages;              // prefetch: pull the CTA's next tile into L2 ahead of time; nstages: ops with tables
    u64 ntiles;
    const double2 *tables;              // stage tables (global memory)
    unsigned char tbit[16];             // tile-local bit -> state bit, ascending
    // Tile enumeration.  A launch may be restricted to a CHUNK of the state: the amplitudes whose `fix` bits (state-index
    // bits outside the tile, at most 4) have a given value (qipb_apply_fused_chunk; the sharded engine pipelines a pass
    // chunk by chunk against the NVLink exchange of the neighbouring chunk).  ebit = tile bits and fix bits merged,
    // ascending: tile number t -> base index by inserting a zero at each of them, then OR fix_value.  To the gates a
    // fix bit is an ordinary outside bit (controls, diagonal targets and stage cells read it from the base index).
    int nexp;
    unsigned char ebit[20];
    u64 fix_value;
    // Dynamic tile scheduling: {next tile, CTAs done} in device memory (null: tile t = blockIdx.x + k * gridDim.x).  The
    // CTAs of a launch are persistent; when some of them cannot be resident from the start -- a remap of the chunk
    // pipeline or an NCCL kernel holds part of an SM -- a static split makes the late CTAs a second wave that doubles the
    // launch (measured: 13 ms instead of 5.5 per chunk pass beside the remap); with a shared counter they just take
    // fewer tiles.
    unsigned int *sched;
    DevGate g[FUSED_MAX_OPS];
};
static_assert(sizeof(FusedArgs) <= 32764, "FusedArgs must fit in the kernel parameter space");

// base index of tile t (all tile-local bits 0)
QIPB_HD u64 fused_tile_base(const FusedArgs &f, u64 t) {
    u64 base = t;
    for (int j = 0; j < f.nexp; ++j) base = insert_zero(base, f.ebit[j]);
    return base | f.fix_value;
}

// ---- device side --------------------------------------------------------------------------------
// Every gate is one sweep over the tile in shared memory.  The sweeps are specialised on the number
// of fixed tile-local positions (NINS: in-tile targets + in-tile controls), so that the index
// expansion is a handful of ALU instructions on masks held in (uniform) registers, and the matrix
// coefficients are read from the kernel-parameter bank with warp-uniform addresses: the compiler
// keeps them in uniform registers / feeds them to DFMA directly, so the inner loops are
// LDS + DFMA + STS and nothing else (the first version re-loaded 24 coefficients per group with
// indexed LDC and was issue bound at ~0.37 of the FP64 / shared-memory limit).

// Sweep loops: `count` work items over the CTA's threads.  UNI (count is a multiple of NT, the number of threads sweeping the tile)
// gives the loop a warp-uniform trip count; with a possibly divergent exit condition ptxas stops
// treating the descriptor index as uniform and falls back to indexed LDC loads inside the loops.
#define QIPB_SWEEP(var, count) \
    for (u32 it_ = 0, nit_ = UNI ? (count) / NT : ((count) + NT - 1) / NT, var = tid; \
         it_ < nit_ && (UNI || var < (count)); ++it_, var += NT)

// matrix coefficient in the amplitude's precision
template <typename A> struct Cf { typename amp_traits<A>::real x, y; };
// coefficient i of a gate descriptor, stored by the host in the amplitude's precision (complex64 launches
// hold float2 over the same bytes: a per-use F2F of a warp-uniform double costs more than the FFMA it feeds)
template <typename A> QIPB_HD Cf<A> coef(const DevGate &g, int i);
template <> QIPB_HD Cf<double2> coef<double2>(const DevGate &g, int i) {
    Cf<double2> c;
    c.x = g.m[i].x;
    c.y = g.m[i].y;
    return c;
}
template <> QIPB_HD Cf<float2> coef<float2>(const DevGate &g, int i) {
    const float2 v = reinterpret_cast<const float2 *>(g.m)[i];
    Cf<float2> c;
    c.x = v.x;
    c.y = v.y;
    return c;
}
template <typename A> QIPB_HD A cmulc(const Cf<A> m, const A a) {
    A r;
    r.x = m.x * a.x - m.y * a.y;
    r.y = m.x * a.y + m.y * a.x;
    return r;
}
template <typename A> QIPB_HD void cfmac(A &acc, const Cf<A> m, const A a) {
    acc.x = fma(m.x, a.x, acc.x);
    acc.x = fma(-m.y, a.y, acc.x);
    acc.y = fma(m.x, a.y, acc.y);
    acc.y = fma(m.y, a.x, acc.y);
}

// insert a zero bit at each of NINS fixed positions (masks ascending), then set the control bits
template <int NINS> struct Expand {
    u32 nm[NINS > 0 ? NINS : 1];
    u32 ior;
    QIPB_HD explicit Expand(const DevGate &g) {
#pragma unroll
        for (int q = 0; q < NINS; ++q) nm[q] = g.nmask[q];
        ior = g.in_or;
    }
    QIPB_HD u32 operator()(u32 w) const {
#pragma unroll
        for (int q = 0; q < NINS; ++q) w += (w & nm[q]);
        return w | ior;
    }
};
// any number of fixed positions (rare shapes: more than four in-tile controls)
struct ExpandAny {
    const DevGate &g;
    QIPB_HD explicit ExpandAny(const DevGate &g_) : g(g_) {}
    QIPB_HD u32 operator()(u32 w) const {
        for (int q = 0; q < g.nins; ++q) w += (w & g.nmask[q]);
        return w | g.in_or;
    }
};

// phase tables of a stage at this tile: value(e) = SL(e & lom) * T_hi[e >> lo]
struct StageRef {
    const double2 *__restrict__ T;     // T_lo at T[0 .. nlo), T_hi at T[nlo ..)
    double2 S;                         // product of the outside-cell tables at this tile's base
    int lo;
    u32 nlo;
    u32 sor;                           // in-tile controls of the stage
};
QIPB_HD StageRef stage_ref(const DevGate &st, const double2 *__restrict__ tables, const double2 S, int tb) {
    const StageInfo &si = *reinterpret_cast<const StageInfo *>(st.m);
    StageRef r;
    r.T = tables + si.tab_off;
    r.lo = tb < FUSED_LO_BITS ? tb : FUSED_LO_BITS;
    r.nlo = 1u << r.lo;
    r.S = S;
    r.sor = st.in_or;
    return r;
}

// The stage scalars of one tile (product of the outside-cell tables at the tile's base index), one op
// per thread, computed while the tile's load is in flight; the sweeps read them from shared memory.
template <int NT>
QIPB_HD void stage_scalars(const FusedArgs &f, u64 base, double2 *stage_S, int tid) {
    const int lo = f.tb < FUSED_LO_BITS ? f.tb : FUSED_LO_BITS;
    for (int op = tid; op < f.ngates; op += NT)
        if (f.g[op].diag >= 2) {
            const StageInfo &si = *reinterpret_cast<const StageInfo *>(f.g[op].m);
            stage_S[op] = stage_scalar(si, f.tables + si.tab_off, base, 1u << lo, 1u << (f.tb - lo));
        }
}

// ---- dense 2-qubit gate: groups of four amplitudes ----
// ---- dense 2-qubit blocks by matrix structure ----
// The sweeps are bound by FP64 issue (a DFMA holds the dispatch port for two cycles), so the number of FP64
// instructions per group is what counts.  Merged blocks of real-world circuits are rarely general: in the
// layered benchmark 26 % are real (H (x) H with CX / Swap), 36 % are a real matrix times column phases
// (an Rm folded in), 10 % are permutations with phases (Swap, CX with Rm's), 38 % general.  The host
// (classify_block) picks the cheapest exact form:
//   MK_GENERAL    16 complex coefficients                         64 FP64 instructions per group
//   MK_REAL       16 real coefficients (first 16 scalars of m)    32
//   MK_REALPHASE  M = R . diag(ph): phases at complex slots 8..11 32 + 4 per phased column
//   MK_MONOMIAL   one non-zero per row/column: coefficients at complex slots 0..3, `perm`    4 per non-unit entry
enum { MK_GENERAL = 0, MK_REAL = 1, MK_REALPHASE = 2, MK_MONOMIAL = 3 };
// exact forms of a dense 1-qubit gate (DevGate::mk of a k == 1 op)
enum { MK1_GENERAL = 0, MK1_REAL = 1 };

template <typename A> QIPB_HD typename amp_traits<A>::real rcoef(const DevGate &g, int i) {
    return reinterpret_cast<const typename amp_traits<A>::real *>(g.m)[i];
}

template <typename A, int MK> struct Block2 {
    typedef typename amp_traits<A>::real R;
    Cf<A> m[MK == MK_GENERAL ? 16 : 4];
    R r[MK == MK_GENERAL ? 1 : 16];
    u32 phmask;
    QIPB_HD explicit Block2(const DevGate &g) {
        phmask = g.phmask;
        if (MK == MK_GENERAL) {
#pragma unroll
            for (int i = 0; i < 16; ++i) m[i] = coef<A>(g, i);
        } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) r[i] = rcoef<A>(g, i);
            if (MK == MK_REALPHASE) {
#pragma unroll
                for (int j = 0; j < 4; ++j) m[j] = coef<A>(g, 8 + j);
            }
        }
    }
    QIPB_HD void apply(A a0, A a1, A a2, A a3, A (&o)[4]) const {
        if (MK == MK_GENERAL) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                o[i] = cmulc<A>(m[4 * i], a0);
                cfmac<A>(o[i], m[4 * i + 1], a1);
                cfmac<A>(o[i], m[4 * i + 2], a2);
                cfmac<A>(o[i], m[4 * i + 3], a3);
            }
        } else {
            if (MK == MK_REALPHASE) {                          // uniform branches: the mask comes from the descriptor
                if (phmask & 1u) a0 = cmulc<A>(m[0], a0);
                if (phmask & 2u) a1 = cmulc<A>(m[1], a1);
                if (phmask & 4u) a2 = cmulc<A>(m[2], a2);
                if (phmask & 8u) a3 = cmulc<A>(m[3], a3);
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                o[i].x = r[4 * i] * a0.x;
                o[i].y = r[4 * i] * a0.y;
                o[i].x = fma(r[4 * i + 1], a1.x, o[i].x);
                o[i].y = fma(r[4 * i + 1], a1.y, o[i].y);
                o[i].x = fma(r[4 * i + 2], a2.x, o[i].x);
                o[i].y = fma(r[4 * i + 2], a2.y, o[i].y);
                o[i].x = fma(r[4 * i + 3], a3.x, o[i].x);
                o[i].y = fma(r[4 * i + 3], a3.y, o[i].y);
            }
        }
    }
};

template <typename A, bool UNI, int NT, int MK, typename EX>
QIPB_HD void sweep_dense2(A *tile, const DevGate &g, const EX ex, u32 ngroups, int tid) {
    const u32 oh = 1u << g.tl[0], ol = 1u << g.tl[1];
    const Block2<A, MK> blk(g);
    if (UNI && (ngroups % (2 * NT)) == 0) {
        // two groups per iteration, all eight loads first: every coefficient fetched from the parameter
        // bank serves both groups, and the second group's loads overlap the first group's arithmetic
#pragma unroll 1
        for (u32 it = 0, nit = ngroups / (2 * NT), w = tid; it < nit; ++it, w += 2 * NT) {
            A *p = tile + ex(w), *q = tile + ex(w + NT);
            const A a0 = p[0], a1 = p[ol], a2 = p[oh], a3 = p[oh + ol];
            const A b0 = q[0], b1 = q[ol], b2 = q[oh], b3 = q[oh + ol];
            A r[4], s[4];
            blk.apply(a0, a1, a2, a3, r);
            blk.apply(b0, b1, b2, b3, s);
            p[0] = r[0];
            p[ol] = r[1];
            p[oh] = r[2];
            p[oh + ol] = r[3];
            q[0] = s[0];
            q[ol] = s[1];
            q[oh] = s[2];
            q[oh + ol] = s[3];
        }
        return;
    }
#pragma unroll 1
    QIPB_SWEEP(w, ngroups) {
        A *p = tile + ex(w);
        const A a0 = p[0], a1 = p[ol], a2 = p[oh], a3 = p[oh + ol];
        A r[4];
        blk.apply(a0, a1, a2, a3, r);
        p[0] = r[0];
        p[ol] = r[1];
        p[oh] = r[2];
        p[oh + ol] = r[3];
    }
}

// ---- dense 2-qubit block (no controls) with the following stage applied to the group while it is in registers ----
// (WIDE kernel.)  The planner sinks the lone diagonal gates of a pass into one table stage; on its own that stage is a
// whole round trip of the tile through shared memory for one table look-up and two complex multiplies per amplitude.  Here
// it rides on the block's sweep like a QFT stage rides on its Hadamard: member m of the group gets S * T_lo * T_hi at its
// own index; a thread's indices keep their low `lo` bits over the sweep, so the four S * T_lo factors are folded once.
template <typename A, int NT, int MK>
QIPB_HD void sweep_dense2_stage(A *tile, const DevGate &g, const StageRef sr, u32 ngroups, int tid) {
    const u32 oh = 1u << g.tl[0], ol = 1u << g.tl[1];
    const u32 nm0 = g.nmask[0], nm1 = g.nmask[1];
    const u32 sor = sr.sor, lom = sr.nlo - 1u;
    const int lo = sr.lo;
    const Block2<A, MK> blk(g);
    const double2 *__restrict__ Th = sr.T + sr.nlo;
    u32 e0 = (u32)tid;
    e0 += e0 & nm0;
    e0 += e0 & nm1;
    const u32 off[4] = {0u, ol, oh, oh + ol};
    double2 SL[4];
#pragma unroll
    for (int m = 0; m < 4; ++m) SL[m] = cmul<double2>(sr.S, sr.T[(e0 | off[m]) & lom]);
#pragma unroll 1
    for (u32 it = 0, nit = ngroups / NT, w = tid; it < nit; ++it, w += NT) {
        u32 e = w;
        e += e & nm0;
        e += e & nm1;
        A *p = tile + e;
        const A a0 = p[0], a1 = p[ol], a2 = p[oh], a3 = p[oh + ol];
        double2 th[4];
#pragma unroll
        for (int m = 0; m < 4; ++m) th[m] = Th[(e | off[m]) >> lo];
        A r[4];
        blk.apply(a0, a1, a2, a3, r);
#pragma unroll
        for (int m = 0; m < 4; ++m)
            if (((e | off[m]) & sor) == sor) r[m] = cmul<A>(cmul<double2>(SL[m], th[m]), r[m]);
        p[0] = r[0];
        p[ol] = r[1];
        p[oh] = r[2];
        p[oh + ol] = r[3];
    }
}

// permutation with phases (Swap, CX / CZ-like blocks with phase gates folded in): pure data movement plus at
// most one complex multiply per amplitude; column j goes to row perm[j]
template <typename A, bool UNI, int NT, typename EX>
QIPB_HD void sweep_mono2(A *tile, const DevGate &g, const EX ex, u32 ngroups, int tid) {
    const u32 oh = 1u << g.tl[0], ol = 1u << g.tl[1];
    const u32 perm = g.perm, phmask = g.phmask;
    const Cf<A> c0 = coef<A>(g, 0), c1 = coef<A>(g, 1), c2 = coef<A>(g, 2), c3 = coef<A>(g, 3);
    u32 dst[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const u32 row = (perm >> (2 * j)) & 3u;
        dst[j] = ((row & 2u) ? oh : 0u) + ((row & 1u) ? ol : 0u);
    }
#pragma unroll 2
    QIPB_SWEEP(w, ngroups) {
        A *p = tile + ex(w);
        A a0 = p[0], a1 = p[ol], a2 = p[oh], a3 = p[oh + ol];
        if (phmask & 1u) a0 = cmulc<A>(c0, a0);
        if (phmask & 2u) a1 = cmulc<A>(c1, a1);
        if (phmask & 4u) a2 = cmulc<A>(c2, a2);
        if (phmask & 8u) a3 = cmulc<A>(c3, a3);
        p[dst[0]] = a0;
        p[dst[1]] = a1;
        p[dst[2]] = a2;
        p[dst[3]] = a3;
    }
}


// ---- two disjoint dense 2-qubit blocks in ONE sweep (WIDE kernel: 128 threads, up to 168 registers) ----
// A sweep is a round trip of the whole tile through shared memory, and with 4-5 dense blocks per pass the
// shared-memory pipe (LDS/STS of the sweeps + the TMA traffic of the tile itself) is what bounds a layered pass
// (~32 B per amplitude and sweep at 128 B/clk: as much as the HBM time of the tile).  Blocks on disjoint target pairs
// commute, so two of them run on a group of 16 amplitudes held in registers: member index k = 4 * ia + ib with ia / ib
// the matrix index of block A / B.  FP64 work is unchanged; shared-memory traffic, index arithmetic and barriers halve.
// A monomial block (Swap, CX, phases folded in) is a phase per row plus a permutation of the STORE offsets.
// Host guarantees (lower_fused): no in-tile controls on either block, the four targets distinct and above the
// bank-conflict bits (so a quarter-warp always touches 8 consecutive amplitudes), ngroups = tile / 16 a multiple of NT.
// PK: the form a block takes inside a pair -- PK_GENERAL, PK_REAL (real matrix, optionally times column phases: MK_REAL /
// MK_REALPHASE, the phase mask is a uniform branch) or PK_MONO.
enum { PK_GENERAL = 0, PK_REAL = 1, PK_MONO = 2 };
QIPB_HD int pair_kind(const DevGate &g) { return g.mk == MK_GENERAL ? PK_GENERAL : g.mk == MK_MONOMIAL ? PK_MONO : PK_REAL; }

template <typename A, int PK, int STRIDE>
QIPB_HD void pair_phase(A (&x)[16], const DevGate &g) {
    // STRIDE 4: the block acts on ia (members s, 4 + s, 8 + s, 12 + s);  STRIDE 1: on ib (members 4 s .. 4 s + 3)
    if (PK == PK_MONO) {
        const u32 phmask = g.phmask;
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (phmask & (1u << j)) {
                const Cf<A> c = coef<A>(g, j);
#pragma unroll
                for (int s = 0; s < 4; ++s) {
                    const int k = STRIDE == 4 ? 4 * j + s : 4 * s + j;
                    x[k] = cmulc<A>(c, x[k]);
                }
            }
    } else {
        const Block2<A, PK == PK_GENERAL ? MK_GENERAL : MK_REALPHASE> blk(g);
#pragma unroll
        for (int s = 0; s < 4; ++s) {
            const int b = STRIDE == 4 ? s : 4 * s;
            A r[4];
            blk.apply(x[b], x[b + STRIDE], x[b + 2 * STRIDE], x[b + 3 * STRIDE], r);
            x[b] = r[0];
            x[b + STRIDE] = r[1];
            x[b + 2 * STRIDE] = r[2];
            x[b + 3 * STRIDE] = r[3];
        }
    }
}

// one loop per combination of forms: the descriptors are read with warp-uniform addresses (ga = f.g[gi], gb = f.g[gi + 1])
// and nothing but LDS / FP64 / STS is left inside
template <typename A, int NT, int PKA, int PKB>
QIPB_HD void sweep_pair2_k(A *tile, const DevGate &ga, const DevGate &gb, u32 ngroups, int tid) {
    const u32 nm0 = ga.nmask[2], nm1 = ga.nmask[3], nm2 = ga.nmask[4], nm3 = ga.nmask[5];
    u32 ldA[4], ldB[4], stA[4], stB[4];
    {
        const u32 oah = 1u << ga.tl[0], oal = 1u << ga.tl[1], obh = 1u << gb.tl[0], obl = 1u << gb.tl[1];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            ldA[j] = ((j & 2) ? oah : 0u) + ((j & 1) ? oal : 0u);
            ldB[j] = ((j & 2) ? obh : 0u) + ((j & 1) ? obl : 0u);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {                         // column j of a monomial block lands in row perm[j]
            const u32 ra = PKA == PK_MONO ? (ga.perm >> (2 * j)) & 3u : (u32)j;
            const u32 rb = PKB == PK_MONO ? (gb.perm >> (2 * j)) & 3u : (u32)j;
            stA[j] = ((ra & 2u) ? oah : 0u) + ((ra & 1u) ? oal : 0u);
            stB[j] = ((rb & 2u) ? obh : 0u) + ((rb & 1u) ? obl : 0u);
        }
    }
#pragma unroll 1
    for (u32 it = 0, nit = ngroups / NT, w = tid; it < nit; ++it, w += NT) {
        u32 e = w;
        e += e & nm0;
        e += e & nm1;
        e += e & nm2;
        e += e & nm3;
        A *p = tile + e;
        A x[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) x[k] = p[ldA[k >> 2] + ldB[k & 3]];
        pair_phase<A, PKA, 4>(x, ga);
        pair_phase<A, PKB, 1>(x, gb);
#pragma unroll
        for (int k = 0; k < 16; ++k) p[stA[k >> 2] + stB[k & 3]] = x[k];
    }
}

template <typename A, int NT>
QIPB_HD void sweep_pair2(A *tile, const DevGate &ga, const DevGate &gb, u32 ngroups, int tid) {
    const int ka = pair_kind(ga), kb = pair_kind(gb);          // uniform: the forms come from the descriptors
    if (ka == PK_GENERAL) {
        if (kb == PK_GENERAL) sweep_pair2_k<A, NT, PK_GENERAL, PK_GENERAL>(tile, ga, gb, ngroups, tid);
        else if (kb == PK_REAL) sweep_pair2_k<A, NT, PK_GENERAL, PK_REAL>(tile, ga, gb, ngroups, tid);
        else sweep_pair2_k<A, NT, PK_GENERAL, PK_MONO>(tile, ga, gb, ngroups, tid);
    } else if (ka == PK_REAL) {
        if (kb == PK_GENERAL) sweep_pair2_k<A, NT, PK_REAL, PK_GENERAL>(tile, ga, gb, ngroups, tid);
        else if (kb == PK_REAL) sweep_pair2_k<A, NT, PK_REAL, PK_REAL>(tile, ga, gb, ngroups, tid);
        else sweep_pair2_k<A, NT, PK_REAL, PK_MONO>(tile, ga, gb, ngroups, tid);
    } else {
        if (kb == PK_GENERAL) sweep_pair2_k<A, NT, PK_MONO, PK_GENERAL>(tile, ga, gb, ngroups, tid);
        else if (kb == PK_REAL) sweep_pair2_k<A, NT, PK_MONO, PK_REAL>(tile, ga, gb, ngroups, tid);
        else sweep_pair2_k<A, NT, PK_MONO, PK_MONO>(tile, ga, gb, ngroups, tid);
    }
}

// ---- a dense 2-qubit block and a lone dense 1-qubit gate on a third tile bit in ONE sweep (WIDE kernel) ----
// After the planner has tensored lone 1-qubit gates in pairs (ops.pack_lone_1q) about one per pass is left over, and it
// costs a whole round trip of the tile through shared memory for 8-16 FP64 instructions per pair.  Block and gate act
// on disjoint bits, so they commute: groups of 8 amplitudes (member = 2 * ia + ic, ia the block's matrix index, ic the
// gate's) are held in registers, the block runs on both halves, the gate on the four pairs.  Host guarantees as for
// the block pairs: no in-tile controls, three distinct targets above the bank-conflict bits, tile / 8 a multiple of NT.
// 32 data registers + one coefficient set at a time: unlike the block pairs this fits the WIDE kernel's budget with the
// coefficients on the uniform datapath.
template <typename A, int NT, int PK, bool REAL1>
QIPB_HD void sweep_trio_k(A *tile, const DevGate &ga, const DevGate &gc, u32 ngroups, int tid) {
    typedef typename amp_traits<A>::real R;
    const u32 nm0 = ga.nmask[2], nm1 = ga.nmask[3], nm2 = ga.nmask[4];
    const u32 oc = 1u << gc.tl[0];
    u32 ld[4], st[4];
    {
        const u32 oh = 1u << ga.tl[0], ol = 1u << ga.tl[1];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            ld[j] = ((j & 2) ? oh : 0u) + ((j & 1) ? ol : 0u);
            const u32 r = PK == PK_MONO ? (ga.perm >> (2 * j)) & 3u : (u32)j;     // column j of a monomial block -> row perm[j]
            st[j] = ((r & 2u) ? oh : 0u) + ((r & 1u) ? ol : 0u);
        }
    }
#pragma unroll 1
    for (u32 it = 0, nit = ngroups / NT, w = tid; it < nit; ++it, w += NT) {
        u32 e = w;
        e += e & nm0;
        e += e & nm1;
        e += e & nm2;
        A *p = tile + e;
        A x[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) x[k] = p[ld[k >> 1] + ((k & 1) ? oc : 0u)];
        if (PK == PK_MONO) {
            const u32 phmask = ga.phmask;
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (phmask & (1u << j)) {
                    const Cf<A> c = coef<A>(ga, j);
                    x[2 * j] = cmulc<A>(c, x[2 * j]);
                    x[2 * j + 1] = cmulc<A>(c, x[2 * j + 1]);
                }
        } else {
            const Block2<A, PK == PK_GENERAL ? MK_GENERAL : MK_REALPHASE> blk(ga);
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                A r[4];
                blk.apply(x[c], x[2 + c], x[4 + c], x[6 + c], r);
                x[c] = r[0];
                x[2 + c] = r[1];
                x[4 + c] = r[2];
                x[6 + c] = r[3];
            }
        }
        if (REAL1) {
            const R m0 = rcoef<A>(gc, 0), m1 = rcoef<A>(gc, 1), m2 = rcoef<A>(gc, 2), m3 = rcoef<A>(gc, 3);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const A a0 = x[2 * j], a1 = x[2 * j + 1];
                x[2 * j].x = fma(m1, a1.x, m0 * a0.x);
                x[2 * j].y = fma(m1, a1.y, m0 * a0.y);
                x[2 * j + 1].x = fma(m3, a1.x, m2 * a0.x);
                x[2 * j + 1].y = fma(m3, a1.y, m2 * a0.y);
            }
        } else {
            const Cf<A> m0 = coef<A>(gc, 0), m1 = coef<A>(gc, 1), m2 = coef<A>(gc, 2), m3 = coef<A>(gc, 3);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const A a0 = x[2 * j], a1 = x[2 * j + 1];
                A r0 = cmulc<A>(m0, a0), r1 = cmulc<A>(m2, a0);
                cfmac<A>(r0, m1, a1);
                cfmac<A>(r1, m3, a1);
                x[2 * j] = r0;
                x[2 * j + 1] = r1;
            }
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) p[st[k >> 1] + ((k & 1) ? oc : 0u)] = x[k];
    }
}

template <typename A, int NT>
QIPB_HD void sweep_trio(A *tile, const DevGate &ga, const DevGate &gc, u32 ngroups, int tid) {
    const int ka = pair_kind(ga);                              // uniform: the forms come from the descriptors
    const bool real1 = gc.mk == MK1_REAL;
    if (ka == PK_GENERAL) {
        if (real1) sweep_trio_k<A, NT, PK_GENERAL, true>(tile, ga, gc, ngroups, tid);
        else sweep_trio_k<A, NT, PK_GENERAL, false>(tile, ga, gc, ngroups, tid);
    } else if (ka == PK_REAL) {
        if (real1) sweep_trio_k<A, NT, PK_REAL, true>(tile, ga, gc, ngroups, tid);
        else sweep_trio_k<A, NT, PK_REAL, false>(tile, ga, gc, ngroups, tid);
    } else {
        if (real1) sweep_trio_k<A, NT, PK_MONO, true>(tile, ga, gc, ngroups, tid);
        else sweep_trio_k<A, NT, PK_MONO, false>(tile, ga, gc, ngroups, tid);
    }
}

// ---- the same gates when a target sits on one of the lowest tile bits ----
// Shared memory serves one 128-byte line per cycle, i.e. 2^LOWB amplitudes (8 complex128 / 16 complex64).
// The lanes of one request phase hold consecutive group numbers, so with c targets below bit LOWB only
// 2^(LOWB-c) different bank groups are hit per request: a 2^c-way bank conflict on every load and store.
// Fix without touching the data layout: lane l visits the members of ITS group in the order j ^ R(l)
// (R = the c bits of the lane number that would otherwise collide, placed on the matrix-index bits of the
// low targets), so one request covers all bank groups.  The 2-qubit sweep swaps the loaded amplitudes back
// into member order with predicated register selects (coefficients stay warp-uniform); the 1-qubit sweep
// applies the correspondingly permuted matrix M'[i][j] = M[i ^ R][j ^ R] instead (8 registers per lane).
template <typename A> struct LowBits { static constexpr int value = sizeof(A) == 16 ? 3 : 4; };

template <typename A> QIPB_HD void cswap(const bool c, A &u, A &v) {
    const A t = c ? v : u;
    v = c ? u : v;
    u = t;
}

template <typename A, bool UNI, int NT, int MK, typename EX>
QIPB_HD void sweep_dense2_low(A *tile, const DevGate &g, const EX ex, u32 ngroups, int tid) {
    constexpr int LOWB = LowBits<A>::value;
    const u32 oh = 1u << g.tl[0], ol = 1u << g.tl[1];
    const int c = (g.tl[0] < LOWB) + (g.tl[1] < LOWB);
    const u32 B = (g.tl[0] < LOWB ? 2u : 0u) | (g.tl[1] < LOWB ? 1u : 0u);
    const u32 rb = ((u32)tid >> (LOWB - c)) & ((1u << c) - 1u);
    const u32 R = c == 2 ? rb : (rb ? B : 0u);
    const bool R0 = (R & 1u) != 0u, R1 = (R & 2u) != 0u;
    const Block2<A, MK> blk(g);       // uniform: the 16 coefficients would not fit per lane (64 registers)
    u32 off[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) off[k] = (((k ^ R) & 2u) ? oh : 0u) + (((k ^ R) & 1u) ? ol : 0u);
#pragma unroll 1
    QIPB_SWEEP(w, ngroups) {
        A *p = tile + ex(w);
        A x[4], r[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) x[k] = p[off[k]];          // x[k] = member k ^ R
        if (B & 1u) {                                           // back to member order in registers
            cswap<A>(R0, x[0], x[1]);                           // (uniform branches: only the levels whose
            cswap<A>(R0, x[2], x[3]);                           //  target really is a low bit cost selects)
        }
        if (B & 2u) {
            cswap<A>(R1, x[0], x[2]);
            cswap<A>(R1, x[1], x[3]);
        }
        blk.apply(x[0], x[1], x[2], x[3], r);
        if (B & 1u) {
            cswap<A>(R0, r[0], r[1]);
            cswap<A>(R0, r[2], r[3]);
        }
        if (B & 2u) {
            cswap<A>(R1, r[0], r[2]);
            cswap<A>(R1, r[1], r[3]);
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) p[off[k]] = r[k];
    }
}

template <typename A, bool UNI, int NT, typename EX>
QIPB_HD void sweep_dense1_low(A *tile, const DevGate &g, const EX ex, u32 ngroups, int tid) {
    constexpr int LOWB = LowBits<A>::value;
    const u32 o1 = 1u << g.tl[0];
    const bool R = (((u32)tid >> (LOWB - 1)) & 1u) != 0u;
    const Cf<A> c0 = coef<A>(g, 0), c1 = coef<A>(g, 1), c2 = coef<A>(g, 2), c3 = coef<A>(g, 3);
    const Cf<A> m0 = R ? c3 : c0, m1 = R ? c2 : c1, m2 = R ? c1 : c2, m3 = R ? c0 : c3;
    const u32 f0 = R ? o1 : 0u, f1 = R ? 0u : o1;
#pragma unroll 2
    QIPB_SWEEP(w, ngroups) {
        A *p = tile + ex(w);
        const A a0 = p[f0], a1 = p[f1];
        A r0 = cmulc<A>(m0, a0), r1 = cmulc<A>(m2, a0);
        cfmac<A>(r0, m1, a1);
        cfmac<A>(r1, m3, a1);
        p[f0] = r0;
        p[f1] = r1;
    }
}

// ---- dense 1-qubit gate: pairs ----
template <typename A, bool UNI, int NT, typename EX>
QIPB_HD void sweep_dense1(A *tile, const DevGate &g, const EX ex, u32 ngroups, int tid) {
    const u32 o1 = 1u << g.tl[0];
    const Cf<A> m0 = coef<A>(g, 0), m1 = coef<A>(g, 1), m2 = coef<A>(g, 2), m3 = coef<A>(g, 3);
#pragma unroll 2
    QIPB_SWEEP(w, ngroups) {
        A *p = tile + ex(w);
        const A a0 = p[0], a1 = p[o1];
        A r0 = cmulc<A>(m0, a0), r1 = cmulc<A>(m2, a0);
        cfmac<A>(r0, m1, a1);
        cfmac<A>(r1, m3, a1);
        p[0] = r0;
        p[o1] = r1;
    }
}

// ---- dense 1-qubit gate (no controls) with the following stage applied to the pair while it is in
// registers: H_k and its controlled phases in a QFT are one sweep.  Every thread visits tile indices
// whose low `lo` bits never change (the stride of the sweep is a multiple of 2^lo), so the T_lo factor
// and the stage scalar are folded once per thread; per pair only T_hi is looked up. ----
template <typename A, bool UNI, int NT>
QIPB_HD void sweep_dense1_stage(A *tile, const DevGate &g, const StageRef sr, u32 ngroups, int tid) {
    const u32 o1 = 1u << g.tl[0];
    const u32 nm = g.nmask[0];
    const u32 sor = sr.sor, lom = sr.nlo - 1u;
    const Cf<A> m0 = coef<A>(g, 0), m1 = coef<A>(g, 1), m2 = coef<A>(g, 2), m3 = coef<A>(g, 3);
    const double2 *__restrict__ Th = sr.T + sr.nlo;
    const u32 e_first = (u32)tid + ((u32)tid & nm);
    const double2 SL0 = cmul<double2>(sr.S, sr.T[e_first & lom]);
    const double2 SL1 = cmul<double2>(sr.S, sr.T[(e_first | o1) & lom]);
    const int lo = sr.lo;
    if (sor == o1 && UNI && (ngroups % (4 * NT)) == 0) {
        // the QFT step proper (stage controlled by exactly the gate's target), four pairs per iteration with
        // every load issued before the arithmetic: enough independent chains to cover the FP64 latency
#pragma unroll 1
        for (u32 it = 0, nit = ngroups / (4 * NT), w = tid; it < nit; ++it, w += 4 * NT) {
            A *p[4];
            A a0[4], a1[4];
            double2 th[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const u32 wq = w + q * NT;
                const u32 e = wq + (wq & nm);
                p[q] = tile + e;
                a0[q] = p[q][0];
                a1[q] = p[q][o1];
                th[q] = Th[(e | o1) >> lo];
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const double2 ph = cmul<double2>(SL1, th[q]);
                A r0 = cmulc<A>(m0, a0[q]), r1 = cmulc<A>(m2, a0[q]);
                cfmac<A>(r0, m1, a1[q]);
                cfmac<A>(r1, m3, a1[q]);
                p[q][0] = r0;
                p[q][o1] = cmul<A>(ph, r1);
            }
        }
    } else if (sor == o1) {                 // the same step on tiles whose sweep is not a multiple of 4 * NT pairs
#pragma unroll 2
        QIPB_SWEEP(w, ngroups) {
            const u32 e = w + (w & nm);
            A *p = tile + e;
            const A a0 = p[0], a1 = p[o1];
            const double2 ph = cmul<double2>(SL1, Th[(e | o1) >> lo]);
            A r0 = cmulc<A>(m0, a0), r1 = cmulc<A>(m2, a0);
            cfmac<A>(r0, m1, a1);
            cfmac<A>(r1, m3, a1);
            p[0] = r0;
            p[o1] = cmul<A>(ph, r1);
        }
    } else {
#pragma unroll 1
        QIPB_SWEEP(w, ngroups) {
            const u32 e = w + (w & nm), e1 = e | o1;
            A *p = tile + e;
            const A a0 = p[0], a1 = p[o1];
            A r0 = cmulc<A>(m0, a0), r1 = cmulc<A>(m2, a0);
            cfmac<A>(r0, m1, a1);
            cfmac<A>(r1, m3, a1);
            if ((e & sor) == sor) r0 = cmul<A>(cmul<double2>(SL0, Th[e >> lo]), r0);
            if ((e1 & sor) == sor) r1 = cmul<A>(cmul<double2>(SL1, Th[e1 >> lo]), r1);
            p[0] = r0;
            p[o1] = r1;
        }
    }
}

// ---- EXT sweeps (opt-in, QIPB_FUSED_EXT=1; run by the EXT instantiation of the kernel only) ----
// Exact cheaper forms of dense 1-qubit gates.  The sweeps are FP64-issue / shared-memory bound, and a QFT is n
// Hadamard sweeps each followed by its controlled phases (qip/qfft.py:33-39), so two things pay:
//   MK1_REAL  a real 2x2 matrix (H, X, Ry): 8 FP64 instructions per pair instead of 16;
//   QFT pair  two consecutive "Hadamard + stage" steps on tile bits a and b in ONE sweep over groups of four
//             amplitudes (a radix-4 butterfly): half the shared-memory traffic of two sweeps and 52 instead of
//             2 x 52 FP64 instructions per group (the Hadamard scales fold into the phase factors).
// Coefficient layout of MK1_REAL: r00, r01, r10, r11 in the first four real slots of m (amplitude precision);
// phmask != 0 marks s * [[1, 1], [1, -1]] exactly (only r00 = s is read by the paired sweep).
template <typename A, bool UNI, int NT, typename EX>
QIPB_HD void sweep_real1(A *tile, const DevGate &g, const EX ex, u32 ngroups, int tid) {
    typedef typename amp_traits<A>::real R;
    const u32 o1 = 1u << g.tl[0];
    const R m0 = rcoef<A>(g, 0), m1 = rcoef<A>(g, 1), m2 = rcoef<A>(g, 2), m3 = rcoef<A>(g, 3);
#pragma unroll 2
    QIPB_SWEEP(w, ngroups) {
        A *p = tile + ex(w);
        const A a0 = p[0], a1 = p[o1];
        A r0, r1;
        r0.x = fma(m1, a1.x, m0 * a0.x);
        r0.y = fma(m1, a1.y, m0 * a0.y);
        r1.x = fma(m3, a1.x, m2 * a0.x);
        r1.y = fma(m3, a1.y, m2 * a0.y);
        p[0] = r0;
        p[o1] = r1;
    }
}

// real 1-qubit gate + the stage controlled by exactly its target bit (host guarantees: sr.sor == 1 << g.tl[0],
// no other fixed position, ngroups a multiple of 4 * NT)
template <typename A, int NT>
QIPB_HD void sweep_real1_stage(A *tile, const DevGate &g, const StageRef sr, u32 ngroups, int tid) {
    typedef typename amp_traits<A>::real R;
    const u32 o1 = 1u << g.tl[0];
    const u32 nm = g.nmask[0];
    const u32 lom = sr.nlo - 1u;
    const R m0 = rcoef<A>(g, 0), m1 = rcoef<A>(g, 1), m2 = rcoef<A>(g, 2), m3 = rcoef<A>(g, 3);
    const double2 *__restrict__ Th = sr.T + sr.nlo;
    const u32 e_first = (u32)tid + ((u32)tid & nm);
    const double2 SL1 = cmul<double2>(sr.S, sr.T[(e_first | o1) & lom]);
    const int lo = sr.lo;
#pragma unroll 1
    for (u32 it = 0, nit = ngroups / (4 * NT), w = tid; it < nit; ++it, w += 4 * NT) {
        A *p[4];
        A a0[4], a1[4];
        double2 th[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const u32 wq = w + q * NT;
            const u32 e = wq + (wq & nm);
            p[q] = tile + e;
            a0[q] = p[q][0];
            a1[q] = p[q][o1];
            th[q] = Th[(e | o1) >> lo];
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const double2 ph = cmul<double2>(SL1, th[q]);
            A r0, r1;
            r0.x = fma(m1, a1[q].x, m0 * a0[q].x);
            r0.y = fma(m1, a1[q].y, m0 * a0[q].y);
            r1.x = fma(m3, a1[q].x, m2 * a0[q].x);
            r1.y = fma(m3, a1[q].y, m2 * a0[q].y);
            p[q][0] = r0;
            p[q][o1] = cmul<A>(ph, r1);
        }
    }
}

// Two QFT steps in one sweep.  Members of a group: x[ab] with a = the bit of the first Hadamard (ga, stage sa),
// b = the bit of the second (gb, stage sb).  With s = s_a * s_b and PA / PB the stage phases at a member's index:
//   u0 = x00 + x10, u1 = x01 + x11, w10 = (x00 - x10) PA(10), w11 = (x01 - x11) PA(11)
//   z00 = s (u0 + u1)   z01 = s PB(01) (u0 - u1)   z10 = s (w10 + w11)   z11 = s PB(11) (w10 - w11)
// which is H_a, stage a (on members with a = 1), H_b, stage b (on members with b = 1) applied in that order.
// Host guarantees: both gates are s * [[1, 1], [1, -1]] without controls, both stages are controlled by exactly
// their gate's target bit and have no outside controls, both targets sit above the bank-conflict bits, and
// ngroups (= tile / 4) is a multiple of NT.  A thread's tile indices keep their low `lo` bits over the sweep, so
// the low-cell table factor, the tile scalar and the Hadamard scales are folded once per thread.
template <typename A, int NT>
QIPB_HD void sweep_qft2(A *tile, const DevGate &ga, const StageRef sa, const DevGate &gb, const StageRef sb, u32 ngroups, int tid) {
    typedef typename amp_traits<A>::real R;
    const u32 pa = ga.tl[0], pb = gb.tl[0];
    const u32 oa = 1u << pa, ob = 1u << pb;
    const u32 nm0 = ~((1u << (pa < pb ? pa : pb)) - 1u), nm1 = ~((1u << (pa < pb ? pb : pa)) - 1u);
    const R s = rcoef<A>(ga, 0) * rcoef<A>(gb, 0);
    const u32 lom = sa.nlo - 1u;
    const int lo = sa.lo;
    const double2 *__restrict__ Tha = sa.T + sa.nlo;
    const double2 *__restrict__ Thb = sb.T + sb.nlo;
    u32 e0 = (u32)tid;
    e0 += e0 & nm0;
    e0 += e0 & nm1;
    const double2 A10 = cmul<double2>(sa.S, sa.T[(e0 | oa) & lom]);
    const double2 A11 = cmul<double2>(sa.S, sa.T[(e0 | oa | ob) & lom]);
    double2 B01 = cmul<double2>(sb.S, sb.T[(e0 | ob) & lom]);
    double2 B11 = cmul<double2>(sb.S, sb.T[(e0 | oa | ob) & lom]);
    B01.x *= (double)s;
    B01.y *= (double)s;
    B11.x *= (double)s;
    B11.y *= (double)s;
#pragma unroll 1
    for (u32 it = 0, nit = ngroups / NT, w = tid; it < nit; ++it, w += NT) {
        u32 e = w;
        e += e & nm0;
        e += e & nm1;
        A *p = tile + e;
        const A x00 = p[0], x01 = p[ob], x10 = p[oa], x11 = p[oa + ob];
        const double2 ta0 = Tha[(e | oa) >> lo], ta1 = Tha[(e | oa | ob) >> lo];
        const double2 tb0 = Thb[(e | ob) >> lo], tb1 = Thb[(e | oa | ob) >> lo];
        A u0, u1, d0, d1;
        u0.x = x00.x + x10.x;
        u0.y = x00.y + x10.y;
        u1.x = x01.x + x11.x;
        u1.y = x01.y + x11.y;
        d0.x = x00.x - x10.x;
        d0.y = x00.y - x10.y;
        d1.x = x01.x - x11.x;
        d1.y = x01.y - x11.y;
        const A w10 = cmul<A>(cmul<double2>(A10, ta0), d0);
        const A w11 = cmul<A>(cmul<double2>(A11, ta1), d1);
        A z00, z10, t01, t11;
        z00.x = s * (u0.x + u1.x);
        z00.y = s * (u0.y + u1.y);
        z10.x = s * (w10.x + w11.x);
        z10.y = s * (w10.y + w11.y);
        t01.x = u0.x - u1.x;
        t01.y = u0.y - u1.y;
        t11.x = w10.x - w11.x;
        t11.y = w10.y - w11.y;
        p[0] = z00;
        p[ob] = cmul<A>(cmul<double2>(B01, tb0), t01);
        p[oa] = z10;
        p[oa + ob] = cmul<A>(cmul<double2>(B11, tb1), t11);
    }
}

// Four QFT steps in one sweep: a radix-16 butterfly on groups of 16 amplitudes held in registers (WIDE kernel).  Member
// index k has bit j = the bit of step j's Hadamard (steps in execution order).  Step j: for every pair (k0, k1 = k0 | 2^j)
//   x[k0] <- x[k0] + x[k1],   x[k1] <- (x[k0] - x[k1]) * phase_j(k1)
// and the four Hadamard scales are applied once at the end.  The stage phase at a member's index factorises into what the
// rest of the index contributes -- S * T_lo * T_hi at the group's base index: one table look-up per step and group, the
// low-cell factor folded once per thread -- times what the member's own bits contribute: M_j[k1 without bit j], eight
// complex constants per step that the host derives from the tables (slots 8..15 of the Hadamard's descriptor, always
// double2) after checking that the tables really factorise over those bits (true for the controlled phases of a QFT:
// every gate of the stage couples the target with ONE other bit).  Same FP64 work as two radix-4 sweeps (about 26
// instructions per amplitude), half the shared-memory traffic, index arithmetic and barriers.
template <typename A, int NT>
QIPB_HD void sweep_qft4(A *tile, const FusedArgs &f, int gi, const double2 *stage_S, u32 ngroups, int tid) {
    typedef typename amp_traits<A>::real R;
    const DevGate &h0 = f.g[gi], &h1 = f.g[gi + 2], &h2 = f.g[gi + 4], &h3 = f.g[gi + 6];
    const StageRef s0 = stage_ref(f.g[gi + 1], f.tables, stage_S[gi + 1], f.tb), s1 = stage_ref(f.g[gi + 3], f.tables, stage_S[gi + 3], f.tb);
    const StageRef s2 = stage_ref(f.g[gi + 5], f.tables, stage_S[gi + 5], f.tb), s3 = stage_ref(f.g[gi + 7], f.tables, stage_S[gi + 7], f.tb);
    const u32 o0 = 1u << h0.tl[0], o1 = 1u << h1.tl[0], o2 = 1u << h2.tl[0], o3 = 1u << h3.tl[0];
    const u32 nm0 = h0.nmask[2], nm1 = h0.nmask[3], nm2 = h0.nmask[4], nm3 = h0.nmask[5];
    const u32 lom = s0.nlo - 1u;
    const int lo = s0.lo;
    const double2 *__restrict__ T0 = s0.T + s0.nlo, *__restrict__ T1 = s1.T + s1.nlo;
    const double2 *__restrict__ T2 = s2.T + s2.nlo, *__restrict__ T3 = s3.T + s3.nlo;
    const double scale = (double)rcoef<A>(h0, 0) * (double)rcoef<A>(h1, 0) * (double)rcoef<A>(h2, 0) * (double)rcoef<A>(h3, 0);
    u32 e0 = (u32)tid;
    e0 += e0 & nm0;
    e0 += e0 & nm1;
    e0 += e0 & nm2;
    e0 += e0 & nm3;
    const double2 SL0 = cmul<double2>(s0.S, s0.T[e0 & lom]), SL1 = cmul<double2>(s1.S, s1.T[e0 & lom]);
    const double2 SL2 = cmul<double2>(s2.S, s2.T[e0 & lom]);
    double2 SL3 = cmul<double2>(s3.S, s3.T[e0 & lom]);
    SL3.x *= scale;                                            // the last step's phased members take the scale with their phase
    SL3.y *= scale;
#pragma unroll 1
    for (u32 it = 0, nit = ngroups / NT, w = tid; it < nit; ++it, w += NT) {
        u32 e = w;
        e += e & nm0;
        e += e & nm1;
        e += e & nm2;
        e += e & nm3;
        A *p = tile + e;
        A x[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) x[k] = p[((k & 1) ? o0 : 0u) + ((k & 2) ? o1 : 0u) + ((k & 4) ? o2 : 0u) + ((k & 8) ? o3 : 0u)];
        const u32 eh = e >> lo;
        const double2 P0 = cmul<double2>(SL0, T0[eh]), P1 = cmul<double2>(SL1, T1[eh]);
        const double2 P2 = cmul<double2>(SL2, T2[eh]), P3 = cmul<double2>(SL3, T3[eh]);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const DevGate &h = j == 0 ? h0 : j == 1 ? h1 : j == 2 ? h2 : h3;
            const double2 P = j == 0 ? P0 : j == 1 ? P1 : j == 2 ? P2 : P3;
#pragma unroll
            for (int k0 = 0; k0 < 16; ++k0) {
                if (k0 & (1 << j)) continue;
                const int k1 = k0 | (1 << j);
                const int idx = (k1 & ((1 << j) - 1)) | ((k1 >> (j + 1)) << j);      // k1 without bit j
                A u, d;
                u.x = x[k0].x + x[k1].x;
                u.y = x[k0].y + x[k1].y;
                d.x = x[k0].x - x[k1].x;
                d.y = x[k0].y - x[k1].y;
                if (j == 3) {
                    u.x *= (R)scale;
                    u.y *= (R)scale;
                }
                x[k0] = u;
                x[k1] = cmul<A>(cmul<double2>(P, h.m[8 + idx]), d);
            }
        }
#pragma unroll
        for (int k = 0; k < 16; ++k) p[((k & 1) ? o0 : 0u) + ((k & 2) ? o1 : 0u) + ((k & 4) ? o2 : 0u) + ((k & 8) ? o3 : 0u)] = x[k];
    }
}

// QFT steps on the LOWEST tile bits (0..2, the bank-conflict bits): up to three consecutive "Hadamard + stage" steps in one
// sweep, butterflies across LANES.  A thread keeps amplitude e = tid + NT * i, so consecutive lanes hold consecutive
// amplitudes (every shared-memory request is conflict free) and the partner of a pair on bit b <= 4 sits in lane ^ 2^b:
// one __shfl_xor of the two components.  With y = the partner's value a lane computes y + x (its bit b clear) or
// y - x (bit set), both as fma(sign, x, y), times a per-lane multiplier: s, or s * S * T_lo * T_hi at its own index where
// the stage's in-tile controls are satisfied (T_lo folded once per thread: the low six bits of e never change).  10 FP64 instructions, 4 shuffles and one table
// look-up per amplitude and step -- against a sweep of its own per step with 2- to 4-way bank conflicts before: the 12-bit
// pass at the low end of a QFT carried three such sweeps (128 ms of a 372 ms QFFT-33, profiles/r02_ab_qft4.txt).
template <typename A, int NT>
QIPB_HD void sweep_qft_low(A *tile, const FusedArgs &f, int gi, int nsteps, const double2 *stage_S, u32 tsize, int tid) {
    typedef typename amp_traits<A>::real R;
#if defined(__CUDA_ARCH__)
    const u32 lane_e = (u32)tid;
    u32 bpos[3], sor[3];
    double sc[3];
    double2 SL[3];
    const double2 *Th[3];
    int lo = 0;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        const int gj = j < nsteps ? gi + 2 * j : gi;            // (unused slots mirror step 0; never applied)
        const DevGate &h = f.g[gj];
        const StageRef sr = stage_ref(f.g[gj + 1], f.tables, stage_S[gj + 1], f.tb);
        bpos[j] = h.tl[0];
        sor[j] = sr.sor;                                       // in-tile controls of the stage (usually just the target bit)
        sc[j] = (double)coef<A>(h, 0).x;
        SL[j] = cmul<double2>(sr.S, sr.T[lane_e & (sr.nlo - 1u)]);
        SL[j].x *= sc[j];
        SL[j].y *= sc[j];
        Th[j] = sr.T + sr.nlo;
        lo = sr.lo;
    }
    constexpr int U = 4;
#pragma unroll 1
    for (u32 i = 0, n = tsize / NT; i < n; i += U) {
        A x[U];
#pragma unroll
        for (int u = 0; u < U; ++u) x[u] = tile[lane_e + NT * (i + u)];
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            if (j >= nsteps) break;
            const u32 b = bpos[j];
            const bool hi = ((lane_e >> b) & 1u) != 0u;
            const double sgn = hi ? -1.0 : 1.0;
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const u32 e = lane_e + NT * (i + u);
                const double2 th = Th[j][e >> lo];
                A y;
                y.x = __shfl_xor_sync(0xffffffffu, x[u].x, 1 << b);
                y.y = __shfl_xor_sync(0xffffffffu, x[u].y, 1 << b);
                A r;
                r.x = fma((R)sgn, x[u].x, y.x);
                r.y = fma((R)sgn, x[u].y, y.y);
                double2 mult = cmul<double2>(SL[j], th);
                if ((e & sor[j]) != sor[j]) mult = make_double2(sc[j], 0.0);
                x[u] = cmul<A>(mult, r);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) tile[lane_e + NT * (i + u)] = x[u];
    }
#else
    // host emulation (tests/csrc/fused_emul.cu runs "threads" one after another: no shuffles): the same arithmetic pair by
    // pair, executed once by the thread that owns the pair's lower member
    for (int j = 0; j < nsteps; ++j) {
        const DevGate &h = f.g[gi + 2 * j];
        const StageRef sr = stage_ref(f.g[gi + 2 * j + 1], f.tables, stage_S[gi + 2 * j + 1], f.tb);
        const u32 b = h.tl[0];
        const double s = (double)coef<A>(h, 0).x;
        const double2 *T_hi = sr.T + sr.nlo;
        // a step is complete over the whole tile before the next one starts on the device (all lanes advance together per
        // batch, and batches are independent); on the host "thread" tid handles its amplitudes for step j only when called
        // with that step -- the emulator calls this function once per thread, so thread 0 does the whole tile
        if (tid != 0) continue;
        for (u32 e = 0; e < tsize; ++e) {
            if ((e >> b) & 1u) continue;
            const u32 e1 = e | (1u << b);
            const A x0 = tile[e], x1 = tile[e1];
            double2 SLl = cmul<double2>(sr.S, sr.T[e1 & (sr.nlo - 1u)]);
            SLl.x *= s;
            SLl.y *= s;
            const double2 ph = cmul<double2>(SLl, T_hi[e1 >> sr.lo]);
            double2 SL0 = cmul<double2>(sr.S, sr.T[e & (sr.nlo - 1u)]);
            SL0.x *= s;
            SL0.y *= s;
            const double2 ph0 = cmul<double2>(SL0, T_hi[e >> sr.lo]);
            A u, d;
            u.x = fma((R)1.0, x0.x, x1.x);
            u.y = fma((R)1.0, x0.y, x1.y);
            d.x = fma((R)-1.0, x1.x, x0.x);
            d.y = fma((R)-1.0, x1.y, x0.y);
            tile[e] = cmul<A>((e & sr.sor) == sr.sor ? ph0 : make_double2(s, 0.0), u);
            tile[e1] = cmul<A>((e1 & sr.sor) == sr.sor ? ph : make_double2(s, 0.0), d);
        }
    }
#endif
}

// Fill mode (EXT kernel, qipb_apply_fused_fill): the pass acts on the all-ones vector instead of the buffer's
// content, and its first op is a stage without controls -- the tile is WRITTEN from the phase tables instead of being
// loaded from HBM.  A product state of one-qubit feeds is exactly that: prod_b diag(v_b[0], v_b[1]) . ones
// (CythonBackend.make_state's kron product, qip/backend.py:88-101, for one-qubit groups), so the initial state of a
// run() never travels through HBM before the first gate pass reads it.
template <typename A, int NT>
QIPB_HD void sweep_stage_fill(A *tile, const StageRef sr, u32 n, int tid) {
    typedef typename amp_traits<A>::real R;
    const double2 *__restrict__ Th = sr.T + sr.nlo;
    const double2 SL = cmul<double2>(sr.S, sr.T[(u32)tid & (sr.nlo - 1u)]);   // NT is a multiple of 2^lo: low bits are sweep-invariant
    const int lo = sr.lo;
#pragma unroll 4
    for (u32 it = 0, nit = n / NT, x = tid; it < nit; ++it, x += NT) {   // warp-uniform trip count (n is a multiple of NT)
        const double2 v = cmul<double2>(SL, Th[x >> lo]);
        tile[x] = make_amp<A>((R)v.x, (R)v.y);
    }
}

// ---- a stage on its own: one phase per element ----
template <typename A, bool UNI, int NT, typename EX>
QIPB_HD void sweep_stage(A *tile, const StageRef sr, const EX ex, u32 n, int tid) {
    const double2 *__restrict__ Th = sr.T + sr.nlo;
    const double2 SL = cmul<double2>(sr.S, sr.T[ex((u32)tid) & (sr.nlo - 1u)]);   // low bits are sweep-invariant
    const int lo = sr.lo;
    if (UNI && (n % (4 * NT)) == 0) {
#pragma unroll 1
        for (u32 it = 0, nit = n / (4 * NT), x = tid; it < nit; ++it, x += 4 * NT) {
            u32 e[4];
            A v[4];
            double2 th[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                e[q] = ex(x + q * NT);
                v[q] = tile[e[q]];
                th[q] = Th[e[q] >> lo];
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) tile[e[q]] = cmul<A>(cmul<double2>(SL, th[q]), v[q]);
        }
        return;
    }
#pragma unroll 1
    QIPB_SWEEP(x, n) {
        const u32 e = ex(x);
        tile[e] = cmul<A>(cmul<double2>(SL, Th[e >> lo]), tile[e]);
    }
}

// ---- a lone diagonal gate (k <= 2 targets, any of them possibly outside the tile) ----
template <typename A, bool UNI, int NT, typename EX>
QIPB_HD void sweep_diag(A *tile, const DevGate &g, const EX ex, u32 ngroups, u64 base, int tid) {
    u32 sel_out = 0;           // matrix-index bits of the targets that lie outside the tile (fixed per tile)
    for (int j = 0; j < g.k; ++j)
        if (g.tl[j] == 0xFF) sel_out |= (u32)((base >> g.tg[j]) & 1ull) << (g.k - 1 - j);
    const int D = 1 << g.k;
    if (g.kin == 0) {
        const Cf<A> c = coef<A>(g, sel_out * D + sel_out);
        if (c.x == 1 && c.y == 0) return;
#pragma unroll 4
        QIPB_SWEEP(w, ngroups) {
            const u32 e = ex(w);
            tile[e] = cmulc<A>(c, tile[e]);
        }
    } else if (g.kin == 1) {
        const int j = (g.tl[0] != 0xFF) ? 0 : 1;
        const u32 o1 = 1u << g.tl[j];
        const u32 s1 = 1u << (g.k - 1 - j);
        const Cf<A> d0 = coef<A>(g, sel_out * D + sel_out), d1 = coef<A>(g, (sel_out | s1) * D + (sel_out | s1));
#pragma unroll 2
        QIPB_SWEEP(w, ngroups) {
            A *p = tile + ex(w);
            p[0] = cmulc<A>(d0, p[0]);
            p[o1] = cmulc<A>(d1, p[o1]);
        }
    } else {
        const u32 oh = 1u << g.tl[0], ol = 1u << g.tl[1];
        const Cf<A> d0 = coef<A>(g, 0), d1 = coef<A>(g, 5), d2 = coef<A>(g, 10), d3 = coef<A>(g, 15);
#pragma unroll 1
        QIPB_SWEEP(w, ngroups) {
            A *p = tile + ex(w);
            p[0] = cmulc<A>(d0, p[0]);
            p[ol] = cmulc<A>(d1, p[ol]);
            p[oh] = cmulc<A>(d2, p[oh]);
            p[oh + ol] = cmulc<A>(d3, p[oh + ol]);
        }
    }
}

// One op on the tile.  All branches here are uniform over the CTA (they depend on the descriptor and
// on the tile's base index only).  UNI: every sweep of this launch has a multiple of the CTA size as
// its item count (tile bits - fixed positions >= 8) and at most four fixed positions -- true for every
// production-size pass; tiny states and gates with many in-tile controls take the generic kernel.
template <typename A, bool UNI, int NT>
QIPB_HD void run_op(A *tile, const DevGate &g, const DevGate &next, const double2 *__restrict__ tables,
                                       const double2 *stage_S, int gi, u64 base, int tb, u32 tsize, int tid) {
    if ((base & g.out_ctrl) != g.out_ctrl) return;
    const u32 ngroups = tsize >> g.nins;
    const int nins = (UNI && (ngroups % NT) == 0) ? g.nins : -1;
    if (g.diag == 2) {
        const StageRef sr = stage_ref(g, tables, stage_S[gi], tb);
        switch (nins) {
        case 0: sweep_stage<A, UNI, NT>(tile, sr, Expand<0>(g), ngroups, tid); break;
        case 1: sweep_stage<A, UNI, NT>(tile, sr, Expand<1>(g), ngroups, tid); break;
        case 2: sweep_stage<A, UNI, NT>(tile, sr, Expand<2>(g), ngroups, tid); break;
        case 3: sweep_stage<A, UNI, NT>(tile, sr, Expand<3>(g), ngroups, tid); break;
        case 4: sweep_stage<A, UNI, NT>(tile, sr, Expand<4>(g), ngroups, tid); break;
        default: sweep_stage<A, false, NT>(tile, sr, ExpandAny(g), ngroups, tid); break;
        }
    } else if (g.diag == 1) {
        switch (nins) {
        case 0: sweep_diag<A, UNI, NT>(tile, g, Expand<0>(g), ngroups, base, tid); break;
        case 1: sweep_diag<A, UNI, NT>(tile, g, Expand<1>(g), ngroups, base, tid); break;
        case 2: sweep_diag<A, UNI, NT>(tile, g, Expand<2>(g), ngroups, base, tid); break;
        case 3: sweep_diag<A, UNI, NT>(tile, g, Expand<3>(g), ngroups, base, tid); break;
        case 4: sweep_diag<A, UNI, NT>(tile, g, Expand<4>(g), ngroups, base, tid); break;
        default: sweep_diag<A, false, NT>(tile, g, ExpandAny(g), ngroups, base, tid); break;
        }
    } else if (g.k == 2) {
        if (nins == 2) {                                        // no in-tile controls: the structured forms
            if (g.mk == MK_MONOMIAL) {
                sweep_mono2<A, UNI, NT>(tile, g, Expand<2>(g), ngroups, tid);
            } else if (g.tl[0] < LowBits<A>::value || g.tl[1] < LowBits<A>::value) {
                if (g.mk == MK_REAL) sweep_dense2_low<A, UNI, NT, MK_REAL>(tile, g, Expand<2>(g), ngroups, tid);
                else if (g.mk == MK_REALPHASE) sweep_dense2_low<A, UNI, NT, MK_REALPHASE>(tile, g, Expand<2>(g), ngroups, tid);
                else sweep_dense2_low<A, UNI, NT, MK_GENERAL>(tile, g, Expand<2>(g), ngroups, tid);
            } else {
                if (g.mk == MK_REAL) sweep_dense2<A, UNI, NT, MK_REAL>(tile, g, Expand<2>(g), ngroups, tid);
                else if (g.mk == MK_REALPHASE) sweep_dense2<A, UNI, NT, MK_REALPHASE>(tile, g, Expand<2>(g), ngroups, tid);
                else sweep_dense2<A, UNI, NT, MK_GENERAL>(tile, g, Expand<2>(g), ngroups, tid);
            }
            return;
        }
        switch (nins) {                                         // the host keeps MK_GENERAL for these
        case 3: sweep_dense2<A, UNI, NT, MK_GENERAL>(tile, g, Expand<3>(g), ngroups, tid); break;
        case 4: sweep_dense2<A, UNI, NT, MK_GENERAL>(tile, g, Expand<4>(g), ngroups, tid); break;
        default: sweep_dense2<A, false, NT, MK_GENERAL>(tile, g, ExpandAny(g), ngroups, tid); break;
        }
    } else {
        if (g.post != 0 && (base & next.out_ctrl) == next.out_ctrl) {     // host guarantees nins == 1, no controls
            sweep_dense1_stage<A, UNI, NT>(tile, g, stage_ref(next, tables, stage_S[gi + 1], tb), ngroups, tid);
            return;
        }
        if (nins == 1 && g.tl[0] < LowBits<A>::value) {
            sweep_dense1_low<A, UNI, NT>(tile, g, Expand<1>(g), ngroups, tid);
            return;
        }
        switch (nins) {
        case 1: sweep_dense1<A, UNI, NT>(tile, g, Expand<1>(g), ngroups, tid); break;
        case 2: sweep_dense1<A, UNI, NT>(tile, g, Expand<2>(g), ngroups, tid); break;
        case 3: sweep_dense1<A, UNI, NT>(tile, g, Expand<3>(g), ngroups, tid); break;
        case 4: sweep_dense1<A, UNI, NT>(tile, g, Expand<4>(g), ngroups, tid); break;
        default: sweep_dense1<A, false, NT>(tile, g, ExpandAny(g), ngroups, tid); break;
        }
    }
}

// The op loop of every executor (the two kernels below, and the host emulation of tests/csrc/fused_emul.cu):
//   for gi: if (!fused_op_is_skipped(g[gi])) { run_fused_op(...); barrier; }
// EXT: the kernel instantiation that also knows the opt-in forms (real 1-qubit gates, paired QFT steps); the host
// launches it only for passes that carry such ops (fused_has_ext), so the default kernels stay as measured.
template <bool EXT>
QIPB_HD bool fused_op_is_skipped(const DevGate &g) {
    // stage applied by the dense gate before it / second Hadamard of a QFT pair / second block of a block pair
    return g.diag == 3 || (EXT && (g.post == 3 || g.post == 5 || g.post == 7));   // (7: the 1-qubit gate of a trio)
}

static inline bool fused_has_ext(const FusedArgs &f) {
    for (int gi = 0; gi < f.ngates; ++gi) {
        const DevGate &g = f.g[gi];
        if (g.post >= 2 || g.diag == 5 || (!g.diag && g.k == 1 && g.mk == MK1_REAL)) return true;   // incl. block pairs (post 4 / 5), trios (6 / 7)
    }
    return false;
}

template <typename A, bool UNI, int NT, bool EXT>
QIPB_HD void run_fused_op(A *tile, const FusedArgs &f, int gi, const double2 *stage_S, u64 base, u32 tsize, int tid) {
    if (EXT) {
        const DevGate &g = f.g[gi];
        if (QIPB_ENABLE_PAIRS && NT == 128 && g.post == 4) {                                      // block pair (WIDE launches only): ops gi and gi + 1
            const DevGate &h = f.g[gi + 1 < FUSED_MAX_OPS ? gi + 1 : gi];
            const bool on_a = (base & g.out_ctrl) == g.out_ctrl, on_b = (base & h.out_ctrl) == h.out_ctrl;
            if (on_a && on_b) {
                sweep_pair2<A, NT>(tile, g, h, tsize >> 4, tid);
            } else if (on_a) {                                  // the other one is switched off on this tile
                run_op<A, UNI, NT>(tile, g, g, f.tables, stage_S, gi, base, f.tb, tsize, tid);
            } else if (on_b) {
                run_op<A, UNI, NT>(tile, h, h, f.tables, stage_S, gi, base, f.tb, tsize, tid);
            }
            return;
        }
        if (NT == 128 && g.post == 8) {                         // dense 2-qubit block + the stage behind it (WIDE launches only)
            const StageRef sr = stage_ref(f.g[gi + 1 < FUSED_MAX_OPS ? gi + 1 : gi], f.tables, stage_S[gi + 1 < FUSED_MAX_OPS ? gi + 1 : gi], f.tb);
            if (g.mk == MK_REAL) sweep_dense2_stage<A, NT, MK_REAL>(tile, g, sr, tsize >> 2, tid);
            else if (g.mk == MK_REALPHASE) sweep_dense2_stage<A, NT, MK_REALPHASE>(tile, g, sr, tsize >> 2, tid);
            else sweep_dense2_stage<A, NT, MK_GENERAL>(tile, g, sr, tsize >> 2, tid);
            return;
        }
        if (QIPB_ENABLE_TRIOS && NT == 128 && g.post == 6) {    // block + lone 1-qubit gate (WIDE launches only): ops gi, gi + 1
            const DevGate &h = f.g[gi + 1 < FUSED_MAX_OPS ? gi + 1 : gi];
            const bool on_a = (base & g.out_ctrl) == g.out_ctrl, on_c = (base & h.out_ctrl) == h.out_ctrl;
            if (on_a && on_c) {
                sweep_trio<A, NT>(tile, g, h, tsize >> 3, tid);
            } else if (on_a) {
                run_op<A, UNI, NT>(tile, g, g, f.tables, stage_S, gi, base, f.tb, tsize, tid);
            } else if (on_c) {
                if (h.mk == MK1_REAL) sweep_real1<A, UNI, NT>(tile, h, Expand<1>(h), tsize >> 1, tid);
                else sweep_dense1<A, UNI, NT>(tile, h, Expand<1>(h), tsize >> 1, tid);
            }
            return;
        }
        if (g.diag == 5) {                                      // fill mode: op 0 writes the tile (no controls, host-checked)
            sweep_stage_fill<A, NT>(tile, stage_ref(g, f.tables, stage_S[gi], f.tb), tsize, tid);
            return;
        }
        if (NT == 128 && g.post == 10) {                        // g.pair consecutive QFT steps on the lowest tile bits (WIDE launches)
            sweep_qft_low<A, NT>(tile, f, gi, (int)g.pair, stage_S, tsize, tid);
            return;
        }
        if (NT == 128 && g.post == 9) {                         // ops gi .. gi+7 = four QFT steps (WIDE launches only)
            sweep_qft4<A, NT>(tile, f, gi, stage_S, tsize >> 4, tid);
            return;
        }
        if (g.post == 2) {                                      // ops gi .. gi+3 = H_a, stage a, H_b, stage b
            sweep_qft2<A, NT>(tile, g, stage_ref(f.g[gi + 1], f.tables, stage_S[gi + 1], f.tb), f.g[gi + 2],
                              stage_ref(f.g[gi + 3], f.tables, stage_S[gi + 3], f.tb), tsize >> 2, tid);
            return;
        }
        if (!g.diag && g.k == 1 && g.mk == MK1_REAL) {         // host: no controls inside the tile, target above the low bits
            if ((base & g.out_ctrl) != g.out_ctrl) return;
            if (g.post == 1 && (base & f.g[gi + 1].out_ctrl) == f.g[gi + 1].out_ctrl)
                sweep_real1_stage<A, NT>(tile, g, stage_ref(f.g[gi + 1], f.tables, stage_S[gi + 1], f.tb), tsize >> 1, tid);
            else
                sweep_real1<A, UNI, NT>(tile, g, Expand<1>(g), tsize >> 1, tid);
            return;
        }
    }
    run_op<A, UNI, NT>(tile, f.g[gi], f.g[gi + 1 < FUSED_MAX_OPS ? gi + 1 : gi], f.tables, stage_S, gi, base, f.tb, tsize, tid);
}

// BULK: tile staging with cp.async.bulk (TMA 1-D bulk copies, one per contiguous run, completion on
// an mbarrier) instead of LDG/STS through registers.  Requires runs of >= 16 bytes.
// NT threads per CTA: 256 for 2^12-amplitude tiles, 128 for 2^11 (twice as many CTAs per SM on the same shared memory)
// WIDE: 2^12-amplitude tiles swept by 128 threads, still 3 (complex128) / 4 (complex64) CTAs per SM -- the register
// budget per thread doubles (168 / 128), which is what the 16-amplitude register groups of the block pairs and the EXT
// forms need (the 256-thread EXT kernel spilled at its 80-register cap).
// SLIM (complex128 WIDE only): the same kernel capped at 128 registers -- the instantiation the CHUNK launches of the
// exchange / compute pipeline take, so that three pass CTAs (3 x 128 x 128 registers) and one persistent 256-thread remap
// CTA (16384) fit an SM's 65536 together.  The 168-register build is 2-4 % faster (the radix-16 QFT sweep needs 168
// registers and spills a little at 128: QFFT-33 372 against 385 ms, layered 242 against 246 ms on one box,
// profiles/r02_ab_regcap.txt) and serves every launch that has the SMs to itself.
template <typename A, bool BULK, bool UNI, int NT, bool EXT, bool WIDE = false, bool SLIM = false>
__global__ void __launch_bounds__(NT, WIDE ? (sizeof(A) == 16 ? (SLIM ? 4 : QIPB_WIDE_MINB) : QIPB_WIDE_MINB_C64) : (sizeof(A) == 16 ? 3 : 4) * (256 / NT)) fused_kernel(A *__restrict__ state, const __grid_constant__ FusedArgs f) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) unsigned long long bar;
    __shared__ __align__(16) double2 stage_S[FUSED_MAX_OPS + 1];
    A *tile = reinterpret_cast<A *>(smem_raw);
    const int tid = threadIdx.x;
    const u32 tsize = 1u << f.tb;
    const u32 lowmask = (1u << f.lowrun) - 1u;
    const u32 run_amps = 1u << f.lowrun;
    const u32 nruns = tsize >> f.lowrun;
    const u32 run_bytes = run_amps * (u32)sizeof(A);
    if (BULK) {
        if (tid == 0) {
            mbar_init(&bar, 1);
            fence_proxy_async();
        }
        __syncthreads();
    }
    u32 parity = 0;
    __shared__ unsigned int next_tile;
    unsigned int *const sched = f.sched;
    u64 t = blockIdx.x;
    if (sched) {
        if (tid == 0) next_tile = atomicAdd(&sched[0], 1u);
        __syncthreads();
        t = (u64)__shfl_sync(0xffffffffu, next_tile, 0);     // (the shuffle keeps the tile number warp-uniform for ptxas)
    }

    for (; t < f.ntiles;) {
        const u64 base = fused_tile_base(f, t);

        // ---- stage the tile: runs of 2^lowrun consecutive amplitudes ----
        if (EXT && f.g[0].diag == 5) {                          // fill mode: nothing is loaded, op 0 writes the tile
            if (tid < 32) stage_scalars<32>(f, base, stage_S, tid);
            __syncthreads();
        } else if (BULK) {
            if (tid < 32) {
                // no __syncwarp here: complete_tx may precede expect_tx (the phase cannot complete before
                // lane 0 arrives), and a warp-level sync makes ptxas give up warp-uniform descriptor loads
                if (tid == 0) mbar_expect_tx(&bar, tsize * (u32)sizeof(A));
                // optional (QIPB_FUSED_PREFETCH=1, off by default): pull this CTA's NEXT tile into L2 while this one
                // is computed and written back.  Measured on B200: no gain for compute-heavy passes and a worse
                // floor (profiles/r01_probe_fused_prefetch.txt), so it stays a profiling knob.
                u64 nbase = t + gridDim.x;
                const bool pf = f.prefetch && !sched && nbase < f.ntiles;
                if (pf) nbase = fused_tile_base(f, nbase);
                for (u32 r = tid; r < nruns; r += 32) {
                    u64 off = 0;
                    for (int j = f.lowrun; j < f.tb; ++j) off |= (u64)((r >> (j - f.lowrun)) & 1u) << f.tbit[j];
                    bulk_g2s(tile + (size_t)r * run_amps, state + base + off, run_bytes, &bar);
                    if (pf) bulk_prefetch_l2(state + nbase + off, run_bytes);
                }
                // stage scalars of this tile, computed by the issuing warp while the copies are in flight
                // (kept inside this warp's branch: a thread-dependent loop in the common path makes ptxas
                // drop the warp-uniform descriptor loads of the sweeps)
                if (f.nstages) stage_scalars<32>(f, base, stage_S, tid);
            }
            mbar_wait(&bar, parity);
            parity ^= 1u;
            if (f.nstages) __syncthreads();
        } else {
            if (f.nstages) stage_scalars<NT>(f, base, stage_S, tid);
            for (u32 e = tid; e < tsize; e += NT) {
                u64 off = e & lowmask;
                for (int j = f.lowrun; j < f.tb; ++j) off |= (u64)((e >> j) & 1u) << f.tbit[j];
                tile[e] = state[base + off];
            }
            __syncthreads();
        }

        // ---- run the gate list on the tile ----
        bool fetched = false, prefetched = false;
        for (int gi = 0; gi < f.ngates; ++gi) {
            if (fused_op_is_skipped<EXT>(f.g[gi])) continue;
            run_fused_op<A, UNI, NT, EXT>(tile, f, gi, stage_S, base, tsize, tid);
            __syncthreads();
            // the next tile of this CTA: fetched behind the first barrier of the tile (every warp has read the current
            // number by then), consumed behind the last one -- the atomic's latency hides under the sweeps
            if (sched && !fetched) {
                if (tid == 0) next_tile = atomicAdd(&sched[0], 1u);
                fetched = true;
            } else if (BULK && sched && f.prefetch && fetched && !prefetched) {
                // optional (QIPB_FUSED_PREFETCH=1): one barrier later every thread may read the next tile's number --
                // warp 0 pulls that tile into L2 while this one is still being swept, so the CTA's next load is an L2 hit
                prefetched = true;
                if (tid < 32) {
                    const u64 tn = (u64)next_tile;
                    if (tn < f.ntiles) {
                        const u64 nbase = fused_tile_base(f, tn);
                        for (u32 r = tid; r < nruns; r += 32) {
                            u64 off = 0;
                            for (int j = f.lowrun; j < f.tb; ++j) off |= (u64)((r >> (j - f.lowrun)) & 1u) << f.tbit[j];
                            bulk_prefetch_l2(state + nbase + off, run_bytes);
                        }
                    }
                }
            }
        }

        // ---- write the tile back ----
        if (BULK) {
            fence_proxy_async();          // generic-proxy writes -> visible to the async proxy
            __syncthreads();
            if (tid < 32) {
                for (u32 r = tid; r < nruns; r += 32) {
                    u64 off = 0;
                    for (int j = f.lowrun; j < f.tb; ++j) off |= (u64)((r >> (j - f.lowrun)) & 1u) << f.tbit[j];
                    bulk_s2g(state + base + off, tile + (size_t)r * run_amps, run_bytes);
                }
                bulk_commit_wait_read();  // smem may be overwritten once the copies have read it
            }
            __syncthreads();
        } else {
            for (u32 e = tid; e < tsize; e += NT) {
                u64 off = e & lowmask;
                for (int j = f.lowrun; j < f.tb; ++j) off |= (u64)((e >> j) & 1u) << f.tbit[j];
                state[base + off] = tile[e];
            }
            __syncthreads();
        }
        t = sched ? (u64)__shfl_sync(0xffffffffu, next_tile, 0) : t + gridDim.x;
    }
    if (sched && tid == 0) {                                  // the last CTA to leave re-arms the counters
        __threadfence();
        if (atomicAdd(&sched[1], 1u) == gridDim.x - 1u) {
            sched[0] = 0u;
            sched[1] = 0u;
        }
    }
}

// ---- persistent ring kernel: one CTA per SM, warp-specialised -------------------------------------
// 16 compute warps sweep ONE tile at a time; one I/O warp keeps a ring of NBUF tile buffers moving with
// TMA bulk copies: while tile k is being computed, tile k+1 (and k+2) are landing and tile k-1 is being
// written back.  full[b]: the load of buffer b has landed (expect_tx / complete_tx); done[b]: the compute
// warps are finished with buffer b.  Lane j of the I/O warp always moves the same runs of a buffer (r = j
// mod 32), so "my stores from this buffer have been read out" (wait_group.read) is all it needs before it
// refills the very same shared-memory bytes.
#define RING_COMPUTE 512
#define RING_THREADS (RING_COMPUTE + 32)

template <typename A, int NBUF>
__global__ void __launch_bounds__(RING_THREADS, 1) fused_ring_kernel(A *__restrict__ state, const __grid_constant__ FusedArgs f) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) unsigned long long full[NBUF], done[NBUF];
    __shared__ __align__(16) double2 stage_S[NBUF][FUSED_MAX_OPS + 1];   // per buffer: the I/O warp fills them at load time
    const int tid = threadIdx.x;
    const u32 tsize = 1u << f.tb;
    if (tid == 0) {
        for (int b = 0; b < NBUF; ++b) {
            mbar_init(&full[b], 32);           // every lane of the I/O warp arrives (lane 0 with the byte count)
            mbar_init(&done[b], 1);
        }
        fence_proxy_async();
    }
    __syncthreads();
    const u64 ntiles = f.ntiles, step = gridDim.x;
    // warp-uniform role id (the shuffle lets ptxas keep the role branch, and with it the descriptor
    // loads of the compute path, on the uniform datapath)
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);

    if (warp >= RING_COMPUTE / 32) {
        // ---------------- I/O warp ----------------
        const u32 lane = (u32)tid - RING_COMPUTE;
        const u32 run_amps = 1u << f.lowrun, nruns = tsize >> f.lowrun, run_bytes = run_amps * (u32)sizeof(A);
        u64 off[4];                        // global offsets of this lane's runs inside a tile (nruns <= 128)
        for (u32 i = 0; i < 4; ++i) {
            const u32 r = lane + 32u * i;
            u64 o = 0;
            for (int j = f.lowrun; j < f.tb; ++j) o |= (u64)((r >> (j - f.lowrun)) & 1u) << f.tbit[j];
            off[i] = o;
        }
        auto tile_base = [&](u64 t) { return fused_tile_base(f, t); };
        auto load_tile = [&](int b, u64 t) {
            A *tile = reinterpret_cast<A *>(smem_raw) + (size_t)b * tsize;
            const u64 base = tile_base(t);
            // stage scalars first: the mbarrier's completion then also publishes them (lane 0's arrive is a
            // release after its own writes; the other lanes' writes are ordered by the __syncwarp-free rule
            // below: every lane arrives on the barrier itself)
            if (f.nstages) stage_scalars<32>(f, base, stage_S[b], (int)lane);
            if (lane == 0) mbar_expect_tx(&full[b], tsize * (u32)sizeof(A));
            else mbar_arrive(&full[b]);
#pragma unroll
            for (u32 i = 0; i < 4; ++i) {
                const u32 r = lane + 32u * i;
                if (r < nruns) bulk_g2s(tile + (size_t)r * run_amps, state + base + off[i], run_bytes, &full[b]);
            }
        };
        {
            u64 t = blockIdx.x;
            for (int b = 0; b < NBUF && t < ntiles; ++b, t += step) load_tile(b, t);
        }
        u32 k = 0;
        for (u64 t = blockIdx.x; t < ntiles; t += step, ++k) {
            const int b = (int)(k % NBUF);
            mbar_wait(&done[b], (k / NBUF) & 1u);
            A *tile = reinterpret_cast<A *>(smem_raw) + (size_t)b * tsize;
            const u64 base = tile_base(t);
#pragma unroll
            for (u32 i = 0; i < 4; ++i) {
                const u32 r = lane + 32u * i;
                if (r < nruns) bulk_s2g(state + base + off[i], tile + (size_t)r * run_amps, run_bytes);
            }
            bulk_commit();
            const u64 tn = t + (u64)NBUF * step;
            if (tn < ntiles) {
                bulk_wait_read_all();      // this lane's stores have left the buffer: refill the same bytes
                load_tile(b, tn);
            }
        }
        bulk_wait_all();
    } else {
        // ---------------- compute warps ----------------
        u32 k = 0;
        for (u64 t = blockIdx.x; t < ntiles; t += step, ++k) {
            const int b = (int)(k % NBUF);
            A *tile = reinterpret_cast<A *>(smem_raw) + (size_t)b * tsize;
            const u64 base = fused_tile_base(f, t);
            mbar_wait(&full[b], (k / NBUF) & 1u);
            bool first = true;
            for (int gi = 0; gi < f.ngates; ++gi) {
                if (fused_op_is_skipped<false>(f.g[gi])) continue;
                if (!first) named_bar_sync(1, RING_COMPUTE);
                first = false;
                run_fused_op<A, true, RING_COMPUTE, false>(tile, f, gi, stage_S[b], base, tsize, tid);
            }
            fence_proxy_async();           // generic-proxy writes -> visible to the bulk store
            named_bar_sync(1, RING_COMPUTE);
            if (tid == 0) mbar_arrive(&done[b]);
        }
    }
}

static bool wide_enabled();
static bool ext_enabled() {
    // real 1-qubit forms and paired QFT steps (EXT sweeps): on with the WIDE kernel (measured on B200 in round 2: QFT
    // -8 %; the 256-thread EXT kernel spills and loses on layered passes, so without WIDE they stay opt-in); read per
    // call so that tests can toggle it.  Numerics are covered on the CPU tier by tests/test_fused_emul.py.
    const char *e = getenv("QIPB_FUSED_EXT");
    return e ? atoi(e) != 0 : wide_enabled();
}

static bool wide_enabled() {
    // default ON: 2^12 tiles run by the WIDE kernel (128 threads, block pairs, EXT forms); QIPB_FUSED_WIDE=0 gives the
    // 256-thread kernels of round 1 back (A/B runs); read per call so that tests can toggle it
    const char *e = getenv("QIPB_FUSED_WIDE");
    return !e || atoi(e) != 0;
}

static bool pair_enabled() {
    if (!QIPB_ENABLE_PAIRS) return false;                      // compiled out (see QIPB_ENABLE_PAIRS above)
    const char *e = getenv("QIPB_FUSED_PAIR");                // A/B knob of builds that carry them, read per call
    return !e || atoi(e) != 0;
}

static bool qftlow_enabled() {
    const char *e = getenv("QIPB_FUSED_QFTLOW");              // QFT steps on the lowest tile bits as lane butterflies; A/B knob
    return !e || atoi(e) != 0;
}

static bool short_runs_enabled() {
    const char *e = getenv("QIPB_FUSED_SHORT_RUNS");          // 1-2 diagonal gates behind a dense gate become a riding stage; A/B knob
    return !e || atoi(e) != 0;
}

static bool qft4_enabled() {
    const char *e = getenv("QIPB_FUSED_QFT4");                // four QFT steps per sweep (WIDE kernel); A/B knob, read per call
    return !e || atoi(e) != 0;
}

static bool ride2_enabled() {
    const char *e = getenv("QIPB_FUSED_RIDE2");               // stage riding on a dense 2-qubit sweep (WIDE kernel); A/B knob
    return !e || atoi(e) != 0;
}

static bool trio_enabled() {
    if (!QIPB_ENABLE_TRIOS) return false;
    const char *e = getenv("QIPB_FUSED_TRIO");                // A/B knob, read per call
    return !e || atoi(e) != 0;
}

static bool ring_enabled() {
    // opt-in (measured on B200, profiles/r01_probe_fused_ring.txt: one tile at a time on 16 warps sweeps
    // slower than three independent CTAs per SM do); read per call so that tests can toggle it
    const char *e = getenv("QIPB_FUSED_RING");
    return e && atoi(e) != 0;
}

static inline u32 tsize_runs(const FusedArgs &f) { return 1u << (f.tb - f.lowrun); }

// BULK: TMA staging (runs of >= 512 bytes, more than one tile).  UNI: the specialised sweeps (tile of 2^12, or
// 2^11 with 128-thread CTAs).  Also decides whether descriptors may carry the structured matrix forms.
static inline bool launch_is_bulk(const FusedArgs &f, size_t amp_bytes) { return (amp_bytes << f.lowrun) >= 512 && f.ntiles >= 2; }
static inline bool launch_is_uni(const FusedArgs &f, size_t amp_bytes) { return launch_is_bulk(f, amp_bytes) && f.tb >= 11; }
static bool wide_enabled();
static bool ring_enabled();
static bool pair_enabled();
static bool trio_enabled();
static bool ride2_enabled();
static bool qft4_enabled();
static bool short_runs_enabled();
static bool qftlow_enabled();
static inline bool launch_is_wide(const FusedArgs &f, size_t amp_bytes) {
    return launch_is_uni(f, amp_bytes) && f.tb == 12 && wide_enabled() && !ring_enabled();
}

// Exact structure of a dense 4x4 block (see Block2): fills the descriptor's coefficient area in the layout of
// the chosen form.  `put_c(slot, re, im)` / `put_r(slot, v)` write in the amplitude's precision.
template <typename PutC, typename PutR>
static void classify_block(const double *mat, DevGate &d, PutC put_c, PutR put_r) {
    cplx M[4][4];
    double big = 0.0;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            M[i][j] = cplx(mat[2 * (4 * i + j)], mat[2 * (4 * i + j) + 1]);
            big = std::max(big, std::abs(M[i][j]));
        }
    // monomial: exactly one non-zero per row and per column
    int row_of[4], rows_hit = 0;
    bool mono = true;
    for (int j = 0; j < 4 && mono; ++j) {
        int cnt = 0;
        for (int i = 0; i < 4; ++i)
            if (M[i][j] != cplx(0.0, 0.0)) { row_of[j] = i; ++cnt; }
        mono = cnt == 1;
        if (mono) rows_hit |= 1 << row_of[j];
    }
    if (mono && rows_hit == 15) {
        d.mk = MK_MONOMIAL;
        d.perm = d.phmask = 0;
        for (int j = 0; j < 4; ++j) {
            d.perm |= (unsigned char)(row_of[j] << (2 * j));
            const cplx c = M[row_of[j]][j];
            if (c != cplx(1.0, 0.0)) d.phmask |= (unsigned char)(1 << j);
            put_c(j, c.real(), c.imag());
        }
        return;
    }
    // real matrix times column phases: M[:, j] = ph_j * (real column); imaginary residue at rounding level only
    const double eps = 4e-16 * big;
    cplx ph[4];
    double R[4][4];
    bool rc = true;
    for (int j = 0; j < 4 && rc; ++j) {
        int k = 0;
        for (int i = 1; i < 4; ++i)
            if (std::abs(M[i][j]) > std::abs(M[k][j])) k = i;
        const double a = std::abs(M[k][j]);
        ph[j] = a > 0.0 ? M[k][j] / a : cplx(1.0, 0.0);
        if (ph[j].imag() == 0.0) ph[j] = cplx(1.0, 0.0);       // a real column keeps its signs in R
        for (int i = 0; i < 4; ++i) {
            const cplx z = M[i][j] * std::conj(ph[j]);
            if (std::fabs(z.imag()) > eps) { rc = false; break; }
            R[i][j] = z.real();
        }
    }
    if (rc) {
        d.phmask = 0;
        for (int j = 0; j < 4; ++j)
            if (ph[j] != cplx(1.0, 0.0)) d.phmask |= (unsigned char)(1 << j);
        d.mk = d.phmask ? MK_REALPHASE : MK_REAL;
        for (int i = 0; i < 4; ++i)
            for (int j = 0; j < 4; ++j) put_r(4 * i + j, R[i][j]);
        for (int j = 0; j < 4; ++j) put_c(8 + j, ph[j].real(), ph[j].imag());
        return;
    }
    d.mk = MK_GENERAL;
}

static bool dynsched_enabled() {
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("QIPB_FUSED_DYNSCHED");        // A/B knob: 0 = static tile split
        v = e ? atoi(e) : 1;
    }
    return v != 0;
}

// The kernels are compiled in two translation units so that `make -j` halves the build: this file as it is carries the
// complex128 instantiations and all host code; compiled with -DQIPB_FUSED_TU_C64 (-> fused_c64.o) it carries nothing but
// launch_fused<float2> and the complex64 kernels behind it.
template <typename A>
int launch_fused(qipb_ctx *ctx, A *state, const FusedArgs &f_in) {
    static thread_local FusedArgs f;                           // (a copy: the scheduling slot is per launch)
    f = f_in;
    f.sched = nullptr;
    if (dynsched_enabled()) {
        if (!ctx->sched_ring) {
            QIPB_CUDA(cudaMalloc(&ctx->sched_ring, QIPB_SCHED_SLOTS * 2 * sizeof(unsigned int)));
            QIPB_CUDA(cudaMemset(ctx->sched_ring, 0, QIPB_SCHED_SLOTS * 2 * sizeof(unsigned int)));
        }
        f.sched = ctx->sched_ring + 2 * (ctx->sched_slot++ % QIPB_SCHED_SLOTS);
    }
    const size_t smem = sizeof(A) << f.tb;
    const bool bulk = launch_is_bulk(f, sizeof(A));
    int per_sm = (int)((224u * 1024u) / (smem + 3072));
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 8) per_sm = 8;
    u64 grid = (u64)ctx->sm_count * per_sm;
    if (grid > f.ntiles) grid = f.ntiles;
    // UNI: all specialised sweeps (<= 4 fixed positions) have a multiple of the CTA size as item count
    const bool half = bulk && f.tb == 11;                      // 2^11 tiles: 128-thread CTAs
    const bool uni = launch_is_uni(f, sizeof(A));
#define QIPB_LAUNCH_FUSED(B, U, T, X, ...)                                                                                                  \
    do {                                                                                                                                    \
        QIPB_CUDA(cudaFuncSetAttribute(fused_kernel<A, B, U, T, X, ##__VA_ARGS__>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        fused_kernel<A, B, U, T, X, ##__VA_ARGS__><<<(unsigned)grid, T, smem, ctx->stream>>>(state, f);                                      \
    } while (0)
#ifdef QIPB_PROBE_WIDE_ONLY   /* compile-time probe (scripts/sass_probe.sh): only the WIDE complex128 kernel is instantiated */
    if constexpr (sizeof(A) == 16) QIPB_LAUNCH_FUSED(true, true, 128, true, true);
#else
    if (uni && !half && ring_enabled() && f.ntiles >= 4ull * (u64)ctx->sm_count && (tsize_runs(f) <= 128)) {
        constexpr int NBUF = sizeof(A) == 16 ? 3 : 6;
        const size_t rsmem = (size_t)NBUF * smem;
        QIPB_CUDA(cudaFuncSetAttribute(fused_ring_kernel<A, NBUF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rsmem));
        fused_ring_kernel<A, NBUF><<<(unsigned)ctx->sm_count, RING_THREADS, rsmem, ctx->stream>>>(state, f);
        ctx->ring_launches++;
    } else if (half) QIPB_LAUNCH_FUSED(true, true, 128, false);
    else if (launch_is_wide(f, sizeof(A))) {
        bool slim = false;
        if constexpr (sizeof(A) == 16) {
            if (f.nexp > f.tb) {                               // a chunk launch: it shares the SMs with a remap
                QIPB_LAUNCH_FUSED(true, true, 128, true, true, true);
                slim = true;
            }
        }
        if (!slim) QIPB_LAUNCH_FUSED(true, true, 128, true, true);
        ctx->ext_launches += fused_has_ext(f) ? 1 : 0;
    } else if (uni && fused_has_ext(f)) {                      // the host marks EXT ops only for this launch shape
        QIPB_LAUNCH_FUSED(true, true, 256, true);
        ctx->ext_launches++;
    } else if (uni) QIPB_LAUNCH_FUSED(true, true, 256, false);
    else if (bulk) QIPB_LAUNCH_FUSED(true, false, 256, false);
    else QIPB_LAUNCH_FUSED(false, false, 256, false);
#endif
#undef QIPB_LAUNCH_FUSED
    ctx->launches++;
    QIPB_CUDA(cudaGetLastError());
    return QIPB_OK;
}


#ifdef QIPB_FUSED_TU_C64
template int launch_fused<float2>(qipb_ctx *ctx, float2 *state, const FusedArgs &f_in);
}  // namespace qipb
#else
#ifndef QIPB_FUSED_SINGLE_TU                                  /* (the test harness tests/csrc/fused_emul.cu is one translation unit) */
extern template int launch_fused<float2>(qipb_ctx *ctx, float2 *state, const FusedArgs &f_in);     // fused_c64.o
#endif

// ---- host side: folding runs of diagonal gates into stages -------------------------------------
static bool post_enabled() {
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("QIPB_FUSED_POST");            // tuning knob for profiling runs
        v = e ? atoi(e) : 1;
    }
    return v != 0;
}

bool stages_enabled() {
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("QIPB_FUSED_STAGES");          // tuning knob for profiling runs
        v = e ? atoi(e) : 1;
    }
    return v != 0;
}

// cell of a state bit: 0 = tile-lo, 1 = tile-hi, 2.. = groups of 7 outside bits (ascending)
struct CellMap {
    int cell_of[64];
    int idx_in_cell[64];
    std::vector<int> bits[2 + FUSED_OUT_CELLS];
    int ncells;
};

static int residual_cell(const qipb_gate &g, u64 common, const CellMap &cm) {
    // -1: no residual bits; -2: spans several cells; else the cell id
    int cell = -1;
    u64 rest = g.ctrl_mask & ~common;
    for (int j = 0; j < g.k; ++j) rest |= 1ull << g.bits[j];
    for (int b = 0; b < 64; ++b)
        if ((rest >> b) & 1ull) {
            const int c = cm.cell_of[b];
            if (c < 0) return -2;
            if (cell == -1) cell = c;
            else if (cell != c) return -2;
        }
    return cell;
}

void build_stages(const qipb_gate *gates, const std::vector<int> &run, int nbits, int tb, const int *local_of,
                  u64 tmask, std::vector<Op> &ops, std::vector<cplx> &tables, int min_run) {
    CellMap cm;
    for (int b = 0; b < 64; ++b) { cm.cell_of[b] = -1; cm.idx_in_cell[b] = 0; }
    const int lo = tb < FUSED_LO_BITS ? tb : FUSED_LO_BITS;
    int nout_bits = 0;
    for (int b = 0; b < nbits; ++b) {
        int c;
        if (local_of[b] >= 0) c = local_of[b] < lo ? 0 : 1;
        else c = 2 + nout_bits++ / 7;
        if (c >= 2 + FUSED_OUT_CELLS) continue;               // beyond the supported outside cells: not table-izable
        cm.cell_of[b] = c;
        cm.idx_in_cell[b] = (int)cm.bits[c].size();
        cm.bits[c].push_back(b);
    }
    // tile cells are indexed by tile-local position (lo: e & 63, hi: e >> 6)
    for (int b = 0; b < nbits; ++b)
        if (local_of[b] >= 0) cm.idx_in_cell[b] = local_of[b] < lo ? local_of[b] : local_of[b] - lo;
    cm.ncells = 2 + (nout_bits + 6) / 7;
    if (cm.ncells > 2 + FUSED_OUT_CELLS) cm.ncells = 2 + FUSED_OUT_CELLS;

    size_t pos = 0;
    while (pos < run.size()) {
        // grow a stage greedily: the common control set may only shrink, every member must stay single-cell
        u64 common = gates[run[pos]].ctrl_mask;
        size_t end = pos + 1;
        {
            // a single gate is "single-cell" w.r.t. its own controls iff its targets are
            if (residual_cell(gates[run[pos]], common, cm) == -2) common = 0;
        }
        while (end < run.size()) {
            const u64 nc = common & gates[run[end]].ctrl_mask;
            bool ok = true;
            for (size_t t = pos; t <= end && ok; ++t) ok = residual_cell(gates[run[t]], nc, cm) != -2;
            if (!ok) break;
            common = nc;
            ++end;
        }
        if ((int)(end - pos) < min_run || residual_cell(gates[run[pos]], common, cm) == -2) {
            // too short to pay for tables (or not table-izable at all): keep the gate as it is
            Op o;
            o.stage = false;
            o.gate = run[pos];
            ops.push_back(o);
            ++pos;
            continue;
        }
        // tables: T_lo, T_hi, then the outside cells that are actually used
        std::vector<std::vector<cplx>> T(cm.ncells);
        std::vector<bool> used(cm.ncells, false);
        used[0] = used[1] = true;
        for (int c = 0; c < cm.ncells; ++c) T[c].assign((size_t)1 << (c == 0 ? lo : c == 1 ? tb - lo : (int)cm.bits[c].size()), cplx(1.0, 0.0));
        for (size_t t = pos; t < end; ++t) {
            const qipb_gate &g = gates[run[t]];
            int c = residual_cell(g, common, cm);
            if (c == -1) c = 0;
            used[c] = true;
            const int D = 1 << g.k;
            const u64 rc = g.ctrl_mask & ~common;
            for (size_t v = 0; v < T[c].size(); ++v) {
                bool on = true;
                for (int b = 0; b < nbits && on; ++b)
                    if ((rc >> b) & 1ull) on = (v >> cm.idx_in_cell[b]) & 1u;
                if (!on) continue;
                int sel = 0;
                for (int j = 0; j < g.k; ++j)
                    if ((v >> cm.idx_in_cell[g.bits[j]]) & 1u) sel |= 1 << (g.k - 1 - j);
                T[c][v] *= cplx(g.mat[2 * (sel * D + sel)], g.mat[2 * (sel * D + sel) + 1]);
            }
        }
        Op o;
        o.stage = true;
        o.gate = -1;
        o.common = common;
        o.support = 0;
        for (size_t t = pos; t < end; ++t) {
            const qipb_gate &g = gates[run[t]];
            o.support |= g.ctrl_mask;
            for (int j = 0; j < g.k; ++j) o.support |= 1ull << g.bits[j];
        }
        o.tab_off = (u32)tables.size();
        o.nout = 0;
        tables.insert(tables.end(), T[0].begin(), T[0].end());
        tables.insert(tables.end(), T[1].begin(), T[1].end());
        for (int c = 2; c < cm.ncells; ++c)
            if (used[c]) {
                o.cells[o.nout] = cm.bits[c];
                tables.insert(tables.end(), T[c].begin(), T[c].end());
                o.nout++;
            }
        ops.push_back(o);
        pos = end;
    }
}

// Stage tables: pinned staging ring -> device buffer, stream ordered.
int upload_tables(qipb_ctx *ctx, const std::vector<cplx> &tables, const double2 **out) {
    *out = nullptr;
    if (tables.empty()) return QIPB_OK;
    const size_t need = tables.size();
    if (ctx->tab_cap < need) {
        QIPB_CUDA(cudaStreamSynchronize(ctx->stream));
        size_t cap = 1u << 16;
        while (cap < need) cap <<= 1;
        if (ctx->tab_dev) QIPB_CUDA(cudaFree(ctx->tab_dev));
        ctx->tab_dev = nullptr;
        QIPB_CUDA(cudaMalloc(&ctx->tab_dev, cap * sizeof(double2)));
        for (int i = 0; i < 4; ++i) {
            if (ctx->tab_host[i]) QIPB_CUDA(cudaFreeHost(ctx->tab_host[i]));
            ctx->tab_host[i] = nullptr;
            QIPB_CUDA(cudaMallocHost(&ctx->tab_host[i], cap * sizeof(double2)));
            if (!ctx->tab_ev[i]) QIPB_CUDA(cudaEventCreateWithFlags(&ctx->tab_ev[i], cudaEventDisableTiming));
        }
        ctx->tab_cap = cap;
    }
    const int slot = ctx->tab_slot;
    ctx->tab_slot = (slot + 1) & 3;
    QIPB_CUDA(cudaEventSynchronize(ctx->tab_ev[slot]));      // the copy that last used this slot is done
    memcpy(ctx->tab_host[slot], tables.data(), need * sizeof(double2));
    QIPB_CUDA(cudaMemcpyAsync(ctx->tab_dev, ctx->tab_host[slot], need * sizeof(double2), cudaMemcpyHostToDevice, ctx->stream));
    QIPB_CUDA(cudaEventRecord(ctx->tab_ev[slot], ctx->stream));
    *out = ctx->tab_dev;
    return QIPB_OK;
}

// Validation and lowering of one fused pass: gate list -> device descriptors, FUSED_MAX_OPS per launch.  No CUDA
// calls in here: `prepare(tables, f)` makes the stage tables reachable through f.tables (the library uploads them,
// tests/csrc/fused_emul.cu points at the host vector) and `launch(f)` consumes one filled FusedArgs.
// fill: the pass acts on the all-ones vector (qipb_apply_fused_fill); the first op must then be a stage without
// controls, executed by the EXT kernel in fill mode -- otherwise QIPB_ERR_UNSUPPORTED and nothing is launched.
struct FusedChunk {            // restriction of a launch to the amplitudes whose `bits` have `value` (index-bit mask)
    int nfix;
    const int *bits;
    u64 value;
};

template <typename Prepare, typename Launch>
static int lower_fused(int nbits, int dtype, int ntile_bits, const int *tile_bits, int ngates, const qipb_gate *gates,
                       Prepare prepare, Launch launch, bool fill = false, FusedChunk chunk = FusedChunk{0, nullptr, 0}) {
    QIPB_REQUIRE(gates && tile_bits, "null argument");
    QIPB_REQUIRE(nbits >= 0 && nbits <= 40, "nbits %d unsupported", nbits);
    QIPB_REQUIRE(ntile_bits >= 0 && ntile_bits <= QIPB_MAX_TILE_BITS && ntile_bits <= nbits, "tile bits %d unsupported", ntile_bits);
    QIPB_REQUIRE(ngates >= 1 && ngates <= QIPB_MAX_FUSED_GATES, "ngates %d unsupported (1..%d)", ngates, QIPB_MAX_FUSED_GATES);
    QIPB_REQUIRE(dtype == QIPB_C128 || dtype == QIPB_C64, "unknown dtype %d", dtype);
    static thread_local FusedArgs f;    // ~29 KiB: keep it off the stack
    memset(&f, 0, sizeof(f));
    f.nbits = nbits;
    f.tb = ntile_bits;
    f.ntiles = 1ull << (nbits - ntile_bits);
    {
        const char *e = getenv("QIPB_FUSED_PREFETCH");        // tuning knob for profiling runs
        f.prefetch = e ? atoi(e) : 0;   // measured: no gain, the floor gets worse (profiles/r01_probe_fused_prefetch.txt)
    }
    int local_of[64];
    for (int b = 0; b < 64; ++b) local_of[b] = -1;
    u64 tmask = 0;
    for (int j = 0; j < ntile_bits; ++j) {
        QIPB_REQUIRE(tile_bits[j] >= 0 && tile_bits[j] < nbits, "tile bit %d out of range", tile_bits[j]);
        QIPB_REQUIRE(j == 0 || tile_bits[j] > tile_bits[j - 1], "tile bits must be ascending and distinct");
        f.tbit[j] = (unsigned char)tile_bits[j];
        local_of[tile_bits[j]] = j;
        tmask |= 1ull << tile_bits[j];
    }
    f.lowrun = 0;
    while (f.lowrun < ntile_bits && tile_bits[f.lowrun] == f.lowrun) f.lowrun++;
    {
        QIPB_REQUIRE(chunk.nfix >= 0 && chunk.nfix <= 4 && (chunk.nfix == 0 || chunk.bits) && ntile_bits + chunk.nfix <= nbits,
                     "chunk: %d fixed bits unsupported (0..4)", chunk.nfix);
        u64 fmask = 0;
        for (int j = 0; j < chunk.nfix; ++j) {
            const int b = chunk.bits[j];
            QIPB_REQUIRE(b >= 0 && b < nbits && !((tmask >> b) & 1ull) && !((fmask >> b) & 1ull), "chunk: bad fixed bit %d", b);
            fmask |= 1ull << b;
        }
        QIPB_REQUIRE((chunk.value & ~fmask) == 0, "chunk: value has bits outside the fixed bits");
        f.nexp = 0;
        for (int b = 0; b < nbits; ++b)
            if (((tmask | fmask) >> b) & 1ull) f.ebit[f.nexp++] = (unsigned char)b;
        f.fix_value = chunk.value;
        f.ntiles = 1ull << (nbits - ntile_bits - chunk.nfix);
    }
    // ---- pass 1: validate, and fold runs of diagonal gates into stages ----
    std::vector<Op> ops;
    std::vector<cplx> tables;
    {
        std::vector<int> run;
        auto flush_run = [&]() {
            // a run of >= 3 diagonal gates pays for its tables as a stage of its own; a shorter run only where it can ride on
            // the sweep of the un-controlled dense gate right in front of it (then it costs no sweep at all)
            int min_run = 3;
            if (!ops.empty() && !ops.back().stage && short_runs_enabled()) {
                const qipb_gate &pg = gates[ops.back().gate];
                if (!(pg.diagonal != 0 || pg.k == 0) && pg.ctrl_mask == 0) min_run = 1;
            }
            if (!run.empty()) build_stages(gates, run, nbits, ntile_bits, local_of, tmask, ops, tables, min_run);
            run.clear();
        };
        for (int gi = 0; gi < ngates; ++gi) {
            const qipb_gate &s = gates[gi];
            QIPB_REQUIRE(s.k >= 0 && s.k <= 2, "fused gate %d: k=%d unsupported", gi, s.k);
            u64 tgt = 0;
            const bool diag = s.diagonal != 0 || s.k == 0;
            for (int j = 0; j < s.k; ++j) {
                const int b = s.bits[j];
                QIPB_REQUIRE(b >= 0 && b < nbits && !((tgt >> b) & 1ull), "fused gate %d: bad target bit %d", gi, b);
                tgt |= 1ull << b;
                QIPB_REQUIRE(diag || local_of[b] >= 0, "fused gate %d: non-diagonal target bit %d is not a tile bit", gi, b);
            }
            QIPB_REQUIRE((s.ctrl_mask & tgt) == 0, "fused gate %d: control mask overlaps targets", gi);
            QIPB_REQUIRE(nbits == 64 || (s.ctrl_mask >> nbits) == 0, "fused gate %d: control outside local bits", gi);
            if (diag && stages_enabled()) {
                run.push_back(gi);
            } else {
                flush_run();
                Op o;
                o.gate = gi;
                o.stage = false;
                ops.push_back(o);
            }
        }
        flush_run();
    }
    if (fill) {
        const bool shape_ok = launch_is_uni(f, dtype == QIPB_C128 ? 16 : 8) && f.tb == 12 && !ring_enabled();
        if (!shape_ok || ops.empty() || !ops[0].stage || ops[0].common != 0) {
            set_error("fill mode needs 2^12-amplitude tiles with runs of >= 512 bytes and a leading run of un-controlled diagonal gates");
            return QIPB_ERR_UNSUPPORTED;
        }
    }
    {
        int rc = prepare(tables, f);
        if (rc) return rc;
    }

    // ---- pass 2: device descriptors, FUSED_MAX_OPS per launch ----
    static const bool structured = []() {
        const char *e = getenv("QIPB_FUSED_STRUCTURED");      // tuning knob for profiling runs
        return e ? atoi(e) != 0 : true;
    }();
    for (size_t first = 0; first < ops.size(); first += FUSED_MAX_OPS) {
        const size_t cnt = ops.size() - first < FUSED_MAX_OPS ? ops.size() - first : FUSED_MAX_OPS;
        f.ngates = (int)cnt;
        memset(f.g, 0, sizeof(f.g));
        for (size_t oi = 0; oi < cnt; ++oi) {
            const Op &o = ops[first + oi];
            DevGate &d = f.g[oi];
            const u64 ctrl_mask = o.stage ? o.common : gates[o.gate].ctrl_mask;
            u64 fixed_local = 0;
            if (o.stage) {
                d.k = 0;
                d.kin = 0;
                d.diag = 2;
                StageInfo &si = *reinterpret_cast<StageInfo *>(d.m);
                si.tab_off = o.tab_off;
                si.nout = (unsigned char)o.nout;
                for (int c = 0; c < o.nout; ++c) {
                    si.cn[c] = (unsigned char)o.cells[c].size();
                    for (size_t j = 0; j < o.cells[c].size(); ++j) si.cb[c][j] = (unsigned char)o.cells[c][j];
                }
            } else {
                const qipb_gate &s = gates[o.gate];
                d.k = (unsigned char)s.k;
                d.diag = (unsigned char)(s.diagonal != 0 || s.k == 0);
                d.kin = 0;
                for (int j = 0; j < s.k; ++j) {
                    const int b = s.bits[j];
                    d.tg[j] = (unsigned char)b;
                    if (local_of[b] >= 0) {
                        d.tl[j] = (unsigned char)local_of[b];
                        fixed_local |= 1ull << local_of[b];
                        d.kin++;
                    } else {
                        d.tl[j] = 0xFF;
                    }
                }
                const int D = 1 << s.k;
                if (dtype == QIPB_C64) {            // the kernel reads float2 over the same bytes (coef<float2>)
                    float2 *mf = reinterpret_cast<float2 *>(d.m);
                    for (int e = 0; e < D * D; ++e) mf[e] = make_float2((float)s.mat[2 * e], (float)s.mat[2 * e + 1]);
                } else {
                    for (int e = 0; e < D * D; ++e) d.m[e] = make_double2(s.mat[2 * e], s.mat[2 * e + 1]);
                }
            }
            d.out_ctrl = ctrl_mask & ~tmask;
            d.in_or = 0;
            for (int b = 0; b < nbits; ++b)
                if (((ctrl_mask & tmask) >> b) & 1ull) {
                    d.in_or |= 1u << local_of[b];
                    fixed_local |= 1ull << local_of[b];
                }
            d.nins = 0;
            for (int j = 0; j < ntile_bits; ++j)
                if ((fixed_local >> j) & 1ull) d.nmask[d.nins++] = ~((1u << j) - 1u);
            // structured forms of a dense 2-qubit block (only the specialised sweeps without in-tile controls read them)
            if (!o.stage && !d.diag && d.k == 2 && d.nins == 2 && structured && launch_is_uni(f, dtype == QIPB_C128 ? 16 : 8)) {
                const double *mat = gates[o.gate].mat;
                if (dtype == QIPB_C64) {
                    float2 *mc = reinterpret_cast<float2 *>(d.m);
                    float *mr = reinterpret_cast<float *>(d.m);
                    float2 keep[16];
                    memcpy(keep, mc, sizeof(keep));
                    classify_block(mat, d, [&](int slot, double re, double im) { mc[slot] = make_float2((float)re, (float)im); },
                                   [&](int slot, double v) { mr[slot] = (float)v; });
                    if (d.mk == MK_GENERAL) memcpy(mc, keep, sizeof(keep));
                } else {
                    double2 keep[16];
                    memcpy(keep, d.m, sizeof(keep));
                    double *mr = reinterpret_cast<double *>(d.m);
                    classify_block(mat, d, [&](int slot, double re, double im) { d.m[slot] = make_double2(re, im); },
                                   [&](int slot, double v) { mr[slot] = v; });
                    if (d.mk == MK_GENERAL) memcpy(d.m, keep, sizeof(keep));
                }
            }
        }
        if (fill && first == 0) f.g[0].diag = 5;               // written, not multiplied (sweep_stage_fill)
        if (post_enabled())
            for (size_t oi = 0; oi + 1 < cnt; ++oi) {
                DevGate &d = f.g[oi], &nx = f.g[oi + 1];
                if (!d.diag && d.k == 1 && d.nins == 1 && d.out_ctrl == 0 && nx.diag == 2) {
                    d.post = 1;
                    nx.diag = 3;
                } else if (ride2_enabled() && launch_is_wide(f, dtype == QIPB_C128 ? 16 : 8) && !d.diag && d.k == 2 && d.nins == 2 &&
                           d.out_ctrl == 0 && d.in_or == 0 && d.mk != MK_MONOMIAL && d.tl[0] != 0xFF && d.tl[1] != 0xFF &&
                           d.tl[0] >= (dtype == QIPB_C128 ? 3 : 4) && d.tl[1] >= (dtype == QIPB_C128 ? 3 : 4) &&
                           nx.diag == 2 && nx.out_ctrl == 0) {
                    d.post = 8;                                // the stage rides on the block's sweep (sweep_dense2_stage)
                    nx.diag = 3;
                }
            }
        // opt-in forms (QIPB_FUSED_EXT=1): real dense 1-qubit gates and paired QFT steps, only for the launch shape
        // of the specialised kernel (2^12 tiles, 256 threads) and only where the EXT sweeps' preconditions hold
        if (ext_enabled() && !ring_enabled() && launch_is_uni(f, dtype == QIPB_C128 ? 16 : 8) && f.tb == 12) {
            const int lowb = dtype == QIPB_C128 ? 3 : 4;       // LowBits<A>: targets below take the bank-conflict-free sweeps
            for (size_t oi = 0; oi < cnt; ++oi) {
                DevGate &d = f.g[oi];
                const Op &o = ops[first + oi];
                if (o.stage || d.diag || d.k != 1 || d.nins != 1 || d.tl[0] < lowb) continue;
                const double *mat = gates[o.gate].mat;         // row-major complex 2x2
                if (mat[1] != 0.0 || mat[3] != 0.0 || mat[5] != 0.0 || mat[7] != 0.0) continue;
                if (d.post == 1 && f.g[oi + 1].in_or != (1u << d.tl[0])) continue;   // stage not controlled by exactly the target
                d.mk = MK1_REAL;
                d.phmask = (mat[0] == mat[2] && mat[0] == mat[4] && mat[0] == -mat[6] && mat[0] != 0.0) ? 1 : 0;
                if (dtype == QIPB_C64) {
                    float *mr = reinterpret_cast<float *>(d.m);
                    for (int e = 0; e < 4; ++e) mr[e] = (float)mat[2 * e];
                } else {
                    double *mr = reinterpret_cast<double *>(d.m);
                    for (int e = 0; e < 4; ++e) mr[e] = mat[2 * e];
                }
            }
            // a QFT step = Hadamard (post == 1, exact s * [[1, 1], [1, -1]]) + its stage (diag == 3, controlled by exactly the
            // Hadamard's target, nothing outside the tile)
            auto qft_step = [&](size_t oi) {
                if (oi + 1 >= cnt) return false;
                const DevGate &a = f.g[oi], &sa = f.g[oi + 1];
                return a.post == 1 && a.mk == MK1_REAL && a.phmask && sa.diag == 3 && sa.out_ctrl == 0 && ops[first + oi + 1].stage;
            };
            const bool wide = launch_is_wide(f, dtype == QIPB_C128 ? 16 : 8);
            const int lo_bits = f.tb < FUSED_LO_BITS ? f.tb : FUSED_LO_BITS;
            // M_j[k1 without bit j] of sweep_qft4 for the stage at op `so`, member positions pos[0..3]; false if the stage's
            // tile tables do not factorise over the member bits
            auto member_factors = [&](size_t so, const unsigned char *pos, int j, cplx *out8) {
                const cplx *Tlo = tables.data() + ops[first + so].tab_off, *Thi = Tlo + ((size_t)1 << lo_bits);
                u32 mlo = 0, mhi = 0;
                for (int t = 0; t < 4; ++t) {
                    if (pos[t] < lo_bits) mlo |= 1u << pos[t];
                    else mhi |= 1u << (pos[t] - lo_bits);
                }
                auto factorises = [&](const cplx *T, u32 size, u32 mask) {
                    if (T[0] == cplx(0.0, 0.0)) return false;
                    for (u32 v = 0; v < size; ++v) {
                        const cplx lhs = T[v] * T[0], rhs = T[v & mask] * T[v & ~mask];
                        if (std::abs(lhs - rhs) > 1e-13 * (std::abs(lhs) + std::abs(rhs) + 1e-300)) return false;
                    }
                    return true;
                };
                if (!factorises(Tlo, 1u << lo_bits, mlo) || !factorises(Thi, 1u << (f.tb - lo_bits), mhi)) return false;
                for (int k1 = 0; k1 < 16; ++k1) {
                    if (!(k1 & (1 << j))) continue;
                    u32 vlo = 0, vhi = 0;
                    for (int t = 0; t < 4; ++t)
                        if (k1 & (1 << t)) {
                            if (pos[t] < lo_bits) vlo |= 1u << pos[t];
                            else vhi |= 1u << (pos[t] - lo_bits);
                        }
                    const int idx = (k1 & ((1 << j) - 1)) | ((k1 >> (j + 1)) << j);
                    out8[idx] = (Tlo[vlo] / Tlo[0]) * (Thi[vhi] / Thi[0]);
                }
                return true;
            };
            // QFT steps on the bank-conflict bits: Hadamard-like dense 1-qubit gate (general descriptor form: the real forms
            // are only marked above those bits) + its stage, up to three in a row -> one lane-butterfly sweep
            auto low_step = [&](size_t oi) {
                if (oi + 1 >= cnt) return false;
                const DevGate &a = f.g[oi], &sa = f.g[oi + 1];
                if (!(a.post == 1 && !a.diag && a.k == 1 && a.nins == 1 && a.tl[0] < lowb && a.out_ctrl == 0 && a.in_or == 0)) return false;
                if (!(sa.diag == 3 && sa.out_ctrl == 0 && ops[first + oi + 1].stage) || ops[first + oi].stage) return false;
                const double *mat = gates[ops[first + oi].gate].mat;
                return mat[1] == 0.0 && mat[3] == 0.0 && mat[5] == 0.0 && mat[7] == 0.0 && mat[0] != 0.0 && mat[0] == mat[2] &&
                       mat[0] == mat[4] && mat[0] == -mat[6];
            };
            if (wide && qftlow_enabled())
                for (size_t oi = 0; oi + 1 < cnt;) {
                    int nst = 0;
                    while (nst < 3 && low_step(oi + 2 * nst)) {
                        bool distinct = true;
                        for (int t = 0; t < nst; ++t) distinct = distinct && f.g[oi + 2 * t].tl[0] != f.g[oi + 2 * nst].tl[0];
                        if (!distinct) break;
                        ++nst;
                    }
                    if (nst == 0) {
                        ++oi;
                        continue;
                    }
                    f.g[oi].post = 10;
                    f.g[oi].pair = (unsigned char)nst;
                    for (int t = 1; t < nst; ++t) f.g[oi + 2 * t].post = 3;
                    oi += 2 * nst;
                }
            for (size_t oi = 0; oi + 3 < cnt;) {
                if (wide && qft4_enabled() && oi + 7 < cnt && qft_step(oi) && qft_step(oi + 2) && qft_step(oi + 4) && qft_step(oi + 6)) {
                    const unsigned char pos[4] = {f.g[oi].tl[0], f.g[oi + 2].tl[0], f.g[oi + 4].tl[0], f.g[oi + 6].tl[0]};
                    bool ok = pos[0] != pos[1] && pos[0] != pos[2] && pos[0] != pos[3] && pos[1] != pos[2] && pos[1] != pos[3] && pos[2] != pos[3];
                    cplx M[4][8];
                    for (int j = 0; j < 4 && ok; ++j) ok = member_factors(oi + 2 * j + 1, pos, j, M[j]);
                    if (ok) {
                        for (int j = 0; j < 4; ++j) {
                            DevGate &h = f.g[oi + 2 * j];
                            for (int t = 0; t < 8; ++t) h.m[8 + t] = make_double2(M[j][t].real(), M[j][t].imag());
                            h.post = j == 0 ? 9 : 3;
                        }
                        unsigned char srt[4] = {pos[0], pos[1], pos[2], pos[3]};
                        for (int x = 1; x < 4; ++x)
                            for (int y = x; y > 0 && srt[y] < srt[y - 1]; --y) std::swap(srt[y], srt[y - 1]);
                        for (int x = 0; x < 4; ++x) f.g[oi].nmask[2 + x] = ~((1u << srt[x]) - 1u);
                        oi += 8;
                        continue;
                    }
                }
                DevGate &a = f.g[oi], &b = f.g[oi + 2];
                if (qft_step(oi) && qft_step(oi + 2) && a.tl[0] != b.tl[0]) {
                    a.post = 2;
                    b.post = 3;
                    oi += 4;
                } else {
                    ++oi;
                }
            }
        }
        // block pairs (WIDE launches): two dense 2-qubit blocks on disjoint targets share one sweep over 16-amplitude
        // register groups (sweep_pair2).  The partner may sit later in the list: it is executed early, at the leader's
        // position, which is exact when it commutes with every op in between -- the non-diagonal targets of either
        // side must avoid everything the other side reads (targets and controls; diagonal ops only read).
        if ((pair_enabled() || trio_enabled()) && launch_is_wide(f, dtype == QIPB_C128 ? 16 : 8)) {
            const int lowb = dtype == QIPB_C128 ? 3 : 4;       // LowBits<A>
            std::vector<int> origin(cnt);                      // position in f.g -> index of its Op (positions move below)
            for (size_t oi = 0; oi < cnt; ++oi) origin[oi] = (int)(first + oi);
            auto support = [&](size_t oi) -> u64 {
                const Op &o = ops[origin[oi]];
                if (o.stage) return o.support;
                const qipb_gate &g = gates[o.gate];
                u64 m = g.ctrl_mask;
                for (int j = 0; j < g.k; ++j) m |= 1ull << g.bits[j];
                return m;
            };
            auto nondiag = [&](size_t oi) -> u64 {
                const Op &o = ops[origin[oi]];
                if (o.stage) return 0;
                const qipb_gate &g = gates[o.gate];
                if (g.diagonal || g.k == 0) return 0;
                u64 m = 0;
                for (int j = 0; j < g.k; ++j) m |= 1ull << g.bits[j];
                return m;
            };
            auto block2 = [&](size_t oi) {                     // a dense 2-qubit block the register-group sweeps can take
                const DevGate &d = f.g[oi];
                return !ops[origin[oi]].stage && !d.diag && d.k == 2 && d.nins == 2 && d.post == 0 && d.in_or == 0 &&
                       d.tl[0] != 0xFF && d.tl[1] != 0xFF && d.tl[0] >= lowb && d.tl[1] >= lowb;
            };
            auto lone1 = [&](size_t oi) {                      // a lone dense 1-qubit gate (no stage rides on it)
                const DevGate &d = f.g[oi];
                return !ops[origin[oi]].stage && !d.diag && d.k == 1 && d.nins == 1 && d.post == 0 && d.in_or == 0 &&
                       d.tl[0] != 0xFF && d.tl[0] >= lowb;
            };
            auto move_next_to = [&](size_t i, size_t j) {      // op j -> position i + 1 (the ops in between shift by one)
                const DevGate moved = f.g[j];
                const int moved_from = origin[j];
                for (size_t t = j; t > i + 1; --t) {
                    f.g[t] = f.g[t - 1];
                    origin[t] = origin[t - 1];
                }
                f.g[i + 1] = moved;
                origin[i + 1] = moved_from;
            };
            auto set_masks = [&](DevGate &a, const unsigned char *pos, int npos) {
                unsigned char srt[4];
                for (int x = 0; x < npos; ++x) srt[x] = pos[x];
                for (int x = 1; x < npos; ++x)
                    for (int y = x; y > 0 && srt[y] < srt[y - 1]; --y) std::swap(srt[y], srt[y - 1]);
                for (int x = 0; x < npos; ++x) a.nmask[2 + x] = ~((1u << srt[x]) - 1u);
            };
            // the partner is executed at the leader's position: exact when it commutes with every op in between -- the
            // non-diagonal targets of either side must avoid everything the other side reads (targets and controls)
            auto find_partner = [&](size_t i, bool want_block) -> size_t {
                u64 mid_support = 0, mid_nondiag = 0;
                for (size_t j = i + 1; j < cnt; ++j) {
                    const bool kind_ok = want_block ? block2(j) : lone1(j);
                    if (kind_ok && !(nondiag(j) & (support(i) | mid_support)) && !(support(j) & (nondiag(i) | mid_nondiag))) return j;
                    mid_support |= support(j);
                    mid_nondiag |= nondiag(j);
                }
                return cnt;
            };
            if (pair_enabled())
                for (size_t i = 0; i + 1 < cnt; ++i) {
                    if (!block2(i)) continue;
                    const size_t j = find_partner(i, true);
                    if (j >= cnt) continue;
                    move_next_to(i, j);
                    DevGate &a = f.g[i], &b = f.g[i + 1];
                    const unsigned char pos4[4] = {a.tl[0], a.tl[1], b.tl[0], b.tl[1]};
                    set_masks(a, pos4, 4);
                    a.post = 4;
                    a.pair = 1;
                    b.post = 5;
                    ++i;                                       // the partner is taken
                }
            if (trio_enabled())
                for (size_t i = 0; i + 1 < cnt; ++i) {
                    const bool is_block = block2(i), is_lone = lone1(i);
                    if (!is_block && !is_lone) continue;
                    const size_t j = find_partner(i, !is_block);
                    if (j >= cnt) continue;
                    move_next_to(i, j);
                    if (!is_block) {                           // the block leads (they commute: disjoint bits, no in-tile controls)
                        std::swap(f.g[i], f.g[i + 1]);
                        std::swap(origin[i], origin[i + 1]);
                    }
                    DevGate &a = f.g[i], &c = f.g[i + 1];
                    const unsigned char pos3[3] = {a.tl[0], a.tl[1], c.tl[0]};
                    set_masks(a, pos3, 3);
                    a.post = 6;
                    a.pair = 1;
                    c.post = 7;
                    ++i;
                }
        }
        f.nstages = 0;
        for (size_t oi = 0; oi < cnt; ++oi) f.nstages += f.g[oi].diag >= 2;
        if (getenv("QIPB_DEBUG")) {
            int nst = 0, npost = 0, ndiag = 0, npair = 0, nqft2 = 0, d1 = 0, d1real = 0, mk[4] = {0, 0, 0, 0}, low2 = 0, ctl2 = 0, sweeps = 0;
            for (size_t oi = 0; oi < cnt; ++oi) {
                const DevGate &g = f.g[oi];
                nst += g.diag >= 2;
                npost += g.post == 1;
                ndiag += g.diag == 1;
                npair += g.post == 4;
                nqft2 += g.post == 2;
                if (!g.diag && g.k == 2) {
                    mk[g.mk & 3]++;
                    low2 += g.nins == 2 && (g.tl[0] < 3 || g.tl[1] < 3);
                    ctl2 += g.nins > 2;
                }
                if (!g.diag && g.k == 1) { d1++; d1real += g.mk == MK1_REAL; }
                sweeps += !(g.diag == 3 || g.post == 3 || g.post == 5);
            }
            if (atoi(getenv("QIPB_DEBUG")) >= 2)
                for (size_t oi = 0; oi < cnt; ++oi)
                    fprintf(stderr, "[qipb]   op %2zu: k=%d diag=%d post=%d mk=%d nins=%d tl=(%d,%d) in_or=%x out_ctrl=%llx pair=%d\n", oi, f.g[oi].k,
                            f.g[oi].diag, f.g[oi].post, f.g[oi].mk, f.g[oi].nins, f.g[oi].tl[0], f.g[oi].tl[1], f.g[oi].in_or,
                            (unsigned long long)f.g[oi].out_ctrl, f.g[oi].pair);
            fprintf(stderr, "[qipb] fused launch: %d gates -> %d ops, %d sweeps | 2q general %d real %d realphase %d mono %d (low-bit %d, in-tile ctrl %d, pairs %d) | "
                            "1q %d (real %d, +stage %d, qft2 %d) | lone diag %d | stages %d | tables %zu\n",
                    ngates, (int)cnt, sweeps, mk[0], mk[1], mk[2], mk[3], low2, ctl2, npair, d1, d1real, npost, nqft2, ndiag, nst, tables.size());
        }
        int rc = launch(f);
        if (rc) return rc;
    }
    return QIPB_OK;
}

}  // namespace qipb

using namespace qipb;

static int apply_fused_impl(qipb_ctx *ctx, void *state, int nbits, int dtype, int ntile_bits, const int *tile_bits,
                            int ngates, const qipb_gate *gates, bool fill, FusedChunk chunk = FusedChunk{0, nullptr, 0}) {
    QIPB_REQUIRE(ctx && state, "null argument");
    QIPB_CUDA(cudaSetDevice(ctx->device));
    return lower_fused(
        nbits, dtype, ntile_bits, tile_bits, ngates, gates,
        [&](const std::vector<cplx> &tables, FusedArgs &f) { return upload_tables(ctx, tables, &f.tables); },
        [&](const FusedArgs &f) {
            return dtype == QIPB_C128 ? launch_fused<double2>(ctx, (double2 *)state, f)
                                      : launch_fused<float2>(ctx, (float2 *)state, f);
        },
        fill, chunk);
}

extern "C" int qipb_apply_fused_chunk(qipb_ctx *ctx, void *state, int nbits, int dtype, int ntile_bits, const int *tile_bits,
                                      int ngates, const qipb_gate *gates, int nfix, const int *fix_bits, uint64_t fix_value) {
    return apply_fused_impl(ctx, state, nbits, dtype, ntile_bits, tile_bits, ngates, gates, false, FusedChunk{nfix, fix_bits, fix_value});
}

extern "C" int qipb_apply_fused(qipb_ctx *ctx, void *state, int nbits, int dtype, int ntile_bits, const int *tile_bits,
                                int ngates, const qipb_gate *gates) {
    return apply_fused_impl(ctx, state, nbits, dtype, ntile_bits, tile_bits, ngates, gates, false);
}

extern "C" int qipb_apply_fused_fill(qipb_ctx *ctx, void *state, int nbits, int dtype, int ntile_bits, const int *tile_bits,
                                     int ngates, const qipb_gate *gates) {
    return apply_fused_impl(ctx, state, nbits, dtype, ntile_bits, tile_bits, ngates, gates, true);
}
#endif  // QIPB_FUSED_TU_C64
