// tests/csrc/fused_emul.cu -- TEST INFRASTRUCTURE ONLY (never linked into libqipb200.so, never on the product path).
//
// Runs one fused pass on a HOST state vector with the product's own code: the real validation + lowering of
// qipb_apply_fused (lower_fused: stages, tables, structured block forms, stage riding, launch splitting) and the
// real sweep functions (run_op and everything below it are __host__ __device__, qip_b200/csrc/common.cuh QIPB_HD),
// driven by a loop that plays the roles of the CTA: tiles one after another, ops in order, "threads" tid = 0..NT-1
// one after another (inside one sweep every group of amplitudes is owned by exactly one thread, and sweeps are
// separated by __syncthreads in the kernel, so sequential execution is equivalent).  What it does NOT cover is
// what only exists on the device: TMA staging, mbarriers, the grid loop.  The CPU tier uses it to check the
// lowering and the sweep arithmetic against the numpy bit simulator (tests/test_fused_emul.py).
#include <math.h>
#include <stdarg.h>
#define QIPB_FUSED_SINGLE_TU
#define QIPB_HD __host__ __device__ inline      /* no forced inlining here: the host compile of this harness drops from minutes to seconds */
#include "../../qip_b200/csrc/fused.cu"

namespace qipb {
static thread_local char g_emul_err[512] = "";
void set_error(const char *fmt, ...) {       // the library's copy lives in api.cu, which is not part of this harness
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_emul_err, sizeof(g_emul_err), fmt, ap);
    va_end(ap);
}
}  // namespace qipb

namespace {

template <typename A, bool UNI, int NT, bool EXT>
void emulate_launch(A *state, const FusedArgs &f) {
    const u32 tsize = 1u << f.tb;
    std::vector<A> tile(tsize);
    std::vector<double2> stage_S(FUSED_MAX_OPS + 1);
    for (u64 t = 0; t < f.ntiles; ++t) {
        const u64 base = fused_tile_base(f, t);
        auto offset_of = [&](u32 e) {
            u64 off = 0;
            for (int j = 0; j < f.tb; ++j) off |= (u64)((e >> j) & 1u) << f.tbit[j];
            return off;
        };
        const bool fill = EXT && f.g[0].diag == 5;             // fill mode: the tile is written by op 0, nothing is loaded
        for (u32 e = 0; e < tsize; ++e)
            tile[e] = fill ? make_amp<A>(NAN, NAN) : state[base + offset_of(e)];
        if (f.nstages)
            for (int tid = 0; tid < NT; ++tid) stage_scalars<NT>(f, base, stage_S.data(), tid);
        for (int gi = 0; gi < f.ngates; ++gi) {
            if (fused_op_is_skipped<EXT>(f.g[gi])) continue;
            for (int tid = 0; tid < NT; ++tid)
                run_fused_op<A, UNI, NT, EXT>(tile.data(), f, gi, stage_S.data(), base, tsize, tid);
        }
        for (u32 e = 0; e < tsize; ++e) state[base + offset_of(e)] = tile[e];
    }
}

template <typename A>
int emulate(A *state, const FusedArgs &f, int *info) {
    const bool bulk = launch_is_bulk(f, sizeof(A));
    const bool half = bulk && f.tb == 11;
    const bool uni = launch_is_uni(f, sizeof(A));
    info[0] += 1;
    info[1] += uni ? 1 : 0;
    for (int gi = 0; gi < f.ngates; ++gi) {
        const DevGate &g = f.g[gi];
        info[2] += g.diag >= 2;
        info[3] += g.post == 1;
        info[4] += (!g.diag && g.k == 2 && g.mk != MK_GENERAL);
        info[5] += g.post == 2;
        info[6] += (!g.diag && g.k == 1 && g.mk == MK1_REAL);
        info[8] += g.post == 4;
        info[10] += g.post == 6;
        info[11] += g.post == 8;
        info[12] += g.post == 9;
        info[13] += g.post == 10 ? g.pair : 0;
    }
    // the same choice of kernel instantiation as launch_fused (qip_b200/csrc/fused.cu)
    if (half) emulate_launch<A, true, 128, false>(state, f);
    else if (launch_is_wide(f, sizeof(A))) {
        info[7] += fused_has_ext(f) ? 1 : 0;
        info[9] += 1;
        emulate_launch<A, true, 128, true>(state, f);
    } else if (uni && fused_has_ext(f)) {
        info[7] += 1;
        emulate_launch<A, true, 256, true>(state, f);
    } else if (uni) emulate_launch<A, true, 256, false>(state, f);
    else emulate_launch<A, false, 256, false>(state, f);
    return QIPB_OK;
}

}  // namespace

// info[0] launches, [1] of which take the specialised (UNI) sweeps, [2] stages, [3] stages riding on a dense 1-qubit
// sweep, [4] structured 2-qubit blocks, [5] paired QFT steps, [6] real 1-qubit gates, [7] launches that carry EXT ops,
// [8] block pairs (two dense 2-qubit blocks in one sweep), [9] launches of the WIDE kernel, [10] trios (block + lone 1-qubit gate), [11] stages riding on a dense 2-qubit sweep,
// [12] radix-16 QFT sweeps (four steps each), [13] QFT steps on the lowest tile bits taken by lane-butterfly sweeps
static int emul_impl(void *host_state, int nbits, int dtype, int ntile_bits, const int *tile_bits, int ngates,
                     const qipb_gate *gates, int *info, bool fill, FusedChunk chunk = FusedChunk{0, nullptr, 0}) {
    QIPB_REQUIRE(host_state && info, "null argument");
    for (int i = 0; i < 16; ++i) info[i] = 0;
    return lower_fused(
        nbits, dtype, ntile_bits, tile_bits, ngates, gates,
        [&](const std::vector<cplx> &tables, FusedArgs &f) {
            f.tables = reinterpret_cast<const double2 *>(tables.data());
            return QIPB_OK;
        },
        [&](const FusedArgs &f) {
            return dtype == QIPB_C128 ? emulate<double2>((double2 *)host_state, f, info) : emulate<float2>((float2 *)host_state, f, info);
        },
        fill, chunk);
}

// qipb_apply_fused_chunk on the host: only the amplitudes whose fix bits have the given value may change
extern "C" int qipb_emul_fused_chunk(void *host_state, int nbits, int dtype, int ntile_bits, const int *tile_bits, int ngates,
                                     const qipb_gate *gates, int *info, int nfix, const int *fix_bits, unsigned long long fix_value) {
    return emul_impl(host_state, nbits, dtype, ntile_bits, tile_bits, ngates, gates, info, false, FusedChunk{nfix, fix_bits, fix_value});
}

extern "C" int qipb_emul_fused(void *host_state, int nbits, int dtype, int ntile_bits, const int *tile_bits, int ngates,
                               const qipb_gate *gates, int *info) {
    return emul_impl(host_state, nbits, dtype, ntile_bits, tile_bits, ngates, gates, info, false);
}

// qipb_apply_fused_fill on the host: the state's previous content must not matter
extern "C" int qipb_emul_fused_fill(void *host_state, int nbits, int dtype, int ntile_bits, const int *tile_bits, int ngates,
                                    const qipb_gate *gates, int *info) {
    return emul_impl(host_state, nbits, dtype, ntile_bits, tile_bits, ngates, gates, info, true);
}

// host cost of the lowering alone (no tile is touched): what every launch of the product pays on the CPU
extern "C" int qipb_emul_lower_only(int nbits, int dtype, int ntile_bits, const int *tile_bits, int ngates, const qipb_gate *gates) {
    std::vector<cplx> keep;
    return lower_fused(
        nbits, dtype, ntile_bits, tile_bits, ngates, gates,
        [&](const std::vector<cplx> &tables, FusedArgs &f) {
            keep = tables;
            f.tables = reinterpret_cast<const double2 *>(keep.data());
            return QIPB_OK;
        },
        [&](const FusedArgs &) { return QIPB_OK; });
}

extern "C" const char *qipb_emul_last_error(void) { return qipb::g_emul_err; }
