/*
 * include/qip_b200.h -- C ABI of libqipb200.so, the B200 (sm_100a) state-vector engine for QIP.
 *
 * This is the drop-in boundary for ONE path of Renmusxd/QIP: the native kernels behind
 * qip/backend.py's StateType (state init -> k-qubit gate apply -> measure).  Each entry point
 * names the reference interface it replaces (paths relative to the reference repository).
 * The reference's FFI for this path is Cython (`qip/ext/*.pyx`, imported at
 * qip/backend.py:1-8); a replacement backend binds the functions below with ctypes -- see
 * INTEGRATION.md for the stub and qip_b200/backend.py for the shipped host side.
 *
 * Conventions
 *   - Plain C: pointers, sizes, ints.  No torch / C++ types.  Every function returns 0 on
 *     success, non-zero on failure; qipb_last_error() returns a thread-local message.
 *   - `state` is a DEVICE pointer to 2^nbits amplitudes owned by the caller (the python host
 *     allocates it with torch; qipb_dev_alloc is offered for hosts without torch).
 *     dtype QIPB_C128 = interleaved (re,im) doubles, QIPB_C64 = interleaved floats.
 *   - All kernels are IN PLACE and are enqueued on the context's stream (qipb_set_stream);
 *     nothing synchronises unless stated.
 *   - Bit positions: `bit b` is bit b of the (local) amplitude index, 0 = least significant.
 *     The reference's qubit index q maps to bit (n-1-q) (qip/ext/kronprod.pyx:168,187); the
 *     host does that mapping, and, when the state is sharded over GPUs by its top qubits,
 *     also resolves everything that refers to a global (rank) bit before calling in.
 *   - Matrix coefficients are always complex128 (interleaved doubles) whatever the state dtype.
 */
#ifndef QIP_B200_H
#define QIP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QIPB_C128 0
#define QIPB_C64 1

#define QIPB_MAX_DENSE_K 4     /* register-blocked dense kernels: 1..4 target bits            */
#define QIPB_MAX_BIG_K 10      /* batched dense kernel (FP64 tensor tiles): up to 10 target bits */
#define QIPB_MAX_TILE_BITS 12  /* fused pass: tiles of 2^12 amplitudes (64 KiB complex128, 32 KiB complex64) */
#define QIPB_MAX_FUSED_GATES 280 /* gates per qipb_apply_fused call (runs of diagonal gates fold into stages) */

typedef struct qipb_ctx qipb_ctx;

/* ---- context ------------------------------------------------------------------------- */
int qipb_version(void);
const char *qipb_last_error(void);
int qipb_create(int device, qipb_ctx **out);
int qipb_destroy(qipb_ctx *ctx);
int qipb_set_stream(qipb_ctx *ctx, void *cuda_stream);       /* cudaStream_t; NULL = default  */
int qipb_sync(qipb_ctx *ctx);                                /* cudaStreamSynchronize          */
unsigned long long qipb_launch_count(qipb_ctx *ctx);         /* kernels launched via this ctx  */
unsigned long long qipb_ring_launch_count(qipb_ctx *ctx);    /* of which: persistent ring kernel of qipb_apply_fused (diagnostic) */
unsigned long long qipb_ext_launch_count(qipb_ctx *ctx);     /* of which: fused launches with the opt-in forms (QIPB_FUSED_EXT, diagnostic) */

/* ---- memory helpers for hosts without torch -------------------------------------------- */
int qipb_dev_alloc(qipb_ctx *ctx, size_t bytes, void **out);
int qipb_dev_free(qipb_ctx *ctx, void *ptr);
int qipb_memcpy_h2d(qipb_ctx *ctx, void *dst_dev, const void *src_host, size_t bytes);
int qipb_memcpy_d2h(qipb_ctx *ctx, void *dst_host, const void *src_dev, size_t bytes);

/* ---- state construction ------------------------------------------------------------------
 * Replaces CythonBackend.make_state (qip/backend.py:73-104) + gen_edit_indices
 * (qip/util.py:108-124).
 * qipb_init_basis: all zeros, and amplitude `index` = 1 if 0 <= index < 2^nbits (pass -1 on
 *   shards that do not own |0...0>) -- the empty-feed case, qip/backend.py:90-91.
 * qipb_init_kron: state[i] = prod_g feeds_g[sub_g(G)] where G = (shard_index << nbits) | i is
 *   the global index, for every G with (G & fixed_mask) == fixed_value, else 0 (un-fed qubits are
 *   fixed to 0; a group fed with a one-hot basis index -- qip/distributed/backend.py:42-45, and the
 *   int `Qubit.default` of qip/pipeline.py:101-109 -- is fixed to that index instead of being
 *   multiplied in, so no 2^k vector is ever built).  ngroups may be 0.  Group g lists
 *   group_len[g] GLOBAL bit positions in group_bits (concatenated, most significant sub-index
 *   bit first, i.e. the order of the reference's index group).  feeds_dev = the groups'
 *   vectors concatenated, complex128, on the device.  The product is taken left to right
 *   starting from 1.0 like qip/backend.py:98-101.                                            */
int qipb_init_basis(qipb_ctx *ctx, void *state, int nbits, int dtype, long long index);
int qipb_init_kron(qipb_ctx *ctx, void *state, int nbits, int dtype, int ngroups,
                   const int *group_len, const int *group_bits, const void *feeds_dev,
                   uint64_t fixed_mask, uint64_t fixed_value, uint64_t shard_index);

/* ---- gate application ----------------------------------------------------------------------
 * Replaces cdot_loop (qip/ext/kronprod.pyx:43-200) for ONE entry of a `mats` dict after the host
 * has unwrapped CMat chains into control bits (kronprod.pyx:215-225) and SwapMat into bit swaps
 * (:227-231).  Entries of one dict act on disjoint targets, so the host applies them one after
 * another.
 * qipb_apply_matrix: for every index whose ctrl_mask bits are all 1, the 2^k amplitudes that
 *   differ only on `bits` are replaced by mat * (those amplitudes).  bits[0] is the most
 *   significant bit of the matrix index (kronprod.pyx:184-189).  mat is row-major 2^k x 2^k
 *   complex128.  k = 0 multiplies the controlled sub-space by the scalar mat[0] (phase gates).
 *   diagonal != 0 promises off-diagonal entries are zero (only the diagonal is read).
 *   k <= QIPB_MAX_DENSE_K runs register-blocked; up to QIPB_MAX_BIG_K as batched matrix products staged in shared
 *   memory and multiplied as FP64 tensor tiles (complex64 states are widened on the way in).
 * qipb_apply_swap: exchanges bit_a and bit_b of the index (SwapMat(1)) under ctrl_mask; only the
 *   amplitudes whose two bits differ move.                                                     */
int qipb_apply_matrix(qipb_ctx *ctx, void *state, int nbits, int dtype, int k, const int *bits,
                      const double *mat, uint64_t ctrl_mask, int diagonal);
int qipb_apply_swap(qipb_ctx *ctx, void *state, int nbits, int dtype, int bit_a, int bit_b,
                    uint64_t ctrl_mask);

/* One gate of a fused pass.  bits/ctrl_mask are positions in the LOCAL index (not tile-relative).
 * Non-diagonal gates must have all their `bits` inside the pass's tile bits; control bits and the
 * bits of diagonal gates may lie anywhere.                                                    */
typedef struct {
    int32_t k;            /* 0..2 target bits                                                 */
    int32_t diagonal;     /* 1: only mat[i*(2^k)+i] is used                                    */
    int32_t bits[2];      /* bits[0] = most significant matrix-index bit                       */
    uint64_t ctrl_mask;
    double mat[32];       /* up to 4x4 complex128, row-major, interleaved                      */
} qipb_gate;

/* qipb_apply_fused: one read-modify-write sweep of the state that applies `ngates` gates in order.
 * The state is cut into tiles of 2^ntile_bits amplitudes spanned by tile_bits (ascending,
 * distinct; the lowest ones must be 0..L-1 with L >= 5 so that every global access is a full
 * coalesced run); each CTA stages a tile in shared memory, runs the gate list on it, writes it
 * back.  Replaces `ngates` separate cdot_loop sweeps (qip/ext/kronprod.pyx:157-197).          */
int qipb_apply_fused(qipb_ctx *ctx, void *state, int nbits, int dtype, int ntile_bits,
                     const int *tile_bits, int ngates, const qipb_gate *gates);

/* qipb_apply_fused_chunk: the same pass restricted to ONE CHUNK of the state -- the amplitudes whose nfix (0..4)
 * index bits fix_bits (none of them a tile bit) equal fix_value (a mask over those bits).  The 2^nfix chunks
 * partition the state and the pass acts on each of them independently (no gate of a pass is non-diagonal on a bit
 * outside the tile), so running every chunk once equals qipb_apply_fused.  The sharded engine uses it to pipeline a
 * pass against the NVLink exchange of the neighbouring chunk (qipb_peer_remap_chunk); it stands where the reference's
 * workers interleave compute and socket traffic per 2048-amplitude message (qip/distributed/worker/worker.py:302-357). */
int qipb_apply_fused_chunk(qipb_ctx *ctx, void *state, int nbits, int dtype, int ntile_bits,
                           const int *tile_bits, int ngates, const qipb_gate *gates,
                           int nfix, const int *fix_bits, uint64_t fix_value);

/* qipb_apply_fused_fill: the same pass applied to the ALL-ONES vector; the buffer's previous content is never
 * read.  A product state of one-qubit feeds v_b is prod_b diag(v_b[0], v_b[1]) . ones, so a caller that leads the
 * gate list with those diagonal gates gets "kron-product init (qip/backend.py:88-101) + first gate pass" in one
 * write-only sweep of HBM.  Returns 3 (unsupported, nothing launched, the caller initialises the buffer and uses
 * qipb_apply_fused) unless the tile has 2^12 amplitudes in runs of >= 512 bytes and the list starts with a run of
 * >= 3 un-controlled diagonal gates.                                                                            */
int qipb_apply_fused_fill(qipb_ctx *ctx, void *state, int nbits, int dtype, int ntile_bits,
                          const int *tile_bits, int ngates, const qipb_gate *gates);

/* ---- func_apply -----------------------------------------------------------------------------
 * Replaces func_apply (qip/ext/func_apply.pyx:11-112): |x>|q>|r> -> |x>|f(x) xor q>|r>, in
 * place (the map is an involution on q for fixed x, so amplitudes are swapped pairwise).
 * reg1_bits / reg2_bits list LOCAL bit positions, most significant register bit first.
 * table_dev[x] = f(x) as int64 on the device, indexed by (x_fixed | gathered x) so a shard can
 * pass the contribution of rank bits in x_fixed.  Only the low n2 bits of f(x) are used
 * (func_apply.pyx:97).                                                                        */
int qipb_func_xor(qipb_ctx *ctx, void *state, int nbits, int dtype, int n1, const int *reg1_bits,
                  int n2, const int *reg2_bits, const long long *table_dev, uint64_t x_fixed);
/* The same with a byte table, table_dev[x] = f(x) & 0xFF, for output registers of n2 <= 8 qubits: only the low n2
 * bits of f(x) are used (func_apply.pyx:97), so the table -- read once per amplitude -- shrinks eightfold.       */
int qipb_func_xor_u8(qipb_ctx *ctx, void *state, int nbits, int dtype, int n1, const int *reg1_bits,
                     int n2, const int *reg2_bits, const unsigned char *table_dev, uint64_t x_fixed);

/* ---- measurement -----------------------------------------------------------------------------
 * qipb_probabilities: out_dev[o] (+)= sum of |a_i|^2 over all i with (i & filter_mask) ==
 *   filter_value, where bit out_bits[j] of o is bit bits[j] of i.  k = 0 gives the (filtered)
 *   total probability in out_dev[0].  Replaces prob_magnitude (qip/ext/kronprod.pyx:320-326),
 *   measure_probabilities (:240-261; out_bits[j] = j), the per-outcome sums of soft_measure
 *   (:371-376) and measure_top_probabilities (:290-296) (out_bits = big-endian over the sorted
 *   qubits).  Deterministic: fixed-order tree, no floating-point atomics.  out_dev is device
 *   memory of 2^k doubles; it is overwritten.
 * qipb_collapse: a_i *= scale where (i & mask) == want, a_i = 0 elsewhere.  Replaces the sweep
 *   of measure (qip/ext/kronprod.pyx:439-444).
 * qipb_reduce: dst[j] = src[i(j)] * scale, i(j) = j's bits deposited on the positions NOT in
 *   mask, OR want.  dst holds 2^(nbits - popcount(mask)) amplitudes and must not alias src.
 *   Replaces the sweep of reduce_measure (qip/ext/kronprod.pyx:478-489).                      */
int qipb_probabilities(qipb_ctx *ctx, const void *state, int nbits, int dtype, int k,
                       const int *bits, const int *out_bits, uint64_t filter_mask,
                       uint64_t filter_value, double *out_dev);
int qipb_collapse(qipb_ctx *ctx, void *state, int nbits, int dtype, uint64_t mask, uint64_t want,
                  double scale);
int qipb_reduce(qipb_ctx *ctx, const void *src, void *dst, int nbits, int dtype, uint64_t mask,
                uint64_t want, double scale);

/* ---- range access (CythonBackend.addto_relative_range, qip/backend.py:174-175) ------------- */
int qipb_add_range(qipb_ctx *ctx, void *state, int dtype, uint64_t start, uint64_t count,
                   const void *data_dev);

/* ---- multi-GPU exchange over NVLink peer memory ------------------------------------------------
 * Replaces the worker<->worker state exchange of qip/distributed/worker/worker.py:302-357 and the
 * manager sync of qip/distributed/manager.py:224-236.  The state is sharded by its top qubits,
 * one process per GPU; peers map each other's shard with CUDA IPC.
 * qipb_ipc_export / qipb_ipc_open / qipb_ipc_close: 64-byte handle of a device allocation made
 *   with qipb_dev_alloc, and the peer-side mapping of it.
 * qipb_peer_swap: in place, exchanges `count` amplitudes starting at local[local_off] with
 *   peer[peer_off] -- one kernel, loads and stores straight over NVLink, no staging buffer.
 * qipb_peer_swap_bit: in place, swaps a GLOBAL (rank) index bit with local bit `lbit`: this rank's
 *   amplitudes whose bit lbit != my_gbit trade places with the partner's amplitudes whose bit lbit
 *   == my_gbit.  The 2^(nbits-1) pairs are indexed by w (the local index with bit lbit removed);
 *   the two ranks of a pair each process a disjoint [w_begin, w_begin+count) half.
 * qipb_peer_remap: g (1..3) rank bits <-> g local bits in ONE kernel over up to 7 peers: peers[b] is the
 *   mapped shard of the rank whose g rank bits have value b (entry my_value is ignored), lbits[t] the
 *   local bit paired with value bit t.  Moves (1 - 2^-g) of a shard per direction.
 * qipb_peer_remap_chunk: the same exchange restricted to the chunk whose nfix (0..4) local bits fix_bits (none of
 *   them exchanged) equal fix_value; every rank must pass the same chunk.  max_ctas > 0 bounds the grid (CTAs per
 *   partner; a persistent, grid-striding launch that shares the SMs with a concurrently running fused pass).
 * qipb_peer_gate1: the fused compute+exchange kernel for a 1-qubit gate whose target is a GLOBAL
 *   (rank) bit: for count amplitudes, (lo, hi) <- mat * (lo, hi) where `lo` lives on the shard
 *   whose rank bit is 0 and `hi` on its partner.  The caller that owns `local` passes
 *   local_is_hi; each rank of the pair processes half of the range.                            */
int qipb_ipc_export(qipb_ctx *ctx, void *dev_ptr, unsigned char handle_out[64]);
int qipb_ipc_open(qipb_ctx *ctx, const unsigned char handle[64], void **peer_ptr_out);
int qipb_ipc_close(qipb_ctx *ctx, void *peer_ptr);
int qipb_peer_swap(qipb_ctx *ctx, void *local, void *peer, int dtype, uint64_t local_off,
                   uint64_t peer_off, uint64_t count);
int qipb_peer_swap_bit(qipb_ctx *ctx, void *local, void *peer, int nbits, int dtype, int lbit,
                       int my_gbit, uint64_t w_begin, uint64_t count);
int qipb_peer_remap(qipb_ctx *ctx, void *local, void *const *peers, int nbits, int dtype, int g,
                    const int *lbits, int my_value);
int qipb_peer_remap_chunk(qipb_ctx *ctx, void *local, void *const *peers, int nbits, int dtype, int g,
                          const int *lbits, int my_value, int nfix, const int *fix_bits,
                          uint64_t fix_value, int max_ctas);
int qipb_peer_gate1(qipb_ctx *ctx, void *local, void *peer, int dtype, uint64_t off,
                    uint64_t count, const double *mat, int local_is_hi, uint64_t ctrl_mask);

#ifdef __cplusplus
}
#endif
#endif /* QIP_B200_H */
