// bt_tile_types.cuh -- parameter blocks of the fused tile kernels (bt_tile.cu) shared with the pass specialiser (bt_jit.cu).
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>

#ifndef TILE_MAXG
#define TILE_MAXG 8      // single-gate slots per pass
#endif
#ifndef TILE_MAXC
#define TILE_MAXC 4      // register-cluster slots per pass
#endif
#define TILE_MAXD 16     // diagonal-gate slots per pass (one run-time-indexed code copy: they carry <= 4 matrix entries)
#define TILE_DBASE 16    // item ids >= TILE_DBASE are diagonal slots
#define TILE_PBASE 48    // item ids >= TILE_PBASE are register programs
#define TILE_MAXP 8      // register programs per pass
#define TILE_MAXITEMS (TILE_PBASE + TILE_MAXP)
#ifndef PROG_BITS
#define PROG_BITS 4      // a register program holds the 2^4 amplitudes spanned by 4 tile bits in each thread (5: 32 amplitudes)
#endif
#define PROG_AMPS (1 << PROG_BITS)
#define PROG_MAXOPS 64
#define PROG_MAXCOEF 208 // doubles
#define TILE_LOWB_MIN 3  // the tensor-copy kernel needs only bits 0..2 (one 128-byte row) in the tile: BT_TILE_LOWB=3..5
#define TILE_LOWB 5      // the 5 lowest index bits are always in the tile: a warp's 32 lanes cover one 512 B run
#define TILE_TMAX 12     // largest tile: 2^12 amplitudes = 64 KB of shared memory
#define TILE_TDEF 12     // default tile (measured best on B200: T=12 single-buffered; T=11 double-buffered is ~12 % slower)
#ifndef TILE_THREADS
#define TILE_THREADS 128 // 3 CTAs x 128 threads: up to 170 registers per thread for the 16-amplitude clusters
#endif
#ifndef TILE_MINB
#define TILE_MINB 3
#endif
#ifndef CL_BITS
#define CL_BITS 3        // cluster width: 3 tile bits = 8 amplitudes per thread, slots (0,1),(1,2),(0,1); 4 = 16 amplitudes,
                         // slots (0,1),(2,3),(1,2), 168 registers: measured 4 % slower on C2
#endif
#define CL_AMPS (1 << CL_BITS)
#ifndef CL_PIPE
#define CL_PIPE 0         // 1 = software-pipeline two cluster groups per thread (measured: no gain, 166 registers)
#endif
#ifndef CL_SLOTS
#define CL_SLOTS 3       // cluster pattern on positions (0,1),(2,3),(1,2) [5 = + (0,1),(2,3): measured 4 % slower, code size]
#endif

struct TileGate {
  int32_t kind;       // 0 dense, 1 diagonal
  int32_t k;          // matrix index bits (<= 2)
  int32_t tloc[2];    // tile-local position of matrix bit t, or -1 when the bit lies outside the tile (diagonal only)
  int32_t text[2];    // physical position when outside the tile
  int32_t ni;         // tile-local inserted positions (local targets + local controls), ascending
  uint32_t bit_sw[8];  // swizzled slot contribution of thread-index bit k (host picks which free tile bit each thread bit walks)
  uint32_t lcmask;    // local controls, forced to 1
  uint32_t niter;     // group-loop trip count = max(1, (2^T >> ni) / TILE_THREADS)
  uint64_t ext_cmask; // controls outside the tile: all must be 1 in the tile's base index
  uint32_t iter_sw[32]; // swizzled slot offset contributed by loop iteration i: sw(expand(i * TILE_THREADS))
  double2 m[16];      // dense: row-major (1<<k)^2 ; diagonal: first 1<<k entries
};

// Diagonal gate (k <= 2 targets, any of them possibly outside the tile, plus controls): cheap, so it gets many slots.
struct TileDiag {
  int32_t k;
  int32_t ni;
  int32_t tloc[2];
  int32_t text[2];
  uint32_t lcmask;
  uint32_t niter;
  uint64_t ext_cmask;
  uint32_t bit_sw[8];
  uint32_t iter_sw[16];
  double2 m[4];
};

// A register-resident cluster: each thread holds the 16 amplitudes spanned by 4 tile bits (positions 0..3) and applies up
// to five dense 4x4 blocks on position pairs (0,1),(2,3),(1,2),(0,1),(2,3) before writing back -- one shared-memory round
// trip for e.g. the brickwork triple (a,b),(c,d),(b,c) instead of three.
struct TileCluster {
  int32_t lp[4];        // tile-local bit of cluster position p
  uint32_t bit_sw[8];   // as in TileGate
  uint32_t use;         // bit s set: pattern slot s carries a block
  uint32_t niter;       // max(1, (2^T >> 4) / TILE_THREADS)
  uint32_t iter_sw[8];
  double2 m[CL_SLOTS][16];  // row-major 4x4, matrix bit 0 <-> lower position of the slot's pair
};

// A register program: each thread loads the 16 amplitudes spanned by 4 tile bits (positions 0..3) and runs a list of
// STRUCTURED micro-ops on them before writing back -- one shared-memory round trip for e.g. two brickwork layers on 4 qubits.
// The micro-ops keep the gates' structure instead of multiplying it away into dense 4x4 blocks (16 FP64 ops / amplitude):
//   U1_GEN  general 2x2                        8 ops / amplitude      U1_DIAG  diag(d0, d1)        4
//   U1_REAL real 2x2 (H, RY, X, Z ...)         4                      U1_PHASE diag(1, d)          2
//   U1_RXL  real diagonal, imaginary off-diag  4 (RX, Y ...)          CX       register moves      0
//   CPHASE  phase on |11> (CZ, CP)             1
// op word: site | (first coefficient << 8); the site fixes kind and cluster positions so that register indices are static.
enum { PK_GEN = 0, PK_REAL = 1, PK_RXL = 2, PK_DIAG = 3, PK_PHASE = 4, PK_CX = 5, PK_CPHASE = 6 };
#define PROG_SITE_U1(kind, p) ((kind) * 8 + (p))                 // 0..39
#define PROG_SITE_CX(pc, pt) (40 + (pc) * 8 + (pt))              // 40..103 (pc != pt)
#define PROG_SITE_CPHASE(pa, pb) (104 + (pa) * 8 + (pb))         // 104..167 (pa < pb)
// Conditional ops: a control / phase bit that is NOT one of the program's positions is a property of the thread's group
// (tile-local bits, mask lm) or of the whole tile (bits outside the tile, mask em) -- "controls ride along for free".
// coefficients: CSCALE, CPH1: re, im, em, lm (masks as raw 64-bit words); CCX1: em, lm
#define PROG_SITE_CSCALE 168                                     // all amplitudes *= d             when the masks match
#define PROG_SITE_CPH1(p) (169 + (p))                            // amplitudes with bit p set *= d  when the masks match
#define PROG_SITE_CCX1(p) (177 + (p))                            // X on position p                 when the masks match
struct TileProg {
  int32_t lp[8];          // tile-local bit of cluster position p < PROG_BITS
  uint32_t bit_sw[8];
  uint32_t niter;
  uint32_t nops;
  uint32_t iter_sw[8];
  uint32_t bit_lin[8];    // as bit_sw / iter_sw but unswizzled: the group's tile-local index, for the conditional ops
  uint32_t iter_lin[8];
  uint16_t op[PROG_MAXOPS + 2];   // + padding entry read by the op-word prefetch
  double coef[PROG_MAXCOEF + 4];  // + slack: four coefficients are always fetched
};

struct TileParams {
  int32_t T;            // tile bits
  int32_t lowb;         // min(TILE_LOWB, T)
  int32_t nitems;
  int32_t stagger_ns;   // < 0: measurement modes (-1 = no gates, -2 = no HBM traffic; results invalid); BT_TILE_STAGGER_NS
  int32_t n_sm;
  int32_t swz_mode;     // 0: TMA-compatible 128-B swizzle, 1: all-digit swizzle
  int32_t tma_coord_shift[5];  // tensor-copy variant: coordinate k of a tile = (base >> shift[k]) & mask[k]
  uint32_t tma_coord_mask[5];
  int32_t tma_ncopy;           // a tile whose bits form more than four runs moves as 2, 4, 8 or 16 boxes: the top tile bits are enumerated
  int32_t tma_c4add[16];        // copy e: added to coordinate 4; its shared-memory chunk is e * (tile bytes / ncopy)
  int32_t tbits[TILE_TMAX];  // physical positions of the tile bits, ascending; tbits[j] = j for j < lowb
  uint8_t item[TILE_MAXITEMS];  // item i: < TILE_MAXG -> gate slot, else cluster slot (item - TILE_MAXG)
  TileGate g[TILE_MAXG];
  TileCluster cl[TILE_MAXC];
  TileDiag d[TILE_MAXD];
  TileProg pr[TILE_MAXP];
};

static_assert(sizeof(TileParams) + 128 <= 32764, "kernel parameters (tensor map + TileParams) must fit the 32 KB parameter space");

