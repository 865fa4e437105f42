// bt_jit.cu -- pass specialiser for the fused tile kernel.
//
// The interpreter in bt_tile.cu (run_prog) pays a dispatch per micro-op: op word, jump table, indirect branch, loop branch,
// coefficient loads -- about as many issue cycles as the arithmetic itself.  A pass that is executed
// again and again (every step of a trajectory batch, a variational loop, a benchmark) is therefore compiled once into a
// straight-line kernel: same tile movement (one TMA tensor copy each way), same micro-op code (bt_prog_ops.cuh, handed to
// NVRTC as text), but every structural quantity -- tile bits, program positions, shared-memory offsets, op sequence,
// condition masks -- is a literal, and the coefficients are a by-value parameter block read through constant-bank
// operands.  Keyed by structure only, so a re-run with other angles reuses the module.
//
// Policy: a structure is handed to the compile workers the BT_TILE_JIT_AFTER-th time it is seen (default 1) on states of at
// least 2^BT_TILE_JIT_MINBITS amplitudes (default 22).  Compilation never stalls the caller: NVRTC runs on a small pool of
// worker threads (BT_JIT_THREADS, default min(8, cores / ranks on the node)) while the pass keeps running on the interpreter; the cubin is loaded
// by the calling thread the next time the structure comes by.  Cubins are also kept on disk (BT_JIT_CACHE_DIR, default
// $XDG_CACHE_HOME/bluetangle_cuda or ~/.cache/bluetangle_cuda; empty string = off), so a later process pays a file read
// instead of 0.2 s of NVRTC per structure.  BT_TILE_JIT=0 switches the specialiser off, BT_TILE_JIT=2 compiles synchronously
// at first sight regardless of size (tests), BT_TILE_JIT_ASYNC=0 compiles synchronously under the normal policy.  bt_jit_wait()
// blocks until the workers are idle (benchmarks call it at the end of their warm-up).  Any failure (no libnvrtc, compile
// error, driver entry point missing) marks the structure as interpreter-only; the result is never different, only slower.
//
// No reference analogue (the reference builds one sparse matrix per op, src/hilbert.jl:505).
#include "bt_internal.cuh"
#include "bt_tile_types.cuh"
#include <cuda.h>
#include <dlfcn.h>
#include <stdarg.h>
#include <stdio.h>
#include <algorithm>
#include <stdlib.h>
#include <unistd.h>
#include <string.h>
#include <chrono>
#include <complex>
#include <math.h>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <string>
#include <sys/stat.h>
#include <thread>
#include <unordered_map>
#include <vector>


namespace {

// ---- NVRTC (dlopen) and the driver entry points ----------------------------------------------------------------------
typedef struct _nvrtcProgram* nvrtcProgram;
struct Nvrtc {
  int (*CreateProgram)(nvrtcProgram*, const char*, const char*, int, const char* const*, const char* const*) = nullptr;
  int (*CompileProgram)(nvrtcProgram, int, const char* const*) = nullptr;
  int (*GetCUBINSize)(nvrtcProgram, size_t*) = nullptr;
  int (*GetCUBIN)(nvrtcProgram, char*) = nullptr;
  int (*GetProgramLogSize)(nvrtcProgram, size_t*) = nullptr;
  int (*GetProgramLog)(nvrtcProgram, char*) = nullptr;
  int (*DestroyProgram)(nvrtcProgram*) = nullptr;
  int (*Version)(int*, int*) = nullptr;
  int major = 0, minor = 0;
  bool ok = false;
};

Nvrtc& nvrtc() {
  static Nvrtc n;
  static bool tried = false;
  if (tried) return n;
  tried = true;
  void* h = nullptr;
  const char* user = getenv("BT_NVRTC_LIB");  // explicit path to libnvrtc.so.12 (its builtins library must sit next to it)
  if (user && *user) h = dlopen(user, RTLD_NOW | RTLD_LOCAL);
  for (const char* name : {"libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so.12", "/usr/local/cuda/lib64/libnvrtc.so"}) {
    if (h) break;
    h = dlopen(name, RTLD_NOW | RTLD_LOCAL);
  }
  if (!h) return n;
#define BT_NVRTC_SYM(f) *(void**)(&n.f) = dlsym(h, "nvrtc" #f); if (!n.f) return n;
  BT_NVRTC_SYM(CreateProgram) BT_NVRTC_SYM(CompileProgram) BT_NVRTC_SYM(GetCUBINSize) BT_NVRTC_SYM(GetCUBIN)
  BT_NVRTC_SYM(GetProgramLogSize) BT_NVRTC_SYM(GetProgramLog) BT_NVRTC_SYM(DestroyProgram)
#undef BT_NVRTC_SYM
  *(void**)(&n.Version) = dlsym(h, "nvrtcVersion");
  if (n.Version) n.Version(&n.major, &n.minor);
  n.ok = true;
  return n;
}

struct Driver {
  CUresult (*ModuleLoadData)(CUmodule*, const void*) = nullptr;
  CUresult (*ModuleGetFunction)(CUfunction*, CUmodule, const char*) = nullptr;
  CUresult (*FuncSetAttribute)(CUfunction, CUfunction_attribute, int) = nullptr;
  CUresult (*LaunchKernel)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, CUstream, void**, void**) = nullptr;
  bool ok = false;
};

Driver& driver() {
  static Driver d;
  static bool tried = false;
  if (tried) return d;
  tried = true;
  cudaDriverEntryPointQueryResult q;
#define BT_DRV_SYM(f, name) if (cudaGetDriverEntryPoint(name, (void**)&d.f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) { cudaGetLastError(); return d; }
  BT_DRV_SYM(ModuleLoadData, "cuModuleLoadData") BT_DRV_SYM(ModuleGetFunction, "cuModuleGetFunction")
  BT_DRV_SYM(FuncSetAttribute, "cuFuncSetAttribute") BT_DRV_SYM(LaunchKernel, "cuLaunchKernel")
#undef BT_DRV_SYM
  d.ok = true;
  return d;
}

int env_i(const char* name, int dflt) {
  const char* v = getenv(name);
  return (v && *v) ? atoi(v) : dflt;
}

// ---- source generation -------------------------------------------------------------------------------------------------
void appf(std::string& s, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  int n = vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (n > 0) s.append(buf, (size_t)std::min<int>(n, (int)sizeof(buf) - 1));
}

uint64_t as_u64(double d) { uint64_t v; memcpy(&v, &d, 8); return v; }

// ---- planning: ops of a pass in specialised form ----------------------------------------------------------------------
// Straight-line code needs none of the interpreter's in-place discipline, so the specialiser also reshapes the arithmetic:
//   * a real / RX-like 2x2 is divided by a pivot entry s (|s| within 10x of the largest entry, chosen to leave the most
//     entries equal to +-1 or 0): a rotation [[c,-s],[s,c]] becomes c [[1,-t],[t,1]] -- one FMA per scalar instead of two
//     instructions, H becomes two additions; diag(d0,d1) becomes d0 diag(1, d1/d0).  The factors s multiply up into ONE
//     complex scalar per pass, applied to every amplitude at the end of the last program (a pass has ~30 ops, so the
//     intermediate scale stays far from the floating-point range limits);
//   * CX is a renaming of the amplitude variables (no instruction); CZ a sign change.
// Which entries are +-1 / 0 and which pivot was taken is part of the structure key (the code differs), their values are not.
enum { JK_LIN2 = 0, JK_GEN, JK_PHASE, JK_CX, JK_CPHASE, JK_CZ, JK_CSCALE, JK_CPH1, JK_CCX1 };

struct JOp {
  int kind = 0;
  int p = 0, q = 0;        // positions (LIN2 / GEN / PHASE / CPH1 / CCX1: p; CX: control p, target q; CPHASE / CZ: p, q)
  int rxl = 0;             // LIN2: 0 = pairs (ax,bx),(ay,by); 1 = pairs (ax,by),(ay,bx) with the RX-like signs
  signed char pat[4] = {2, 2, 2, 2};  // LIN2 entries after the pivot division: 0 zero, +1 / -1 unit, 2 general (takes a coefficient)
  int c0 = 0;              // first coefficient of this op in the coefficient block
  uint64_t em = 0, lm = 0; // condition masks
  int merged = 0;          // member of a merged run of diagonal factors (its complex factor is multiplied up per thread first)
  int shear = 0;           // unit-modulus phase applied as three shears (3 FMAs instead of 2 MUL + 2 FMA): coefficients are
                           // (t, s) = (-tan(theta/2), sin(theta)); 2: theta was reduced by pi, the result is negated
  int deferred = 0;        // CSCALE whose factor joins the program's per-thread scalar (applied once, after the last op)
  int sgn = 0;             // conditional factor -1: a sign-bit flip (integer pipe), no coefficients
};

struct Plan {
  std::vector<std::vector<JOp>> prog;  // per item
  std::vector<double> coef;
  int scale_at = -1;                   // coefficient index of the pass scalar (re, im)
  bool complex_scale = false;
  int scale_prog = -1;                 // the program that applies the pass scalar (the last one unless a per-thread scalar carries it)
};

typedef std::complex<double> cd;

// divide the real 2x2 (m0 m1; m2 m3) by a pivot; returns the pivot.  pat/out describe the quotient.
double pivot_real(const double m[4], signed char pat[4], double out[4]) {
  double big = 0;
  for (int i = 0; i < 4; ++i) big = std::max(big, fabs(m[i]));
  int best = -1, best_cost = 99;
  for (int c = 0; c < 4; ++c) {
    if (fabs(m[c]) < 0.1 * big || m[c] == 0.0) continue;
    int cost = 0;
    for (int i = 0; i < 4; ++i) {
      const double r = m[i] / m[c];
      if (!(r == 0.0 || r == 1.0 || r == -1.0)) cost++;
    }
    if (cost < best_cost || (cost == best_cost && fabs(m[c]) > fabs(m[best]))) { best = c; best_cost = cost; }
  }
  if (best < 0) { for (int i = 0; i < 4; ++i) { out[i] = m[i]; pat[i] = m[i] == 0.0 ? 0 : (m[i] == 1.0 ? 1 : (m[i] == -1.0 ? -1 : 2)); } return 1.0; }
  for (int i = 0; i < 4; ++i) {
    const double r = m[i] / m[best];
    out[i] = r;
    pat[i] = r == 0.0 ? 0 : (r == 1.0 ? 1 : (r == -1.0 ? -1 : 2));
  }
  return m[best];
}

int jit_variant();

// Arithmetic reshaping on top of the code shape (BT_JIT_OPT, bit flags; default 7, 0 = the round-2 generator text byte for byte).
// The specialised passes are bound by the FP64 pipe (ncu: 77 % pipe activity), so every FP64 instruction that
// becomes an integer instruction or disappears is time:
//   1  per-thread scalars are deferred: a conditional scalar (a phase whose bits all lie outside the program -- T / RZ / CP on tile
//      bits that are not program positions) multiplies ALL amplitudes a thread holds, so it commutes with every op of the program;
//      the factors of a program are multiplied up per thread (4 instructions each) and applied once behind the last op (64), in
//      the last program together with the pass scalar -- before: 48-64 instructions per factor or per run of factors;
//   2  negations are sign-bit flips on the integer pipe (the compiler turns -x on a double into DADD -RZ, -x): CZ inside a program,
//      lone -a terms of a 2x2, and the pi-reduced shear rotation folds its sign into the operand modifiers of its FMAs;
//   4  a conditional factor -1 (CZ with one bit outside the program) flips the sign bit under a selected mask instead of a selected
//      complex multiplication (4 FP64 instructions per amplitude).
// Counted on the 65 passes of C2 (cuobjdump -sass, DFMA + DADD + DMUL): 119 278 -> see profiles/r2_jit_fp64_counts.txt.
int jit_opt() { return env_i("BT_JIT_OPT", 7); }

bool make_plan(const TileParams& P, int np, Plan& pl) {
  if (P.swz_mode != 0 || np <= 0 || P.nitems != np) return false;
  pl.prog.assign((size_t)P.nitems, {});
  pl.coef.clear();
  cd gs(1.0, 0.0);
  for (int it = 0; it < P.nitems; ++it) {
    const int pi = (int)P.item[it] - TILE_PBASE;
    if (pi < 0 || pi >= TILE_MAXP) return false;
    const TileProg& G = P.pr[pi];
    if (G.niter > 8 || G.iter_sw[0] != 0) return false;
    for (uint32_t k = 0; k < G.nops; ++k) {
      const uint32_t site = G.op[k] & 0xffu;
      const double* c = G.coef + (G.op[k] >> 8);
      JOp o;
      o.c0 = (int)pl.coef.size();
      if (site < 40) {
        const int kind = site / 8;
        o.p = site % 8;
        if (kind == PK_REAL || kind == PK_RXL) {
          o.kind = JK_LIN2; o.rxl = kind == PK_RXL;
          double q[4];
          gs *= pivot_real(c, o.pat, q);
          for (int i = 0; i < 4; ++i) if (o.pat[i] == 2) pl.coef.push_back(q[i]);
        } else if (kind == PK_DIAG) {
          // diag(d0, d1) = d0 diag(1, d1/d0)  (or d1 diag(d0/d1, 1) when d0 is the small one: phase on the bit = 0 half)
          const cd d0(c[0], c[1]), d1(c[2], c[3]);
          if (std::abs(d0) >= 0.1 * std::abs(d1) && d0 != cd(0, 0)) { gs *= d0; const cd r = d1 / d0; o.kind = JK_PHASE; o.q = 1; pl.coef.push_back(r.real()); pl.coef.push_back(r.imag()); }
          else if (d1 != cd(0, 0)) { gs *= d1; const cd r = d0 / d1; o.kind = JK_PHASE; o.q = 0; pl.coef.push_back(r.real()); pl.coef.push_back(r.imag()); }
          else return false;
        } else if (kind == PK_PHASE) {
          o.kind = JK_PHASE; o.q = 1; pl.coef.push_back(c[0]); pl.coef.push_back(c[1]);
        } else {  // PK_GEN: divide by the largest entry; the quotient keeps 8 coefficients (the unit entry is simply 1 + 0i)
          int best = 0;
          for (int i = 1; i < 4; ++i) if (std::abs(cd(c[2 * i], c[2 * i + 1])) > std::abs(cd(c[2 * best], c[2 * best + 1]))) best = i;
          const cd pv(c[2 * best], c[2 * best + 1]);
          if (pv == cd(0, 0)) return false;
          gs *= pv;
          o.kind = JK_GEN; o.q = best;
          for (int i = 0; i < 4; ++i) { const cd r = cd(c[2 * i], c[2 * i + 1]) / pv; pl.coef.push_back(r.real()); pl.coef.push_back(r.imag()); }
        }
      } else if (site < 104) {
        o.kind = JK_CX; o.p = (site - 40) / 8; o.q = (site - 40) % 8;
      } else if (site < 168) {
        o.p = (site - 104) / 8; o.q = (site - 104) % 8;
        if (c[0] == -1.0 && c[1] == 0.0) o.kind = JK_CZ;
        else { o.kind = JK_CPHASE; pl.coef.push_back(c[0]); pl.coef.push_back(c[1]); }
      } else if (site == PROG_SITE_CSCALE || (site >= 169 && site < 177)) {
        o.kind = site == PROG_SITE_CSCALE ? JK_CSCALE : JK_CPH1;
        o.p = site == PROG_SITE_CSCALE ? 0 : (int)site - 169;
        o.em = as_u64(c[2]); o.lm = as_u64(c[3]);
        if (o.kind == JK_CSCALE && o.em == 0 && o.lm == 0) { gs *= cd(c[0], c[1]); continue; }  // unconditional scalar: joins the pass scalar
        if ((jit_opt() & 4) && c[0] == -1.0 && c[1] == 0.0) o.sgn = 1;
        else { pl.coef.push_back(c[0]); pl.coef.push_back(c[1]); }
      } else if (site >= 177 && site < 185) {
        o.kind = JK_CCX1; o.p = (int)site - 177; o.em = as_u64(c[0]); o.lm = as_u64(c[1]);
      } else {
        return false;
      }
      pl.prog[(size_t)it].push_back(o);
    }
  }
  // Runs of diagonal ops: the factors of a run that hit the same amplitudes are merged when there are two or more of them
  // (generate() multiplies them up per thread).  Every other phase with a unit-modulus factor becomes a rotation by three shears
  //   r += t i;  i += s r;  r += t i      t = -tan(theta/2), s = sin(theta)
  // (3 FMAs per amplitude instead of 2 MUL + 2 FMA; |theta| <= pi/2 after an optional reduction by pi, so |t| <= 1).  A factor whose
  // modulus is not 1 to rounding (T is rounded to 10 digits in the reference's table) keeps the complex multiplication.
  const bool merge = env_i("BT_JIT_MERGE_DIAG", 1) != 0, use_shear = env_i("BT_JIT_SHEAR", 1) != 0;
  const int variant = jit_variant();
  pl.scale_prog = (int)pl.prog.size() - 1;
  if (jit_opt() & 1) {
    // deferred per-thread scalars: worth it when a program carries two or more factors, or in the program that applies the pass
    // scalar -- the scalar commutes with everything, so it rides with the per-thread scalar of the program that has the most
    // factors (a single factor elsewhere keeps its own form: three shears cost 48, the deferred application 64)
    const int defer_min = std::max(1, env_i("BT_JIT_DEFER_MIN", 2));
    int best = 0;
    for (size_t it = 0; it < pl.prog.size(); ++it) {
      int n = 0;
      for (const JOp& o : pl.prog[it]) if (o.kind == JK_CSCALE) n++;
      if (n > 0 && n >= best) { best = n; pl.scale_prog = (int)it; }
    }
    for (size_t it = 0; it < pl.prog.size(); ++it) {
      int nfac = 0, nsgn = 0;
      for (const JOp& o : pl.prog[it]) if (o.kind == JK_CSCALE) { if (o.sgn) nsgn++; else nfac++; }
      if (nfac >= defer_min || (nfac + nsgn >= 1 && (int)it == pl.scale_prog))
        for (JOp& o : pl.prog[it]) if (o.kind == JK_CSCALE) o.deferred = 1;
    }
  }
  for (std::vector<JOp>& L : pl.prog) {
    auto is_diag = [](const JOp& o) { return o.kind == JK_PHASE || o.kind == JK_CPH1 || o.kind == JK_CSCALE || o.kind == JK_CPHASE || o.kind == JK_CZ; };
    size_t i = 0;
    while (i < L.size()) {
      if (!is_diag(L[i])) { ++i; continue; }
      size_t j = i;
      while (j < L.size() && is_diag(L[j])) ++j;
      if (merge)
        for (int cls = -1; cls < PROG_BITS; ++cls) {
          std::vector<size_t> mem;
          for (size_t t = i; t < j; ++t) {
            const JOp& o = L[t];
            if (o.sgn || o.deferred) continue;
            if (cls < 0 ? o.kind == JK_CSCALE : ((o.kind == JK_CPH1 && o.p == cls) || (o.kind == JK_PHASE && o.p == cls && o.q == 1))) mem.push_back(t);
          }
          if (mem.size() >= 2)
            for (size_t t : mem) L[t].merged = 1;
        }
      i = j;
    }
    if (!use_shear) continue;
    for (JOp& o : L) {
      if (o.merged || o.sgn || o.deferred || !(o.kind == JK_PHASE || o.kind == JK_CPH1 || o.kind == JK_CSCALE || o.kind == JK_CPHASE)) continue;
      const bool conditional_select = (o.kind == JK_CPH1 || o.kind == JK_CSCALE) && o.lm != 0 && (variant & 2);
      const bool conditional_branch = (o.kind == JK_CPH1 || o.kind == JK_CSCALE) && !conditional_select;
      (void)conditional_branch;
      double& cr = pl.coef[(size_t)o.c0];
      double& ci = pl.coef[(size_t)o.c0 + 1];
      if (fabs(cr * cr + ci * ci - 1.0) > 4.5e-16) continue;
      double th = atan2(ci, cr);
      int shape = 1;
      if (fabs(th) > 1.5707963267948966) {
        if (conditional_select && !(jit_opt() & 4)) continue;  // a selected identity cannot carry the sign (BT_JIT_OPT & 4: the sign is a selected mask)
        th -= th > 0 ? 3.141592653589793 : -3.141592653589793;
        shape = 2;
      }
      o.shear = shape;
      cr = -tan(0.5 * th);
      ci = sin(th);
    }
  }
  pl.scale_at = (int)pl.coef.size();
  pl.complex_scale = gs.imag() != 0.0;
  pl.coef.push_back(gs.real());
  pl.coef.push_back(gs.imag());
  return true;
}

// Code shape of the generated kernels (BT_JIT_VARIANT; bit flags):
//   0  a condition that depends on the thread (a control / phase bit inside the tile but outside the program) is a branch around the
//      arithmetic -- the fastest form (C2: 160.6 ms per step);
//   2  such a condition SELECTS the coefficient (identity when off) and the arithmetic runs unconditionally: no thread-dependent
//      control flow is left inside the group loop (C2: 171.8 ms);
//   1  the group loop is fully unrolled (179 ms); 4 select only for lane-dependent conditions; 16 opaque group index (experiments).
// Default: 0 with NVRTC <= 12.8, 2 with anything newer or unknown.  Measured on B200 (profiles/r2_jit_nvrtc129.txt): the cubins NVRTC
// 12.9.86 produces for form 0 drop individual conditional phases in QFT-like passes (dozens of thread-dependent branches inside the
// two-iteration group loop; amplitudes off by 1e-4) -- the same text is correct when compiled by NVRTC 12.8, with -Xptxas=-O0, when
// executed on the host (tests/test_jit_codegen_cpu.py), and in forms 1 and 2.  BT_JIT_VERIFY=1 cross-checks every launch on the device.
// resident CTAs per SM the register allocation aims at (BT_JIT_MINB; 3 = 168 registers, 4 = 128, 5 = 102): more CTAs need tiles of
// <= 2^11 amplitudes (shared memory)
int jit_minb() { return std::max(1, std::min(8, env_i("BT_JIT_MINB", TILE_MINB))); }

int jit_variant() {
  const char* v = getenv("BT_JIT_VARIANT");
  if (v && *v) return atoi(v);
  Nvrtc& n = nvrtc();
  const bool known_good = n.ok && n.major > 0 && (n.major < 12 || (n.major == 12 && n.minor <= 8));
  return known_good ? 0 : 2;
}

// structure key: everything the generator turns into literals or code shape (numeric coefficients excluded)
void make_key(const TileParams& P, const Plan& pl, int device, std::string& key, int K = 1) {
  key.clear();
  key.append((const char*)&K, 4);
  auto put = [&](const void* p, size_t n) { key.append((const char*)p, n); };
  put(&device, 4); put(&P.T, 4); put(&P.lowb, 4); put(&P.nitems, 4); put(&P.swz_mode, 4);
  put(P.tma_coord_shift, sizeof(P.tma_coord_shift)); put(P.tma_coord_mask, sizeof(P.tma_coord_mask));
  put(&P.tma_ncopy, 4); put(P.tma_c4add, sizeof(int32_t) * (size_t)std::max(1, P.tma_ncopy));
  put(P.tbits, sizeof(int32_t) * (size_t)P.T);
  put(P.item, (size_t)P.nitems);
  const char cs = pl.complex_scale ? 1 : 0;
  put(&cs, 1);
  put(&pl.scale_prog, 4);
  const int variant = jit_variant();
  put(&variant, 4);
  const int minb = jit_minb();
  put(&minb, 4);
  const int opt = jit_opt();
  put(&opt, 4);
  for (int it = 0; it < P.nitems; ++it) {
    const TileProg& G = P.pr[(int)P.item[it] - TILE_PBASE];
    put(G.lp, sizeof(int32_t) * PROG_BITS); put(G.bit_sw, sizeof(G.bit_sw)); put(&G.niter, 4);
    put(G.iter_sw, sizeof(G.iter_sw)); put(G.bit_lin, sizeof(G.bit_lin)); put(G.iter_lin, sizeof(G.iter_lin));
    const uint32_t n = (uint32_t)pl.prog[(size_t)it].size();
    put(&n, 4);
    for (const JOp& o : pl.prog[(size_t)it]) {
      const int32_t h[8] = {o.kind, o.p, o.q, o.rxl, o.merged, o.shear, o.deferred, o.sgn};
      put(h, sizeof(h)); put(o.pat, 4); put(&o.em, 8); put(&o.lm, 8);
    }
  }
}

// "t = alpha * a + beta * b" with alpha, beta given as pattern + coefficient reference; sa / sb flip the sign of a term
std::string lin(const char* out, signed char pa, int ka, bool na, const std::string& a, signed char pb, int kb, bool nb, const std::string& b, bool xneg = false) {
  auto term_coef = [](signed char pat, int k, bool neg) {  // textual coefficient of a general term
    char buf[48];
    snprintf(buf, sizeof(buf), neg ? "(-C.c[%d])" : "C.c[%d]", k);
    return std::string(buf);
  };
  auto unit_sign = [](signed char pat, bool neg) { return (pat > 0) != neg; };  // true: +, false: -
  std::string e;
  const bool za = pa == 0, zb = pb == 0, ua = pa == 1 || pa == -1, ub = pb == 1 || pb == -1;
  if (za && zb) e = "0.0";
  else if (za) e = ub ? (unit_sign(pb, nb) ? b : (xneg ? "bt_neg(" + b + ")" : "-" + b)) : term_coef(pb, kb, nb) + " * " + b;
  else if (zb) e = ua ? (unit_sign(pa, na) ? a : (xneg ? "bt_neg(" + a + ")" : "-" + a)) : term_coef(pa, ka, na) + " * " + a;
  else if (ua && ub) e = std::string(unit_sign(pa, na) ? "" : "-") + a + (unit_sign(pb, nb) ? " + " : " - ") + b;
  else if (ua) e = "fma(" + term_coef(pb, kb, nb) + ", " + b + ", " + (unit_sign(pa, na) ? "" : "-") + a + ")";
  else if (ub) e = "fma(" + term_coef(pa, ka, na) + ", " + a + ", " + (unit_sign(pb, nb) ? "" : "-") + b + ")";
  else e = "fma(" + term_coef(pb, kb, nb) + ", " + b + ", " + term_coef(pa, ka, na) + " * " + a + ")";
  return std::string("const double ") + out + " = " + e + ";";
}

std::string cond_open(const JOp& o) {
  char buf[160];
  snprintf(buf, sizeof(buf), "      if ((base & 0x%llxull) == 0x%llxull && (gl & 0x%llxull) == 0x%llxull) {\n", (unsigned long long)o.em, (unsigned long long)o.em,
           (unsigned long long)o.lm, (unsigned long long)o.lm);
  return buf;
}

std::string cond_expr(const JOp& o) {
  char buf[160];
  snprintf(buf, sizeof(buf), "((base & 0x%llxull) == 0x%llxull && (gl & 0x%llxull) == 0x%llxull)", (unsigned long long)o.em, (unsigned long long)o.em, (unsigned long long)o.lm,
           (unsigned long long)o.lm);
  return buf;
}

// BT_JIT_VARIANT (bit flags, experiments): 1 = the group loop is fully unrolled, 2 = conditions that depend on the thread (a control bit
// inside the tile) select the coefficient (identity when off) instead of branching around the arithmetic

// Emits the kernel for the planned pass.
// variant 128 ("wide"): 256 threads per CTA, 2 CTAs per SM; the second iteration of the two-iteration group loop becomes thread bit 7,
// so a program is one straight-line block without any loop (16 instead of 12 warps per SM, two tiles in flight instead of three)
bool wide_ok(const TileParams& P) {
  if (!(jit_variant() & 128)) return false;
  for (int it = 0; it < P.nitems; ++it)
    if (P.pr[(int)P.item[it] - TILE_PBASE].niter != 2) return false;
  return true;
}

bool generate(const TileParams& P, const Plan& pl, std::string& s, int K = 1) {
  const int variant = jit_variant();
  const int opt = jit_opt();
  const bool xneg = (opt & 2) != 0;
  const bool wide = wide_ok(P) && K <= 1;
  const int T = P.T;
  const uint32_t nloc = 1u << T, ng = nloc >> PROG_BITS;
  std::string body;
  for (int it = 0; it < P.nitems; ++it) {
    const TileProg& G = P.pr[(int)P.item[it] - TILE_PBASE];
    const bool last = it == pl.scale_prog;  // the program that applies the pass scalar
    bool need_gl = false;
    for (const JOp& o : pl.prog[(size_t)it]) if (o.lm) need_gl = true;
    int perm[PROG_AMPS];  // logical amplitude (index bits = program positions) -> variable
    for (int j = 0; j < PROG_AMPS; ++j) perm[j] = j;
    uint64_t lane_mask = 0;  // tile-local bits that vary across the lanes of a warp
    for (int k = 0; k < 5; ++k) lane_mask |= G.bit_lin[k];
    auto use_select = [&](const JOp& o) { return o.lm != 0 && ((variant & 2) || ((variant & 4) && (o.lm & lane_mask))); };
    std::string ops;
    auto X = [&](int logical, char part) { char b[24]; snprintf(b, sizeof(b), "x%c[%d]", part, perm[logical]); return std::string(b); };
    auto emit_op = [&](const JOp& o) {
      int k = o.c0;
      if (o.deferred) return;  // joins the per-thread scalar behind the last op
      if (o.sgn) {
        // conditional factor -1 on the amplitudes with position bit p set (CPH1) or on all of them (CSCALE): sign-bit flips
        const bool sel = use_select(o);
        if (sel) ops += "      { const uint32_t sm_ = " + cond_expr(o) + " ? 0x80000000u : 0u;\n";
        else ops += cond_open(o);
        for (int j = 0; j < PROG_AMPS; ++j) {
          if (o.kind == JK_CPH1 && !((j >> o.p) & 1)) continue;
          const std::string r = X(j, 'r'), i = X(j, 'i');
          if (sel) ops += "        " + r + " = bt_xsign(" + r + ", sm_); " + i + " = bt_xsign(" + i + ", sm_);\n";
          else ops += "        " + r + " = bt_neg(" + r + "); " + i + " = bt_neg(" + i + ");\n";
        }
        ops += "      }\n";
        return;
      }
      if (o.kind == JK_LIN2) {
        int kk[4];
        for (int i = 0; i < 4; ++i) kk[i] = o.pat[i] == 2 ? k++ : -1;
        for (int r = 0; r < PROG_AMPS / 2; ++r) {
          const int i0 = ((r >> o.p) << (o.p + 1)) | (r & ((1 << o.p) - 1)), i1 = i0 | (1 << o.p);
          ops += "      { ";
          if (!o.rxl) {
            for (char part : {'r', 'i'}) {
              const std::string a = X(i0, part), b = X(i1, part);
              ops += lin(part == 'r' ? "ta" : "tc", o.pat[0], kk[0], false, a, o.pat[1], kk[1], false, b, xneg) + " ";
              ops += lin(part == 'r' ? "tb" : "td", o.pat[2], kk[2], false, a, o.pat[3], kk[3], false, b, xneg) + " ";
            }
            ops += X(i0, 'r') + " = ta; " + X(i1, 'r') + " = tb; " + X(i0, 'i') + " = tc; " + X(i1, 'i') + " = td; }\n";
          } else {
            // a' = m0 a + i m1 b, b' = i m2 a + m3 b:  (ar, bi) <- [[m0, -m1], [m2, m3]],  (ai, br) <- [[m0, m1], [-m2, m3]]
            const std::string ar = X(i0, 'r'), ai = X(i0, 'i'), br = X(i1, 'r'), bi = X(i1, 'i');
            ops += lin("ta", o.pat[0], kk[0], false, ar, o.pat[1], kk[1], true, bi, xneg) + " ";
            ops += lin("tb", o.pat[2], kk[2], false, ar, o.pat[3], kk[3], false, bi, xneg) + " ";
            ops += lin("tc", o.pat[0], kk[0], false, ai, o.pat[1], kk[1], false, br, xneg) + " ";
            ops += lin("td", o.pat[2], kk[2], true, ai, o.pat[3], kk[3], false, br, xneg) + " ";
            ops += ar + " = ta; " + bi + " = tb; " + ai + " = tc; " + br + " = td; }\n";
          }
        }
      } else if (o.kind == JK_GEN) {
        for (int r = 0; r < PROG_AMPS / 2; ++r) {
          const int i0 = ((r >> o.p) << (o.p + 1)) | (r & ((1 << o.p) - 1)), i1 = i0 | (1 << o.p);
          const std::string ar = X(i0, 'r'), ai = X(i0, 'i'), br = X(i1, 'r'), bi = X(i1, 'i');
          char buf[640];
          // entry o.q was the pivot: it equals 1 + 0i exactly, the compiler folds the multiplications by the literal away
          snprintf(buf, sizeof(buf),
                   "      { const double ta = C.c[%d] * %s - C.c[%d] * %s + C.c[%d] * %s - C.c[%d] * %s, tc = C.c[%d] * %s + C.c[%d] * %s + C.c[%d] * %s + C.c[%d] * %s,\n"
                   "          tb = C.c[%d] * %s - C.c[%d] * %s + C.c[%d] * %s - C.c[%d] * %s, td = C.c[%d] * %s + C.c[%d] * %s + C.c[%d] * %s + C.c[%d] * %s;\n",
                   k, ar.c_str(), k + 1, ai.c_str(), k + 2, br.c_str(), k + 3, bi.c_str(), k, ai.c_str(), k + 1, ar.c_str(), k + 2, bi.c_str(), k + 3, br.c_str(),  //
                   k + 4, ar.c_str(), k + 5, ai.c_str(), k + 6, br.c_str(), k + 7, bi.c_str(), k + 4, ai.c_str(), k + 5, ar.c_str(), k + 6, bi.c_str(), k + 7, br.c_str());
          ops += buf;
          ops += "        " + ar + " = ta; " + ai + " = tc; " + br + " = tb; " + bi + " = td; }\n";
        }
      } else if (o.shear && (o.kind == JK_PHASE || o.kind == JK_CPH1 || o.kind == JK_CSCALE || o.kind == JK_CPHASE)) {
        // unit-modulus phase as three shears; coefficients (t, s); a condition selects (t, s) or (0, 0) = identity, or branches
        const bool cond = o.kind == JK_CPH1 || o.kind == JK_CSCALE;
        const bool sel = cond && use_select(o);
        if (cond && !sel) ops += cond_open(o);
        else ops += "      {\n";
        char hdr[320];
        if (sel && o.shear == 2)
          snprintf(hdr, sizeof(hdr), "        const bool on = %s; const double st = on ? C.c[%d] : 0.0, ss = on ? C.c[%d] : 0.0; const uint32_t sm_ = on ? 0x80000000u : 0u;\n", cond_expr(o).c_str(), k, k + 1);
        else if (sel) snprintf(hdr, sizeof(hdr), "        const bool on = %s; const double st = on ? C.c[%d] : 0.0, ss = on ? C.c[%d] : 0.0;\n", cond_expr(o).c_str(), k, k + 1);
        else snprintf(hdr, sizeof(hdr), "        const double st = C.c[%d], ss = C.c[%d];\n", k, k + 1);
        ops += hdr;
        for (int j = 0; j < PROG_AMPS; ++j) {
          bool hit;
          if (o.kind == JK_CSCALE) hit = true;
          else if (o.kind == JK_CPHASE) hit = ((j >> o.p) & 1) && ((j >> o.q) & 1);
          else hit = ((j >> o.p) & 1) == (o.kind == JK_CPH1 ? 1 : o.q);
          if (!hit) continue;
          const std::string r = X(j, 'r'), i = X(j, 'i');
          if (o.shear == 2 && sel)  // rotation by theta - pi, then the sign under the selected mask
            ops += "        { const double r1 = fma(st, " + i + ", " + r + "), i1 = fma(ss, r1, " + i + "); " + r + " = bt_xsign(fma(st, i1, r1), sm_); " + i + " = bt_xsign(i1, sm_); }\n";
          else if (o.shear == 2 && xneg)  // -(rotation): i1n = -i1 and r' = -(r1 + st i1) through the operand signs of the FMAs
            ops += "        { const double r1 = fma(st, " + i + ", " + r + "), i1n = fma(-ss, r1, -" + i + "); " + r + " = fma(st, i1n, -r1); " + i + " = i1n; }\n";
          else if (o.shear == 2)
            ops += "        { const double r1 = fma(st, " + i + ", " + r + "), i1 = fma(ss, r1, " + i + "); " + r + " = -fma(st, i1, r1); " + i + " = -i1; }\n";
          else
            ops += "        { const double r1 = fma(st, " + i + ", " + r + "), i1 = fma(ss, r1, " + i + "); " + r + " = fma(st, i1, r1); " + i + " = i1; }\n";
        }
        ops += "      }\n";
      } else if (o.kind == JK_PHASE || o.kind == JK_CPH1) {
        const bool sel = o.kind == JK_CPH1 && use_select(o);
        if (o.kind == JK_CPH1 && !sel) ops += cond_open(o);
        if (sel) { char b2[320]; snprintf(b2, sizeof(b2), "      { const bool on = %s; const double cr = on ? C.c[%d] : 1.0, ci = on ? C.c[%d] : 0.0;\n", cond_expr(o).c_str(), k, k + 1); ops += b2; }
        const int half = o.kind == JK_CPH1 ? 1 : o.q;
        for (int j = 0; j < PROG_AMPS; ++j)
          if (((j >> o.p) & 1) == half) {
            const std::string r = X(j, 'r'), i = X(j, 'i');
            char buf[256];
            if (sel) snprintf(buf, sizeof(buf), "      { const double tr = cr * %s - ci * %s, ti = cr * %s + ci * %s; %s = tr; %s = ti; }\n", r.c_str(), i.c_str(), i.c_str(), r.c_str(), r.c_str(), i.c_str());
            else
            snprintf(buf, sizeof(buf), "      { const double tr = C.c[%d] * %s - C.c[%d] * %s, ti = C.c[%d] * %s + C.c[%d] * %s; %s = tr; %s = ti; }\n", k, r.c_str(), k + 1, i.c_str(), k,
                     i.c_str(), k + 1, r.c_str(), r.c_str(), i.c_str());
            ops += buf;
          }
        if (o.kind == JK_CPH1) ops += "      }\n";
      } else if (o.kind == JK_CX) {
        for (int j = 0; j < PROG_AMPS; ++j)
          if (((j >> o.p) & 1) && !((j >> o.q) & 1)) std::swap(perm[j], perm[j | (1 << o.q)]);
      } else if (o.kind == JK_CPHASE || o.kind == JK_CZ) {
        for (int j = 0; j < PROG_AMPS; ++j)
          if (((j >> o.p) & 1) && ((j >> o.q) & 1)) {
            const std::string r = X(j, 'r'), i = X(j, 'i');
            char buf[256];
            if (o.kind == JK_CZ && xneg) snprintf(buf, sizeof(buf), "      %s = bt_neg(%s); %s = bt_neg(%s);\n", r.c_str(), r.c_str(), i.c_str(), i.c_str());
            else if (o.kind == JK_CZ) snprintf(buf, sizeof(buf), "      %s = -%s; %s = -%s;\n", r.c_str(), r.c_str(), i.c_str(), i.c_str());
            else
              snprintf(buf, sizeof(buf), "      { const double tr = C.c[%d] * %s - C.c[%d] * %s, ti = C.c[%d] * %s + C.c[%d] * %s; %s = tr; %s = ti; }\n", k, r.c_str(), k + 1, i.c_str(), k,
                       i.c_str(), k + 1, r.c_str(), r.c_str(), i.c_str());
            ops += buf;
          }
      } else if (o.kind == JK_CSCALE) {
        const bool sel = use_select(o);
        if (!sel) ops += cond_open(o);
        else { char b2[320]; snprintf(b2, sizeof(b2), "      { const bool on = %s; const double cr = on ? C.c[%d] : 1.0, ci = on ? C.c[%d] : 0.0;\n", cond_expr(o).c_str(), k, k + 1); ops += b2; }
        for (int j = 0; j < PROG_AMPS; ++j) {
          char buf[256];
          if (sel) snprintf(buf, sizeof(buf), "      { const double tr = cr * xr[%d] - ci * xi[%d], ti = cr * xi[%d] + ci * xr[%d]; xr[%d] = tr; xi[%d] = ti; }\n", j, j, j, j, j, j);
          else
          snprintf(buf, sizeof(buf), "      { const double tr = C.c[%d] * xr[%d] - C.c[%d] * xi[%d], ti = C.c[%d] * xi[%d] + C.c[%d] * xr[%d]; xr[%d] = tr; xi[%d] = ti; }\n", k, j, k + 1, j, k, j,
                   k + 1, j, j, j);
          ops += buf;
        }
        ops += "      }\n";
      } else if (o.kind == JK_CCX1) {
        const bool sel = use_select(o);
        if (!sel) ops += cond_open(o);
        else ops += "      { const bool on = " + cond_expr(o) + ";\n";
        for (int j = 0; j < PROG_AMPS; ++j)
          if (!((j >> o.p) & 1)) {
            const std::string ar = X(j, 'r'), ai = X(j, 'i'), br = X(j | (1 << o.p), 'r'), bi = X(j | (1 << o.p), 'i');
            if (sel)
              ops += "      { const double tr = " + ar + ", ti = " + ai + "; " + ar + " = on ? " + br + " : tr; " + ai + " = on ? " + bi + " : ti; " + br + " = on ? tr : " + br + "; " + bi +
                     " = on ? ti : " + bi + "; }\n";
            else
              ops += "      { const double tr = " + ar + ", ti = " + ai + "; " + ar + " = " + br + "; " + ai + " = " + bi + "; " + br + " = tr; " + bi + " = ti; }\n";
          }
        ops += "      }\n";
      }
    };
    // Diagonal ops commute, so inside a maximal run of them (no 2x2 block, no CX in between) the conditional / unconditional factors
    // that hit the SAME set of amplitudes are multiplied up per thread first -- one complex multiplication per op -- and applied
    // once: all 16 amplitudes for conditional scalars (a phase whose bits all lie outside the program), the 8 amplitudes with
    // position bit p set for phases on p.  A QFT pass carries runs of ~40 such factors (64 FP64 instructions each before).
    {
      const std::vector<JOp>& L = pl.prog[(size_t)it];
      const bool merge = env_i("BT_JIT_MERGE_DIAG", 1) != 0;
      auto is_diag = [](const JOp& o) { return o.kind == JK_PHASE || o.kind == JK_CPH1 || o.kind == JK_CSCALE || o.kind == JK_CPHASE || o.kind == JK_CZ; };
      size_t i = 0;
      while (i < L.size()) {
        if (!is_diag(L[i]) || !merge) { emit_op(L[i]); ++i; continue; }
        size_t j = i;
        while (j < L.size() && is_diag(L[j])) ++j;
        std::vector<char> done(j - i, 0);
        for (int cls = -1; cls < PROG_BITS; ++cls) {  // -1: all amplitudes (CSCALE); p >= 0: amplitudes with position bit p set
          std::vector<size_t> mem;
          for (size_t t = i; t < j; ++t) {
            const JOp& o = L[t];
            if (!o.merged) continue;  // decided in make_plan (the coefficients of merged members stay complex factors)
            if (cls < 0 ? o.kind == JK_CSCALE : ((o.kind == JK_CPH1 && o.p == cls) || (o.kind == JK_PHASE && o.p == cls && o.q == 1))) mem.push_back(t);
          }
          if (mem.empty()) continue;
          const bool assign_first = (opt & 1) != 0;  // the first factor is an assignment (1.0 * c - 0.0 * d does not fold without fast-math)
          if (!assign_first) ops += "      { double fr = 1.0, fi = 0.0;\n";
          bool first_member = true;
          for (size_t t : mem) {
            const JOp& o = L[t];
            char buf[512];
            if (assign_first && first_member) {
              if (o.kind == JK_PHASE) snprintf(buf, sizeof(buf), "      { double fr = C.c[%d], fi = C.c[%d];\n", o.c0, o.c0 + 1);
              else snprintf(buf, sizeof(buf), "      { const bool on0 = %s; double fr = on0 ? C.c[%d] : 1.0, fi = on0 ? C.c[%d] : 0.0;\n", cond_expr(o).c_str(), o.c0, o.c0 + 1);
              ops += buf;
              done[t - i] = 1;
              first_member = false;
              continue;
            }
            if (o.kind == JK_PHASE)
              snprintf(buf, sizeof(buf), "        { const double t0 = fr * C.c[%d] - fi * C.c[%d]; fi = fr * C.c[%d] + fi * C.c[%d]; fr = t0; }\n", o.c0, o.c0 + 1, o.c0 + 1, o.c0);
            else
              snprintf(buf, sizeof(buf), "        { const bool on = %s; const double cr = on ? C.c[%d] : 1.0, ci = on ? C.c[%d] : 0.0; const double t0 = fr * cr - fi * ci; fi = fr * ci + fi * cr; fr = t0; }\n",
                       cond_expr(o).c_str(), o.c0, o.c0 + 1);
            ops += buf;
            done[t - i] = 1;
          }
          for (int a = 0; a < PROG_AMPS; ++a)
            if (cls < 0 || ((a >> cls) & 1)) {
              const std::string r = X(a, 'r'), im = X(a, 'i');
              ops += "        { const double tr = fr * " + r + " - fi * " + im + ", ti = fr * " + im + " + fi * " + r + "; " + r + " = tr; " + im + " = ti; }\n";
            }
          ops += "      }\n";
        }
        for (size_t t = i; t < j; ++t)
          if (!done[t - i]) emit_op(L[t]);
        i = j;
      }
    }
    // the program's per-thread scalar: product of its deferred conditional factors (and, in the last program, of the pass scalar)
    bool scalar_applied = false;
    {
      std::vector<const JOp*> dq;
      for (const JOp& o : pl.prog[(size_t)it]) if (o.deferred) dq.push_back(&o);
      if (!dq.empty()) {
        ops += "      {\n";
        bool have = false;
        char buf[512];
        if (last) {
          if (pl.complex_scale) snprintf(buf, sizeof(buf), "        double fr = C.c[%d], fi = C.c[%d];\n", pl.scale_at, pl.scale_at + 1);
          else snprintf(buf, sizeof(buf), "        double fr = C.c[%d], fi = 0.0;\n", pl.scale_at);
          ops += buf;
          have = true;
          scalar_applied = true;
        }
        for (const JOp* o : dq) {
          if (o->sgn) {
            if (!have) { ops += "        double fr = 1.0, fi = 0.0;\n"; have = true; }
            ops += "        { const uint32_t sm_ = " + cond_expr(*o) + " ? 0x80000000u : 0u; fr = bt_xsign(fr, sm_); fi = bt_xsign(fi, sm_); }\n";
          } else if (!have) {
            snprintf(buf, sizeof(buf), "        const bool on0 = %s; double fr = on0 ? C.c[%d] : 1.0, fi = on0 ? C.c[%d] : 0.0;\n", cond_expr(*o).c_str(), o->c0, o->c0 + 1);
            ops += buf;
            have = true;
          } else {
            snprintf(buf, sizeof(buf), "        { const bool on = %s; const double cr = on ? C.c[%d] : 1.0, ci = on ? C.c[%d] : 0.0; const double t0 = fr * cr - fi * ci; fi = fr * ci + fi * cr; fr = t0; }\n",
                     cond_expr(*o).c_str(), o->c0, o->c0 + 1);
            ops += buf;
          }
        }
        for (int j = 0; j < PROG_AMPS; ++j) {
          snprintf(buf, sizeof(buf), "        { const double tr = fr * xr[%d] - fi * xi[%d], ti = fr * xi[%d] + fi * xr[%d]; xr[%d] = tr; xi[%d] = ti; }\n", j, j, j, j, j, j);
          ops += buf;
        }
        ops += "      }\n";
      }
    }
    if (last && !scalar_applied) {  // the pass scalar: product of all pivots and unconditional scalars
      const int k = pl.scale_at;
      for (int j = 0; j < PROG_AMPS; ++j) {
        char buf[256];
        if (pl.complex_scale)
          snprintf(buf, sizeof(buf), "      { const double tr = C.c[%d] * xr[%d] - C.c[%d] * xi[%d], ti = C.c[%d] * xi[%d] + C.c[%d] * xr[%d]; xr[%d] = tr; xi[%d] = ti; }\n", k, j, k + 1, j, k, j,
                   k + 1, j, j, j);
        else
          snprintf(buf, sizeof(buf), "      xr[%d] *= C.c[%d]; xi[%d] *= C.c[%d];\n", j, k, j, k);
        ops += buf;
      }
    }
    // program frame
    appf(body, "  if (tid < %uu) {  // program %d: tile bits %d %d %d %d\n", ng, it, G.lp[0], G.lp[1], G.lp[2], G.lp[3]);
    body += "    uint32_t s0 = 0u;\n";
    for (int k = 0; k < 8; ++k)
      if (G.bit_sw[k]) appf(body, "    if (tid & %uu) s0 ^= %uu;\n", 1u << k, G.bit_sw[k]);
    if (need_gl) {
      body += "    uint32_t g0 = 0u;\n";
      for (int k = 0; k < 8; ++k)
        if (G.bit_lin[k]) appf(body, "    if (tid & %uu) g0 ^= %uu;\n", 1u << k, G.bit_lin[k]);
    }
    uint32_t o[PROG_BITS];
    for (int q = 0; q < PROG_BITS; ++q) { uint32_t c = 1u << G.lp[q]; o[q] = c ^ ((c >> 3) & 7u); }
    if (wide) {
      appf(body, "    if (tid & 128u) s0 ^= %uu;\n", G.iter_sw[1]);
      if (need_gl) appf(body, "    if (tid & 128u) g0 ^= %uu;\n", G.iter_lin[1]);
      body += "    {\n";
    } else {
      appf(body, "%s\n    for (uint32_t it = 0; it < %uu; ++it) {\n", (variant & 1) ? "#pragma unroll" : "#pragma unroll 1", G.niter);
    }
    body += "      uint32_t b = s0;\n";
    for (uint32_t i = 1; i < G.niter && !wide; ++i) appf(body, "      if (it == %uu) b = s0 ^ %uu;\n", i, G.iter_sw[i]);
    if (need_gl) {
      body += "      uint64_t gl = g0;\n";
      for (uint32_t i = 1; i < G.niter && !wide; ++i) appf(body, "      if (it == %uu) gl = g0 ^ %uu;\n", i, G.iter_lin[i]);
      // variant 16: the group index is opaque to the compiler in every iteration, so no condition on it can be hoisted out of the loop
      if (variant & 16) body += "      asm volatile(\"mov.b64 %0, %0;\" : \"+l\"(gl));\n";
    } else {
      body += "      const uint64_t gl = 0ull;\n";
    }
    // variant 32: the coefficient block is addressed through an offset the compiler cannot see through (always 0), fetched once per
    // loop iteration: without it every coefficient load is hoisted out of the two-iteration group loop, the hundreds of live values
    // overflow the uniform register file and come back through local memory and R2UR moves (30 % of the stall samples, ncu)
    if (variant & 32) body += "      uint32_t zoff; asm volatile(\"mov.u32 %0, 0;\" : \"=r\"(zoff));\n      const double* __restrict__ CC = C.c + zoff;\n";
    body += "      double xr[PROG_AMPS], xi[PROG_AMPS];\n";
    for (int j = 0; j < PROG_AMPS; ++j) {
      uint32_t off = 0;
      for (int q = 0; q < PROG_BITS; ++q) if ((j >> q) & 1) off ^= o[q];
      appf(body, "      { const double2 v = sm[b ^ %uu]; xr[%d] = v.x; xi[%d] = v.y; }\n", off, j, j);
    }
    if (variant & 32) {
      std::string o2;
      o2.reserve(ops.size());
      for (size_t q = 0; q < ops.size(); ++q) {
        if (ops.compare(q, 4, "C.c[") == 0) { o2 += "CC["; q += 3; }
        else o2 += ops[q];
      }
      ops.swap(o2);
    }
    body += ops;
    for (int j = 0; j < PROG_AMPS; ++j) {  // logical amplitude j lives in variable perm[j] (CX renamings)
      uint32_t off = 0;
      for (int q = 0; q < PROG_BITS; ++q) if ((j >> q) & 1) off ^= o[q];
      appf(body, "      sm[b ^ %uu] = make_double2(xr[%d], xi[%d]);\n", off, perm[j], perm[j]);
    }
    body += "      (void)gl;\n    }\n  }\n  __syncthreads();\n";
  }
  const int ncoef = (int)pl.coef.size();

  // ---- prelude + kernel frame
  s.clear();
  s.reserve(body.size() + 8192);
  s += "typedef unsigned int uint32_t;\ntypedef unsigned long long uint64_t;\ntypedef int int32_t;\n";
  appf(s, "#define PROG_AMPS %d\n", PROG_AMPS);
  s += "struct __align__(64) BtTensorMap { unsigned long long opaque[16]; };\n";
  appf(s, "struct BtCoefs { double c[%d]; };\n", ncoef);
  s += "__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }\n";
  if (opt & 7)  // sign-bit flips on the integer pipe (the high word of the double XOR a mask)
    s += "__device__ __forceinline__ double bt_xsign(double x, uint32_t m) { return __hiloint2double(__double2hiint(x) ^ (int)m, __double2loint(x)); }\n"
         "__device__ __forceinline__ double bt_neg(double x) { return bt_xsign(x, 0x80000000u); }\n";
  if (K > 1) {
    // ---- pipelined frame (BT_TILE_PIPE = K, tiles of <= 2^11 amplitudes): a CTA owns K consecutive tiles and two tile buffers; the load of
    // tile i+1 is in flight while the programs of tile i run, the store of tile i drains while tile i+1 computes, and the buffer is
    // refilled with tile i+2 as soon as the store has read it.  Same programs, emitted once inside the tile loop.
    const unsigned tb = (unsigned)(sizeof(double2) << T), ch = tb / (unsigned)P.tma_ncopy;
    s += "__device__ __forceinline__ uint64_t tile_base(uint64_t t) {\n";
    appf(s, "  uint64_t base = t << %d;\n", P.lowb);
    for (int j = P.lowb; j < T; ++j) {
      const int b = P.tbits[j];
      appf(s, "  base = ((base >> %d) << %d) | (base & 0x%llxull);\n", b, b + 1, (unsigned long long)((1ull << b) - 1ull));
    }
    s += "  return base;\n}\n";
    s += "__device__ __forceinline__ void issue_load(uint64_t tm, uint32_t mb, uint32_t dst, uint64_t base) {\n";
    for (int k = 1; k <= 4; ++k) appf(s, "  const int32_t c%d = (int32_t)((base >> %d) & 0x%xu);\n", k, P.tma_coord_shift[k], P.tma_coord_mask[k]);
    appf(s, "  asm volatile(\"mbarrier.arrive.expect_tx.shared::cta.b64 _, [%%0], %%1;\" ::\"r\"(mb), \"r\"(%uu) : \"memory\");\n", tb);
    for (int e = 0; e < P.tma_ncopy; ++e)
      appf(s,
           "  asm volatile(\"cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%%0], [%%1, {%%2, %%3, %%4, %%5, %%6}], [%%7];\" "
           "::\"r\"(dst + %uu), \"l\"(tm), \"r\"(0), \"r\"(c1), \"r\"(c2), \"r\"(c3), \"r\"(c4 + %d), \"r\"(mb) : \"memory\");\n",
           e * ch, P.tma_c4add[e]);
    s += "}\n";
    appf(s, "extern \"C\" __global__ void __launch_bounds__(%d, %d) bt_jit_pass(const __grid_constant__ BtTensorMap tmap, const __grid_constant__ BtCoefs C) {\n", TILE_THREADS,
         jit_minb());
    s += "  extern __shared__ unsigned char smem_raw[];\n  const uint32_t raw = smem_u32(smem_raw);\n  const uint32_t pad = (1024u - (raw & 1023u)) & 1023u;\n";
    s += "  double2* buf0 = reinterpret_cast<double2*>(smem_raw + pad);\n  const uint32_t tid = threadIdx.x;\n";
    appf(s, "  const uint32_t dst0 = smem_u32(buf0), mb0 = dst0 + %uu;\n  const uint64_t tm = reinterpret_cast<uint64_t>(&tmap);\n", 2 * tb);
    appf(s, "  const uint64_t t0 = (uint64_t)blockIdx.x * %du;\n", K);
    s += "  if (tid == 0) {\n    asm volatile(\"mbarrier.init.shared::cta.b64 [%0], 1;\" ::\"r\"(mb0));\n    asm volatile(\"mbarrier.init.shared::cta.b64 [%0], 1;\" ::\"r\"(mb0 + 8u));\n"
         "    asm volatile(\"fence.mbarrier_init.release.cluster;\" ::: \"memory\");\n  }\n  __syncthreads();\n";
    appf(s, "  if (tid == 0) {\n    issue_load(tm, mb0, dst0, tile_base(t0));\n    issue_load(tm, mb0 + 8u, dst0 + %uu, tile_base(t0 + 1));\n  }\n", tb);
    appf(s, "#pragma unroll 1\n  for (uint32_t ti = 0; ti < %du; ++ti) {\n", K);
    appf(s, "    const uint32_t mb = mb0 + 8u * (ti & 1u), dst = dst0 + (ti & 1u) * %uu;\n    double2* sm = buf0 + (ti & 1u) * %uu;\n    const uint64_t base = tile_base(t0 + ti);\n", tb, 1u << T);
    s += "    {\n      const uint32_t parity = (ti >> 1) & 1u;\n      uint32_t done = 0;\n      for (uint32_t spin = 0; !done; ++spin) {\n"
         "        asm volatile(\"{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }\" : \"=r\"(done) : \"r\"(mb), \"r\"(parity) : \"memory\");\n"
         "        if (spin > (1u << 22)) __trap();\n      }\n    }\n";
    s += body;
    s += "    asm volatile(\"fence.proxy.async.shared::cta;\" ::: \"memory\");\n    __syncthreads();\n    if (tid == 0) {\n";
    for (int k = 1; k <= 4; ++k) appf(s, "      const int32_t c%d = (int32_t)((base >> %d) & 0x%xu);\n", k, P.tma_coord_shift[k], P.tma_coord_mask[k]);
    for (int e = 0; e < P.tma_ncopy; ++e)
      appf(s,
           "      asm volatile(\"cp.async.bulk.tensor.5d.global.shared::cta.tile.bulk_group [%%0, {%%1, %%2, %%3, %%4, %%5}], [%%6];\" ::\"l\"(tm), \"r\"(0), \"r\"(c1), \"r\"(c2), "
           "\"r\"(c3), \"r\"(c4 + %d), \"r\"(dst + %uu) : \"memory\");\n",
           P.tma_c4add[e], e * ch);
    s += "      asm volatile(\"cp.async.bulk.commit_group;\" ::: \"memory\");\n";
    appf(s, "      if (ti + 2u < %du) {\n        asm volatile(\"cp.async.bulk.wait_group.read 0;\" ::: \"memory\");\n        issue_load(tm, mb, dst, tile_base(t0 + ti + 2u));\n      }\n    }\n  }\n", K);
    s += "  if (tid == 0) asm volatile(\"cp.async.bulk.wait_group.read 0;\" ::: \"memory\");\n}\n";
    return true;
  }
  appf(s, "extern \"C\" __global__ void __launch_bounds__(%d, %d) bt_jit_pass(const __grid_constant__ BtTensorMap tmap, const __grid_constant__ BtCoefs C) {\n",
       wide ? 2 * TILE_THREADS : TILE_THREADS, wide ? 2 : jit_minb());
  s += "  extern __shared__ unsigned char smem_raw[];\n  const uint32_t raw = smem_u32(smem_raw);\n  const uint32_t pad = (1024u - (raw & 1023u)) & 1023u;\n";
  s += "  double2* sm = reinterpret_cast<double2*>(smem_raw + pad);\n  const uint32_t tid = threadIdx.x;\n";
  appf(s, "  uint64_t base = (uint64_t)blockIdx.x << %d;\n", P.lowb);
  for (int j = P.lowb; j < T; ++j) {
    const int b = P.tbits[j];
    appf(s, "  base = ((base >> %d) << %d) | (base & 0x%llxull);\n", b, b + 1, (unsigned long long)((1ull << b) - 1ull));
  }
  for (int k = 1; k <= 4; ++k)
    appf(s, "  const int32_t c%d = (int32_t)((base >> %d) & 0x%xu);\n", k, P.tma_coord_shift[k], P.tma_coord_mask[k]);
  const unsigned tile_bytes = (unsigned)(sizeof(double2) << T), chunk = tile_bytes / (unsigned)P.tma_ncopy;
  appf(s, "  const uint32_t dst = smem_u32(sm), mb = dst + %uu;\n  const uint64_t tm = reinterpret_cast<uint64_t>(&tmap);\n", tile_bytes);
  s += "  if (tid == 0) {\n    asm volatile(\"mbarrier.init.shared::cta.b64 [%0], 1;\" ::\"r\"(mb));\n    asm volatile(\"fence.mbarrier_init.release.cluster;\" ::: \"memory\");\n  }\n  __syncthreads();\n";
  s += "  if (tid == 0) {\n";
  appf(s, "    asm volatile(\"mbarrier.arrive.expect_tx.shared::cta.b64 _, [%%0], %%1;\" ::\"r\"(mb), \"r\"(%uu) : \"memory\");\n", tile_bytes);
  for (int e = 0; e < P.tma_ncopy; ++e)
    appf(s,
         "    asm volatile(\"cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%%0], [%%1, {%%2, %%3, %%4, %%5, %%6}], [%%7];\" "
         "::\"r\"(dst + %uu), \"l\"(tm), \"r\"(0), \"r\"(c1), \"r\"(c2), \"r\"(c3), \"r\"(c4 + %d), \"r\"(mb) : \"memory\");\n",
         e * chunk, P.tma_c4add[e]);
  if (variant & 64) {
    // L2 prefetch of the tile a later CTA of this SM slot will load (blockIdx + BT_JIT_PREFETCH_DIST, default 3 CTAs x 148 SMs): the pass is
    // compute-bound, 12 % of the warp samples wait for the tile load -- from L2 the wait is shorter than from HBM
    const int dist = env_i("BT_JIT_PREFETCH_DIST", 444);
    appf(s, "    if (blockIdx.x + %du < gridDim.x) {\n      uint64_t b2 = (uint64_t)(blockIdx.x + %du) << %d;\n", dist, dist, P.lowb);
    for (int j = P.lowb; j < T; ++j) {
      const int b = P.tbits[j];
      appf(s, "      b2 = ((b2 >> %d) << %d) | (b2 & 0x%llxull);\n", b, b + 1, (unsigned long long)((1ull << b) - 1ull));
    }
    for (int e = 0; e < P.tma_ncopy; ++e)
      appf(s,
           "      asm volatile(\"cp.async.bulk.prefetch.tensor.5d.L2.global.tile [%%0, {%%1, %%2, %%3, %%4, %%5}];\" ::\"l\"(tm), \"r\"(0), \"r\"((int32_t)((b2 >> %d) & 0x%xu)), "
           "\"r\"((int32_t)((b2 >> %d) & 0x%xu)), \"r\"((int32_t)((b2 >> %d) & 0x%xu)), \"r\"((int32_t)((b2 >> %d) & 0x%xu) + %d) : \"memory\");\n",
           P.tma_coord_shift[1], P.tma_coord_mask[1], P.tma_coord_shift[2], P.tma_coord_mask[2], P.tma_coord_shift[3], P.tma_coord_mask[3], P.tma_coord_shift[4],
           P.tma_coord_mask[4], P.tma_c4add[e]);
    s += "    }\n";
  }
  s += "  }\n  {\n    uint32_t done = 0;\n    for (uint32_t spin = 0; !done; ++spin) {\n"
       "      asm volatile(\"{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }\" : \"=r\"(done) : \"r\"(mb) : \"memory\");\n"
       "      if (spin > (1u << 22)) __trap();\n    }\n  }\n";
  s += body;
  s += "  asm volatile(\"fence.proxy.async.shared::cta;\" ::: \"memory\");\n  __syncthreads();\n  if (tid == 0) {\n";
  for (int e = 0; e < P.tma_ncopy; ++e)
    appf(s,
         "    asm volatile(\"cp.async.bulk.tensor.5d.global.shared::cta.tile.bulk_group [%%0, {%%1, %%2, %%3, %%4, %%5}], [%%6];\" ::\"l\"(tm), \"r\"(0), \"r\"(c1), \"r\"(c2), "
         "\"r\"(c3), \"r\"(c4 + %d), \"r\"(dst + %uu) : \"memory\");\n",
         P.tma_c4add[e], e * chunk);
  s += "    asm volatile(\"cp.async.bulk.commit_group;\" ::: \"memory\");\n    asm volatile(\"cp.async.bulk.wait_group.read 0;\" ::: \"memory\");\n  }\n}\n";
  return true;
}

struct Entry {
  int seen = 0;
  int state = 0;  // 0 = not compiled, 1 = ready, -1 = interpreter only, 2 = with the compile workers, 3 = cubin ready, module not loaded yet
  CUfunction fn = nullptr;
  int ncoef = 0;
  int smem_bytes = 0;
  int threads = TILE_THREADS;
  int tiles_per_cta = 1;
  std::vector<char> cubin;  // state 3 only
};

// Never destroyed (references to leaked heap objects): the detached compile workers wait on the condition variable for the life of
// the process, and destroying a condition variable that has waiters blocks in pthread_cond_destroy -- a static object here would hang
// every process at exit once the workers exist.
std::mutex& g_mu = *new std::mutex();
std::condition_variable& g_cv_jobs = *new std::condition_variable();
std::condition_variable& g_cv_idle = *new std::condition_variable();
std::unordered_map<std::string, Entry>& g_cache = *new std::unordered_map<std::string, Entry>();  // node-based: Entry addresses are stable, the workers keep pointers
uint64_t g_compiled = 0, g_launches = 0, g_failed = 0;
double g_compile_seconds = 0.0;
std::string& g_last_log = *new std::string();

// ---- optional on-disk cubin cache (BT_JIT_CACHE_DIR; off when unset) -----------------------------------------------------
// A pass structure costs ~0.2 s of NVRTC time; a process that runs the same circuits as an earlier one (the next job of a
// parameter sweep, the next rank of a node) can take the cubin from disk instead.  File name = two independent 64-bit FNV-1a
// hashes of the generated source + the compile options; written to a temporary name and renamed, so a reader never sees a
// partial file; any I/O or load failure simply falls through to a fresh compilation.
uint64_t fnv1a(const std::string& s, uint64_t h) {
  for (unsigned char c : s) { h ^= c; h *= 1099511628211ull; }
  return h;
}

const char* const kJitOptions[] = {"--gpu-architecture=sm_100a", "-std=c++17", "-lineinfo"};

void mkdir_p(const std::string& d) {
  for (size_t i = 1; i <= d.size(); ++i)
    if (i == d.size() || d[i] == '/') mkdir(d.substr(0, i).c_str(), 0755);
}

// BT_JIT_CACHE_DIR set: that directory ("" = no cache); unset: $XDG_CACHE_HOME/bluetangle_cuda, else ~/.cache/bluetangle_cuda
std::string cache_dir() {
  const char* dir = getenv("BT_JIT_CACHE_DIR");
  std::string d;
  if (dir) d = dir;
  else if (const char* x = getenv("XDG_CACHE_HOME"); x && *x) d = std::string(x) + "/bluetangle_cuda";
  else if (const char* h = getenv("HOME"); h && *h) d = std::string(h) + "/.cache/bluetangle_cuda";
  if (d.empty()) return d;
  static std::mutex& mu = *new std::mutex();        // leaked on purpose: the compile workers may still be running at exit
  static std::string& made = *new std::string();
  std::lock_guard<std::mutex> lk(mu);
  if (made != d) { mkdir_p(d); made = d; }
  return d;
}

std::string cache_path(const std::string& src) {
  const std::string dir = cache_dir();
  if (dir.empty()) return std::string();
  std::string salted = src;
  for (const char* o : kJitOptions) { salted += '\n'; salted += o; }
  if (const char* x = getenv("BT_JIT_EXTRA_OPTS")) { salted += '\n'; salted += x; }
  char name[80];
  snprintf(name, sizeof(name), "/btjit_%016llx%016llx.cubin", (unsigned long long)fnv1a(salted, 14695981039346656037ull),
           (unsigned long long)fnv1a(salted, 0x9e3779b97f4a7c15ull));
  return dir + name;
}

bool cache_read(const std::string& path, std::vector<char>& cubin) {
  if (path.empty()) return false;
  FILE* f = fopen(path.c_str(), "rb");
  if (!f) return false;
  bool ok = false;
  if (fseek(f, 0, SEEK_END) == 0) {
    const long n = ftell(f);
    if (n > 0 && n < (64l << 20) && fseek(f, 0, SEEK_SET) == 0) {
      cubin.resize((size_t)n);
      ok = fread(cubin.data(), 1, (size_t)n, f) == (size_t)n;
    }
  }
  fclose(f);
  return ok && cubin.size() > 4 && memcmp(cubin.data(), "\x7f" "ELF", 4) == 0;
}

void cache_write(const std::string& path, const std::vector<char>& cubin) {
  if (path.empty() || cubin.empty()) return;
  char tmp[64];
  snprintf(tmp, sizeof(tmp), ".tmp%ld_%llx", (long)getpid(), (unsigned long long)std::hash<std::thread::id>()(std::this_thread::get_id()));
  const std::string t = path + tmp;
  FILE* f = fopen(t.c_str(), "wb");
  if (!f) return;
  const bool ok = fwrite(cubin.data(), 1, cubin.size(), f) == cubin.size();
  if (fclose(f) != 0 || !ok || rename(t.c_str(), path.c_str()) != 0) remove(t.c_str());
}

uint64_t g_cache_hits = 0;

bool load_cubin(const std::vector<char>& cubin, Entry& e, int smem_bytes) {
  Driver& d = driver();
  CUmodule mod = nullptr;
  if (d.ModuleLoadData(&mod, cubin.data()) != CUDA_SUCCESS) { g_last_log = "cuModuleLoadData failed"; return false; }
  if (d.ModuleGetFunction(&e.fn, mod, "bt_jit_pass") != CUDA_SUCCESS) { g_last_log = "cuModuleGetFunction failed"; return false; }
  if (d.FuncSetAttribute(e.fn, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, smem_bytes) != CUDA_SUCCESS) { g_last_log = "cuFuncSetAttribute failed"; return false; }
  return true;
}

// source text -> cubin through NVRTC (no device needed; safe to call from several threads, one program each)
bool nvrtc_to_cubin(const std::string& src, std::vector<char>& cubin, std::string& log) {
  Nvrtc& n = nvrtc();
  if (!n.ok) { log = "libnvrtc is not available"; return false; }
  nvrtcProgram prog = nullptr;
  if (n.CreateProgram(&prog, src.c_str(), "bt_jit_pass.cu", 0, nullptr, nullptr) != 0) return false;
  // BT_JIT_EXTRA_OPTS: space-separated additional NVRTC options (experiments: --fmad=false, -Xptxas=-O0 ...)
  std::vector<std::string> extra;
  if (const char* x = getenv("BT_JIT_EXTRA_OPTS")) {
    std::string cur;
    for (const char* c = x;; ++c) {
      if (*c == ' ' || *c == '\0') { if (!cur.empty()) extra.push_back(cur); cur.clear(); if (!*c) break; }
      else cur += *c;
    }
  }
  std::vector<const char*> opts(kJitOptions, kJitOptions + 3);
  for (const std::string& e : extra) opts.push_back(e.c_str());
  const int rc = n.CompileProgram(prog, (int)opts.size(), opts.data());
  if (rc != 0) {
    size_t ls = 0;
    n.GetProgramLogSize(prog, &ls);
    log.assign(ls, '\0');
    if (ls) n.GetProgramLog(prog, &log[0]);
    n.DestroyProgram(&prog);
    return false;
  }
  size_t cs = 0;
  n.GetCUBINSize(prog, &cs);
  cubin.resize(cs);
  n.GetCUBIN(prog, cubin.data());
  n.DestroyProgram(&prog);
  return cs > 0;
}

// cubin for a source text: from the disk cache or through NVRTC (and then into the cache).  No CUDA context needed.
bool obtain_cubin(const std::string& src, std::vector<char>& cubin, std::string& log, bool* from_cache) {
  const std::string path = cache_path(src);
  *from_cache = false;
  if (cache_read(path, cubin)) { *from_cache = true; return true; }
  cubin.clear();
  if (!nvrtc_to_cubin(src, cubin, log)) return false;
  cache_write(path, cubin);
  return true;
}

}  // namespace
bool bt_jit_source_for(const TileParams& P, int np, std::string& src);
namespace {
// ---- compile workers -----------------------------------------------------------------------------------------------------
struct Job { std::string src; Entry* e; };
std::deque<Job>* g_jobs = nullptr;   // heap objects that are never destroyed: the detached workers may outlive static destructors
int g_workers = 0, g_pending = 0;

void worker_main() {
  for (;;) {
    Job j;
    {
      std::unique_lock<std::mutex> lk(g_mu);
      g_cv_jobs.wait(lk, [] { return !g_jobs->empty(); });
      j = std::move(g_jobs->front());
      g_jobs->pop_front();
    }
    const auto t0 = std::chrono::steady_clock::now();
    std::vector<char> cubin;
    std::string log;
    bool hit = false;
    const bool ok = obtain_cubin(j.src, cubin, log, &hit);
    const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    {
      std::lock_guard<std::mutex> lk(g_mu);
      g_compile_seconds += dt;
      if (ok) { j.e->cubin = std::move(cubin); j.e->state = 3; if (hit) g_cache_hits++; }
      else { j.e->state = -1; g_failed++; g_last_log = log; }
      if (--g_pending == 0) g_cv_idle.notify_all();
    }
  }
}

// g_mu held
void enqueue_compile(std::string&& src, Entry* e) {
  if (!g_jobs) g_jobs = new std::deque<Job>();
  // default: min(8, cores / processes of this job on the node) -- eight ranks with eight compile threads each on a 16-core host starve
  // the threads that enqueue the kernels (LOCAL_WORLD_SIZE is set by torchrun; MPI launchers set OMPI_COMM_WORLD_LOCAL_SIZE)
  int local_world = std::max(1, std::max(env_i("LOCAL_WORLD_SIZE", 1), env_i("OMPI_COMM_WORLD_LOCAL_SIZE", 1)));
  const int cores = (int)std::max(1u, std::thread::hardware_concurrency());
  const int want = std::max(1, std::min(env_i("BT_JIT_THREADS", 8), std::max(1, cores / local_world)));
  while (g_workers < want) { std::thread(worker_main).detach(); ++g_workers; }
  e->state = 2;
  g_pending++;
  g_jobs->push_back(Job{std::move(src), e});
  g_cv_jobs.notify_one();
}

}  // namespace

// debugging aid: only the k-th eligible pass since the call (0-based) may run specialised, every other one stays on the interpreter
// (k < 0: no filter); with dump != 0 the CUDA text of that pass goes to stderr
static int g_only = -1, g_only_seq = 0, g_only_dump = 0;
extern "C" int bt_jit_debug_only(int k, int dump) {
  std::lock_guard<std::mutex> lk(g_mu);
  g_only = k; g_only_seq = 0; g_only_dump = dump;
  return BT_OK;
}

// 1: the pass was launched as a specialised kernel; 0: not (the caller runs the interpreter).  Never fails the pass.
int bt_jit_try_launch(bt_sv* s, const TileParams& P, const CUtensorMap& tmap, uint64_t ntiles, size_t tile_bytes, int np) {
  const int mode = env_i("BT_TILE_JIT", 1);
  if (mode == 0 || np <= 0 || P.nitems != np || P.swz_mode != 0) return 0;
  if (g_only >= 0) {
    const int seq = g_only_seq++;
    if (seq != g_only) return 0;
    if (g_only_dump) {
      std::string src;
      if (bt_jit_source_for(P, np, src)) fprintf(stderr, "// ===== specialised pass %d =====\n%s\n", seq, src.c_str());
    }
  }
  if (mode != 2 && s->len < (1ull << env_i("BT_TILE_JIT_MINBITS", 22))) return 0;
  Plan pl;
  if (!make_plan(P, np, pl)) return 0;
  // pipelined frame: K tiles per CTA with two tile buffers (tiles of <= 2^11 amplitudes: 2 x 32 KB, still three CTAs per SM)
  int K = env_i("BT_TILE_PIPE", 0);
  if (P.T > 11 || K < 2 || (K & (K - 1)) != 0) K = 1;
  while (K > 1 && (ntiles % (uint64_t)K != 0 || ntiles / (uint64_t)K < 6ull * 148ull)) K >>= 1;
  std::string key;
  make_key(P, pl, s->device, key, K);
  std::lock_guard<std::mutex> lk(g_mu);
  if (g_cache.size() > 8192 && g_cache.find(key) == g_cache.end()) return 0;  // bounded: a long-running host with ever-new passes keeps interpreting
  Entry& e = g_cache[key];
  e.seen++;
  if (e.state < 0 || e.state == 2) return 0;
  if (e.state == 0) {
    if (mode != 2 && e.seen < env_i("BT_TILE_JIT_AFTER", 1)) return 0;
    if (!driver().ok || !nvrtc().ok) { e.state = -1; g_failed++; g_last_log = "NVRTC or the driver entry points are not available"; return 0; }
    std::string src;
    if (!generate(P, pl, src, K)) { e.state = -1; g_failed++; return 0; }
    e.ncoef = (int)pl.coef.size();
    e.smem_bytes = (int)((K > 1 ? 2 : 1) * tile_bytes + 1024 + 64);
    e.threads = (wide_ok(P) && K <= 1) ? 2 * TILE_THREADS : TILE_THREADS;
    e.tiles_per_cta = K;
    if (mode != 2 && env_i("BT_TILE_JIT_ASYNC", 1) != 0) {
      enqueue_compile(std::move(src), &e);  // the interpreter runs this pass; the cubin is picked up at a later launch
      return 0;
    }
    const auto t0 = std::chrono::steady_clock::now();
    bool hit = false;
    const bool ok = obtain_cubin(src, e.cubin, g_last_log, &hit);
    g_compile_seconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (!ok) {
      e.state = -1;
      g_failed++;
      if (env_i("BT_TILE_DEBUG", 0)) fprintf(stderr, "[jit] pass not specialised: %s\n", g_last_log.c_str());
      return 0;
    }
    if (hit) g_cache_hits++;
    e.state = 3;
  }
  if (e.state == 3) {  // module load needs the caller's CUDA context: done here, not on the workers
    const bool ok = load_cubin(e.cubin, e, e.smem_bytes);
    std::vector<char>().swap(e.cubin);
    if (!ok) { e.state = -1; g_failed++; return 0; }
    e.state = 1;
    g_compiled++;
  }
  if ((int)pl.coef.size() != e.ncoef) return 0;
  std::vector<double>& coef = pl.coef;
  void* args[2] = {(void*)&tmap, (void*)coef.data()};
  if (driver().LaunchKernel(e.fn, (unsigned)(ntiles / (uint64_t)e.tiles_per_cta), 1, 1, (unsigned)e.threads, 1, 1, (unsigned)e.smem_bytes, (CUstream)s->stream, args, nullptr) != CUDA_SUCCESS) {
    e.state = -1;
    return 0;
  }
  g_launches++;
  return 1;
}

// host-only debugging aid: the CUDA text the specialiser would compile for this pass (planning dry runs: BT_JIT_DUMP=1)
bool bt_jit_source_for(const TileParams& P, int np, std::string& src) {
  Plan pl;
  if (!make_plan(P, np, pl)) return false;
  if (!generate(P, pl, src)) return false;
  char buf[64];
  {
    std::string key;
    make_key(P, pl, 0, key);
    unsigned long long h = 1469598103934665603ull;
    for (unsigned char ch : key) { h ^= ch; h *= 1099511628211ull; }
    snprintf(buf, sizeof(buf), "// key %016llx\n", h);
    src += buf;
  }
  src += "// coefficients:";
  for (size_t i = 0; i < pl.coef.size(); ++i) { snprintf(buf, sizeof(buf), " [%zu]=%.17g", i, pl.coef[i]); src += buf; }
  src += "\n";
  return true;
}

// blocks until the compile workers are idle; *pending_before = structures that were still being compiled at the call
extern "C" int bt_jit_wait(uint64_t* pending_before) {
  std::unique_lock<std::mutex> lk(g_mu);
  if (pending_before) *pending_before = (uint64_t)g_pending;
  g_cv_idle.wait(lk, [] { return g_pending == 0; });
  return BT_OK;
}

// modules taken from the on-disk cache instead of NVRTC, and the cache directory in use ("" = off)
extern "C" int bt_jit_cache_info(uint64_t* disk_hits, char* dir, uint64_t cap) {
  const std::string d = cache_dir();
  std::lock_guard<std::mutex> lk(g_mu);
  if (disk_hits) *disk_hits = g_cache_hits;
  if (dir && cap) { strncpy(dir, d.c_str(), cap - 1); dir[cap - 1] = 0; }
  return BT_OK;
}

extern "C" int bt_jit_stats(uint64_t* compiled, uint64_t* launches, uint64_t* failed, double* compile_seconds) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (compiled) *compiled = g_compiled;
  if (launches) *launches = g_launches;
  if (failed) *failed = g_failed;
  if (compile_seconds) *compile_seconds = g_compile_seconds;
  return BT_OK;
}

extern "C" int bt_jit_config(int* nvrtc_major, int* nvrtc_minor, int* variant, int* opt) {
  Nvrtc& n = nvrtc();
  if (nvrtc_major) *nvrtc_major = n.ok ? n.major : 0;
  if (nvrtc_minor) *nvrtc_minor = n.ok ? n.minor : 0;
  if (variant) *variant = jit_variant();
  if (opt) *opt = jit_opt();
  return BT_OK;
}

// synthetic pass that uses every micro-op site kind, as CUDA text (host only)
static bool selftest_source(std::string& src) {
  TileParams* Pp = new TileParams();
  TileParams& P = *Pp;
  memset(&P, 0, sizeof(P));
  P.T = 12; P.lowb = 5; P.swz_mode = 0; P.tma_ncopy = 2; P.tma_c4add[1] = 4;
  for (int j = 0; j < 12; ++j) P.tbits[j] = j < 5 ? j : 2 * j;
  for (int k = 1; k <= 4; ++k) { P.tma_coord_shift[k] = 3 * k; P.tma_coord_mask[k] = 0xff; }
  P.nitems = 1; P.item[0] = TILE_PBASE;
  TileProg& G = P.pr[0];
  for (int q = 0; q < PROG_BITS; ++q) G.lp[q] = 2 * q + 1;
  for (int k = 0; k < 7; ++k) { G.bit_sw[k] = 1u << (2 * k); G.bit_lin[k] = 1u << (2 * k); }
  G.niter = 2; G.iter_sw[1] = 1u << 10; G.iter_lin[1] = 1u << 10;
  int ko = 0, kc = 0;
  auto emit = [&](int site, int ncf) { G.op[ko++] = (uint16_t)(site | (kc << 8)); for (int e = 0; e < ncf; ++e) G.coef[kc++] = (e == 3 ? 0.25 : 0.25 * (e + 1)); };
  for (int kind = 0; kind < 5; ++kind) emit(PROG_SITE_U1(kind, kind % PROG_BITS), kind == 0 ? 8 : 4);
  emit(PROG_SITE_CX(0, 2), 0); emit(PROG_SITE_CPHASE(1, 3), 2);
  { double m; uint64_t em = 1ull << 20, lm = 1ull << 6; emit(PROG_SITE_CSCALE, 2); memcpy(&m, &em, 8); G.coef[kc++] = m; memcpy(&m, &lm, 8); G.coef[kc++] = m; }
  { double m; uint64_t em = 0, lm = 1ull << 8; emit(PROG_SITE_CPH1(2), 2); memcpy(&m, &em, 8); G.coef[kc++] = m; memcpy(&m, &lm, 8); G.coef[kc++] = m; }
  { double m; uint64_t em = 1ull << 22, lm = 0; emit(PROG_SITE_CCX1(1), 0); memcpy(&m, &em, 8); G.coef[kc++] = m; memcpy(&m, &lm, 8); G.coef[kc++] = m; }
  G.nops = (uint32_t)ko;
  Plan pl;
  const bool ok = make_plan(P, 1, pl) && generate(P, pl, src, std::max(1, env_i("BT_JIT_SELFTEST_PIPE", 1)));  // BT_JIT_SELFTEST_PIPE=K: the pipelined frame
  delete Pp;
  return ok;
}

// Host-only self test (no device needed): specialise a synthetic pass that uses every micro-op site kind and compile it with
// NVRTC.  Returns 0 on success; `source` (optional) receives the generated text.
extern "C" int bt_jit_selftest(char* source, uint64_t cap) {
  std::string src;
  int rc = selftest_source(src) ? 0 : -1;
  if (rc == 0 && source && cap) { strncpy(source, src.c_str(), cap - 1); source[cap - 1] = 0; }
  if (rc == 0) {
    if (!nvrtc().ok) rc = -2;
    else {
      std::vector<char> cubin;
      std::string log;
      if (!nvrtc_to_cubin(src, cubin, log)) {
        rc = -4;
        bt_set_error("%s", log.substr(0, 900).c_str());
      } else {
        // the on-disk cache, when BT_JIT_CACHE_DIR is set: what was written must come back byte for byte
        const std::string path = cache_path(src);
        if (!path.empty()) {
          std::vector<char> back;
          cache_write(path, cubin);
          if (!cache_read(path, back) || back != cubin) rc = -5;
        }
      }
    }
  }
  return rc;
}

// Host-only check of the compile workers: `jobs` variants of the synthetic pass go through the worker pool (disk cache or NVRTC),
// bt_jit_wait() must see them all finish.  0 = ok, -2 = no libnvrtc.  tests/test_build_tools.py also checks that a process that has
// started the workers still EXITS (the workers are detached and wait on a condition variable for the life of the process).
extern "C" int bt_jit_selftest_workers(int jobs) {
  std::string src;
  if (!selftest_source(src)) return -1;
  if (!nvrtc().ok) return -2;
  std::vector<Entry*> es;
  {
    std::lock_guard<std::mutex> lk(g_mu);
    for (int i = 0; i < jobs; ++i) {
      char key[64];
      snprintf(key, sizeof(key), "selftest-worker#%d#%p", i, (void*)&es);
      Entry& e = g_cache[key];
      es.push_back(&e);
      char tag[64];
      snprintf(tag, sizeof(tag), "\n// selftest variant %d\n", i);
      enqueue_compile(src + tag, &e);
    }
  }
  bt_jit_wait(nullptr);
  std::lock_guard<std::mutex> lk(g_mu);
  for (Entry* e : es) {
    if (e->state != 3 || e->cubin.size() < 1000) { bt_set_error("%s", g_last_log.substr(0, 900).c_str()); return -4; }
    std::vector<char>().swap(e->cubin);
    e->state = -1;
  }
  return 0;
}
