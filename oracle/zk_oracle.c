/* oracle/zk_oracle.c -- TEST INFRASTRUCTURE (CPU oracle), not product code.  See zk_oracle.h.
 * Part 1: field exports, Poseidon2, sponge/compression, DFT/LDE, MMCS, Pcs::commit. */
#include "zk_oracle.h"
#include "kb31.h"
#include "../include/zk_poseidon2_rc.h"
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------------------------------ */
/* field exports                                                                              */
/* ------------------------------------------------------------------------------------------ */
uint32_t ork_to_monty(uint32_t c) { return kb_from_u32(c); }
uint32_t ork_from_monty(uint32_t m) { return kb_from_monty(m); }
uint32_t ork_mul(uint32_t a, uint32_t b) { return kb_mul(a, b); }
uint32_t ork_inv(uint32_t a) { return kb_inv(a); }
uint32_t ork_two_adic_generator(uint32_t bits) { return kb_two_adic_generator(bits); }
void ork_ext_mul(const uint32_t a[4], const uint32_t b[4], uint32_t out[4]) {
  kb4_t x, y; memcpy(x.c, a, 16); memcpy(y.c, b, 16);
  kb4_t r = kb4_mul(x, y); memcpy(out, r.c, 16);
}
void ork_ext_inv(const uint32_t a[4], uint32_t out[4]) {
  kb4_t x; memcpy(x.c, a, 16);
  kb4_t r = kb4_inv(x); memcpy(out, r.c, 16);
}
void ork_to_monty_vec(const uint32_t* c, uint32_t* out, uint64_t n) {
#pragma omp parallel for schedule(static)
  for (uint64_t i = 0; i < n; i++) out[i] = kb_from_u32(c[i]);
}
void ork_from_monty_vec(const uint32_t* m, uint32_t* out, uint64_t n) {
#pragma omp parallel for schedule(static)
  for (uint64_t i = 0; i < n; i++) out[i] = kb_from_monty(m[i]);
}
int32_t ork_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
/* overrides OMP_NUM_THREADS (torchrun exports OMP_NUM_THREADS=1 to its workers) */
void ork_set_num_threads(int32_t n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

/* ------------------------------------------------------------------------------------------ */
/* Poseidon2 width 16, x^3, 8 external + 13 internal rounds                                   */
/*   linear layers: crates/recursion/core/include/poseidon2.hpp:21-71                          */
/*   round schedule: crates/recursion/core/include/poseidon2_wide.hpp:10-146,                   */
/*                   crates/recursion/gnark-ffi/go/zkm/poseidon2/poseidon2_koalabear.go:26-55  */
/*   constants: crates/primitives/src/lib.rs:563-1121 (via include/zk_poseidon2_rc.h)          */
/* ------------------------------------------------------------------------------------------ */
static const uint32_t P2_EXT_RC[8][16] = ZK_P2_EXT_RC_MONTY;
static const uint32_t P2_INT_RC[13] = ZK_P2_INT_RC_MONTY;

/* V = [-2, 1, 2, 1/2, 3, 4, -1/2, -3, -4, 1/2^8, 1/8, 1/2^24, -1/2^8, -1/8, -1/16, -1/2^24]
 * given as canonical representatives exactly as poseidon2_constants.hpp:1083-1100 writes them. */
static const uint32_t P2_DIAG_CANON[16] = {
    KB_P - 2, 1, 2, (KB_P + 1) >> 1, 3, 4, (KB_P - 1) >> 1, KB_P - 3, KB_P - 4,
    KB_P - ((KB_P - 1) >> 8), KB_P - ((KB_P - 1) >> 3), KB_P - 127,
    (KB_P - 1) >> 8, (KB_P - 1) >> 3, (KB_P - 1) >> 4, 127};

static void p2_m4(kb_t* x) { /* poseidon2.hpp:21-31 */
  kb_t t01 = kb_add(x[0], x[1]);
  kb_t t23 = kb_add(x[2], x[3]);
  kb_t t0123 = kb_add(t01, t23);
  kb_t t01123 = kb_add(t0123, x[1]);
  kb_t t01233 = kb_add(t0123, x[3]);
  kb_t n3 = kb_add(t01233, kb_dbl(x[0]));
  kb_t n1 = kb_add(t01123, kb_dbl(x[2]));
  kb_t n0 = kb_add(t01123, t01);
  kb_t n2 = kb_add(t01233, t23);
  x[0] = n0; x[1] = n1; x[2] = n2; x[3] = n3;
}
static void p2_external(kb_t* s) { /* poseidon2.hpp:34-50 */
  for (int i = 0; i < 16; i += 4) p2_m4(s + i);
  kb_t sums[4] = {0, 0, 0, 0};
  for (int k = 0; k < 4; k++)
    for (int j = 0; j < 16; j += 4) sums[k] = kb_add(sums[k], s[j + k]);
  for (int j = 0; j < 16; j++) s[j] = kb_add(s[j], sums[j % 4]);
}
static void p2_internal(kb_t* s, const kb_t* diag) { /* poseidon2.hpp:53-71 */
  kb_t sum = 0;
  for (int i = 0; i < 16; i++) sum = kb_add(sum, s[i]);
  for (int i = 0; i < 16; i++) s[i] = kb_add(kb_mul(s[i], diag[i]), sum);
}
static inline kb_t p2_cube(kb_t x) { return kb_mul(kb_mul(x, x), x); }

void ork_poseidon2_permute(uint32_t* s) {
  kb_t diag[16];
  for (int i = 0; i < 16; i++) diag[i] = kb_to_monty(P2_DIAG_CANON[i]);
  p2_external(s);
  for (int r = 0; r < 4; r++) {
    for (int i = 0; i < 16; i++) s[i] = p2_cube(kb_add(s[i], P2_EXT_RC[r][i]));
    p2_external(s);
  }
  for (int r = 0; r < 13; r++) {
    s[0] = p2_cube(kb_add(s[0], P2_INT_RC[r]));
    p2_internal(s, diag);
  }
  for (int r = 4; r < 8; r++) {
    for (int i = 0; i < 16; i++) s[i] = p2_cube(kb_add(s[i], P2_EXT_RC[r][i]));
    p2_external(s);
  }
}
/* ------------------------------------------------------------------------------------------ */
/* Trace fillers (restated; checked row for row against the reference's own C++ in oracle/_ref) */
/* ------------------------------------------------------------------------------------------ */
/* Poseidon2WideChip<DEGREE>::populate_perm (crates/recursion/core/src/chips/poseidon2_wide/trace.rs:271-330,
 * populate_external_round :332-379, populate_internal_rounds :381-420; C++ twin poseidon2_wide.hpp:10-197).
 * Row layout = PermutationState then PermutationSBoxState (columns/permutation.rs:20-35):
 *   external_rounds_state[8][16] | internal_rounds_state[16] | internal_rounds_s0[12] | output_state[16]
 *   | external_rounds_sbox[8][16] | internal_rounds_sbox[13]           (172 words, 313 with the S-box columns) */
static void p2_wide_row(const kb_t in[16], kb_t* row, int sbox) {
  kb_t diag[16], st[16];
  for (int i = 0; i < 16; i++) diag[i] = kb_to_monty(P2_DIAG_CANON[i]);
  kb_t* ext = row;             /* [8][16] */
  kb_t* ist = row + 128;       /* [16]    */
  kb_t* s0 = row + 144;        /* [12]    */
  kb_t* out = row + 156;       /* [16]    */
  kb_t* xsb = row + 172;       /* [8][16] */
  kb_t* isb = row + 300;       /* [13]    */
  memcpy(ext, in, 64);
  for (int r = 0; r < 8; r++) {
    memcpy(st, ext + 16 * r, 64);
    if (r == 0) p2_external(st);
    for (int i = 0; i < 16; i++) {
      st[i] = p2_cube(kb_add(st[i], P2_EXT_RC[r][i]));
      if (sbox) xsb[16 * r + i] = st[i];
    }
    p2_external(st);
    if (r == 3) {
      memcpy(ist, st, 64);
      for (int k = 0; k < 13; k++) {
        st[0] = p2_cube(kb_add(st[0], P2_INT_RC[k]));
        if (sbox) isb[k] = st[0];
        p2_internal(st, diag);
        if (k < 12) s0[k] = st[0];
      }
      memcpy(ext + 64, st, 64);
    } else if (r == 7) {
      memcpy(out, st, 64);
    } else {
      memcpy(ext + 16 * (r + 1), st, 64);
    }
  }
}
/* generate_trace (trace.rs:76-108): one row per event, padding rows = the row of the all-zero input */
void ork_poseidon2_wide_trace(const uint32_t* inputs, uint64_t n_events, uint64_t rows, int32_t sbox, uint32_t* out) {
  uint64_t w = sbox ? 313 : 172;
  kb_t zero[16];
  memset(zero, 0, sizeof zero);
#pragma omp parallel for schedule(static)
  for (uint64_t r = 0; r < rows; r++) p2_wide_row(r < n_events ? inputs + 16 * r : zero, out + r * w, sbox);
}
/* generate_preprocessed_trace (trace.rs:183-216): instr = input addrs[16], output addrs[16], mults[16] ->
 * input[16], output[16] x {addr, mult}, is_real_neg = -1; padding rows are zero */
void ork_poseidon2_wide_prep(const uint32_t* instrs, uint64_t n, uint64_t rows, uint32_t* out) {
  memset(out, 0, rows * 49 * sizeof(uint32_t));
  for (uint64_t r = 0; r < n; r++) {
    const uint32_t* in = instrs + 48 * r;
    uint32_t* o = out + 49 * r;
    for (int i = 0; i < 16; i++) {
      o[i] = in[i];
      o[16 + 2 * i] = in[16 + i];
      o[17 + 2 * i] = in[32 + i];
    }
    o[48] = kb_neg(KB_ONE);
  }
}
/* AddSubChip::event_to_row + AddOperation::populate (crates/core/machine/src/alu/add_sub/mod.rs:150-172,
 * operations/add.rs:26-60; C++ twin crates/core/machine/include/add_sub.hpp).  event = AluEvent #[repr(C)]
 * {pc, next_pc, opcode:u8 (+3 pad), hi, a, b, c} = 7 words; row = AddSubCols (19 words); padding rows zero. */
void ork_add_sub_trace(const uint32_t* events, uint64_t n_events, uint64_t rows, uint32_t* out) {
  memset(out, 0, rows * 19 * sizeof(uint32_t));
  for (uint64_t r = 0; r < n_events; r++) {
    const uint32_t* e = events + 7 * r;
    uint32_t* o = out + 19 * r;
    int is_add = (e[2] & 0xFF) == 0; /* Opcode::ADD = 0 */
    uint32_t op1 = is_add ? e[5] : e[4], op2 = e[6], val = op1 + op2;
    o[0] = kb_from_u32(e[0]);
    o[1] = kb_from_u32(e[1]);
    uint32_t carry = 0;
    for (int k = 0; k < 4; k++) {
      o[2 + k] = kb_to_monty((val >> (8 * k)) & 0xFF);
      o[9 + k] = kb_to_monty((op1 >> (8 * k)) & 0xFF);
      o[13 + k] = kb_to_monty((op2 >> (8 * k)) & 0xFF);
      if (k < 3) {
        carry = (((op1 >> (8 * k)) & 0xFF) + ((op2 >> (8 * k)) & 0xFF) + carry) > 0xFF;
        o[6 + k] = carry ? KB_ONE : 0;
      }
    }
    o[17] = is_add ? KB_ONE : 0;
    o[18] = is_add ? 0 : KB_ONE;
  }
}

void ork_poseidon2_permute_canonical(uint32_t* s) {
  for (int i = 0; i < 16; i++) s[i] = kb_from_u32(s[i]);
  ork_poseidon2_permute(s);
  for (int i = 0; i < 16; i++) s[i] = kb_from_monty(s[i]);
}

/* PaddingFreeSponge<Perm,16,8,8>: overwrite-mode absorption, no padding; semantics restated at
 * crates/recursion/circuit/src/hash.rs:40-49 (and crates/zkvm/lib/src/poseidon2.rs:44-61). */
void ork_hash(const uint32_t* in, uint64_t n, uint32_t out[8]) {
  kb_t st[16];
  memset(st, 0, sizeof st);
  for (uint64_t off = 0; off < n; off += 8) {
    uint64_t len = n - off < 8 ? n - off : 8;
    for (uint64_t i = 0; i < len; i++) st[i] = in[off + i];
    ork_poseidon2_permute(st);
  }
  memcpy(out, st, 32);
}
/* TruncatedPermutation<Perm,2,8,16>: recursion/circuit/src/hash.rs:76-81 */
void ork_compress(const uint32_t l[8], const uint32_t r[8], uint32_t out[8]) {
  kb_t st[16];
  memcpy(st, l, 32);
  memcpy(st + 8, r, 32);
  ork_poseidon2_permute(st);
  memcpy(out, st, 32);
}
void ork_hash_rows(const uint32_t* mat, uint64_t h, uint64_t w, uint32_t* digests) {
#pragma omp parallel for schedule(static)
  for (uint64_t r = 0; r < h; r++) ork_hash(mat + r * w, w, digests + r * 8);
}

/* ------------------------------------------------------------------------------------------ */
/* DFT / coset LDE                                                                            */
/* ------------------------------------------------------------------------------------------ */
/* In-place decimation-in-frequency NTT of one contiguous column: natural order in, bit-reversed out.
 * tw[j] = w^j for j < n/2, w a primitive n-th root (forward: g_n, inverse: g_n^{-1}). */
static void dif_inplace(kb_t* a, uint64_t n, const kb_t* tw) {
  for (uint64_t m = n >> 1, step = 1; m >= 1; m >>= 1, step <<= 1) {
    for (uint64_t base = 0; base < n; base += 2 * m) {
      for (uint64_t j = 0; j < m; j++) {
        kb_t u = a[base + j], v = a[base + j + m];
        a[base + j] = kb_add(u, v);
        a[base + j + m] = kb_mul(kb_sub(u, v), tw[j * step]);
      }
    }
    if (m == 1) break;
  }
}
static kb_t* make_twiddles(uint64_t n, kb_t w) {
  uint64_t half = n > 1 ? n / 2 : 1;
  kb_t* t = (kb_t*)malloc(half * sizeof(kb_t));
  kb_t x = KB_ONE;
  for (uint64_t i = 0; i < half; i++) { t[i] = x; x = kb_mul(x, w); }
  return t;
}

void ork_dft_batch(const uint32_t* in, uint64_t h, uint64_t w, uint32_t* out) {
  unsigned n = log2_exact(h);
  kb_t* tw = make_twiddles(h, kb_two_adic_generator(n));
#pragma omp parallel
  {
    kb_t* col = (kb_t*)malloc(h * sizeof(kb_t));
#pragma omp for schedule(dynamic, 1)
    for (uint64_t c = 0; c < w; c++) {
      for (uint64_t r = 0; r < h; r++) col[r] = in[r * w + c];
      if (h > 1) dif_inplace(col, h, tw);
      for (uint64_t r = 0; r < h; r++) out[(uint64_t)bitrev32((uint32_t)r, n) * w + c] = col[r];
    }
    free(col);
  }
  free(tw);
}

/* Radix2DitParallel::coset_lde_batch(evals, added_bits, shift).bit_reverse_rows(), the call made
 * by TwoAdicFriPcs::commit (configured at crates/stark/src/kb31_poseidon2.rs:179-181).  Semantics
 * (SURVEY A.7): interpolate over the subgroup of order h, scale coefficient i by shift^i, zero-pad to
 * h << log_blowup coefficients, forward DFT; rows stored bit-reversed.  The result is mathematically
 * unique, so the butterfly order chosen here is irrelevant to parity.  Columns are processed in
 * blocks gathered into contiguous scratch so the CPU baseline is cache-friendly. */
void ork_coset_lde(const uint32_t* in, uint64_t h, uint64_t w, uint32_t log_blowup, uint32_t shift,
                   uint32_t* out) {
  unsigned n = log2_exact(h);
  uint64_t H = h << log_blowup;
  unsigned N = n + log_blowup;
  kb_t ginv = kb_inv(kb_two_adic_generator(n));
  kb_t* tw_inv = make_twiddles(h, ginv);
  kb_t* tw_fwd = make_twiddles(H, kb_two_adic_generator(N));
  /* scale[i] = shift^i / h */
  kb_t* scale = (kb_t*)malloc(h * sizeof(kb_t));
  {
    kb_t x = kb_inv(kb_from_u32((uint32_t)(h % KB_P)));
    for (uint64_t i = 0; i < h; i++) { scale[i] = x; x = kb_mul(x, shift); }
  }
  enum { CB = 8 };
#pragma omp parallel
  {
    kb_t* buf = (kb_t*)malloc((size_t)CB * H * sizeof(kb_t));
    kb_t* tmp = (kb_t*)malloc(h * sizeof(kb_t));
#pragma omp for schedule(dynamic, 1)
    for (uint64_t c0 = 0; c0 < w; c0 += CB) {
      uint64_t nc = w - c0 < CB ? w - c0 : CB;
      for (uint64_t r = 0; r < h; r++)
        for (uint64_t c = 0; c < nc; c++) buf[c * H + r] = in[r * w + c0 + c];
      for (uint64_t c = 0; c < nc; c++) {
        kb_t* col = buf + c * H;
        if (h > 1) dif_inplace(col, h, tw_inv);              /* coefficients, bit-reversed */
        for (uint64_t r = 0; r < h; r++) tmp[bitrev32((uint32_t)r, n)] = col[r];
        for (uint64_t i = 0; i < h; i++) col[i] = kb_mul(tmp[i], scale[i]);
        for (uint64_t i = h; i < H; i++) col[i] = 0;
        if (H > 1) dif_inplace(col, H, tw_fwd);              /* evaluations, bit-reversed = storage order */
      }
      for (uint64_t r = 0; r < H; r++)
        for (uint64_t c = 0; c < nc; c++) out[r * w + c0 + c] = buf[c * H + r];
    }
    free(buf);
    free(tmp);
  }
  free(scale);
  free(tw_inv);
  free(tw_fwd);
}

/* ------------------------------------------------------------------------------------------ */
/* MMCS: MerkleTreeMmcs<_, _, MyHash, MyCompress, 8> (kb31_poseidon2.rs:176-177)              */
/* layout pinned by verify_batch, recursion/circuit/src/fri.rs:363-405           */
/* ------------------------------------------------------------------------------------------ */
struct ork_tree {
  uint32_t n_mats;
  uint64_t* heights;
  uint64_t* widths;
  kb_t** mats;
  int owns;
  uint32_t log_max_height;
  kb_t** layers; /* log_max_height + 1 layers of 8-word digests */
  uint32_t* order; /* matrix indices sorted by height descending, stable */
};

static void sort_by_height_desc(uint32_t n, const uint64_t* heights, uint32_t* order) {
  for (uint32_t i = 0; i < n; i++) order[i] = i;
  for (uint32_t i = 1; i < n; i++) { /* stable insertion sort */
    uint32_t k = order[i];
    uint32_t j = i;
    while (j > 0 && heights[order[j - 1]] < heights[k]) { order[j] = order[j - 1]; j--; }
    order[j] = k;
  }
}

/* digest of the concatenation of row `r` over the matrices order[lo..hi) */
static void hash_concat_rows(const ork_tree* t, uint32_t lo, uint32_t hi, uint64_t r, uint32_t out[8]) {
  kb_t st[16];
  memset(st, 0, sizeof st);
  unsigned pos = 0;
  for (uint32_t k = lo; k < hi; k++) {
    uint32_t m = t->order[k];
    const kb_t* row = t->mats[m] + r * t->widths[m];
    for (uint64_t c = 0; c < t->widths[m]; c++) {
      st[pos++] = row[c];
      if (pos == 8) { ork_poseidon2_permute(st); pos = 0; }
    }
  }
  if (pos) ork_poseidon2_permute(st);
  memcpy(out, st, 32);
}

int32_t ork_mmcs_commit(uint32_t n_mats, const uint32_t* const* mats, const uint64_t* heights,
                        const uint64_t* widths, int32_t copy, uint32_t root[8], ork_tree** out) {
  if (n_mats == 0) return -1;
  for (uint32_t i = 0; i < n_mats; i++)
    if (heights[i] == 0 || (heights[i] & (heights[i] - 1))) return -2;
  ork_tree* t = (ork_tree*)calloc(1, sizeof *t);
  t->n_mats = n_mats;
  t->heights = (uint64_t*)malloc(n_mats * sizeof(uint64_t));
  t->widths = (uint64_t*)malloc(n_mats * sizeof(uint64_t));
  t->mats = (kb_t**)malloc(n_mats * sizeof(kb_t*));
  t->order = (uint32_t*)malloc(n_mats * sizeof(uint32_t));
  t->owns = copy;
  for (uint32_t i = 0; i < n_mats; i++) {
    t->heights[i] = heights[i];
    t->widths[i] = widths[i];
    if (copy) {
      size_t bytes = (size_t)heights[i] * widths[i] * 4;
      t->mats[i] = (kb_t*)malloc(bytes ? bytes : 4);
      memcpy(t->mats[i], mats[i], bytes);
    } else {
      t->mats[i] = (kb_t*)mats[i];
    }
  }
  sort_by_height_desc(n_mats, heights, t->order);
  uint64_t hmax = heights[t->order[0]];
  t->log_max_height = log2_exact(hmax);
  t->layers = (kb_t**)calloc(t->log_max_height + 1, sizeof(kb_t*));
  uint32_t next = 0; /* first not-yet-absorbed entry of order[] */
  uint32_t hi = next;
  while (hi < n_mats && heights[t->order[hi]] == hmax) hi++;
  t->layers[0] = (kb_t*)malloc((size_t)hmax * 32);
#pragma omp parallel for schedule(static)
  for (uint64_t r = 0; r < hmax; r++) hash_concat_rows(t, next, hi, r, t->layers[0] + r * 8);
  next = hi;
  for (uint32_t l = 1; l <= t->log_max_height; l++) {
    uint64_t len = hmax >> l;
    t->layers[l] = (kb_t*)malloc((size_t)len * 32);
    hi = next;
    while (hi < n_mats && heights[t->order[hi]] == len) hi++;
    const kb_t* prev = t->layers[l - 1];
    kb_t* cur = t->layers[l];
    uint32_t lo = next;
#pragma omp parallel for schedule(static)
    for (uint64_t i = 0; i < len; i++) {
      uint32_t d[8];
      ork_compress(prev + 16 * i, prev + 16 * i + 8, d);
      if (hi > lo) {
        uint32_t inj[8];
        hash_concat_rows(t, lo, hi, i, inj);
        ork_compress(d, inj, cur + 8 * i);
      } else {
        memcpy(cur + 8 * i, d, 32);
      }
    }
    next = hi;
  }
  memcpy(root, t->layers[t->log_max_height], 32);
  *out = t;
  return 0;
}

void ork_tree_free(ork_tree* t) {
  if (!t) return;
  if (t->owns)
    for (uint32_t i = 0; i < t->n_mats; i++) free(t->mats[i]);
  for (uint32_t l = 0; l <= t->log_max_height; l++) free(t->layers[l]);
  free(t->layers); free(t->mats); free(t->heights); free(t->widths); free(t->order);
  free(t);
}
uint32_t ork_tree_num_matrices(const ork_tree* t) { return t->n_mats; }
uint64_t ork_tree_height(const ork_tree* t, uint32_t i) { return t->heights[i]; }
uint64_t ork_tree_width(const ork_tree* t, uint32_t i) { return t->widths[i]; }
const uint32_t* ork_tree_matrix(const ork_tree* t, uint32_t i) { return t->mats[i]; }
uint32_t ork_tree_log_max_height(const ork_tree* t) { return t->log_max_height; }
const uint32_t* ork_tree_layer(const ork_tree* t, uint32_t l) { return t->layers[l]; }

void ork_tree_open(const ork_tree* t, uint64_t index, uint32_t* opened, uint32_t* proof) {
  uint64_t off = 0;
  for (uint32_t m = 0; m < t->n_mats; m++) { /* original matrix order */
    unsigned lh = log2_exact(t->heights[m]);
    uint64_t r = index >> (t->log_max_height - lh);
    memcpy(opened + off, t->mats[m] + r * t->widths[m], t->widths[m] * 4);
    off += t->widths[m];
  }
  for (uint32_t l = 0; l < t->log_max_height; l++)
    memcpy(proof + 8 * l, t->layers[l] + 8 * ((index >> l) ^ 1), 32);
}

/* verify_batch: recursion/circuit/src/fri.rs:363-405 */
int32_t ork_mmcs_verify(const uint32_t root[8], uint32_t n_mats, const uint64_t* heights,
                        const uint64_t* widths, uint64_t index, const uint32_t* opened,
                        const uint32_t* proof, uint32_t proof_len) {
  uint32_t* order = (uint32_t*)malloc(n_mats * sizeof(uint32_t));
  uint64_t* offs = (uint64_t*)malloc(n_mats * sizeof(uint64_t));
  uint64_t tot = 0;
  for (uint32_t i = 0; i < n_mats; i++) { offs[i] = tot; tot += widths[i]; }
  sort_by_height_desc(n_mats, heights, order);
  kb_t* buf = (kb_t*)malloc((tot ? tot : 1) * 4);
  uint64_t cur = heights[order[0]];
  uint32_t k = 0;
  uint64_t n = 0;
  while (k < n_mats && heights[order[k]] == cur) {
    memcpy(buf + n, opened + offs[order[k]], widths[order[k]] * 4);
    n += widths[order[k]];
    k++;
  }
  uint32_t node[8];
  ork_hash(buf, n, node);
  for (uint32_t l = 0; l < proof_len; l++) {
    uint32_t bit = (uint32_t)((index >> l) & 1);
    if (bit) ork_compress(proof + 8 * l, node, node); else ork_compress(node, proof + 8 * l, node);
    cur >>= 1;
    if (k < n_mats && heights[order[k]] == cur) {
      n = 0;
      while (k < n_mats && heights[order[k]] == cur) {
        memcpy(buf + n, opened + offs[order[k]], widths[order[k]] * 4);
        n += widths[order[k]];
        k++;
      }
      uint32_t inj[8];
      ork_hash(buf, n, inj);
      ork_compress(node, inj, node);
    }
  }
  int ok = memcmp(node, root, 32) == 0 && k == n_mats;
  free(order); free(offs); free(buf);
  return ok;
}

/* TwoAdicFriPcs::commit (call sites crates/stark/src/prover.rs:277,403,497). */
int32_t ork_pcs_commit(uint32_t n_mats, const uint32_t* const* mats, const uint64_t* heights,
                       const uint64_t* widths, const uint32_t* domain_shifts, uint32_t log_blowup,
                       uint32_t root[8], ork_tree** out) {
  kb_t** ldes = (kb_t**)malloc(n_mats * sizeof(kb_t*));
  uint64_t* lh = (uint64_t*)malloc(n_mats * sizeof(uint64_t));
  for (uint32_t i = 0; i < n_mats; i++) {
    lh[i] = heights[i] << log_blowup;
    size_t bytes = (size_t)lh[i] * widths[i] * 4;
    ldes[i] = (kb_t*)malloc(bytes ? bytes : 4);
    kb_t shift = kb_mul(kb_generator(), kb_inv(domain_shifts[i])); /* GENERATOR / domain.shift */
    ork_coset_lde(mats[i], heights[i], widths[i], log_blowup, shift, ldes[i]);
  }
  int32_t rc = ork_mmcs_commit(n_mats, (const uint32_t* const*)ldes, lh, widths, 0, root, out);
  if (rc == 0) (*out)->owns = 1; /* the tree now owns the LDE buffers */
  else for (uint32_t i = 0; i < n_mats; i++) free(ldes[i]);
  free(ldes);
  free(lh);
  return rc;
}
