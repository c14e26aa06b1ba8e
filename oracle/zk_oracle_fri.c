/* oracle/zk_oracle_fri.c -- TEST INFRASTRUCTURE (CPU oracle), not product code.  See zk_oracle.h.
 * Part 2: duplex challenger, TwoAdicFriPcs::open (opening reduction + FRI prover) and the in-repo
 * verifier transliterated.
 *
 * Prover side: the algorithm lives in ProjectZKM/Plonky3 @ faa24ca (p3-fri two_adic_pcs.rs / prover.rs,
 * p3-challenger duplex_challenger.rs), absent from /root/reference; it is restated here from SURVEY
 * A.6/A.10 and is pinned by ork_pcs_verify below, a line-by-line transliteration of the reference's own
 * verifier crates/recursion/circuit/src/fri.rs:34-405 (+ challenger.rs:90-233): a proof that it accepts
 * has, by uniqueness of field results, the same opened values, commitments and query openings as the
 * reference prover given the same pow_witness.  PARITY STATUS: unpinned against stored vectors (the
 * reference has none for FRI); structurally pinned by the verifier.
 */
#include "zk_oracle.h"
#include "kb31.h"
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------------------------------ */
/* DuplexChallenger<Val, Perm, 16, 8>  (crates/stark/src/kb31_poseidon2.rs:180;               */
/* semantics: crates/recursion/circuit/src/challenger.rs:90-114,201-233)                      */
/* ------------------------------------------------------------------------------------------ */
void ork_ch_init(ork_challenger* c) { memset(c, 0, sizeof *c); }

static void ch_duplex(ork_challenger* c) { /* challenger.rs:221-232 */
  for (uint32_t i = 0; i < c->n_in; i++) c->state[i] = c->in[i];
  c->n_in = 0;
  ork_poseidon2_permute(c->state);
  for (int i = 0; i < 8; i++) c->out[i] = c->state[i];
  c->n_out = 8;
}
void ork_ch_observe(ork_challenger* c, const uint32_t* v, uint32_t n) { /* challenger.rs:90-98 */
  for (uint32_t i = 0; i < n; i++) {
    c->n_out = 0;
    c->in[c->n_in++] = v[i];
    if (c->n_in == 8) ch_duplex(c);
  }
}
uint32_t ork_ch_sample(ork_challenger* c) { /* challenger.rs:100-106: pops from the END */
  if (c->n_in != 0 || c->n_out == 0) ch_duplex(c);
  return c->out[--c->n_out];
}
void ork_ch_sample_ext(ork_challenger* c, uint32_t out[4]) { /* challenger.rs:201-207 */
  for (int i = 0; i < 4; i++) out[i] = ork_ch_sample(c);
}
uint32_t ork_ch_sample_bits(ork_challenger* c, uint32_t bits) { /* challenger.rs:108-114 */
  uint32_t v = kb_from_monty(ork_ch_sample(c));
  return bits >= 32 ? v : (v & ((1u << bits) - 1));
}
int32_t ork_ch_check_witness(ork_challenger* c, uint32_t bits, uint32_t witness) { /* challenger.rs:209-219 */
  ork_ch_observe(c, &witness, 1);
  return ork_ch_sample_bits(c, bits) == 0;
}
/* grind: Plonky3 searches 0..p in parallel and takes ANY hit (non-deterministic); the oracle takes the
 * smallest canonical witness so that runs are reproducible.  Mutates c like check_witness. */
uint32_t ork_ch_grind(ork_challenger* c, uint32_t bits) {
  for (uint32_t w = 0; w < KB_P; w++) {
    ork_challenger t = *c;
    uint32_t wm = kb_to_monty(w);
    if (ork_ch_check_witness(&t, bits, wm)) {
      *c = t;
      return wm;
    }
  }
  return 0xffffffffu;
}

/* ------------------------------------------------------------------------------------------ */
/* helpers                                                                                    */
/* ------------------------------------------------------------------------------------------ */
static kb4_t ld4(const uint32_t* p) { kb4_t r; memcpy(r.c, p, 16); return r; }
static void st4(uint32_t* p, kb4_t v) { memcpy(p, v.c, 16); }

/* x_r = shift * g_L^{bitrev_L(r)} for r < 2^L */
static kb_t* coset_points_bitrev(unsigned L, kb_t shift) {
  uint64_t H = 1ull << L;
  kb_t* xs = (kb_t*)malloc(H * sizeof(kb_t));
  kb_t g = kb_two_adic_generator(L);
  kb_t x = shift;
  for (uint64_t j = 0; j < H; j++) { xs[bitrev32((uint32_t)j, L)] = x; x = kb_mul(x, g); }
  return xs;
}

/* Evaluate every column of `low` (N rows = evaluations over shift*H_N in bit-reversed row order) at the
 * extension point z: barycentric formula  p(z) = (z^N - s^N)/(N s^N) * sum_i x_i y_i / (z - x_i). */
static void interpolate_coset_bitrev(const kb_t* low, uint64_t N, uint64_t w, kb_t shift, kb4_t z, kb4_t* ys) {
  unsigned n = log2_exact(N);
  kb_t* xs = coset_points_bitrev(n, shift);
  kb4_t* wts = (kb4_t*)malloc(N * sizeof(kb4_t));
#pragma omp parallel for schedule(static)
  for (uint64_t i = 0; i < N; i++) {
    kb4_t d = kb4_sub_base(z, xs[i]);
    wts[i] = kb4_mul_base(kb4_inv(d), xs[i]);
  }
  kb_t sN = kb_pow(shift, N);
  kb4_t zN = kb4_pow(z, N);
  kb4_t scale = kb4_mul_base(kb4_sub_base(zN, sN), kb_inv(kb_mul(kb_from_u32((uint32_t)(N % KB_P)), sN)));
#pragma omp parallel for schedule(static)
  for (uint64_t c = 0; c < w; c++) {
    kb4_t acc = kb4_zero();
    for (uint64_t i = 0; i < N; i++) acc = kb4_add(acc, kb4_mul_base(wts[i], low[i * w + c]));
    ys[c] = kb4_mul(acc, scale);
  }
  free(xs);
  free(wts);
}

/* ------------------------------------------------------------------------------------------ */
/* flat proof layout shared with libzkgpu (include/zkgpu.h, "proof layout")                  */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
  uint32_t n_rounds;
  uint32_t n_mats_total;
  uint32_t log_max;     /* log of the tallest LDE over all rounds */
  uint32_t n_layers;    /* log_max - log_blowup */
  uint64_t opened_words;
  uint64_t query_words; /* words per query */
  uint64_t total_words;
} proof_shape;

static proof_shape shape_of(uint32_t n_rounds, const uint32_t* n_mats, const uint64_t* lde_heights,
                            const uint32_t* widths, const uint32_t* n_points, uint32_t log_blowup,
                            uint32_t num_queries) {
  proof_shape s;
  memset(&s, 0, sizeof s);
  s.n_rounds = n_rounds;
  uint32_t k = 0;
  uint64_t per_query = 0;
  for (uint32_t r = 0; r < n_rounds; r++) {
    uint32_t lmax = 0;
    uint64_t sumw = 0;
    for (uint32_t m = 0; m < n_mats[r]; m++, k++) {
      unsigned L = log2_exact(lde_heights[k]);
      if (L > lmax) lmax = L;
      sumw += widths[k];
      s.opened_words += (uint64_t)n_points[k] * widths[k] * 4;
    }
    if (lmax > s.log_max) s.log_max = lmax;
    per_query += sumw + (uint64_t)lmax * 8;
  }
  s.n_mats_total = k;
  s.n_layers = s.log_max - log_blowup;
  for (uint32_t i = 0; i < s.n_layers; i++) per_query += 4 + (uint64_t)(s.log_max - i - 1) * 8;
  s.query_words = per_query;
  s.total_words = s.opened_words + (uint64_t)s.n_layers * 8 + 4 + 1 + (uint64_t)num_queries * per_query;
  return s;
}

uint64_t ork_pcs_proof_words(uint32_t n_rounds, const uint32_t* n_mats, const uint64_t* lde_heights,
                             const uint32_t* widths, const uint32_t* n_points, uint32_t log_blowup,
                             uint32_t num_queries) {
  return shape_of(n_rounds, n_mats, lde_heights, widths, n_points, log_blowup, num_queries).total_words;
}

/* ------------------------------------------------------------------------------------------ */
/* TwoAdicFriPcs::open  (call site crates/stark/src/prover.rs:546-556)           */
/* ------------------------------------------------------------------------------------------ */
int32_t ork_pcs_open(uint32_t n_rounds, const ork_tree* const* trees, const uint32_t* n_points,
                     const uint32_t* points, uint32_t log_blowup, uint32_t num_queries, uint32_t pow_bits,
                     ork_challenger* ch, int64_t inject_witness, uint32_t* proof, uint64_t proof_cap) {
  /* shapes */
  uint32_t total = 0;
  for (uint32_t r = 0; r < n_rounds; r++) total += ork_tree_num_matrices(trees[r]);
  uint32_t* n_mats = (uint32_t*)malloc(n_rounds * sizeof(uint32_t));
  uint64_t* hs = (uint64_t*)malloc(total * sizeof(uint64_t));
  uint32_t* ws = (uint32_t*)malloc(total * sizeof(uint32_t));
  for (uint32_t r = 0, k = 0; r < n_rounds; r++) {
    n_mats[r] = ork_tree_num_matrices(trees[r]);
    for (uint32_t m = 0; m < n_mats[r]; m++, k++) {
      hs[k] = ork_tree_height(trees[r], m);
      ws[k] = (uint32_t)ork_tree_width(trees[r], m);
    }
  }
  proof_shape S = shape_of(n_rounds, n_mats, hs, ws, n_points, log_blowup, num_queries);
  if (proof_cap < S.total_words) { free(n_mats); free(hs); free(ws); return -1; }
  uint32_t* out = proof;

  kb4_t alpha; { uint32_t a[4]; ork_ch_sample_ext(ch, a); alpha = ld4(a); }

  kb4_t* ro[33];
  uint64_t num_reduced[33];
  for (int i = 0; i < 33; i++) { ro[i] = NULL; num_reduced[i] = 0; }
  kb_t gen = kb_generator();

  const uint32_t* pt = points;
  for (uint32_t r = 0, k = 0; r < n_rounds; r++) {
    for (uint32_t m = 0; m < n_mats[r]; m++, k++) {
      uint64_t H = hs[k], w = ws[k];
      unsigned L = log2_exact(H);
      uint64_t N = H >> log_blowup;
      const kb_t* mat = ork_tree_matrix(trees[r], m);
      if (!ro[L]) ro[L] = (kb4_t*)calloc(H, sizeof(kb4_t));
      kb_t* xs = coset_points_bitrev(L, gen);
      /* alpha powers for this matrix */
      kb4_t* apow = (kb4_t*)malloc((w ? w : 1) * sizeof(kb4_t));
      { kb4_t a = kb4_one(); for (uint64_t j = 0; j < w; j++) { apow[j] = a; a = kb4_mul(a, alpha); } }
      /* row reductions sum_j alpha^j m[r][j] (dot_ext_powers), shared by all points of the matrix */
      kb4_t* rowred = (kb4_t*)malloc(H * sizeof(kb4_t));
#pragma omp parallel for schedule(static)
      for (uint64_t i = 0; i < H; i++) {
        kb4_t acc = kb4_zero();
        const kb_t* row = mat + i * w;
        for (uint64_t j = 0; j < w; j++) acc = kb4_add(acc, kb4_mul_base(apow[j], row[j]));
        rowred[i] = acc;
      }
      kb4_t* ys = (kb4_t*)malloc((w ? w : 1) * sizeof(kb4_t));
      for (uint32_t p = 0; p < n_points[k]; p++, pt += 4) {
        kb4_t z = ld4(pt);
        interpolate_coset_bitrev(mat, N, w, gen, z, ys);
        for (uint64_t j = 0; j < w; j++, out += 4) st4(out, ys[j]);
        kb4_t red_ys = kb4_zero();
        for (uint64_t j = 0; j < w; j++) red_ys = kb4_add(red_ys, kb4_mul(apow[j], ys[j]));
        kb4_t aoff = kb4_pow(alpha, num_reduced[L]);
        kb4_t* roL = ro[L];
#pragma omp parallel for schedule(static)
        for (uint64_t i = 0; i < H; i++) {
          kb4_t inv_den = kb4_inv(kb4_neg(kb4_sub_base(z, xs[i]))); /* 1 / (x - z) */
          roL[i] = kb4_add(roL[i], kb4_mul(aoff, kb4_mul(kb4_sub(rowred[i], red_ys), inv_den)));
        }
        num_reduced[L] += w;
      }
      free(ys); free(rowred); free(apow); free(xs);
    }
  }

  /* ---- FRI commit phase (p3-fri prover::commit_phase; mirror fri.rs:257-358) ---- */
  unsigned Lmax = S.log_max;
  uint64_t len = 1ull << Lmax;
  kb4_t* folded = (kb4_t*)malloc(len * sizeof(kb4_t));
  memcpy(folded, ro[Lmax], len * sizeof(kb4_t));
  ork_tree** layer_trees = (ork_tree**)calloc(S.n_layers ? S.n_layers : 1, sizeof(ork_tree*));
  uint32_t* commits = out; out += (uint64_t)S.n_layers * 8;
  kb_t half = kb_inv(kb_to_monty(2));
  for (uint32_t i = 0; i < S.n_layers; i++) {
    unsigned Li = Lmax - i; /* log of current length */
    uint64_t hh = len >> 1;
    const uint32_t* mp = (const uint32_t*)folded;
    uint64_t hgt = hh, wid = 8;
    uint32_t root[8];
    ork_mmcs_commit(1, &mp, &hgt, &wid, 1, root, &layer_trees[i]);
    memcpy(commits + 8 * i, root, 32);
    ork_ch_observe(ch, root, 8);
    kb4_t beta; { uint32_t b[4]; ork_ch_sample_ext(ch, b); beta = ld4(b); }
    /* f'[k] = (1/2 + beta/(2 x_k)) e0 + (1/2 - beta/(2 x_k)) e1,  x_k = g_{Li}^{bitrev_{Li-1}(k)} */
    kb_t gi = kb_inv(kb_two_adic_generator(Li));
    kb_t* inv_x = (kb_t*)malloc(hh * sizeof(kb_t));
    { kb_t x = KB_ONE; for (uint64_t j = 0; j < hh; j++) { inv_x[bitrev32((uint32_t)j, Li - 1)] = x; x = kb_mul(x, gi); } }
    kb4_t* nxt = (kb4_t*)malloc(hh * sizeof(kb4_t));
    kb4_t half_beta = kb4_mul_base(beta, half);
#pragma omp parallel for schedule(static)
    for (uint64_t k2 = 0; k2 < hh; k2++) {
      kb4_t pw = kb4_mul_base(half_beta, inv_x[k2]);
      kb4_t a = kb4_add_base(pw, half);
      kb4_t b = kb4_neg(kb4_sub_base(pw, half)); /* 1/2 - pw */
      nxt[k2] = kb4_add(kb4_mul(a, folded[2 * k2]), kb4_mul(b, folded[2 * k2 + 1]));
    }
    free(inv_x);
    free(folded);
    folded = nxt;
    len = hh;
    if (ro[Li - 1]) {
      kb4_t b2 = kb4_sqr(beta);
      for (uint64_t k2 = 0; k2 < len; k2++) folded[k2] = kb4_add(folded[k2], kb4_mul(b2, ro[Li - 1][k2]));
    }
  }
  int32_t rc = 0;
  for (uint64_t k2 = 1; k2 < len; k2++)
    if (!kb4_eq(folded[k2], folded[0])) rc = -2; /* not a constant: inputs were not low degree */
  kb4_t final_poly = folded[0];
  st4(out, final_poly); out += 4;
  ork_ch_observe(ch, final_poly.c, 4);
  uint32_t witness;
  if (inject_witness >= 0) {
    witness = (uint32_t)inject_witness;
    if (!ork_ch_check_witness(ch, pow_bits, witness)) rc = rc ? rc : -3;
  } else {
    witness = ork_ch_grind(ch, pow_bits);
  }
  *out++ = witness;

  /* ---- query phase (prover::prove + answer_query) ---- */
  for (uint32_t q = 0; q < num_queries; q++) {
    uint64_t index = ork_ch_sample_bits(ch, Lmax);
    for (uint32_t r = 0; r < n_rounds; r++) {
      unsigned lr = ork_tree_log_max_height(trees[r]);
      uint64_t ridx = index >> (Lmax - lr);
      uint64_t sumw = 0;
      for (uint32_t m = 0; m < n_mats[r]; m++) sumw += ork_tree_width(trees[r], m);
      ork_tree_open(trees[r], ridx, out, out + sumw);
      out += sumw + (uint64_t)lr * 8;
    }
    for (uint32_t i = 0; i < S.n_layers; i++) {
      uint64_t index_i = index >> i;
      uint64_t pair = index_i >> 1;
      uint32_t row[8];
      uint32_t lh = ork_tree_log_max_height(layer_trees[i]);
      ork_tree_open(layer_trees[i], pair, row, out + 4);
      memcpy(out, row + 4 * ((index_i ^ 1) & 1), 16); /* sibling value */
      out += 4 + (uint64_t)lh * 8;
    }
  }
  if ((uint64_t)(out - proof) != S.total_words) rc = rc ? rc : -4;

  for (uint32_t i = 0; i < S.n_layers; i++) ork_tree_free(layer_trees[i]);
  free(layer_trees);
  free(folded);
  for (int i = 0; i < 33; i++) free(ro[i]);
  free(n_mats); free(hs); free(ws);
  return rc;
}

/* ------------------------------------------------------------------------------------------ */
/* verifier: crates/recursion/circuit/src/fri.rs:34-361 transliterated                         */
/* ------------------------------------------------------------------------------------------ */
int32_t ork_pcs_verify(uint32_t n_rounds, const uint32_t* roots, const uint32_t* n_mats,
                       const uint64_t* lde_heights, const uint32_t* widths, const uint32_t* n_points,
                       const uint32_t* points, uint32_t log_blowup, uint32_t num_queries, uint32_t pow_bits,
                       ork_challenger* ch, const uint32_t* proof, uint64_t proof_words) {
  proof_shape S = shape_of(n_rounds, n_mats, lde_heights, widths, n_points, log_blowup, num_queries);
  if (proof_words != S.total_words) return -1;
  const uint32_t* opened = proof;
  const uint32_t* commits = proof + S.opened_words;
  const uint32_t* final_poly_p = commits + (uint64_t)S.n_layers * 8;
  uint32_t witness = final_poly_p[4];
  const uint32_t* qp = final_poly_p + 5;

  /* fri.rs:78 */
  kb4_t alpha; { uint32_t a[4]; ork_ch_sample_ext(ch, a); alpha = ld4(a); }
  /* verify_shape_and_sample_challenges, fri.rs:34-69 */
  kb4_t* betas = (kb4_t*)malloc((S.n_layers ? S.n_layers : 1) * sizeof(kb4_t));
  for (uint32_t i = 0; i < S.n_layers; i++) {
    ork_ch_observe(ch, commits + 8 * i, 8);
    uint32_t b[4]; ork_ch_sample_ext(ch, b); betas[i] = ld4(b);
  }
  kb4_t final_poly = ld4(final_poly_p);
  ork_ch_observe(ch, final_poly_p, 4);
  if (!ork_ch_check_witness(ch, pow_bits, witness)) { free(betas); return -2; }
  unsigned Lmax = S.n_layers + log_blowup; /* log_global_max_height, fri.rs:62 */
  if (Lmax != S.log_max) { free(betas); return -3; }

  int32_t rc = 1;
  kb_t gen = kb_generator();
  for (uint32_t q = 0; q < num_queries && rc == 1; q++) {
    uint64_t index = ork_ch_sample_bits(ch, Lmax);
    kb4_t ro[33];
    uint64_t log_height_pow[33];
    for (int i = 0; i < 33; i++) { ro[i] = kb4_zero(); log_height_pow[i] = 0; }
    const uint32_t* op = opened;
    const uint32_t* pt = points;
    for (uint32_t r = 0, k = 0; r < n_rounds && rc == 1; r++) {
      /* fri.rs:106-128: batch dims, reduced index, verify_batch */
      unsigned lr = 0;
      uint64_t sumw = 0;
      for (uint32_t m = 0; m < n_mats[r]; m++) {
        unsigned L = log2_exact(lde_heights[k + m]);
        if (L > lr) lr = L;
        sumw += widths[k + m];
      }
      uint64_t ridx = index >> (Lmax - lr);
      const uint32_t* rows = qp;
      const uint32_t* path = qp + sumw;
      qp += sumw + (uint64_t)lr * 8;
      uint64_t* w64 = (uint64_t*)malloc(n_mats[r] * sizeof(uint64_t));
      for (uint32_t m = 0; m < n_mats[r]; m++) w64[m] = widths[k + m];
      if (!ork_mmcs_verify(roots + 8 * r, n_mats[r], lde_heights + k, w64, ridx, rows, path, lr)) rc = -10 - (int32_t)r;
      free(w64);
      /* fri.rs:130-203 */
      const uint32_t* row = rows;
      for (uint32_t m = 0; m < n_mats[r]; m++, k++) {
        unsigned L = log2_exact(lde_heights[k]);
        uint64_t rix = index >> (Lmax - L);
        kb_t x = kb_mul(gen, kb_pow(kb_two_adic_generator(L), bitrev32((uint32_t)rix, L)));
        for (uint32_t p = 0; p < n_points[k]; p++, pt += 4) {
          kb4_t z = ld4(pt);
          kb4_t acc = kb4_zero();
          kb4_t ap = kb4_pow(alpha, log_height_pow[L]);
          for (uint32_t j = 0; j < widths[k]; j++, op += 4) {
            kb4_t p_at_z = ld4(op);
            acc = kb4_add(acc, kb4_mul(ap, kb4_sub_base(p_at_z, row[j]))); /* alpha^pow (p(z) - p(x)) */
            ap = kb4_mul(ap, alpha);
          }
          log_height_pow[L] += widths[k];
          ro[L] = kb4_add(ro[L], kb4_mul(acc, kb4_inv(kb4_sub_base(z, x)))); /* acc / (z - x) */
        }
        row += widths[k];
      }
    }
    if (rc != 1) break;
    /* fri.rs:206 */
    if (!kb4_eq(ro[log_blowup], kb4_zero())) { rc = -4; break; }
    /* verify_query, fri.rs:243-361 */
    kb4_t folded = ro[Lmax];
    kb_t x = kb_pow(kb_two_adic_generator(Lmax), bitrev32((uint32_t)index, Lmax));
    kb_t g1 = kb_two_adic_generator(1);
    for (uint32_t i = 0; i < S.n_layers; i++) {
      unsigned log_folded_height = Lmax - 1 - i;
      uint32_t bit = (uint32_t)((index >> i) & 1);
      uint64_t pair = index >> (i + 1);
      kb4_t sib = ld4(qp);
      const uint32_t* path = qp + 4;
      qp += 4 + (uint64_t)log_folded_height * 8;
      kb4_t e0 = bit ? sib : folded, e1 = bit ? folded : sib;
      uint32_t leaf[8];
      st4(leaf, e0); st4(leaf + 4, e1);
      uint64_t h1 = 1ull << log_folded_height, w1 = 8;
      if (!ork_mmcs_verify(commits + 8 * i, 1, &h1, &w1, pair, leaf, path, log_folded_height)) { rc = -100 - (int32_t)i; break; }
      kb_t xs_new = kb_mul(x, g1);
      kb_t x0 = bit ? xs_new : x, x1 = bit ? x : xs_new;
      kb4_t t3 = kb4_mul_base(kb4_sub(e1, e0), kb_inv(kb_sub(x1, x0)));
      kb4_t t5 = kb4_mul(kb4_sub_base(betas[i], x0), t3); /* (beta - x0) * t3 */
      folded = kb4_add(kb4_add(e0, t5), kb4_mul(kb4_sqr(betas[i]), ro[log_folded_height]));
      x = kb_mul(x, x);
    }
    if (rc != 1) break;
    if (!kb4_eq(folded, final_poly)) rc = -5;
  }
  free(betas);
  return rc;
}
