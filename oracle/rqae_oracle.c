/*
 * Plain-C CPU oracle for the RQAE residual-quantization hot path.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): used by tests/, by
 * __graft_entry__.smoke() and by bench.py's cpu_baseline leg as the checker.  It is
 * never linked into, imported by or called from the shipped CUDA path.
 *
 * It restates the algorithm of the reference rqae/model.py (harish-kamath/rqae):
 *   forward  rqae/model.py:199-230   per layer: z = W_in r + b_in (:211), cosine-similarity
 *            argmax over the layer codebook (:187-193, :180-182), straight-through value
 *            c' = z + (c - z) (:218-220), o = W_out c' + b_out (:221), r -= o (:223),
 *            q += o (:224)
 *   decode   rqae/model.py:232-252   codewords from codebook[0], out-projections summed in
 *            ascending layer order
 *
 * The reference computes with PyTorch/MKL kernels whose K=2304 summation order is not
 * defined by the reference, so the fp32 entry point takes the order as a
 * parameter: order_nt = 0 sums sequentially over d; order_nt = NT > 0 uses NT interleaved
 * partial sums (lane t takes d = t, t+NT, ...), combines each aligned group of 32 lanes
 * with a pairwise tree (tree = 0: lanes paired by strides 16,8,4,2,1; tree = 1: strides
 * 1,2,16,8,4), then adds the groups sequentially (gtree = 0) or as a pairwise tree
 * ((g0+g1)+(g2+g3))+... (gtree = 1) -- the shapes a 32-wide SIMT machine produces.  seg_blocks > 0 cuts
 * the d axis into segments of seg_blocks * NT elements that are summed separately in that order and then
 * added in segment order (a thread-block cluster whose CTAs own slices of the d axis).  Where the reference's fp32 arithmetic IS defined (the
 * K=4 contractions and the norm; probed against torch 2.11 CPU) this file
 * uses exactly that arithmetic:
 *   cos  = fmaf(z3,c3, fmaf(z2,c2, fmaf(z1,c1, z0*c0)))        (matmul, model.py:190)
 *   nrm  = sqrtf(((z0*z0 + z1*z1) + z2*z2) + z3*z3)            (norm,   model.py:188)
 *   o    = fmaf(c3,w3, fmaf(c2,w2, fmaf(c1,w1, c0*w0))) + b    (linear, model.py:221)
 * fold_bias = 1 selects the one-rounding-fewer variant o = fmaf(c3,w3,...fmaf(c0,w0,b)).
 *
 * The fp64 entry point evaluates the same recurrence in double precision and reports, per
 * (token, layer), the cos-sim margin between the winner and the best competitor whose
 * codeword differs from the winner's: the quantity the parity protocol uses to classify a
 * disagreement as a near-tie.
 *
 * Build: gcc -O2 -ffp-contract=off -mavx2 -mfma -fopenmp -shared -fPIC (see oracle/build.py).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define RQO_OK 0
#define RQO_EINVAL 1
#define RQO_ENOMEM 2

int rqo_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* ---- fp32 pieces -------------------------------------------------------------------- */

/* z[k] = sum_d w_in[k][d] * r[d] in the requested order (bias added by the caller). */
static void inproj_f32(const float *w_in /*[cd][D]*/, const float *r, int D, int cd, int nt, int tree, int gtree,
                       int seg_blocks, float *scratch /*[nt]*/, float *z) {
  static const int strides[2][5] = {{16, 8, 4, 2, 1}, {1, 2, 16, 8, 4}};
  if (nt > 0 && seg_blocks > 0 && (long)seg_blocks * nt < D) {
    /* D-split order (the cluster variant of the CUDA kernel): the d axis is cut into segments of seg_blocks * nt
     * elements, each segment is summed on its own in the order below, and the segment sums are added in order */
    const int seg = seg_blocks * nt;
    for (int k = 0; k < cd; k++) z[k] = 0.0f;
    for (int s0 = 0, si = 0; s0 < D; s0 += seg, si++) {
      const int len = (D - s0 < seg) ? (D - s0) : seg;
      float zs[16];
      /* recursion depth 1: the segment itself is not split again */
      float *wseg = (float *)malloc(sizeof(float) * (size_t)cd * (size_t)len);
      for (int k = 0; k < cd; k++) memcpy(wseg + (size_t)k * len, w_in + (size_t)k * D + s0, sizeof(float) * (size_t)len);
      inproj_f32(wseg, r + s0, len, cd, nt, tree, gtree, 0, scratch, zs);
      free(wseg);
      for (int k = 0; k < cd; k++) z[k] = (si == 0) ? zs[k] : z[k] + zs[k];
    }
    return;
  }
  for (int k = 0; k < cd; k++) {
    const float *w = w_in + (size_t)k * D;
    if (nt <= 0) {
      float acc = 0.0f;
      for (int d = 0; d < D; d++) acc = fmaf(w[d], r[d], acc);
      z[k] = acc;
      continue;
    }
    for (int t = 0; t < nt; t++) scratch[t] = 0.0f;
    int d0 = 0;
    for (; d0 + nt <= D; d0 += nt)
      for (int t = 0; t < nt; t++) scratch[t] = fmaf(w[d0 + t], r[d0 + t], scratch[t]);
    for (int t = 0; d0 + t < D; t++) scratch[t] = fmaf(w[d0 + t], r[d0 + t], scratch[t]);
    float total = 0.0f;
    float gsum[64];
    int ng = 0;
    for (int g = 0; g < nt; g += 32) {
      float v[32];
      for (int i = 0; i < 32; i++) v[i] = (g + i < nt) ? scratch[g + i] : 0.0f;
      int done = 0; /* lane bits already folded: only lanes with those bits clear stay live */
      for (int si = 0; si < 5; si++) {
        const int s = strides[tree][si];
        for (int i = 0; i < 32; i++)
          if ((i & done) == 0 && (i & s) == 0) v[i] = v[i] + v[i | s];
        done |= s;
      }
      total = (g == 0) ? v[0] : total + v[0];
      if (ng < 64) gsum[ng++] = v[0];
    }
    if (gtree) {
      for (int step = 1; step < ng; step *= 2)
        for (int i = 0; i + step < ng; i += 2 * step) gsum[i] = gsum[i] + gsum[i + step];
      total = gsum[0];
    }
    z[k] = total;
  }
}

static int argmax_cos_f32(const float *zn, const float *cb /*[K][cd]*/, int K, int cd) {
  int best = 0;
  float bv = 0.0f;
  for (int k = 0; k < K; k++) {
    const float *c = cb + (size_t)k * cd;
    float v = zn[0] * c[0];
    for (int i = 1; i < cd; i++) v = fmaf(zn[i], c[i], v);
    if (v != v) return k; /* torch.argmax: the first NaN wins */
    if (k == 0 || v > bv) { bv = v; best = k; }
  }
  return best;
}

int rqo_forward_f32(const float *w_in, const float *b_in, const float *w_out, const float *b_out,
                    const float *codebook, int cb_shared, int nq_run, int D, int cd, int K,
                    const float *x, long n_tokens, int order_nt, int tree, int gtree, int seg_blocks, int fold_bias,
                    int recon_mode, const int32_t *teacher, int32_t *codes, float *q_out) {
  if (D <= 0 || cd <= 0 || cd > 16 || K <= 0 || nq_run < 0 || n_tokens < 0 || order_nt < 0 || seg_blocks < 0) return RQO_EINVAL;
  if (tree < 0 || tree > 1 || (gtree && order_nt > 64 * 32)) return RQO_EINVAL;
  int err = 0;
#pragma omp parallel
  {
    float *r = (float *)malloc(sizeof(float) * (size_t)D);
    float *q = (float *)malloc(sizeof(float) * (size_t)D);
    float *scr = (float *)malloc(sizeof(float) * (size_t)(order_nt > 0 ? order_nt : 1));
    if (!r || !q || !scr) {
#pragma omp atomic write
      err = RQO_ENOMEM;
    } else {
#pragma omp for schedule(dynamic, 1)
      for (long t = 0; t < n_tokens; t++) {
        const float *xt = x + (size_t)t * D;
        memcpy(r, xt, sizeof(float) * (size_t)D);
        for (int l = 0; l < nq_run; l++) {
          float z[16], zn[16], c2[16];
          inproj_f32(w_in + (size_t)l * cd * D, r, D, cd, order_nt, tree, gtree, seg_blocks, scr, z);
          const float *bi = b_in + (size_t)l * cd;
          for (int k = 0; k < cd; k++) z[k] = z[k] + bi[k];
          float ss = z[0] * z[0];
          for (int k = 1; k < cd; k++) ss = ss + z[k] * z[k];
          float nrm = sqrtf(ss);
          for (int k = 0; k < cd; k++) zn[k] = z[k] / nrm;
          const float *cb = codebook + (cb_shared ? 0 : (size_t)l * K * cd);
          int idx = argmax_cos_f32(zn, cb, K, cd);
          codes[(size_t)t * nq_run + l] = idx;
          if (teacher) idx = teacher[(size_t)t * nq_run + l];
          const float *c = cb + (size_t)idx * cd;
          for (int k = 0; k < cd; k++) c2[k] = z[k] + (c[k] - z[k]); /* STE, model.py:218-220 */
          const float *wo = w_out + (size_t)l * D * cd;
          const float *bo = b_out + (size_t)l * D;
          for (int d = 0; d < D; d++) {
            const float *w = wo + (size_t)d * cd;
            float o;
            if (fold_bias) {
              o = bo[d];
              for (int k = 0; k < cd; k++) o = fmaf(c2[k], w[k], o);
            } else {
              o = c2[0] * w[0];
              for (int k = 1; k < cd; k++) o = fmaf(c2[k], w[k], o);
              o = o + bo[d];
            }
            r[d] = r[d] - o;
            if (recon_mode == 0) q[d] = (l == 0) ? (0.0f + o) : (q[d] + o);
          }
        }
        if (q_out) {
          float *qt = q_out + (size_t)t * D;
          if (recon_mode == 0) {
            if (nq_run == 0) memset(qt, 0, sizeof(float) * (size_t)D);
            else memcpy(qt, q, sizeof(float) * (size_t)D);
          } else {
            for (int d = 0; d < D; d++) qt[d] = xt[d] - r[d];
          }
        }
      }
    }
    free(r); free(q); free(scr);
  }
  return err;
}

/* ---- fp64 with margins -------------------------------------------------------------- */

int rqo_forward_f64(const float *w_in, const float *b_in, const float *w_out, const float *b_out,
                    const float *codebook, int cb_shared, int nq_run, int D, int cd, int K,
                    const float *x, long n_tokens, const int32_t *teacher, int32_t *codes,
                    double *q_out, float *margins) {
  if (D <= 0 || cd <= 0 || cd > 16 || K <= 0 || nq_run < 0 || n_tokens < 0) return RQO_EINVAL;
  int err = 0;
#pragma omp parallel
  {
    double *r = (double *)malloc(sizeof(double) * (size_t)D);
    double *q = (double *)malloc(sizeof(double) * (size_t)D);
    double *cos = (double *)malloc(sizeof(double) * (size_t)K);
    if (!r || !q || !cos) {
#pragma omp atomic write
      err = RQO_ENOMEM;
    } else {
#pragma omp for schedule(dynamic, 1)
      for (long t = 0; t < n_tokens; t++) {
        for (int d = 0; d < D; d++) { r[d] = x[(size_t)t * D + d]; q[d] = 0.0; }
        for (int l = 0; l < nq_run; l++) {
          double z[16], zn[16], c2[16];
          const float *wi = w_in + (size_t)l * cd * D;
          for (int k = 0; k < cd; k++) {
            double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
            const float *w = wi + (size_t)k * D;
            int d = 0;
            for (; d + 4 <= D; d += 4) {
              a0 += (double)w[d] * r[d]; a1 += (double)w[d + 1] * r[d + 1];
              a2 += (double)w[d + 2] * r[d + 2]; a3 += (double)w[d + 3] * r[d + 3];
            }
            for (; d < D; d++) a0 += (double)w[d] * r[d];
            z[k] = ((a0 + a1) + (a2 + a3)) + (double)b_in[(size_t)l * cd + k];
          }
          double ss = 0;
          for (int k = 0; k < cd; k++) ss += z[k] * z[k];
          double nrm = sqrt(ss);
          for (int k = 0; k < cd; k++) zn[k] = z[k] / nrm;
          const float *cb = codebook + (cb_shared ? 0 : (size_t)l * K * cd);
          int best = 0, saw_nan = 0;
          for (int k = 0; k < K; k++) {
            double v = 0;
            for (int i = 0; i < cd; i++) v += zn[i] * (double)cb[(size_t)k * cd + i];
            cos[k] = v;
            if (v != v) { if (!saw_nan) { best = k; saw_nan = 1; } }
            else if (!saw_nan && (k == 0 || v > cos[best])) best = k;
          }
          codes[(size_t)t * nq_run + l] = best;
          if (margins) {
            double second = -INFINITY;
            const float *cw = cb + (size_t)best * cd;
            for (int k = 0; k < K; k++) {
              if (memcmp(cb + (size_t)k * cd, cw, sizeof(float) * (size_t)cd) == 0) continue;
              int same = 1; /* +0.0 / -0.0 rows count as equal values */
              for (int i = 0; i < cd; i++) same &= (cb[(size_t)k * cd + i] == cw[i]);
              if (same) continue;
              if (cos[k] > second) second = cos[k];
            }
            margins[(size_t)t * nq_run + l] = saw_nan ? 0.0f : (float)(cos[best] - second);
          }
          int idx = teacher ? teacher[(size_t)t * nq_run + l] : best;
          for (int k = 0; k < cd; k++) { double c = cb[(size_t)idx * cd + k]; c2[k] = z[k] + (c - z[k]); }
          const float *wo = w_out + (size_t)l * D * cd;
          const float *bo = b_out + (size_t)l * D;
          for (int d = 0; d < D; d++) {
            double o = bo[d];
            for (int k = 0; k < cd; k++) o += c2[k] * (double)wo[(size_t)d * cd + k];
            r[d] -= o; q[d] += o;
          }
        }
        if (q_out) memcpy(q_out + (size_t)t * D, q, sizeof(double) * (size_t)D);
      }
    }
    free(r); free(q); free(cos);
  }
  return err;
}

/* ---- decode (rqae/model.py:232-252), reference fp32 arithmetic ------------------------- */

int rqo_decode_f32(const float *w_out, const float *b_out, const float *codebook0, int nq, int D,
                   int cd, int K, const int32_t *codes /*[n][nq]*/, const float *cv /*nullable [n][nq][cd]*/,
                   const uint8_t *layer_mask /*nullable [nq]*/, long n_tokens, float *q_out) {
  if (D <= 0 || cd <= 0 || cd > 16 || nq < 0 || n_tokens < 0 || (!codes && !cv)) return RQO_EINVAL;
#pragma omp parallel for schedule(static)
  for (long t = 0; t < n_tokens; t++) {
    float *q = q_out + (size_t)t * D;
    int first = 1;
    for (int l = 0; l < nq; l++) {
      if (layer_mask && !layer_mask[l]) continue;
      const float *c = cv ? cv + ((size_t)t * nq + l) * cd
                          : codebook0 + (size_t)codes[(size_t)t * nq + l] * cd;
      const float *wo = w_out + (size_t)l * D * cd;
      const float *bo = b_out + (size_t)l * D;
      for (int d = 0; d < D; d++) {
        const float *w = wo + (size_t)d * cd;
        float o = c[0] * w[0];
        for (int k = 1; k < cd; k++) o = fmaf(c[k], w[k], o);
        o = o + bo[d];
        q[d] = first ? o : (q[d] + o);
      }
      first = 0;
    }
    if (first) memset(q, 0, sizeof(float) * (size_t)D); /* reference returns None; caller handles */
  }
  (void)K;
  return RQO_OK;
}
